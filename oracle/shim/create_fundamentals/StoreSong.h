#pragma once
#include <create_fundamentals/DiffDrive.h>
