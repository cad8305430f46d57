#pragma once
#include <pink_fundamentals/PID_drive.h>
