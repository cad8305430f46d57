// shim_prelude.h — force-included (-include) before the reference translation unit.
// TEST INFRASTRUCTURE: lets /root/reference/pink_fundamentals/src/monte_carlo.cpp compile unmodified
// without ROS, and makes its random_device-seeded engines reproducible:
//   std::random_device         -> a queue of seeds the harness controls
//   std::default_random_engine -> minstd_rand0 that registers itself so the harness can reseed the
//                                 function-local statics at MC:411 and MC:452
#pragma once
#include <math.h>      // tf/LinearMath/Scalar.h includes <math.h>; this is what makes cos(float) pick cosf (Q6)
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <vector>
#include <deque>
#include <ctime>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <utility>

namespace mclshim {
struct SeedQueue {
    std::deque<unsigned> q;
    unsigned fallback = 12345u;
    std::vector<unsigned> history;
    unsigned next() { unsigned v = fallback; if (!q.empty()) { v = q.front(); q.pop_front(); } history.push_back(v); return v; }
};
inline SeedQueue& seeds() { static SeedQueue s; return s; }
struct fixed_random_device {
    typedef unsigned result_type;
    unsigned operator()() { return seeds().next(); }
};
struct registered_engine;
inline std::vector<registered_engine*>& engines() { static std::vector<registered_engine*> v; return v; }
struct registered_engine : std::minstd_rand0 {
    explicit registered_engine(unsigned s) : std::minstd_rand0(s) { engines().push_back(this); }
};
}  // namespace mclshim
namespace std {
typedef ::mclshim::fixed_random_device mclshim_random_device;
typedef ::mclshim::registered_engine mclshim_default_engine;
}
// std::time(nullptr) seeds srand at MC:808 (k-means): the harness sets the value it returns
namespace mclshim { inline time_t& fake_time() { static time_t t = 1; return t; } }
namespace std { inline time_t mclshim_time(time_t*) { return ::mclshim::fake_time(); } }
#define time mclshim_time
#define random_device mclshim_random_device
#define default_random_engine mclshim_default_engine
