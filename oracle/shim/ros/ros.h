// Minimal stand-in for roscpp: just enough surface for monte_carlo.cpp to compile. No behaviour.
#pragma once
#include <boost_shim.h>
#include <cstdio>
#include <string>
namespace mclshim { struct Log { int last_injected = -1; }; inline Log& log() { static Log l; return l; } 
  inline void info(const char* fmt, int v) { if (std::string(fmt).find("New Injected Particles") != std::string::npos) log().last_injected = v; }
  template <class... A> inline void info(const char*, A...) {}
  // the last message of each type handed to any ros::Publisher (so the harness can read what the reference published)
  template <class M> inline M& last_published() { static M m; return m; }
  template <class M> inline int& publish_count() { static int n = 0; return n; }
}
#define ROS_INFO(...) ::mclshim::info(__VA_ARGS__)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
namespace ros {
struct Time { static Time now() { return Time(); } };
struct Duration { Duration(double = 0) {} void sleep() {} };
struct Rate { Rate(double) {} void sleep() {} };
struct TimerEvent {};
struct Timer { void stop() {} };
struct Publisher { template <class M> void publish(const M& m) const { ::mclshim::last_published<M>() = m; ++::mclshim::publish_count<M>(); } };
struct Subscriber {};
struct ServiceClient { template <class S> bool call(S&) { return true; } };
struct NodeHandle {
    template <class F> Subscriber subscribe(const char*, int, F) { return Subscriber(); }
    template <class S> ServiceClient serviceClient(const char*) { return ServiceClient(); }
    template <class M> Publisher advertise(const char*, int) { return Publisher(); }
    template <class F> Timer createTimer(Duration, F) { return Timer(); }
};
inline void init(int&, char**, const char*) {}
inline bool ok() { return false; }
inline void spin() {}
inline void spinOnce() {}
inline void shutdown() {}
}  // namespace ros
