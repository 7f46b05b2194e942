// TEST INFRASTRUCTURE: stand-in for the slice of boost::posix_time video_framework/video_unit.{h,cpp} uses.
#pragma once
#include <chrono>
namespace boost { namespace posix_time {
struct time_duration {
  long long us = 0;
  long long total_microseconds() const { return us; }
};
inline time_duration microseconds(long long n) { time_duration d; d.us = n; return d; }
struct ptime {
  std::chrono::system_clock::time_point tp;
  ptime operator+(const time_duration& d) const { ptime r; r.tp = tp + std::chrono::microseconds(d.us); return r; }
};
struct microsec_clock { static ptime local_time() { ptime p; p.tp = std::chrono::system_clock::now(); return p; } };
struct time_period {
  ptime a, b;
  time_period(const ptime& a_, const ptime& b_) : a(a_), b(b_) {}
  time_duration length() const { time_duration d; d.us = std::chrono::duration_cast<std::chrono::microseconds>(b.tp - a.tp).count(); return d; }
};
} }  // namespace boost::posix_time
