// TEST INFRASTRUCTURE: stand-in for the one boost::thread member video_unit.cpp uses (sleep until a time).
#pragma once
#include <thread>
#include <boost/thread/mutex.hpp>
#include <boost/thread/thread_time.hpp>
namespace boost {
class thread {
 public:
  static void sleep(const posix_time::ptime& until) { std::this_thread::sleep_until(until.tp); }
};
}  // namespace boost
