// TEST INFRASTRUCTURE: stand-in for boost::mutex / scoped_lock.
#pragma once
#include <mutex>
namespace boost {
class mutex {
 public:
  class scoped_lock {
   public:
    explicit scoped_lock(mutex& m) : l_(m.m_) {}
   private:
    std::unique_lock<std::mutex> l_;
  };
 private:
  std::mutex m_;
};
}  // namespace boost
