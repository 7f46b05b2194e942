// TEST INFRASTRUCTURE: stand-in for boost::circular_buffer (boost is not installed in this image) with the members
// video_framework/video_unit.{h,cpp} use for its rate statistics.  Not product code.
#pragma once
#include <cstddef>
#include <deque>
namespace boost {
template <class T> class circular_buffer {
 public:
  typedef typename std::deque<T>::const_iterator const_iterator;
  circular_buffer() {}
  explicit circular_buffer(std::size_t cap) : cap_(cap) {}
  void set_capacity(std::size_t cap) { cap_ = cap; while (d_.size() > cap_) d_.pop_front(); }
  std::size_t capacity() const { return cap_; }
  std::size_t size() const { return d_.size(); }
  bool empty() const { return d_.empty(); }
  void push_back(const T& v) { if (!cap_) return; if (d_.size() == cap_) d_.pop_front(); d_.push_back(v); }
  const T& back() const { return d_.back(); }
  const_iterator begin() const { return d_.begin(); }
  const_iterator end() const { return d_.end(); }
 private:
  std::size_t cap_ = 0;
  std::deque<T> d_;
};
}  // namespace boost
