// TEST INFRASTRUCTURE ONLY.  Stand-in for <google/protobuf/repeated_field.h>: RepeatedField / RepeatedPtrField with the
// accessor names the reference uses, backed by std::vector (pointer-stable elements for RepeatedPtrField).
#ifndef VSO_REF_SHIM_PROTOBUF_REPEATED_FIELD_H_
#define VSO_REF_SHIM_PROTOBUF_REPEATED_FIELD_H_
#include <algorithm>
#include <iterator>
#include <memory>
#include <vector>
namespace google {
namespace protobuf {
typedef int int32;
typedef unsigned int uint32;
template <class T> class RepeatedField {
 public:
  typedef typename std::vector<T>::iterator iterator;
  typedef typename std::vector<T>::const_iterator const_iterator;
  typedef T value_type;
  int size() const { return (int)v_.size(); }
  bool empty() const { return v_.empty(); }
  const T& Get(int i) const { return v_[i]; }
  T* Mutable(int i) { return &v_[i]; }
  void Set(int i, const T& x) { v_[i] = x; }
  void Add(const T& x) { v_.push_back(x); }
  T* Add() { v_.push_back(T()); return &v_.back(); }
  void Clear() { v_.clear(); }
  void Reserve(int n) { v_.reserve(n); }
  void RemoveLast() { v_.pop_back(); }
  void Truncate(int n) { v_.resize(n); }
  void Swap(RepeatedField* o) { v_.swap(o->v_); }
  void CopyFrom(const RepeatedField& o) { v_ = o.v_; }
  void MergeFrom(const RepeatedField& o) { v_.insert(v_.end(), o.v_.begin(), o.v_.end()); }
  const T* data() const { return v_.data(); }
  T* mutable_data() { return v_.data(); }
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  iterator erase(const_iterator a, const_iterator b) { return v_.erase(a, b); }
  const T& operator[](int i) const { return v_[i]; }
  T& operator[](int i) { return v_[i]; }
 private:
  std::vector<T> v_;
};
template <class T> class RepeatedFieldBackInsertIterator {
 public:
  typedef std::output_iterator_tag iterator_category;
  typedef T value_type;
  typedef void difference_type;
  typedef void pointer;
  typedef void reference;
  explicit RepeatedFieldBackInsertIterator(RepeatedField<T>* f) : f_(f) {}
  RepeatedFieldBackInsertIterator& operator=(const T& v) { f_->Add(v); return *this; }
  RepeatedFieldBackInsertIterator& operator*() { return *this; }
  RepeatedFieldBackInsertIterator& operator++() { return *this; }
  RepeatedFieldBackInsertIterator& operator++(int) { return *this; }
 private:
  RepeatedField<T>* f_;
};
template <class T> RepeatedFieldBackInsertIterator<T> RepeatedFieldBackInserter(RepeatedField<T>* f) { return RepeatedFieldBackInsertIterator<T>(f); }
template <class T> class RepeatedPtrField {
  typedef std::vector<std::unique_ptr<T>> Store;
  template <class Base, class Ref, class Ptr> struct It {
    typedef std::random_access_iterator_tag iterator_category;
    typedef T value_type;
    typedef std::ptrdiff_t difference_type;
    typedef Ptr pointer;
    typedef Ref reference;
    Base b;
    It() {}
    It(Base b_) : b(b_) {}
    template <class B2, class R2, class P2> It(const It<B2, R2, P2>& o) : b(o.b) {}
    Ref operator*() const { return **b; }
    Ptr operator->() const { return b->get(); }
    Ref operator[](difference_type n) const { return *b[n]; }
    It& operator++() { ++b; return *this; }
    It operator++(int) { It t(*this); ++b; return t; }
    It& operator--() { --b; return *this; }
    It operator--(int) { It t(*this); --b; return t; }
    It& operator+=(difference_type n) { b += n; return *this; }
    It& operator-=(difference_type n) { b -= n; return *this; }
    It operator+(difference_type n) const { return It(b + n); }
    It operator-(difference_type n) const { return It(b - n); }
    difference_type operator-(const It& o) const { return b - o.b; }
    bool operator==(const It& o) const { return b == o.b; }
    bool operator!=(const It& o) const { return b != o.b; }
    bool operator<(const It& o) const { return b < o.b; }
    bool operator>(const It& o) const { return b > o.b; }
    bool operator<=(const It& o) const { return b <= o.b; }
    bool operator>=(const It& o) const { return b >= o.b; }
  };
 public:
  typedef It<typename Store::iterator, T&, T*> iterator;
  typedef It<typename Store::const_iterator, const T&, const T*> const_iterator;
  typedef T value_type;
  // iterators over the element POINTERS (protobuf's pointer_begin / pointer_end, used to sort in place)
  struct PtrIt {
    typedef std::random_access_iterator_tag iterator_category;
    typedef T* value_type;
    typedef std::ptrdiff_t difference_type;
    typedef T** pointer;
    typedef T*& reference;
    typename Store::iterator b;
    PtrIt() {}
    PtrIt(typename Store::iterator b_) : b(b_) {}
    // a unique_ptr<T> slot is layout-compatible with a T* slot
    T*& operator*() const { return *reinterpret_cast<T**>(&*b); }
    T*& operator[](difference_type n) const { return *reinterpret_cast<T**>(&b[n]); }
    PtrIt& operator++() { ++b; return *this; }
    PtrIt operator++(int) { PtrIt t(*this); ++b; return t; }
    PtrIt& operator--() { --b; return *this; }
    PtrIt operator--(int) { PtrIt t(*this); --b; return t; }
    PtrIt& operator+=(difference_type n) { b += n; return *this; }
    PtrIt& operator-=(difference_type n) { b -= n; return *this; }
    PtrIt operator+(difference_type n) const { return PtrIt(b + n); }
    PtrIt operator-(difference_type n) const { return PtrIt(b - n); }
    difference_type operator-(const PtrIt& o) const { return b - o.b; }
    bool operator==(const PtrIt& o) const { return b == o.b; }
    bool operator!=(const PtrIt& o) const { return b != o.b; }
    bool operator<(const PtrIt& o) const { return b < o.b; }
    bool operator>(const PtrIt& o) const { return b > o.b; }
    bool operator<=(const PtrIt& o) const { return b <= o.b; }
    bool operator>=(const PtrIt& o) const { return b >= o.b; }
  };
  typedef PtrIt pointer_iterator;
  RepeatedPtrField() {}
  RepeatedPtrField(const RepeatedPtrField& o) { CopyFrom(o); }
  RepeatedPtrField& operator=(const RepeatedPtrField& o) { if (this != &o) CopyFrom(o); return *this; }
  int size() const { return (int)v_.size(); }
  bool empty() const { return v_.empty(); }
  const T& Get(int i) const { return *v_[i]; }
  T* Mutable(int i) { return v_[i].get(); }
  T* Add() { v_.emplace_back(new T()); return v_.back().get(); }
  void Clear() { v_.clear(); }
  void Reserve(int) {}
  void RemoveLast() { v_.pop_back(); }
  void Swap(RepeatedPtrField* o) { v_.swap(o->v_); }
  void SwapElements(int a, int b) { v_[a].swap(v_[b]); }
  void CopyFrom(const RepeatedPtrField& o) { v_.clear(); MergeFrom(o); }
  void MergeFrom(const RepeatedPtrField& o) { for (const auto& p : o.v_) v_.emplace_back(new T(*p)); }
  void DeleteSubrange(int start, int num) { v_.erase(v_.begin() + start, v_.begin() + start + num); }
  iterator begin() { return iterator(v_.begin()); }
  iterator end() { return iterator(v_.end()); }
  const_iterator begin() const { return const_iterator(v_.begin()); }
  const_iterator end() const { return const_iterator(v_.end()); }
  pointer_iterator pointer_begin() { return PtrIt(v_.begin()); }
  pointer_iterator pointer_end() { return PtrIt(v_.end()); }
  const T& operator[](int i) const { return *v_[i]; }
  T& operator[](int i) { return *v_[i]; }
 private:
  Store v_;
};
}  // namespace protobuf
}  // namespace google
#endif
