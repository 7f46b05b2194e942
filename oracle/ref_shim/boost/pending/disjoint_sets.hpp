// TEST INFRASTRUCTURE ONLY.  Stand-in for <boost/pending/disjoint_sets.hpp>: union by rank with path compression over
// caller-provided rank / parent arrays (the interface ConnectedComponents uses, segmentation_util.cpp:1039-1072).
// Only the partition matters to the caller (components are emitted in first-seen order), not the representatives.
#ifndef VSO_REF_SHIM_BOOST_DISJOINT_SETS_HPP_
#define VSO_REF_SHIM_BOOST_DISJOINT_SETS_HPP_
namespace boost {
template <class RankPA, class ParentPA> class disjoint_sets {
 public:
  disjoint_sets(RankPA r, ParentPA p) : rank_(r), parent_(p) {}
  template <class E> void make_set(E x) { parent_[x] = x; rank_[x] = 0; }
  template <class E> E find_set(E x) {
    E r = x;
    while (parent_[r] != r) r = parent_[r];
    while (parent_[x] != r) { E n = parent_[x]; parent_[x] = r; x = n; }
    return r;
  }
  template <class E> void union_set(E x, E y) { link(find_set(x), find_set(y)); }
  template <class E> void link(E x, E y) {
    if (x == y) return;
    if (rank_[x] > rank_[y]) parent_[y] = x;
    else { parent_[x] = y; if (rank_[x] == rank_[y]) ++rank_[y]; }
  }
  template <class It> int count_sets(It first, It last) {
    int n = 0;
    for (; first != last; ++first) if (parent_[*first] == *first) ++n;
    return n;
  }
 private:
  RankPA rank_;
  ParentPA parent_;
};
}  // namespace boost
#endif
