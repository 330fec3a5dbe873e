// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Stand-in for boost::container::flat_map (Boost is not installed here): a
// sorted std::vector of pairs offering the handful of members the reference's
// src/rcpp_cx_report.cpp uses (reserve, begin/end, clear, try_emplace with a
// hint).  It is only a container: iteration is in ascending key order, like
// Boost's, so results are identical.  The hinted insert is O(1) when the key
// follows the hint (the reference's access pattern), which if anything
// flatters the CPU baseline relative to real Boost.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

namespace boost { namespace container {

template <class K, class V>
class flat_map {
 public:
  typedef std::pair<K, V> value_type;
  typedef typename std::vector<value_type>::iterator iterator;
  typedef typename std::vector<value_type>::const_iterator const_iterator;

  flat_map() {}
  void reserve(size_t n) { d_.reserve(n); }
  iterator begin() { return d_.begin(); }
  iterator end() { return d_.end(); }
  void clear() { d_.clear(); }
  size_t size() const { return d_.size(); }

  iterator try_emplace(iterator hint, const K& k, const V& v) {
    iterator e = d_.end();
    if (hint != e) {
      if (hint->first == k) return hint;
      if (hint->first < k) {
        iterator nx = hint + 1;
        if (nx == e) { d_.emplace_back(k, v); return d_.end() - 1; }
        if (nx->first == k) return nx;
        if (k < nx->first) return d_.insert(nx, value_type(k, v));
        iterator lb = lower_bound_(nx, e, k);
        if (lb != e && lb->first == k) return lb;
        return d_.insert(lb, value_type(k, v));
      }
      iterator lb = lower_bound_(d_.begin(), hint, k);
      if (lb != hint && lb->first == k) return lb;
      return d_.insert(lb, value_type(k, v));
    }
    if (d_.empty() || d_.back().first < k) { d_.emplace_back(k, v); return d_.end() - 1; }
    iterator lb = lower_bound_(d_.begin(), e, k);
    if (lb != e && lb->first == k) return lb;
    return d_.insert(lb, value_type(k, v));
  }

 private:
  static iterator lower_bound_(iterator a, iterator b, const K& k) {
    return std::lower_bound(a, b, k, [](const value_type& x, const K& key) { return x.first < key; });
  }
  std::vector<value_type> d_;
};

}}  // namespace boost::container
