// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for <Rcpp.h>, just large enough for the reference's three
// hot-path translation units (src/rcpp_threshold_reads.cpp,
// src/rcpp_get_xm_beta.cpp, src/rcpp_cx_report.cpp) and the BED matcher
// (src/rcpp_match_target.cpp) and the lMHL report
// (src/rcpp_mhl_report.cpp, 8f rank 1) to compile UNMODIFIED from
// /root/reference/src (see oracle/Makefile).  The surface mirrored here is the
// one listed in SURVEY.md Appendix A.  Nothing from the reference is copied:
// this is an independent shim of the Rcpp API those files touch.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef void* SEXP;
#ifndef NA_INTEGER
#define NA_INTEGER INT_MIN  /* R's integer NA (R_ext/Arith.h: R_NaInt = INT_MIN) */
#endif
#ifndef NA_REAL
#define NA_REAL (std::nan("1954"))  /* a NaN stands in for R's NA_real_; the driver reports NA as NaN */
#endif

namespace Rcpp {

// Attribute proxy: remembers a pointer when one is stored, swallows the rest
// (factor class / levels are irrelevant to the integer payload we compare).
struct AttrProxy {
  SEXP* slot = nullptr;
  SEXP held = nullptr;
  AttrProxy() {}
  explicit AttrProxy(SEXP* s) : slot(s), held(s ? *s : nullptr) {}
  AttrProxy& operator=(const AttrProxy& o) { if (slot) *slot = o.held; held = o.held; return *this; }
  template <class T> AttrProxy& operator=(const T&) { return *this; }
  AttrProxy& operator=(SEXP p) { if (slot) *slot = p; held = p; return *this; }
  operator SEXP() const { return held; }
};

template <class T>
class XPtr {
 public:
  explicit XPtr(SEXP p) : p_(static_cast<T*>(p)) {}
  XPtr(T* p, bool) : p_(p) {}
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  operator SEXP() const { return static_cast<SEXP>(p_); }
 private:
  T* p_;
};

class IntegerVector {
 public:
  IntegerVector() : v_(std::make_shared<std::vector<int>>()) {}
  explicit IntegerVector(std::shared_ptr<std::vector<int>> v) : v_(std::move(v)) {}
  long size() const { return static_cast<long>(v_->size()); }
  int& operator[](long i) { return (*v_)[static_cast<size_t>(i)]; }
  const int& operator[](long i) const { return (*v_)[static_cast<size_t>(i)]; }
  AttrProxy attr(const char*) { return AttrProxy(); }
  static bool is_na(int x) { return x == NA_INTEGER; }
  std::shared_ptr<std::vector<int>> shared() const { return v_; }
 private:
  std::shared_ptr<std::vector<int>> v_;
};
typedef IntegerVector LogicalVector;

// column-major double matrix, zero-initialised (R's NumericMatrix(nrow, ncol))
class NumericMatrix {
 public:
  NumericMatrix() : r_(0), c_(0), v_(std::make_shared<std::vector<double>>()) {}
  NumericMatrix(long nrow, long ncol) : r_(nrow), c_(ncol), v_(std::make_shared<std::vector<double>>(static_cast<size_t>(nrow * ncol), 0.0)) {}
  double& operator()(long i, long j) { return (*v_)[static_cast<size_t>(j * r_ + i)]; }
  long nrow() const { return r_; }
  long ncol() const { return c_; }
  std::shared_ptr<std::vector<double>> shared() const { return v_; }
 private:
  long r_, c_;
  std::shared_ptr<std::vector<double>> v_;
};

class CharacterVector {
 public:
  template <class... A> static CharacterVector create(A...) { return CharacterVector(); }
};

struct NamedColumn { std::string name; std::vector<int> values; };
struct NamedDoubleColumn { std::string name; std::vector<double> values; };
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  NamedColumn operator=(const std::vector<int>& v) const { return NamedColumn{name, v}; }
  NamedDoubleColumn operator=(const std::vector<double>& v) const { return NamedDoubleColumn{name, v}; }
};
inline void stop(const char*) { std::abort(); }

class DataFrame {
 public:
  DataFrame() {}
  IntegerVector operator[](const char* name) {
    auto it = cols_.find(name);
    if (it == cols_.end()) it = cols_.emplace(name, IntegerVector()).first;
    return it->second;
  }
  IntegerVector operator[](const std::string& name) { return (*this)[name.c_str()]; }
  AttrProxy attr(const char* name) { return AttrProxy(&attrs_[name]); }
  void set_column(const char* name, std::shared_ptr<std::vector<int>> v) {
    cols_[name] = IntegerVector(std::move(v));
  }
  template <class... A> static DataFrame create(A&&... a) {
    DataFrame df;
    (df.add(std::forward<A>(a)), ...);
    return df;
  }
 private:
  void add(NamedColumn&& c) {
    cols_[c.name] = IntegerVector(std::make_shared<std::vector<int>>(std::move(c.values)));
  }
  void add(NamedDoubleColumn&& c) { dcols_[c.name] = std::make_shared<std::vector<double>>(std::move(c.values)); }
 public:
  std::shared_ptr<std::vector<double>> double_column(const char* name) { return dcols_[name]; }
 private:
  std::map<std::string, std::shared_ptr<std::vector<double>>> dcols_;
  std::map<std::string, IntegerVector> cols_;
  std::map<std::string, SEXP> attrs_;
};

inline void checkUserInterrupt() {}

}  // namespace Rcpp
