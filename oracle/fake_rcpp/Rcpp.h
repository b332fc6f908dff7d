// Rcpp.h -- TEST INFRASTRUCTURE ONLY: the handful of Rcpp names /root/reference/utilities/cpiu.cpp uses, so that the
// UNMODIFIED file compiles into oracle/_ref/libcpiu_ref.so (oracle/Makefile `cpiu`).  Nothing here ships.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace Rcpp {

template <typename T>
class Vec {
 public:
  Vec() {}
  explicit Vec(long n) : v_(static_cast<size_t>(n), T(0)) {}       // Rcpp vectors start zero-filled
  Vec(const T* p, long n) : v_(p, p + n) {}
  long size() const { return static_cast<long>(v_.size()); }
  T& operator[](long i) { return v_[static_cast<size_t>(i)]; }
  const T& operator[](long i) const { return v_[static_cast<size_t>(i)]; }
  const T* data() const { return v_.data(); }
 private:
  std::vector<T> v_;
};
typedef Vec<double> NumericVector;
typedef Vec<int> IntegerVector;

class List {
 public:
  void set(const std::string& k, const NumericVector& v) { m_[k] = v; }
  const NumericVector& operator[](const char* k) const { return m_.at(k); }
 private:
  std::map<std::string, NumericVector> m_;
};

template <typename T> T as(const NumericVector& v) { return v; }
inline void stop(const char* msg) { throw std::runtime_error(msg); }

}  // namespace Rcpp
