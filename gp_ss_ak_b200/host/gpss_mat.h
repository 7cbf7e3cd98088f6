// Minimal column-major dense matrix used by the host classes when Armadillo is not installed.
//
// north_star keeps Armadillo only for I/O and the symmetric-standardisation preprocessing; this image has no
// Armadillo, so the host code is written against the SUBSET of the arma::mat interface declared here
// (same member names and semantics: column-major storage, operator()(i,j), linear operator()(i)/[i],
// resize() that preserves elements and zero-fills, csv_ascii save/load).  With -DGPSS_USE_ARMADILLO and
// Armadillo on the include path the very same host sources compile against the real arma::mat.
#pragma once

#if defined(GPSS_USE_ARMADILLO)
#include <armadillo>
using arma::mat;
using arma::csv_ascii;
using arma::accu;
#else

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

enum gpss_file_type { csv_ascii };

class mat {
 public:
  typedef unsigned long long uword;
  uword n_rows, n_cols, n_elem;

  mat() : n_rows(0), n_cols(0), n_elem(0) {}
  mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), v_(r * c) {}   // like arma: contents unspecified (here 0)

  double& operator()(uword i, uword j) { return v_[j * n_rows + i]; }
  const double& operator()(uword i, uword j) const { return v_[j * n_rows + i]; }
  double& operator()(uword i) { return v_[i]; }
  const double& operator()(uword i) const { return v_[i]; }
  double& operator[](uword i) { return v_[i]; }
  const double& operator[](uword i) const { return v_[i]; }
  double* memptr() { return v_.data(); }
  const double* memptr() const { return v_.data(); }
  double* colptr(uword j) { return v_.data() + j * n_rows; }
  const double* colptr(uword j) const { return v_.data() + j * n_rows; }

  mat& zeros() { std::fill(v_.begin(), v_.end(), 0.0); return *this; }
  mat& zeros(uword r, uword c) { set_size(r, c); return zeros(); }
  mat& ones() { std::fill(v_.begin(), v_.end(), 1.0); return *this; }
  mat& fill(double x) { std::fill(v_.begin(), v_.end(), x); return *this; }
  void set_size(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; v_.assign(n_elem, 0.0); }
  // arma::Mat::resize: keeps the overlapping block in place, new elements are zero
  void resize(uword r, uword c)
  {
    if (r == n_rows && c == n_cols) return;
    std::vector<double> nv(r * c, 0.0);
    const uword rr = std::min(r, n_rows), cc = std::min(c, n_cols);
    for (uword j = 0; j < cc; j++)
      for (uword i = 0; i < rr; i++) nv[j * r + i] = v_[j * n_rows + i];
    v_.swap(nv);
    n_rows = r; n_cols = c; n_elem = r * c;
  }
  double min() const { return *std::min_element(v_.begin(), v_.end()); }
  double max() const { return *std::max_element(v_.begin(), v_.end()); }
  bool has_nan() const
  {
    for (double x : v_) if (x != x) return true;
    return false;
  }

  // csv_ascii as arma writes it: scientific notation, comma separated, one row per line
  bool save(const std::string& name, gpss_file_type) const
  {
    std::ofstream f(name.c_str());
    if (!f) return false;
    f.setf(std::ios::scientific);
    f.precision(16);
    for (uword i = 0; i < n_rows; i++) {
      for (uword j = 0; j < n_cols; j++) {
        f << (*this)(i, j);
        if (j + 1 < n_cols) f << ',';
      }
      f << '\n';
    }
    return f.good();
  }
  bool load(const std::string& name, gpss_file_type)
  {
    std::ifstream f(name.c_str());
    if (!f) return false;
    std::vector<std::vector<double> > rows;
    std::string line;
    while (std::getline(f, line)) {
      if (line.empty()) continue;
      std::vector<double> r;
      std::stringstream ss(line);
      std::string tok;
      while (std::getline(ss, tok, ',')) r.push_back(std::atof(tok.c_str()));
      rows.push_back(r);
    }
    const uword r = rows.size(), c = r ? rows[0].size() : 0;
    set_size(r, c);
    for (uword i = 0; i < r; i++)
      for (uword j = 0; j < c && j < rows[i].size(); j++) (*this)(i, j) = rows[i][j];
    return true;
  }

 private:
  std::vector<double> v_;
};

inline double accu(const mat& A)
{
  double s = 0.0;
  for (mat::uword i = 0; i < A.n_elem; i++) s += A[i];
  return s;
}

#endif  // GPSS_USE_ARMADILLO
