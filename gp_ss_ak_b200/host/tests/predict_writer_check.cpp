// CPU-only check of Control::writePredictTable (tests/test_host_cpu.py): the bytes written with 1 worker, with 7 workers and by the
// reference's own loop (value by value into an ofstream, gp_ss_ak.cpp:470-481) must be identical; prints the three wall times.
//   predict_writer_check ROWS OUTDIR
#include "../Control.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>

using namespace arma;

static std::string slurp(const std::string& p)
{
  std::ifstream f(p.c_str(), std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

int main(int argc, char** argv)
{
  if (argc < 3) return 2;
  const size_t rows = (size_t)std::atol(argv[1]);
  const std::string dir = argv[2];
  mat regr(rows, 7);
  unsigned long long s = 88172645463325252ull;
  for (size_t i = 0; i < rows; i++) {
    regr(i, 0) = (double)(i + 1);
    for (size_t j = 1; j < 7; j++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double u = (double)(s >> 11) / 9007199254740992.0;
      // magnitudes from 1e-9 to 1e9, both signs, some exact integers: every branch of "%g"
      regr(i, j) = (u - 0.5) * std::pow(10.0, (double)((int)(s % 19) - 9));
      if (s % 97 == 0) regr(i, j) = std::floor(regr(i, j));
    }
  }
  if (rows > 3) {
    regr(1, 3) = 0.0;
    regr(2, 3) = -0.0;
    regr(3, 2) = std::numeric_limits<double>::infinity();
    regr(3, 3) = std::numeric_limits<double>::quiet_NaN();
  }
  typedef std::chrono::steady_clock clk;
  {
    // Control::sortedOrder against sort_index (stable) on a column with many ties
    mat yv(rows, 1);
    for (size_t i = 0; i < rows; i++) yv[i] = std::floor(regr(i, 4) * 1e-3) + (double)(i % 7 == 0);
    auto s0 = clk::now();
    const uvec ref = sort_index(yv, "ascend");
    auto s1 = clk::now();
    const uvec o1 = Control::sortedOrder(yv, 1), o5 = Control::sortedOrder(yv, 5);
    auto s2 = clk::now();
    const uvec o0 = Control::sortedOrder(yv, 0);
    auto s3 = clk::now();
    bool same_order = ref.n_elem == o1.n_elem;
    for (size_t i = 0; same_order && i < rows; i++) same_order = ref[i] == o1[i] && ref[i] == o5[i] && ref[i] == o0[i];
    std::printf("sorted order identical %d  sort_index %.3f s  1 + 5 workers %.3f s  default workers %.3f s\n", (int)same_order,
                std::chrono::duration<double>(s1 - s0).count(), std::chrono::duration<double>(s2 - s1).count(), std::chrono::duration<double>(s3 - s2).count());
    if (!same_order) return 4;
  }
  auto t0 = clk::now();
  {
    std::ofstream outputs((dir + "/ref_loop.txt").c_str());
    outputs << "# SampleNo, Y,  Yh, StdYh, Inputs" << std::endl;
    for (size_t i = 0; i < rows; i++) {
      for (size_t j = 0; j < 7; j++) outputs << regr(i, j) << "\t";
      outputs << std::endl;
    }
  }
  auto t1 = clk::now();
  if (!Control::writePredictTable(dir + "/w1.txt", regr, 1)) return 3;
  auto t2 = clk::now();
  if (!Control::writePredictTable(dir + "/w7.txt", regr, 7)) return 3;
  auto t3 = clk::now();
  if (!Control::writePredictTable(dir + "/w0.txt", regr, 0)) return 3;
  auto t4 = clk::now();
  const std::string a = slurp(dir + "/ref_loop.txt"), b = slurp(dir + "/w1.txt"), c = slurp(dir + "/w7.txt"), d = slurp(dir + "/w0.txt");
  const bool same = a == b && b == c && c == d;
  std::printf("rows %zu bytes %zu identical %d  ostream loop %.3f s  1 worker %.3f s  7 workers %.3f s  default workers %.3f s\n", rows, a.size(), (int)same,
              std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(),
              std::chrono::duration<double>(t3 - t2).count(), std::chrono::duration<double>(t4 - t3).count());
  return same ? 0 : 1;
}
