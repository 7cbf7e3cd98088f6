// See Control.h.  File:line citations are into /root/reference.
#include "Control.h"
#include "DistHost.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <cstdio>
#include <iostream>
#include <thread>
#include <utility>
#include <vector>

using namespace arma;
using std::cerr;
using std::cout;
using std::endl;
using std::string;

Control::Control(int arc, char** arv) : flgs(false), verbose(0), argNo(1), prepareM(1), argc(arc), argv(arv) {}

bool Control::isArg(string shortName, string longName) { return getArg() == shortName || getArg() == longName; }
void Control::UnkFlg() { ErrorTermination("Unknown flag: " + getArg() + " provided."); }
void Control::Helping() { cout << "To get more information use help command." << endl; }
void Control::ErrorTermination(const string error)
{
  cerr << error << endl << endl;
  Helping();
  exit(1);
}
void Control::NormalTermination() { exit(0); }

namespace {
// Streaming reader (at 10 M test points the text I/O is what is left once prediction is
// fast).  The file is read in one piece and scanned in place -- no per-field std::string -- with the reference's rules
// (Control.cpp:27-141): lines starting with '#' are comments; fields are separated by a tab or a comma; EMPTY fields are
// skipped; a field's value is atof() of its text; every non-comment line is a row (an empty line too, Control.cpp:100-103);
// a final segment without a newline is a line if it is non-empty (std::getline).
struct FileText {
  std::string buf;
  bool ok;
  explicit FileText(const string& name) : ok(false)
  {
    std::ifstream in(name.c_str(), std::ios::binary);
    if (!in.is_open()) return;
    in.seekg(0, std::ios::end);
    const std::streamoff len = in.tellg();
    in.seekg(0, std::ios::beg);
    buf.resize((size_t)len);
    if (len > 0) in.read(&buf[0], len);
    ok = true;
  }
};

// atof of the field [b, e): the text is copied to a small NUL-terminated buffer so that strtod can neither run into the next
// field (it skips leading white space, and a tab IS white space) nor past the end of the file
inline double field_value(const char* b, const char* e)
{
  // Fast path (Clinger): [sign] digits [. digits] [e|E [sign] digits] and nothing else, with a decimal significand below 2^53
  // and a power of ten within 10^+-22, is m * 10^k or m / 10^k with BOTH factors exact doubles, hence one correctly rounded
  // operation -- the same double atof returns.  Anything else (white space, trailing text, 17-digit values, inf/nan, hex)
  // takes atof itself below.
  {
    static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) neg = (*p++ == '-');
    unsigned long long m = 0;
    int digits = 0, frac = 0;
    bool ok = true;
    while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { m = m * 10 + (unsigned)(*p - '0'); digits++; } else ok = false; p++; }
    const bool int_part = p > b + ((b < e && (*b == '-' || *b == '+')) ? 1 : 0);
    bool frac_part = false;
    if (p < e && *p == '.') {
      p++;
      while (p < e && *p >= '0' && *p <= '9') { if (digits < 19) { m = m * 10 + (unsigned)(*p - '0'); digits++; frac++; } else ok = false; p++; frac_part = true; }
    }
    int ex = 0;
    if (ok && (int_part || frac_part) && p < e && (*p == 'e' || *p == 'E')) {
      const char* q = p + 1;
      bool eneg = false;
      if (q < e && (*q == '-' || *q == '+')) eneg = (*q++ == '-');
      if (q < e && *q >= '0' && *q <= '9') {
        int v = 0;
        while (q < e && *q >= '0' && *q <= '9' && v < 10000) v = v * 10 + (*q++ - '0');
        ex = eneg ? -v : v;
        p = q;
      }
    }
    if (ok && (int_part || frac_part) && p == e && m < (1ULL << 53)) {
      const int k = ex - frac;
      if (k >= -22 && k <= 22) {
        const double v = k >= 0 ? (double)m * p10[k] : (double)m / p10[-k];
        return neg ? -v : v;
      }
    }
  }
  char tmp[64];
  size_t len = (size_t)(e - b);
  if (len < sizeof tmp) {
    std::memcpy(tmp, b, len);
    tmp[len] = 0;
    return std::atof(tmp);
  }
  return std::atof(string(b, e).c_str());
}

// calls row(line_begin, line_end) for every non-comment line
template <class F>
void for_each_row(const std::string& text, F row)
{
  const char* p = text.data();
  const char* const end = p + text.size();
  while (p < end) {
    const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    if (!(le > p && *p == '#')) row(p, le);
    p = nl ? nl + 1 : end;
  }
}

// calls field(b, e) for every NON-EMPTY field of the line
template <class F>
inline void for_each_field(const char* p, const char* le, F field)
{
  while (p < le) {
    const char* q = p;
    while (q < le && *q != '\t' && *q != ',') q++;
    if (q > p) field(p, q);
    p = q + 1;
  }
}
}  // namespace

// Rows = every non-comment line; D = (largest number of non-empty fields on a line) - 1, the last field being the target
// (Control.cpp:118-130).
int* Control::readDataSize(const string fileName)
{
  int* data_size = new int[2];
  FileText file(fileName);
  if (!file.ok) ErrorTermination("File is " + fileName + " not readable");
  int rows = 0, maxD = 0;
  for_each_row(file.buf, [&](const char* b, const char* e) {
    rows++;
    int fields = 0;
    for_each_field(b, e, [&](const char*, const char*) { fields++; });
    if (fields - 1 > maxD) maxD = fields - 1;
  });
  data_size[0] = rows;
  data_size[1] = maxD;
  if (verbose > 0) {
    cout << "Number of features in the input file are: " << maxD << endl;
    cout << "Number of readable data are: " << rows << endl;
  }
  return data_size;
}

// The first D non-empty fields of a line go to X, every later one overwrites y (Control.cpp:61-77); atof semantics.
void Control::readDataFile(mat& X, mat& y, int* data_size, const string fileName)
{
  FileText file(fileName);
  if (!file.ok) ErrorTermination("File is " + fileName + " not readable");
  const int rows = data_size[0], D = data_size[1];
  double* Xp = X.memptr();
  double* yp = y.memptr();
  const size_t ldx = X.n_rows;
  int row = 0;
  for_each_row(file.buf, [&](const char* b, const char* e) {
    int col = 0;
    for_each_field(b, e, [&](const char* fb, const char* fe) {
      if (row < 0 || row >= rows) ErrorTermination("Erro while reading" + fileName);
      const double v = field_value(fb, fe);
      if (col < D) Xp[(size_t)(col++) * ldx + row] = v;
      else yp[row] = v;
    });
    row++;
  });
}

// per-column min / max / mean / sample standard deviation, plus the global extremes of X and y (Control.h:46-73)
void Control::StatisticsCalc(mat& Xtr, mat& Ytr)
{
  const int D = Xtr.n_cols, n = Xtr.n_rows;
  MaxTotalin = Xtr.max();
  MinTotalin = Xtr.min();
  MaxTotalo = Ytr.max();
  MinTotalo = Ytr.min();
  for (int i = 0; i <= D; i++) {
    const mat column = (i == 0) ? Ytr : mat(Xtr.col(i - 1));
    MinData[i] = (i == 0) ? MinTotalo : column.min();
    MaxData[i] = (i == 0) ? MaxTotalo : column.max();
    MeanData[i] = accu(column) / n;
    StData[i] = std::sqrt(accu(pow(column - MeanData[i], 2)) / (n - 1));
  }
}

// <model>_Statistics.txt: csv, one row per variable (y first): centre, scale, min, max, mean, std (Control.cpp:151-163)
void Control::loadStatistics(const string& ModelN, uword nInputs)
{
  mat Statistics;
  Statistics.load(ModelN + "_Statistics.txt", csv_ascii);
  params = Statistics.cols(0, 1);
  MinData = Statistics.col(2);
  MaxData = Statistics.col(3);
  MeanData = Statistics.col(4);
  StData = Statistics.col(5);
  MaxTotalin = mat(MaxData.submat(1, 0, nInputs, 0)).max();
  MaxTotalo = MaxData[0];
  MinTotalin = mat(MinData.submat(1, 0, nInputs, 0)).min();
  MinTotalo = MinData[0];
}

void Control::prepareData(mat& X, mat& y, int& Data_mode, bool& yscale, string ModelN)
{
  const uword nv = X.n_cols + y.n_cols;
  params.zeros(nv, 2);
  MinData.resize(nv, 1);
  MaxData.resize(nv, 1);
  MeanData.resize(nv, 1);
  StData.resize(nv, 1);
  if (getMode() == "train") StatisticsCalc(X, y);
  else if (getMode() == "test") loadStatistics(ModelN, X.n_cols);
  switch (prepareM) {
    case 0:
      MeanStd(X, y, Data_mode, yscale);
      if (verbose > 0) cout << "Preparation method is between mean and standardDev and y scale is " << yscale << endl;
      break;
    case 1:
      prep_symmetric(X, y, Data_mode, yscale);
      if (verbose > 0) cout << "Preparation method is symmetric and y scale is " << yscale << endl;
      break;
    case 2:
      zeroandone(X, y, Data_mode, yscale);
      if (verbose > 0) cout << "Preparation method is between 0 and 1 and y scale is " << yscale << endl;
      break;
    default:
      ErrorTermination("Unrecognised preparation method.");
  }
  if (getMode() == "train") {
    mat Statistics = join_horiz(params, MinData);
    Statistics = join_horiz(Statistics, MaxData);
    Statistics = join_horiz(Statistics, MeanData);
    Statistics = join_horiz(Statistics, StData);
    Statistics.save(gpss_host::out_path(ModelN + "_Statistics.txt"), csv_ascii);
  }
}

namespace {
void apply_params(mat& X, mat& y, const mat& params, bool yscale)
{
  for (uword j = 0; j < X.n_cols; j++) X.col(j) = (X.col(j) - params(j + 1, 0)) / params(j + 1, 1);
  if (yscale) y = (y - params(0, 0)) / params(0, 1);
}
}  // namespace

void Control::MeanStd(mat& X, mat& y, int&, bool& yscale)
{
  if (getMode() == "train")
    for (uword j = 0; j <= X.n_cols; j++) { params(j, 0) = MeanData[j]; params(j, 1) = StData[j]; }
  apply_params(X, y, params, yscale);
}

// [quirk] the "0..1" method centres on min/2 and is recomputed in test mode too (Control.cpp:276-296)
void Control::zeroandone(mat& X, mat& y, int&, bool& yscale)
{
  for (uword j = 0; j <= X.n_cols; j++) {
    params(j, 0) = 0.5 * MinData[j];
    params(j, 1) = 0.5 * (MaxData[j] - MinData[j]);
  }
  for (uword i = 0; i < X.n_rows; i++) {
    for (uword j = 0; j < X.n_cols; j++) X(i, j) = (X(i, j) - params(j + 1, 0)) / params(j + 1, 1);
    if (yscale) y[i] = (y[i] - params(0, 0)) / params(0, 1);
  }
}

// Symmetric standardisation to [-1, 1]: the first THREE input columns share one centre / half-range taken from the
// global extremes of all inputs, later columns use their own (Control.cpp:299-324).
void Control::prep_symmetric(mat& X, mat& y, int&, bool& yscale)
{
  if (getMode() == "train") {
    params(0, 0) = 0.5 * (MaxTotalo + MinTotalo);
    params(0, 1) = 0.5 * (MaxTotalo - MinTotalo);
    for (int j = 0; j < 3; j++) {
      params(j + 1, 0) = 0.5 * (MaxTotalin + MinTotalin);
      params(j + 1, 1) = 0.5 * (MaxTotalin - MinTotalin);
    }
    for (uword j = 3; j < X.n_cols; j++) {
      params(j + 1, 0) = 0.5 * (MaxData[j + 1] + MinData[j + 1]);
      params(j + 1, 1) = 0.5 * (MaxData[j + 1] - MinData[j + 1]);
    }
  }
  apply_params(X, y, params, yscale);
}

void Control::postData(mat& X, mat& y, bool& yscale, string ModelN)
{
  if (getMode() == "test") loadStatistics(ModelN, X.n_cols);
  for (uword j = 0; j < X.n_cols; j++) X.col(j) = (X.col(j) * params(j + 1, 1)) + params(j + 1, 0);
  if (yscale) y = (y * params(0, 1)) + params(0, 0);
}

void Control::postData(mat& X, bool&, string ModelN)
{
  if (getMode() == "test") loadStatistics(ModelN, X.n_cols);
  X = X * params(0, 1) + params(0, 0);
}

// variance -> standard deviation in the units of y (Control.cpp:238-255)
void Control::postData_var(mat& X, bool& yscale, string ModelN)
{
  if (getMode() == "test") loadStatistics(ModelN, X.n_cols);
  if (yscale) X = sqrt(X * std::pow(params(0, 1), 2));
}


// ---------------------------------------------------------------------------------------------------
// Prediction table writer.  The reference inserts value by value into an ofstream (gp_ss_ak.cpp:470-481): at 10 M rows x 7 columns
// that is 70 M conversions on one core -- 10 s of the 65 s of BASELINE configs[3] on 8 GPUs (profiles/r02_predict_10m_8gpu.log), more than
// reading and standardising the input.  The conversions are independent per row: rows are cut into contiguous chunks, one per worker,
// each formatted with snprintf("%g\t") (= the default ostream format) into its own buffer; the buffers are written in order.
// ---------------------------------------------------------------------------------------------------
bool Control::writePredictTable(const std::string& path, const arma::mat& regr, int threads)
{
  FILE* out = std::fopen(path.c_str(), "w");
  if (!out) return false;
  std::fputs("# SampleNo, Y,  Yh, StdYh, Inputs\n", out);
  const size_t rows = regr.n_rows, cols = regr.n_cols;
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 16) threads = 16;
  }
  if ((size_t)threads > rows / 4096 + 1) threads = (int)(rows / 4096 + 1);      // small tables: no point in spawning workers
  const size_t piece = (size_t)1 << 13;                                         // rows per work item (bounds the buffers: ~2 MB each, allocated once)
  std::vector<std::vector<char> > bufs((size_t)threads);
  std::vector<size_t> used((size_t)threads, 0);
  for (size_t base = 0; base < rows; base += piece * (size_t)threads) {
    auto work = [&](int t) {
      const size_t r0 = base + piece * (size_t)t, r1 = (r0 + piece < rows) ? r0 + piece : rows;
      used[t] = 0;
      if (r0 >= rows) return;
      std::vector<char>& b = bufs[t];
      if (b.size() < piece * (32 * cols + 1)) b.resize(piece * (32 * cols + 1));
      size_t u = 0;
      for (size_t i = r0; i < r1; i++) {
        for (size_t j = 0; j < cols; j++) u += (size_t)std::snprintf(b.data() + u, 32, "%g\t", regr(i, j));
        b[u++] = '\n';
      }
      used[t] = u;
    };
    if (threads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    for (int t = 0; t < threads; t++)
      if (used[t]) std::fwrite(bufs[t].data(), 1, used[t], out);
  }
  return std::fclose(out) == 0;
}


// Sorting 10 M observed values through an index comparator is 23 passes of random reads; (value, row) pairs sort with sequential
// access, have a total order (so the result does not depend on how the work is cut) and the chunks sort in parallel.
arma::uvec Control::sortedOrder(const arma::mat& y, int threads)
{
  typedef std::pair<double, arma::uword> Item;
  const size_t n = y.n_elem;
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 16) threads = 16;
  }
  if ((size_t)threads > n / 65536 + 1) threads = (int)(n / 65536 + 1);
  std::vector<Item> v(n);
  bool has_nan = false;
  for (size_t i = 0; i < n; i++) { v[i] = Item(y[i], (arma::uword)i); has_nan = has_nan || (y[i] != y[i]); }
  if (has_nan) return arma::sort_index(y, "ascend");          // no total order on the pairs: leave it to the library call the reference makes
  std::vector<size_t> cut((size_t)threads + 1);
  for (int t = 0; t <= threads; t++) cut[t] = n * (size_t)t / (size_t)threads;
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back([&v, &cut, t] { std::sort(v.begin() + cut[t], v.begin() + cut[t + 1]); });
    for (auto& th : pool) th.join();
  }
  // pairwise merges of neighbouring runs, each round in parallel, ping-pong between two buffers (std::inplace_merge would allocate and
  // fault in a temporary of half the range at every call)
  std::vector<Item> w(threads > 1 ? n : 0);
  std::vector<Item>* src = &v;
  std::vector<Item>* dst = &w;
  for (int width = 1; width < threads; width *= 2) {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t += 2 * width) {
      const size_t a = cut[t], m = cut[(t + width < threads) ? t + width : threads], b = cut[(t + 2 * width < threads) ? t + 2 * width : threads];
      pool.emplace_back([src, dst, a, m, b] { std::merge(src->begin() + a, src->begin() + m, src->begin() + m, src->begin() + b, dst->begin() + a); });
    }
    for (auto& th : pool) th.join();
    std::swap(src, dst);
  }
  if (src != &v) v.swap(*src);
  arma::uvec order(n);
  for (size_t i = 0; i < n; i++) order[i] = v[i].second;
  return order;
}
