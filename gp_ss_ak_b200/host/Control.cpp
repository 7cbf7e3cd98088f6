// See Control.h.  File:line citations are into /root/reference.
#include "Control.h"
#include "DistHost.h"

#include <cmath>
#include <fstream>
#include <iostream>

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
// next field of a line: everything up to the next tab or comma; the cursor moves past the separator (Control.cpp:52-59)
string next_field(const string& line, size_t& pos)
{
  string token;
  while (pos < line.size() && line[pos] != '\t' && line[pos] != ',') token += line[pos++];
  pos++;
  return token;
}
bool is_comment(const string& line) { return !line.empty() && line[0] == '#'; }
}  // namespace

// Rows = every non-comment line (an empty line counts, Control.cpp:100-103); D = (largest number of non-empty
// fields on a line) - 1, the last field being the target (Control.cpp:118-130).
int* Control::readDataSize(const string fileName)
{
  int* data_size = new int[2];
  std::ifstream in(fileName.c_str());
  if (!in.is_open()) ErrorTermination("File is " + fileName + " not readable");
  string line;
  int rows = 0, maxD = 0;
  while (std::getline(in, line)) {
    if (is_comment(line)) continue;
    rows++;
    size_t pos = 0;
    int fields = 0;
    while (pos < line.size())
      if (next_field(line, pos).size() > 0) fields++;
    if (fields - 1 > maxD) maxD = fields - 1;
  }
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
  std::ifstream in(fileName.c_str());
  if (!in.is_open()) ErrorTermination("File is " + fileName + " not readable");
  const int rows = data_size[0], D = data_size[1];
  string line;
  int row = 0;
  while (std::getline(in, line)) {
    if (is_comment(line)) continue;
    size_t pos = 0;
    int col = 0;
    while (pos < line.size()) {
      const string field = next_field(line, pos);
      if (field.empty()) continue;
      if (col < D) {
        if (row < 0 || row >= rows) ErrorTermination("Erro while reading" + fileName);
        X(row, col++) = std::atof(field.c_str());
      } else {
        y[row] = std::atof(field.c_str());
      }
    }
    row++;
  }
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
