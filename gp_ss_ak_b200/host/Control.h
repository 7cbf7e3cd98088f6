// Command-line cursor, data-file reader and the (host-side, O(n)) standardisation of GP_SS_AK: the reference's
// Control class (Control.h:23-170, Control.cpp in /root/reference) with the same member names and file formats.
// north_star keeps this part on the host with Armadillo; nothing here touches the GPU.
#ifndef GPSS_HOST_CONTROL_H
#define GPSS_HOST_CONTROL_H

#include <armadillo>
#include <cstdlib>
#include <cstring>
#include <string>

#include "StreamInt.h"

class Control {
 public:
  Control(int argc, char** argv);
  virtual ~Control() {}

  // ---- argv cursor (reference Control.h:77-139) ----
  bool isArg(std::string shortName, std::string longName);
  void UnkFlg();
  bool isFlgs() const { return flgs && getArgNo() < argc; }
  void setFlgs(bool val) { flgs = val; }
  int getVerbose() const { return verbose; }
  void setVerbose(int val) { verbose = val; }
  int getprepM() const { return prepareM; }
  void setprepM(int val) { prepareM = val; }
  void incArg() { argNo++; }
  int getArgNo() const { return argNo; }
  void setArgNo(int val) { argNo = val; }
  std::string getArg() const { return argv[argNo]; }
  int getIntArg() const { return (int)std::atol(argv[argNo]); }
  double getArgLen() const { return (double)std::strlen(argv[argNo]); }
  bool isArgFlg() const { return argv[argNo][0] == '-'; }
  void setMode(std::string val) { mode = val; }
  std::string getMode() const { return mode; }

  // ---- data files: tab- or comma-separated, '#' comment lines, columns x_1..x_D then y (Control.cpp:27-141) ----
  int* readDataSize(const std::string fileName);                       // new int[2] = {rows, D}
  void readDataFile(arma::mat& X, arma::mat& y, int* data_size, const std::string fileName);

  // ---- standardisation (Control.cpp:142-324) ----
  void StatisticsCalc(arma::mat& Xtr, arma::mat& Ytr);
  void prepareData(arma::mat& X, arma::mat& y, int& Data_mode, bool& yscale, std::string ModelN);
  void MeanStd(arma::mat& X, arma::mat& y, int& Data_mode, bool& yscale);
  void zeroandone(arma::mat& X, arma::mat& y, int& Data_mode, bool& yscale);
  void prep_symmetric(arma::mat& X, arma::mat& y, int& Data_mode, bool& yscale);
  void postData(arma::mat& X, arma::mat& y, bool& yscale, std::string ModelN);
  void postData(arma::mat& X, bool& yscale, std::string ModelN);
  void postData_var(arma::mat& X, bool& yscale, std::string ModelN);

  // ---- <model>_predict.txt (gp_ss_ak.cpp:470-481): one row per line, every value followed by a tab, formatted like `ostream << double`
  //      with the default format ("%g").  Rows are formatted by `threads` workers into private buffers and written in row order
  //      (0 = as many as the host has cores, capped at 16); the bytes do not depend on the thread count.  false: file not writable.
  static bool writePredictTable(const std::string& path, const arma::mat& regr, int threads = 0);
  // Row numbers of y in ascending order of the value, ties by row number -- what `sort_index(y, "ascend")` gives with a stable sort, one of
  // the orders Armadillo's sort_index may produce (gp_ss_ak.cpp:434-436).  (value, row) pairs are sorted in `threads` chunks and merged.
  static arma::uvec sortedOrder(const arma::mat& y, int threads = 0);

  void ErrorTermination(const std::string error);
  void Helping();
  void NormalTermination();

  mutable arma::mat MinData, MaxData, MeanData, StData;     // row 0 = y, rows 1..D = the X columns
  mutable double MaxTotalin, MinTotalin, MaxTotalo, MinTotalo;
  mutable arma::mat params;                                 // (D+1) x 2: centre, scale

 private:
  void loadStatistics(const std::string& ModelN, arma::uword nInputs);
  bool flgs;
  int verbose, argNo, prepareM;
  std::string mode;

 protected:
  int argc;
  char** argv;
};

#endif
