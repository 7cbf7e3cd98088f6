// gp_ss_ak -- command line of the B200 exact-GP path.  Same commands, flags, prompts, files and printed lines as the
// reference's gp_ss_ak.cpp (/root/reference; file:line citations below):
//   gp_ss_ak [-v N] [-pm M] train [-k ExpAns] [-kn 1] [-o BFGS|LBFGS|SCG] [-# iters] train.txt [model]
//   gp_ss_ak [-v N] [-pm M] test  test.txt model train.txt [predictions.txt]
// Everything numerical happens in GP_utils (device-resident); this file is argument handling and file I/O.
#include "gp_ss_ak.h"
#include "DistHost.h"

#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

using namespace arma;
using std::cin;
using std::cout;
using std::endl;
using std::string;

// GPSS_TIMING=1: wall-clock of the phases of train / test on stderr (rank 0) -- where the time of a 10 M-point test run goes
// (reader, standardisation, factorisation, prediction, sort, writer).  No effect on any output file.
namespace {
struct PhaseClock {
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  PhaseClock() : on(std::getenv("GPSS_TIMING") != nullptr && gpss_host::rank() == 0), t0(std::chrono::steady_clock::now()), last(t0) {}
  void mark(const char* what)
  {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[gpss timing] %-34s %9.3f s   (total %9.3f s)\n", what, std::chrono::duration<double>(now - last).count(),
                 std::chrono::duration<double>(now - t0).count());
    last = now;
  }
};
}  // namespace

int main(int argc, char* argv[])
{
  gpss_host::silence_other_ranks();      // multi-GPU launch: rank 0 alone prints and writes files (DistHost.h)
  gpss_host::clear_stale_ids();          // rank 0: remove rendezvous files a crashed earlier job with the same key left behind
  GP_Cntrl cmd(argc, argv);
  cmd.setFlgs(true);
  cmd.setprepM(1);          // 0: mean/std, 1: symmetric, 2: "0..1"
  cmd.setVerbose(0);
  cmd.setMode("gp");
  while (cmd.isFlgs()) {
    const string arg = cmd.getArg();
    if (arg[0] == '-') {
      if (cmd.isArg("-?", "--?") || cmd.isArg("-h", "--help")) {
        cmd.Help();
        exit(0);
      } else if (cmd.isArg("-v", "--verboseL")) {
        cmd.incArg();
        cmd.setVerbose(cmd.getIntArg());
      } else if (cmd.isArg("-pm", "--prepMethod")) {
        cmd.incArg();
        cmd.setprepM(cmd.getIntArg());
      } else {
        cmd.UnkFlg();
      }
    } else if (arg == "train") {
      cmd.train();
    } else if (arg == "test") {
      cmd.test();
    } else {
      cmd.ErrorTermination("Invalid Commad.");
    }
    cmd.incArg();
  }
  cmd.ErrorTermination("No Command.");
}

GP_Cntrl::GP_Cntrl(int arc, char** arv) : Control(arc, arv) {}

namespace {
// mean squared error and the variance of y, the two figures both commands print (gp_ss_ak.cpp:312-325, 417-430)
void report_errors(const mat& y, const mat& est, uword rows, int verbose, const char* what_mse, const char* what_var)
{
  const double mse = accu(pow(y - est, 2)) / rows;
  mat dY = y - accu(y) / y.n_rows;
  dY = pow(dY, 2);
  const double var_y = accu(dY) / y.n_rows;
  if (verbose > 0) {
    cout << what_mse << mse << "\n";
    cout << what_var << var_y << "\n";
  } else {
    cout << mse << "\n";
    cout << var_y << "\n";
  }
}

// optional interactive override of one starting value (gp_ss_ak.cpp:250-262, 274-283); EOF on stdin keeps the default
long double prompt_value(long double current)
{
  long double d = current;
  if (cin.peek() != '\n') cin >> d;
  cin.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  return d;
}
}  // namespace

void GP_Cntrl::train()
{
  incArg();
  setMode("train");
  bool yscale = true;
  string optimiser = "BFGS";
  int Data_mode = 0;
  int numhyper = 0;
  std::vector<string> KernT;
  bool Knoise = true;                    // `bool Knoise = "true"` in the reference
  string mean_functionstr = "mean_zero", likelihoodf = "Gauss", Inference = "inf_laplace";
  int iters = 100;
  string modelName = "gp_model";
  while (isFlgs()) {
    if (!isArgFlg()) { setFlgs(false); continue; }
    if (isArg("-h", "--help")) { Help(); exit(0); }
    else if (isArg("-mf", "--meanfunction")) { incArg(); mean_functionstr = getArg(); }
    else if (isArg("-lf", "--likefunction")) { incArg(); likelihoodf = getArg(); }
    else if (isArg("-k", "--kernel")) { incArg(); KernT.push_back(getArg()); }
    else if (isArg("-o", "--optimiser")) { incArg(); optimiser = getArg(); }
    else if (isArg("-#", "--iterations")) { incArg(); iters = getIntArg(); }
    else if (isArg("-kn", "--Knoise")) { incArg(); Knoise = getIntArg(); }
    else UnkFlg();
    incArg();
  }
  if (getArgNo() >= argc) ErrorTermination("There are not enough input parameters.");
  const string trainFileName = getArg();
  if ((getArgNo() + 1) < argc) modelName = argv[getArgNo() + 1];

  int* data_size = readDataSize(trainFileName);
  mat X(data_size[0], data_size[1]);
  mat y(data_size[0], 1);
  readDataFile(X, y, data_size, trainFileName);
  prepareData(X, y, Data_mode, yscale, modelName);

  // covariance function: the -k members in order, then the bias term unless -kn 0 (gp_ss_ak.cpp:146-190)
  HybKerns Kerns(X);
  for (size_t i = 0; i < KernT.size(); i++) {
    Kernels* k = 0;
    if (KernT[i] == "ExpAns") k = new Kern_ExpAnisotropic(X);
    else if (KernT[i] == "Bias") k = new Kern_Bias(X);
    else if (KernT[i] == "RBF") k = new Kern_RBF(X);
    else if (KernT[i] == "Exp") k = new Kern_Exponential(X);
    else if (KernT[i] == "White") k = new Kern_White(X);
    else ErrorTermination("Unknown covariance function: " + KernT[i]);
    Kerns.addNewKernel(k);
    delete k;
  }
  if (Kerns.getNumKerns() == 0) {
    Kern_ExpAnisotropic k(X);
    Kerns.addNewKernel(&k);
    KernT.push_back("ExpAn");
  }
  if (Knoise) {
    Kern_Bias k(X);
    Kerns.addNewKernel(&k);
  }
  int likeLtype = -1, Inf_type = -1, mean_type = -1, numlik_par = 0, numMF_par = 0;
  if (likelihoodf == "Gauss") { likeLtype = GP_utils::likeL_Gaussian; numlik_par = 1; }
  if (Inference == "inf_laplace") Inf_type = GP_utils::inf_laplace;
  if (mean_functionstr == "mean_zero") { mean_type = GP_utils::mean_zero; numMF_par = 0; }
  else ErrorTermination("Unrecognised mean function");
  // only stored; the reference compares against "ExpAn", so "-k ExpAns" leaves it unset there (gp_ss_ak.cpp:210-219)
  numhyper = (KernT[0] == "ExpAn") ? 8 : 2;

  GP_utils* GPModel = new GP_utils(&Kerns, X, y, Inf_type, likeLtype, mean_type, numhyper, numlik_par, numMF_par, getVerbose());

  cout << "The inital value of the kernel parameters are as follows :" << endl;
  cout << "There are " << GPModel->KerenlW->getNPars() << " parameters to be optimized" << endl;
  for (unsigned int i = 0; i < GPModel->KerenlW->getNPars(); i++)
    cout << GPModel->KerenlW->getParamName(i) << " : " << GPModel->KerenlW->getParam(i) << endl;
  cout << "Do you want to change the defult kernel parameters (Yes|Y|y or press any key)?" << endl;
  string res = "No";
  cin >> res;
  if (res == "Yes" || res == "Y" || res == "y") {
    for (unsigned int i = 0; i < GPModel->KerenlW->getNPars(); i++) {
      if (GPModel->KerenlW->getParamName(i) == "InversewidthR_ExpAns" && X.n_cols == 3) continue;
      cout << " Please input an initial value for " << GPModel->KerenlW->getParamName(i) << " (Default value was "
           << GPModel->KerenlW->getParam(i) << ") : " << endl;
      GPModel->KerenlW->setParam(prompt_value(GPModel->KerenlW->getParam(i)), i);
    }
  }
  cout << "The inital value of the likelihood function are as follows :" << endl;
  cout << "likelihood hyperparameter : " << GPModel->getHyperlfVal(0) << endl;
  cout << "Do you want to change the defult likelihood function parameters (Yes|Y|y or press any key)?" << endl;
  res = "No";
  cin >> res;
  if (res == "Yes" || res == "Y" || res == "y") {
    cout << "Please input an initial value for Gauss likelihood function : " << endl;
    GPModel->setHyperlfVal(prompt_value(GPModel->getHyperlfVal(0)), 0);
  }
  if (optimiser == "SCG") GPModel->setOptimiser(GP_utils::SCG);
  else if (optimiser == "LBFGS") GPModel->setOptimiser(GP_utils::LBFGS);
  else if (optimiser == "BFGS") GPModel->setOptimiser(GP_utils::BFGS);
  else ErrorTermination("Unrecognised optimiser type: " + optimiser);

  GPModel->OptimisePars(iters);
  writeGPFile(*GPModel, modelName, "# GP_SS_AK Model File ");

  // fitted values on the training set (gp_ss_ak.cpp:301-325)
  mat EstVals(X.n_rows, GPModel->getOutDim()), EstVals_Var(X.n_rows, GPModel->getOutDim());
  GPModel->Calc_Out(EstVals, EstVals_Var, X);
  postData(X, EstVals, yscale, modelName);
  postData_var(EstVals_Var, yscale, modelName);
  postData(y, yscale, modelName);
  EstVals_Var = sqrt(EstVals_Var);
  report_errors(y, EstVals, X.n_rows, getVerbose(), "Mean Square Error of training: ", "Var MSE Train: ");
  exit(0);
}

void GP_Cntrl::test()
{
  incArg();
  setMode("test");
  int Data_mode = 1;
  bool yscale = true;
  string data_File_NameTr, modelName = "model";
  while (isFlgs()) {
    if (!isArgFlg()) { setFlgs(false); continue; }
    if (getArgLen() != 2) UnkFlg();
    else if (isArg("-?", "--?") || isArg("-h", "--help")) { Help(); exit(0); }
    else UnkFlg();
    incArg();
  }
  if (getArgNo() >= argc) ErrorTermination("There are not enough input parameters.");
  const string data_File_Name = getArg();
  if ((getArgNo() + 1) < argc) modelName = argv[getArgNo() + 1];
  if ((getArgNo() + 2) < argc) data_File_NameTr = argv[getArgNo() + 2];
  else cout << "Please provide training data \n";
  string PredictOut = modelName + "_predict.txt";
  if ((getArgNo() + 3) < argc) PredictOut = argv[getArgNo() + 3];

  PhaseClock clk;
  int* data_size = readDataSize(data_File_Name);
  mat X(data_size[0], data_size[1]), y(data_size[0], 1);
  readDataFile(X, y, data_size, data_File_Name);
  clk.mark("read test file");
  prepareData(X, y, Data_mode, yscale, modelName);
  GP_utils* GPModel = readGpFromFile(modelName, getVerbose());
  clk.mark("standardise test data, read model");

  // the model file holds parameters only: the training set is read and standardised again (gp_ss_ak.cpp:384-395)
  data_size = readDataSize(data_File_NameTr);
  mat Xtr(data_size[0], data_size[1]), ytr(data_size[0], 1);
  readDataFile(Xtr, ytr, data_size, data_File_NameTr);
  prepareData(Xtr, ytr, Data_mode, yscale, modelName);
  GPModel->yTarg = ytr;
  GPModel->Xinp = Xtr;
  GPModel->setNumData(Xtr.n_rows);
  GPModel->initialize_vars();
  clk.mark("read + standardise training file");
  GPModel->logLikelihood();
  clk.mark("factorisation (logLikelihood)");
  if (X.n_cols != GPModel->getInpDim()) ErrorTermination("Incorrect dimension of input data.");

  mat EstVals(y.n_rows, y.n_cols), EstVals_Var(y.n_rows, y.n_cols);
  GPModel->Calc_Out(EstVals, EstVals_Var, X);
  clk.mark("prediction (Calc_Out)");
  postData(X, EstVals, yscale, modelName);
  postData_var(EstVals_Var, yscale, modelName);
  postData(y, yscale, modelName);
  report_errors(y, EstVals, X.n_rows, getVerbose(), "Mean Square Error of testing: ", "Var MSE Test: ");

  clk.mark("de-standardise, error report");
  // <model>_predict.txt: rows sorted by the observed value (gp_ss_ak.cpp:434-481)
  const uvec order = Control::sortedOrder(y);                // = sort_index(y, "ascend") with a stable sort, in parallel chunks
  mat regr(y.n_rows, 4 + X.n_cols);
  for (uword i = 0; i < y.n_rows; i++) {
    const uword s = order[i];
    regr(i, 0) = (double)(i + 1);
    regr(i, 1) = y[s];
    regr(i, 2) = EstVals[s];
    regr(i, 3) = EstVals_Var[s];
    for (uword j = 0; j < X.n_cols; j++) regr(i, 4 + j) = X(s, j);
  }
  // [quirk] std::string::find returns 0 only when the name STARTS with the word, so "test.txt" gets "_train" appended
  // and "train.txt" gets "_test" (gp_ss_ak.cpp:450-468); only the plot-script name depends on it
  string base = data_File_Name;
  for (size_t slash = base.find("/"); slash != string::npos && slash > 0 && slash + 1 < base.size(); slash = base.find("/"))
    base = base.substr(slash + 1);
  string plotName = modelName;
  if (base.find("train")) plotName += "_train";
  if (base.find("test")) plotName += "_test";

  // "%g" is exactly what `ostream << double` prints with the default format (6 significant digits), which is what the reference's loop
  // does value by value (gp_ss_ak.cpp:470-481); here the rows are formatted by several host threads (Control::writePredictTable)
  if (!Control::writePredictTable(gpss_host::out_path(PredictOut), regr)) ErrorTermination("File is " + PredictOut + " not writable");
  clk.mark("sort by y + write predict file");

  // gnuplot script (gp_ss_ak.cpp:482-505); gnuplot itself is run only when GPSS_RUN_GNUPLOT is set
  const double hi = std::max(regr.col(1).max(), mat(regr.col(2) + regr.col(3)).max());
  const double lo = std::min(regr.col(1).min(), mat(regr.col(2) - regr.col(3)).min());
  const string script = plotName + "_gnu.plt";
  std::ofstream plt(gpss_host::out_path(script).c_str());
  plt << "#gnuplot -persist output.plt\n set termoption enhanced\n set term wxt background rgb \"white\"\n set term pdf transparent enhanced \n ";
  plt << "set output '" + plotName + "_predict.pdf" + "'  \n";
  plt << "set style fill transparent solid 0.70 noborder\n set grid nopolar\n set key inside left top vertical Right noreverse enhanced "
         "autotitle box lt black linewidth 1.000 dashtype solid\n set title \"Observed vs Estimated\" textcolor  \"black\" font "
         "\"Bold-Times-Roman,20\"\n";
  plt << "set ylabel \"Copper grade\" offset 0.1,0.1 textcolor  \"black\" font \"Bold-Times-Roman,10\" \n";
  plt << "set xlabel \"Sample\" offset 0.1,0.1 textcolor  \"black\" font \"Bold-Times-Roman,10\" \n";
  plt << "set colorbox vertical origin screen 0.9, 0.2, 0 size screen 0.05, 0.6, 0 front bdefault \n";
  plt << "plot [0.9:" + std::to_string((double)y.n_rows + 0.01) + "] [" + std::to_string(lo - 0.02) + ":" + std::to_string(hi + 0.02) + "] \"" +
             PredictOut +
             "\" using 1:($3 + $4):($3 - $4) with filledcurve fc rgb \"green\" title '95% CI', \"\" using 1:3 with lines ls 2 lw 1 lc rgb "
             "\"red\" t \"Estimated\", \"\" using 1:2 ls 1 lw 1 lc rgb \"blue\" t \"Observed\" with lines \n";
  plt.close();
  if (std::getenv("GPSS_RUN_GNUPLOT")) {
    const string cmd = "gnuplot -persist " + script;
    if (system(cmd.c_str()) != 0) cout << "gnuplot could not be run\n";
  }
  exit(0);
}

void GP_Cntrl::Help()
{
  const string m = getMode();
  cout << endl << "GP_SS_AK Code: Version 0.5 (B200 exact-GP path)" << endl;
  if (m == "gp") {
    cout << "Command:\n \t ./gp_ss_ak [options] Command [Comnd-options] modelName TrainDataFile.txt" << endl;
    cout << "Commands:" << endl;
    cout << "train :\n \t To find hyperparameter by maxmizing likelihood." << endl;
    cout << "test :\n \t To estimate test data set and plot the results." << endl;
    cout << "To get more information about each command please type command with --h" << endl << endl;
    cout << "Options:" << endl;
    cout << "-?, -h, --help\n \t To get help on options" << endl;
    cout << "-v, --verbose\n \t Verbosity level (default 0)." << endl;
    cout << "-pm, --prepMethod\n \t preparation methos (between mean and std [0], symmetric [1], ...)" << endl;
  } else if (m == "train") {
    cout << "gp [options] learn example_file [model_file]" << endl;
    cout << "Arguments:" << endl;
    cout << "-mf ,--meanfunction\n \t GP mean function name (Default: zero [mean_zero])" << endl;
    cout << "-lf, --likefunction\n \t likelihood function name (default:  [Gauss]" << endl;
    cout << "-k, --kernel\n \t Kernel name (Exponential Anisotropic [ExpAns], Exponential [Exp], Radial Basis Function [RBF])" << endl;
    cout << "-o, --optimiser\n \t Optimization algorithm (Broyden-Fletcher-Goldfarb-Shanno [BFGS] (default), limited-memory BFGS [LBFGS], scaled conjugate gradient [SCG])" << endl;
    cout << "-kn, --Knoise\n \t Bias kernel (default: true [1], other options false [0])" << endl;
    cout << "trainFileName\n \t File containg trainig data (comma delimitted or tab delimitted file)." << endl;
    cout << "modelName\n \t File to store the model." << endl;
    cout << "-#, --iterations\n \t Number of iterations for optimisation algorithm (takes effect with -v 3)." << endl;
  } else if (m == "test") {
    cout << "[options] gp test [test_file] [model_file] [train_file]" << endl;
    cout << "Arguments:" << endl;
    cout << "test_file\n \t The test data file (comma delimitted or tab delimitted file); the last column is the observed value." << endl;
    cout << "model_file\n \t The model you want to use in estimation." << endl;
    cout << "train_file\n \t The file which has been used in training." << endl;
  }
}
