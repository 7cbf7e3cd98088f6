// See GP_Utils.h.  File:line citations are into /root/reference.
#include "GP_Utils.h"
#include "DistHost.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>

using namespace arma;
using std::cout;
using std::endl;
using std::string;

GP_utils::GP_utils() : Modeling(), Main_Opt_Algs(), KerenlW(0) { _init(); }

GP_utils::GP_utils(Kernels* kernel, mat Xin, mat Yin, int Inf_type, int likeLtype, int mean_type, unsigned int numhyper,
                   unsigned int numlik_par, unsigned int numMF_par, int verbos)
    : Modeling(), Main_Opt_Algs(), Xinp(Xin), yTarg(Yin), KerenlW(kernel)
{
  setNumMFpar(numMF_par);
  setNumlikfpar(numlik_par);
  setNumCovpar(numhyper);
  _init();
  setInferenceType(Inf_type);
  setLikelihoodType(likeLtype);
  setMeanType(mean_type);
  setVerbose(verbos);
  g_param.resize(1, KerenlW->getNPars());
  setOutDim(yTarg.n_cols);
  setInpDim(Xin.n_cols);
  setNumData(yTarg.n_rows);
  initialize_vars();
  // hard-coded starting values of the reference (GP_Utils.cpp:33-44)
  if (numMF_par > 0) { hypermf.resize(numMF_par, 1); hypermf.fill(0.12); }
  if (numlik_par > 0) { hyperlf.resize(numlik_par, 1); hyperlf.fill(0.016); }
}

GP_utils::~GP_utils()
{
  if (handle) gpss_destroy(handle);
}

void GP_utils::_init()
{
  setName("Gaussian process");
  setInf("Lapalce");            // sic: this spelling is part of the model-file format (GP_Utils.cpp:50)
  setlik("Gaussian");
  setMean("Zero");
  likelihoodType_ = likeL_Gaussian;
  InferenceType_ = inf_laplace;
  MeanType_ = mean_zero;
  handle = 0;
  handle_n = -1;
  kind_dev = -1;
  kind2_dev = -1;
  for (int i = 0; i < 8; i++) theta2_dev[i] = 0.0;
  white_cross = 0;
  data_stale = true;
  dirty = true;
  Chol_fail = false;
  device = gpss_host::device();      // GPSS_DEVICE (or LOCAL_RANK in a multi-process launch), default 0
  L.zeros(1, 1);
}

void GP_utils::initialize_vars()
{
  const unsigned int n = getNumData();
  Alpha.resize(n, getOutDim());     // zero-filled growth, like arma::Mat::resize in the reference (GP_Utils.cpp:69)
  yhat.resize(n, getOutDim());
  data_stale = true;
  dirty = true;
}

string GP_utils::getLiklihoodStr() const
{
  if (likelihoodType_ == likeL_Gaussian) return "likeL_Gaussian";
  if (likelihoodType_ == likeL_WarpGauss) return "likeL_WarpGauss";
  cout << "Unknown approximation type \n";
  exit(1);
}
string GP_utils::getInferenceStr() const
{
  if (InferenceType_ == inf_laplace) return "inf_laplace";
  if (InferenceType_ == inf_EP) return "inf_EP";
  cout << "Unknown approximation type \n";
  exit(1);
}
string GP_utils::getMeanTypeStr() const
{
  if (MeanType_ == mean_zero) return "mean_zero";
  if (MeanType_ == mean_sum) return "mean_sum";
  cout << "Unknown approximation type \n";
  exit(1);
}
void GP_utils::setLikelihoodType(const int val)
{
  likelihoodType_ = val;
  if (val == likeL_Gaussian) {
    hyperlf.zeros(1, 1);
    g_hyperlf.zeros(1, 1);
  }
}

// ---------------------------------------------------------------------------------------------------
// parameter vector: kernel parameters, then likelihood, then mean-function parameters (GP_Utils.cpp:101-157)
// ---------------------------------------------------------------------------------------------------
unsigned int GP_utils::getNumPars() const { return KerenlW->getNPars() + getNumMFpar() + getNumlikfpar(); }

void GP_utils::get_GP_Pars(mat& param) const
{
  unsigned int c = 0;
  for (unsigned int i = 0; i < KerenlW->getNPars(); i++) param(c++) = KerenlW->getParam(i);
  for (unsigned int i = 0; i < getNumlikfpar(); i++) param(c++) = hyperlf(i);
  for (unsigned int i = 0; i < getNumMFpar(); i++) param(c++) = hypermf(i);
}

void GP_utils::set_GP_Pars(mat& param) const
{
  unsigned int c = 0;
  for (unsigned int i = 0; i < KerenlW->getNPars(); i++) KerenlW->setParam(param(c++), i);
  for (unsigned int i = 0; i < getNumlikfpar(); i++) hyperlf(i) = param(c++);
  for (unsigned int i = 0; i < getNumMFpar(); i++) hypermf(i) = param(c++);
  dirty = true;                    // setKUpdateStat(false): everything is recomputed at the next evaluation
}

// ---------------------------------------------------------------------------------------------------
// device plumbing
// ---------------------------------------------------------------------------------------------------
void GP_utils::device_failure(const char* what) const
{
  cout << what << " failed: " << gpss_last_error() << "\n";
  exit(1);
}

// Members of the Hyb covariance as the device sees them: at most one distance-based member (the "main" kernel of the C ABI),
// any number of Bias members (their parameters add up in the Sigma_Bias slot) and of White members (gpss_set_white).
struct MemberMap {
  int main_member;          // index of the first ExpAns | Exp | RBF member, -1 if none
  int main_first;           // its first parameter in the concatenated vector
  int kind;                 // GPSS_KERNEL_* of it (ExpAns with Sigma = 0 stands in when there is none)
  int second_member;        // index of a second distance-based member, -1 if none (gpss_set_kernel2)
  int second_first;
  int kind2;
  bool ok;                  // false: something this build does not evaluate (three distance members, an unknown member)
};
static MemberMap map_members(const Kernels* K)
{
  MemberMap m = {-1, 0, GPSS_KERNEL_EXPANS, -1, 0, -1, true};
  const mainKernel* hyb = dynamic_cast<const mainKernel*>(K);
  if (!hyb || K->getKerName() != "Hyb" || hyb->getNumKerns() < 1) { m.ok = false; return m; }
  int first = 0;
  for (unsigned int i = 0; i < hyb->getNumKerns(); i++) {
    const string name = hyb->getKern(i)->getKerName();
    const int kind = name == "ExpAns" ? GPSS_KERNEL_EXPANS : name == "Exp" ? GPSS_KERNEL_EXP : name == "RBF" ? GPSS_KERNEL_RBF : -1;
    if (kind >= 0) {
      if (m.main_member < 0) { m.main_member = (int)i; m.main_first = first; m.kind = kind; }
      else if (m.second_member < 0) { m.second_member = (int)i; m.second_first = first; m.kind2 = kind; }
      else m.ok = false;                                     // a sum of three distance-based kernels
    } else if (name != "Bias" && name != "White Noise") {
      m.ok = false;
    }
    first += (int)hyb->getKern(i)->getNPars();
  }
  return m;
}

void GP_utils::check_supported() const
{
  const char* why = 0;
  if (!KerenlW) why = "no kernel";
  else if (!map_members(KerenlW).ok)
    why = "the kernel must be Hyb{at most two of ExpAns | Exp | RBF, any number of Bias and White members} (-k ExpAns|Exp|RBF|Bias|White, -kn 0|1); "
          "a sum of three distance-based kernels is not evaluated by this build";
  else if (Xinp.n_cols != 3 && Xinp.n_cols != 4) why = "inputs must have 3 columns, or 4 with the rock-type column";
  else if (yTarg.n_cols != 1 || getOutDim() != 1) why = "exactly one output column is supported";
  else if (likelihoodType_ != likeL_Gaussian || getNumlikfpar() != 1) why = "only the Gaussian likelihood is supported";
  else if (MeanType_ != mean_zero || getNumMFpar() != 0) why = "only the zero mean function is supported";
  else if (Xinp.n_rows != yTarg.n_rows || Xinp.n_rows < 2) why = "inconsistent training data";
  if (why) {
    cout << "GP_utils: configuration outside the B200 hot path (" << why << "); there is no CPU fallback.\n";
    exit(1);
  }
}

// GPSS_KERNEL_* of the Hyb kernel's distance-based member (ExpAns when there is none: it then runs with Sigma = 0), -1 if unsupported
int GP_utils::kernel_kind() const
{
  const MemberMap m = map_members(KerenlW);
  return m.ok ? m.kind : -1;
}

// sum of the White members' Sigma_White
double GP_utils::white_now() const
{
  const mainKernel* hyb = dynamic_cast<const mainKernel*>(KerenlW);
  double w = 0.0;
  for (unsigned int i = 0; i < hyb->getNumKerns(); i++)
    if (hyb->getKern(i)->getKerName() == "White Noise") w += hyb->getKern(i)->getParam(0);
  return w;
}

// The C ABI's slot layout (include/gpss.h, gpss_set_kernel): main-kernel parameters, Sigma_Bias (the Bias members' sum; 0 without
// one), sn2.  Without a distance-based member the ExpAns slots hold the class defaults with Sigma = 0: K = bias (+ white).
void GP_utils::theta_now(double theta[GPSS_NPAR]) const
{
  const mainKernel* hyb = dynamic_cast<const mainKernel*>(KerenlW);
  const MemberMap m = map_members(KerenlW);
  for (int i = 0; i < GPSS_NPAR; i++) theta[i] = 0.0;
  unsigned int nk;
  if (m.main_member >= 0) {
    nk = hyb->getKern(m.main_member)->getNPars();
    for (unsigned int i = 0; i < nk; i++) theta[i] = KerenlW->getParam(m.main_first + i);
  } else {
    Kern_ExpAnisotropic dflt;
    nk = dflt.getNPars();
    for (unsigned int i = 0; i < nk; i++) theta[i] = dflt.getParam(i);
    theta[6] = 0.0;                                          // Sigma_ExpAns
  }
  double bias = 0.0;
  for (unsigned int i = 0; i < hyb->getNumKerns(); i++)
    if (hyb->getKern(i)->getKerName() == "Bias") bias += hyb->getKern(i)->getParam(0);
  theta[nk] = bias;
  theta[nk + 1] = hyperlf(0);
}

void GP_utils::sync_device() const
{
  check_supported();
  const int n = (int)Xinp.n_rows;
  if (handle && handle_n != n) {
    gpss_destroy(handle);
    handle = 0;
  }
  if (!handle) {
    if (gpss_create(device, n, (int)Xinp.n_cols, Xinp.memptr(), yTarg.memptr(), &handle) != GPSS_OK) device_failure("gpss_create");
    if (gpss_host::world() > 1) {
      // one process per GPU (DistHost.h): every rank holds the same data; from here on the objective / gradient /
      // prediction calls are collective and return identical results everywhere
      static int communicators = 0;
      unsigned char id[128];
      const std::string idf = gpss_host::id_file(communicators++);
      if (gpss_host::rank() == 0) {
        if (gpss_nccl_unique_id(id) != GPSS_OK) device_failure("gpss_nccl_unique_id");
        if (!gpss_host::publish_id(idf, id)) { cout << "GP_utils: cannot write the rendezvous file " << idf << "\n"; exit(1); }
      } else if (!gpss_host::fetch_id(idf, id)) {
        cout << "GP_utils: rank 0 did not publish " << idf << "\n";
        exit(1);
      }
      if (gpss_dist_init(handle, gpss_host::rank(), gpss_host::world(), id) != GPSS_OK) device_failure("gpss_dist_init");
      if (gpss_host::rank() == 0) std::remove(idf.c_str());
    }
    handle_n = n;
    kind_dev = GPSS_KERNEL_EXPANS;       // a new handle starts with the default kernel
    kind2_dev = -1;
    data_stale = false;
    dirty = true;
  } else if (data_stale) {
    if (gpss_set_data(handle, Xinp.memptr(), yTarg.memptr()) != GPSS_OK) device_failure("gpss_set_data");
    data_stale = false;
    dirty = true;
  }
  if (kernel_kind() != kind_dev) {
    if (gpss_set_kernel(handle, kernel_kind()) != GPSS_OK) device_failure("gpss_set_kernel");
    kind_dev = kernel_kind();
    dirty = true;
  }
  if (gpss_set_white(handle, white_now(), white_cross) != GPSS_OK) device_failure("gpss_set_white");   // invalidates only on change
  {
    // second distance-based member (gpss_set_kernel2 / gpss_set_theta2): sent when it changes
    const MemberMap mm = map_members(KerenlW);
    const mainKernel* hyb = dynamic_cast<const mainKernel*>(KerenlW);
    double t2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (mm.second_member >= 0)
      for (unsigned int i = 0; i < hyb->getKern(mm.second_member)->getNPars() && i < 8; i++) t2[i] = KerenlW->getParam(mm.second_first + i);
    if (mm.kind2 != kind2_dev) {
      if (gpss_set_kernel2(handle, mm.kind2) != GPSS_OK) device_failure("gpss_set_kernel2");
      kind2_dev = mm.kind2;
      dirty = true;
    }
    if (mm.second_member >= 0 && (dirty || std::memcmp(t2, theta2_dev, sizeof t2) != 0)) {
      if (gpss_set_theta2(handle, t2) != GPSS_OK) device_failure("gpss_set_theta2");
      std::memcpy(theta2_dev, t2, sizeof t2);
      dirty = true;
    }
  }
  double theta[GPSS_NPAR];
  theta_now(theta);
  if (dirty || std::memcmp(theta, theta_dev, sizeof theta) != 0) {
    if (gpss_set_theta(handle, theta) != GPSS_OK) device_failure("gpss_set_theta");
    std::memcpy(theta_dev, theta, sizeof theta);
    dirty = false;
  }
}

// GPSS_TIMING: how many evaluations an optimiser run asked of the device, and their device time (reported by OptimisePars on rank 0)
namespace {
struct EvalStats { unsigned long objective = 0, gradient = 0; double device_ms = 0.0; } g_eval_stats;
void count_eval(gpss_handle h, bool gradient)
{
  double ms = 0.0;
  if (h) gpss_get_last_call_ms(h, &ms);
  (gradient ? g_eval_stats.gradient : g_eval_stats.objective)++;
  g_eval_stats.device_ms += ms;
}
}  // namespace

double GP_utils::lastDeviceMs() const
{
  double ms = 0.0;
  if (handle) gpss_get_last_call_ms(handle, &ms);
  return ms;
}

// ---------------------------------------------------------------------------------------------------
// objective and gradient
// ---------------------------------------------------------------------------------------------------
double GP_utils::logLikelihood() const
{
  sync_device();
  double nlml = 0.0;
  const int rc = gpss_nlml(handle, &nlml);
  if (rc < 0) device_failure("gpss_nlml");
  count_eval(handle, false);
  Chol_fail = (rc == GPSS_NOT_POSDEF);
  if (Chol_fail) return std::numeric_limits<double>::quiet_NaN();       // GP_Utils.cpp:1145-1146, 1155-1158
  L.zeros(1, 1);
  L[0] = nlml;
  return nlml;
}

void GP_utils::updateAlpha() const
{
  sync_device();
  Alpha.set_size(Xinp.n_rows, 1);
  yhat.set_size(Xinp.n_rows, 1);
  const int rc = gpss_get_alpha(handle, Alpha.memptr());
  if (rc < 0) device_failure("gpss_get_alpha");
  Chol_fail = (rc == GPSS_NOT_POSDEF);
  if (!Chol_fail && gpss_get_yhat(handle, yhat.memptr()) < 0) device_failure("gpss_get_yhat");
}

double GP_utils::GradLL(mat& g) const
{
  sync_device();
  double nlml = 0.0, gv[GPSS_NPAR];
  const int rc = gpss_nlml_grad(handle, &nlml, gv);
  if (rc < 0) device_failure("gpss_nlml_grad");
  count_eval(handle, true);
  Chol_fail = (rc == GPSS_NOT_POSDEF);
  if (Chol_fail) return std::numeric_limits<double>::quiet_NaN();       // g is left untouched, as in the reference (:1175-1176)
  L.zeros(1, 1);
  L[0] = nlml;
  // same packing as GP_Utils.cpp:1243-1260: kernel entries member by member (HybKerns::getGradients, Kernel.cpp:156-169), likelihood
  // entry, mean entries.  Device slots: main-kernel entries, then trace(QW) (every Bias member's entry), then the sn2 entry.
  const mainKernel* hyb = dynamic_cast<const mainKernel*>(KerenlW);
  const MemberMap mm = map_members(KerenlW);
  const unsigned int nk = (mm.main_member >= 0) ? hyb->getKern(mm.main_member)->getNPars() : Kern_ExpAnisotropic().getNPars();
  double gv2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (mm.second_member >= 0 && gpss_get_grad2(handle, gv2) != GPSS_OK) device_failure("gpss_get_grad2");
  g_param.set_size(1, KerenlW->getNPars());
  unsigned int at = 0;
  for (unsigned int i = 0; i < hyb->getNumKerns(); i++) {
    const Kernels* k = hyb->getKern(i);
    if ((int)i == mm.main_member) for (unsigned int q = 0; q < nk; q++) g_param(at + q) = gv[q];
    else if ((int)i == mm.second_member) for (unsigned int q = 0; q < k->getNPars(); q++) g_param(at + q) = gv2[q];
    else if (k->getKerName() == "Bias") g_param(at) = gv[nk];
    else g_param(at) = 0.0;                                  // White: getGradParam (Kernel.cpp:265-269); see Kernel.h on the reference's crash
    at += k->getNPars();
  }
  g_hyperlf.zeros(1, 1);
  g_hyperlf(0) = gv[nk + 1];                                                      // the slot after Sigma_Bias (include/gpss.h)
  unsigned int c = 0;
  for (unsigned int i = 0; i < KerenlW->getNPars(); i++) g(0, c++) = g_param(0, i);
  for (unsigned int i = 0; i < getNumlikfpar(); i++) g(0, c++) = g_hyperlf(0, i);
  return nlml;
}

// ---------------------------------------------------------------------------------------------------
// prediction
// ---------------------------------------------------------------------------------------------------
void GP_utils::posteriorMeanVar(mat& mu, mat& varSigma, const mat& X) const
{
  if (X.n_cols != Xinp.n_cols) { cout << "GP_utils: test inputs must have as many columns as the training inputs\n"; exit(1); }
  white_cross = (X.n_rows == Xinp.n_rows && X(0) == Xinp(0)) ? 1 : 0;        // Kern_White::computeK's test on (X_train, X_test), Kernel.cpp:261-262
  sync_device();
  mu.set_size(X.n_rows, 1);
  varSigma.set_size(X.n_rows, 1);
  const int rc = gpss_predict(handle, (long)X.n_rows, X.memptr(), mu.memptr(), varSigma.memptr());
  if (rc == GPSS_NOT_POSDEF) {                     // the reference carries NaN through here
    mu.fill(std::numeric_limits<double>::quiet_NaN());
    varSigma.fill(std::numeric_limits<double>::quiet_NaN());
  } else if (rc != GPSS_OK) {
    device_failure("gpss_predict");
  }
}

void GP_utils::posteriorMean(mat& mu, const mat& X) const
{
  if (X.n_cols != Xinp.n_cols) { cout << "GP_utils: test inputs must have as many columns as the training inputs\n"; exit(1); }
  white_cross = (X.n_rows == Xinp.n_rows && X(0) == Xinp(0)) ? 1 : 0;
  sync_device();
  mu.set_size(X.n_rows, 1);
  const int rc = gpss_predict(handle, (long)X.n_rows, X.memptr(), mu.memptr(), 0);
  if (rc == GPSS_NOT_POSDEF) mu.fill(std::numeric_limits<double>::quiet_NaN());
  else if (rc != GPSS_OK) device_failure("gpss_predict");
}

// [quirk kept] the one-output overload also computes the variance and throws it away (GP_Utils.cpp:159-165)
void GP_utils::Calc_Out(mat& yPred, const mat& Xin) const
{
  mat var;
  posteriorMeanVar(yPred, var, Xin);
}
void GP_utils::Calc_Out(mat& yPred, mat& yVar, const mat& Xin) const { posteriorMeanVar(yPred, yVar, Xin); }
void GP_utils::Calc_Out(mat& yPred, mat& /*probPred*/, mat& yVar, const mat& Xin) const { posteriorMeanVar(yPred, yVar, Xin); }

// ---------------------------------------------------------------------------------------------------
// optimisation entry point and report (GP_Utils.cpp:1288-1322)
// ---------------------------------------------------------------------------------------------------
void GP_utils::OptimisePars(unsigned int iters)
{
  if (getVerbose() > 2) {
    cout << "Initial model:" << endl;
    ShowKernelPars(cout);
  }
  // [quirk] the iteration count only takes effect at verbosity > 2: a dangling `if` guards setMaxIters (GP_Utils.cpp:1295-1296)
  if (getVerbose() > 2 && getNumPars() < 40) setMaxIters(iters);
  const EvalStats before = g_eval_stats;
  Optimise();
  if (std::getenv("GPSS_TIMING") && gpss_host::rank() == 0)
    std::fprintf(stderr, "[gpss timing] optimiser: %lu objective + %lu objective-and-gradient calls to the device, %.3f s of device time\n",
                 g_eval_stats.objective - before.objective, g_eval_stats.gradient - before.gradient,
                 (g_eval_stats.device_ms - before.device_ms) * 1e-3);
  if (getVerbose() > 0) ShowKernelPars(cout);
}

void GP_utils::ShowKernelPars(std::ostream& os) const
{
  cout << "Standard GP Model: " << endl;
  cout << "Optimiser: " << getDefaultOptimiserStr() << endl;
  cout << "Inference: " << getInferenceStr() << endl;
  cout << "likelihood function: " << getLiklihoodStr() << endl;
  cout << "Mean function: " << getMeanTypeStr() << endl;
  cout << "Data Set Size: " << getNumData() << endl;
  cout << "Kernel Type: " << endl;
  KerenlW->ShowKernelPars(os);
  for (unsigned int i = 0; i < getNumlikfpar(); i++) cout << "likelihood hyperparmeters : " << hyperlf(i) << endl;
  for (unsigned int i = 0; i < getNumMFpar(); i++) cout << "likelihood hyperparmeters : " << std::exp(hypermf(i)) << endl;
  if (getVerbose()) cout << "Log likelihood: " << logLikelihood() << endl;
}

// ---------------------------------------------------------------------------------------------------
// train_model file (GP_Utils.cpp:1324-1425): key=value lines, default stream precision (6 significant digits)
// ---------------------------------------------------------------------------------------------------
void GP_utils::ToFile_GP_Params(std::ostream& out) const
{
  out << "Inference=" << getInf() << endl;
  out << "likelihood=" << getLikelihoodType() << endl;
  out << "MeanFunction=" << getMean() << endl;
  out << "numData=" << getNumData() << endl;
  out << "outputDim=" << getOutDim() << endl;
  out << "inputDim=" << getInpDim() << endl;
  out << "NumHyperKernel=" << KerenlW->getNPars() << endl;
  out << "NumHyperLik=" << getNumlikfpar() << endl;
  out << "NumHyperMean=" << getNumMFpar() << endl;
  KerenlW->StrmOut(out);
  for (unsigned int i = 0; i < getNumlikfpar(); i++) out << "Hyperparams_likelihood=" << getHyperlfVal(i) << endl;
  for (unsigned int i = 0; i < getNumMFpar(); i++) out << "Hyperparams_meanfunction=" << std::exp(getHypermfVal(i)) << endl;
}

void GP_utils::FromFile_GP_Params(std::istream& in)
{
  setInf(ReadStrStrm(in, "Inference"));
  setLikelihoodType(ReadIntStrm(in, "likelihood"));
  setMean(ReadStrStrm(in, "MeanFunction"));
  setNumData(ReadIntStrm(in, "numData"));
  setOutDim(ReadIntStrm(in, "outputDim"));
  setInpDim(ReadIntStrm(in, "inputDim"));
  setNumCovpar(ReadIntStrm(in, "NumHyperKernel"));
  setNumlikfpar(ReadIntStrm(in, "NumHyperLik"));
  setNumMFpar(ReadIntStrm(in, "NumHyperMean"));
  initialize_vars();
  KerenlW = ReadKerFromFile(in);       // owned by nobody, as in the reference (GP_Utils.cpp:1336)
  g_param.resize(1, KerenlW->getNPars());
  if (getNumlikfpar() > 0) {
    hyperlf.resize(getNumlikfpar(), 1);
    for (unsigned int i = 0; i < getNumlikfpar(); i++) setHyperlfVal(ReadDoubleStrm(in, "Hyperparams_likelihood"), i);
  }
  if (getNumMFpar() > 0) {
    hypermf.resize(getNumMFpar(), 1);
    for (unsigned int i = 0; i < getNumMFpar(); i++) setHypermfVal(std::log((double)ReadIntStrm(in, "Hyperparams_meanfunction")), i);
  }
}

void writeGpToStream(const GP_utils& model, std::ostream& out) { model.StrmOut(out); }
void writeGPFile(const GP_utils& model, const string modelFileName, const string comment) { model.WFile(modelFileName, comment); }

GP_utils* readGpFromStream(std::istream& in)
{
  GP_utils* m = new GP_utils();
  m->StrmIn(in);
  return m;
}

GP_utils* readGpFromFile(const string modelFileName, int verbosity)
{
  if (verbosity > 0) cout << "Loading model file." << endl;
  std::ifstream in(modelFileName.c_str());
  if (!in.is_open()) {
    cout << "Error in reading file name. \n";
    exit(1);
  }
  GP_utils* m = readGpFromStream(in);
  if (verbosity > 0) cout << "Model Info has been read.\n";
  in.close();
  m->setVerbose(verbosity);
  return m;
}
