// See Kernel.h.  File:line citations are into /root/reference.
#include "Kernel.h"

#include <cmath>
#include <cstdlib>

#include "../../include/gpss.h"

using namespace arma;
using std::cout;
using std::endl;
using std::string;

namespace {
[[noreturn]] void fatal(const string& msg)
{
  cout << msg << "\n";
  exit(1);
}
void check(int rc, const char* what)
{
  if (rc != GPSS_OK) fatal(string(what) + " failed: " + gpss_last_error());
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// Kernels: text form in the model file (Kernel.cpp:20-40, 1310-1338)
//   KernelName=<name> / inputDim=<d> / numParams=<p> / one line of space-separated values, each followed by a space;
//   values equal to their integer part are printed as integers; default ostream precision (6 significant digits).
// ---------------------------------------------------------------------------------------------------
std::ostream& Kernels::ShowKernelPars(std::ostream& os) const
{
  os << getKerName() << " kernel:" << endl;
  for (unsigned int i = 0; i < nParams; i++) os << getParamName(i) << ": " << getParam(i) << endl;
  return os;
}

void Kernels::ToFile_GP_Params(std::ostream& out) const
{
  out << "KernelName=" << getKerName() << endl;
  out << "inputDim=" << getInputDim() << endl;
  out << "numParams=" << getNPars() << endl;
  for (unsigned int i = 0; i < getNPars(); i++) {
    const double val = getParam(i);
    if ((val - (int)val) == 0.0) out << (int)val << " ";
    else out << val << " ";
  }
  out << endl;
}

void Kernels::FromFile_GP_Params(std::istream& in)
{
  setInputDim(ReadIntStrm(in, "inputDim"));
  const unsigned int count = ReadIntStrm(in, "numParams");
  string line;
  if (!std::getline(in, line)) fatal("Can not read " + getKerName() + " kernel parameters. ");
  mat values = zeros<mat>(1, count);
  for (unsigned int i = 0; i < count; i++) {
    if (line.size() <= 0) fatal("The nember of Hyper-parameters of " + getKerName() + " are not sufficient. ");
    const size_t pos = line.find(" ");
    values[i] = std::atof(line.substr(0, pos + 1).c_str());
    line.erase(0, pos + 1);
  }
  setParams(values);
}

// ---------------------------------------------------------------------------------------------------
// mainKernel
// ---------------------------------------------------------------------------------------------------
bool mainKernel::locate(unsigned int paramNo, size_t& member, unsigned int& local) const
{
  unsigned int first = 0;
  for (size_t i = 0; i < MainKEl.size(); i++) {
    const unsigned int count = MainKEl[i]->getNPars();
    if (paramNo < first + count) {
      member = i;
      local = paramNo - first;
      return true;
    }
    first += count;
  }
  return false;
}

void mainKernel::setParam(double val, unsigned int paramNo)
{
  size_t m;
  unsigned int l;
  if (locate(paramNo, m, l)) MainKEl[m]->setParam(val, l);
}
double mainKernel::getParam(unsigned int paramNo) const
{
  size_t m;
  unsigned int l;
  return locate(paramNo, m, l) ? MainKEl[m]->getParam(l) : -1;
}
string mainKernel::getParamName(unsigned int paramNo) const
{
  size_t m;
  unsigned int l;
  return locate(paramNo, m, l) ? MainKEl[m]->getParamName(l) : "";
}

void mainKernel::FromFile_GP_Params(std::istream& in)
{
  const unsigned int count = ReadIntStrm(in, "NumberOfKernels");
  for (unsigned int i = 0; i < count; i++) {
    Kernels* k = ReadKerFromFile(in);
    addNewKernel(k);           // stores a clone
    delete k;
  }
}

void mainKernel::ToFile_GP_Params(std::ostream& out) const
{
  out << "KernelName=" << getKerName() << endl;
  out << "NumberOfKernels=" << getNumKerns() << endl;
  for (size_t i = 0; i < MainKEl.size(); i++) MainKEl[i]->StrmOut(out);
}

// ---------------------------------------------------------------------------------------------------
// HybKerns
// ---------------------------------------------------------------------------------------------------
HybKerns::HybKerns() : mainKernel() { _init(); }
HybKerns::HybKerns(unsigned int inDim) : mainKernel(inDim) { _init(); setInputDim(inDim); }
HybKerns::HybKerns(const mat& X) : mainKernel(X) { _init(); setInputDim(X.n_cols); }
HybKerns::HybKerns(const HybKerns& other) : mainKernel()
{
  _init();
  setInputDim(other.getInputDim());
  for (size_t i = 0; i < other.MainKEl.size(); i++) addNewKernel(other.MainKEl[i]);
}
HybKerns::~HybKerns()
{
  for (size_t i = 0; i < MainKEl.size(); i++) delete MainKEl[i];
}
void HybKerns::_init()
{
  nParams = 0;
  setKerName("Hyb");
}

double HybKerns::Diag_Kernel(const mat& X, unsigned int index) const
{
  double y = 0.0;
  for (size_t i = 0; i < MainKEl.size(); i++) y += MainKEl[i]->Diag_Kernel(X, index);
  return y;
}

void HybKerns::diag_Compute(mat& d, const mat& X) const
{
  d.zeros();
  mat part = zeros<mat>(d.n_rows, d.n_cols);
  for (size_t i = 0; i < MainKEl.size(); i++) {
    MainKEl[i]->diag_Compute(part, X);
    d += part;
  }
}

void HybKerns::computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const
{
  mat Kpart = zeros<mat>(K.n_rows, K.n_cols), Dpart = zeros<mat>(D2.n_rows, D2.n_cols);
  D2.zeros();
  K.zeros();
  for (size_t i = 0; i < MainKEl.size(); i++) {
    MainKEl[i]->computeK(X1, X2, Kpart, Dpart);
    D2 += Dpart;
    K += Kpart;
  }
}

void HybKerns::getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const
{
  unsigned int first = 0;
  for (size_t i = 0; i < MainKEl.size(); i++) {
    const unsigned int count = MainKEl[i]->getNPars();
    mat part(1, count);
    MainKEl[i]->getGradients(part, X, X2, D2, QW);
    g.submat(0, first, 0, first + count - 1) = part;
    first += count;
  }
}

// ---------------------------------------------------------------------------------------------------
// Kern_Bias
// ---------------------------------------------------------------------------------------------------
void Kern_Bias::_init()
{
  nParams = 1;
  setKerName("Bias");
  setParamName("Sigma_Bias", 0);
  setInitPars();
}
void Kern_Bias::setParam(double val, unsigned int paramNo)
{
  if (paramNo != 0) fatal("Requested parameter doesn't exist.");
  Sigma_Bias = val;
}
double Kern_Bias::getParam(unsigned int paramNo) const
{
  if (paramNo != 0) fatal("Requested parameter doesn't exist.");
  return Sigma_Bias;
}
void Kern_Bias::computeK(const mat&, const mat&, mat& K, mat& D2) const
{
  D2.zeros();
  K.fill(Sigma_Bias);
}
// vec(QW) . vec(I) = trace(QW)  (Kernel.cpp:370-377)
void Kern_Bias::getGradients(mat& g, const mat&, const mat&, const mat&, const mat& QW) const
{
  double tr = 0.0;
  for (uword i = 0; i < QW.n_rows && i < QW.n_cols; i++) tr += QW(i, i);
  g[0] = tr;
}

// ---------------------------------------------------------------------------------------------------
// Kern_White (Kernel.cpp:180-270)
// ---------------------------------------------------------------------------------------------------
void Kern_White::_init()
{
  nParams = 1;
  setKerName("White Noise");            // what the model file then carries -- and what ReadKerFromFile does NOT accept (it wants "white")
  setParamName("Sigma_White", 0);
  setInitPars();
}
void Kern_White::setParam(double val, unsigned int paramNo)
{
  if (paramNo != 0) fatal("Requested parameter doesn't exist.");
  Sigma_White = val;
}
double Kern_White::getParam(unsigned int paramNo) const
{
  if (paramNo != 0) fatal("Requested parameter doesn't exist.");
  return Sigma_White;
}
void Kern_White::computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const
{
  D2.zeros();
  K.zeros();
  if (X1(0) == X2(0) && X1.n_rows == X2.n_rows)                                   // Kernel.cpp:261-262
    for (uword i = 0; i < K.n_rows && i < K.n_cols; i++) K(i, i) = Sigma_White;
}

// ---------------------------------------------------------------------------------------------------
// Kern_ExpAnisotropic
// ---------------------------------------------------------------------------------------------------
void Kern_ExpAnisotropic::_init()
{
  nParams = 8;
  setKerName("ExpAns");
  static const char* names[8] = {"AngleX_ExpAns", "inverseWidthx_ExpAns", "AngleY_ExpAns", "inverseWidthy_ExpAns",
                                 "AngleZ_ExpAns", "inverseWidthz_ExpAns", "Sigma_ExpAns",  "InversewidthR_ExpAns"};
  for (unsigned int i = 0; i < 8; i++) setParamName(names[i], i);
  setInitPars();
}

// the reference's hard-coded starting point (Kernel.cpp:763-773)
void Kern_ExpAnisotropic::setInitPars()
{
  par[0] = M_PI / 3.1;
  par[1] = 1.5;
  par[2] = M_PI / 3.1;
  par[3] = 1.5;
  par[4] = M_PI / 3.1;
  par[5] = 1.3;
  par[6] = 0.9;
  par[7] = 0.6;
}
void Kern_ExpAnisotropic::setParam(double val, unsigned int paramNo)
{
  if (paramNo >= 8) fatal("Requested parameter doesn't exist.");
  par[paramNo] = val;
}
double Kern_ExpAnisotropic::getParam(unsigned int paramNo) const
{
  if (paramNo >= 8) fatal("Requested parameter doesn't exist.");
  return par[paramNo];
}

namespace {
void require_3d(const mat& X, const char* who)
{
  if (X.n_cols != 3 && X.n_cols != 4) fatal(string(who) + ": inputs must have 3 columns, or 4 with the rock-type column (Kernel.cpp:864-878)");
}
// theta in the C ABI's order with the ExpAns block filled in and no bias / unit noise
void theta_of(const double par[8], double theta[GPSS_NPAR])
{
  for (int i = 0; i < 8; i++) theta[i] = par[i];
  theta[8] = 0.0;
  theta[9] = 1.0;
}
}  // namespace

void Kern_ExpAnisotropic::computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const
{
  require_3d(X1, "Kern_ExpAnisotropic::computeK");
  require_3d(X2, "Kern_ExpAnisotropic::computeK");
  double theta[GPSS_NPAR];
  theta_of(par, theta);
  K.set_size(X1.n_rows, X2.n_rows);
  D2.set_size(X1.n_rows, X2.n_rows);
  if (X1.n_cols != X2.n_cols) fatal("Kern_ExpAnisotropic::computeK: X1 and X2 must have the same number of columns");
  check(gpss_compute_K(0, GPSS_KERNEL_EXPANS, theta, (int)X1.n_cols, (int)X1.n_rows, X1.memptr(), (int)X2.n_rows, X2.memptr(), K.memptr(), D2.memptr()), "gpss_compute_K");
}

void Kern_ExpAnisotropic::getGradients(mat& g, const mat& X, const mat& X2, const mat&, const mat& QW) const
{
  require_3d(X, "Kern_ExpAnisotropic::getGradients");
  if (X2.n_rows != X.n_rows || QW.n_rows != X.n_rows || QW.n_cols != X.n_rows)
    fatal("Kern_ExpAnisotropic::getGradients: the B200 path implements the X2 == X (training) case only");
  double theta[GPSS_NPAR], g8[8];
  theta_of(par, theta);
  check(gpss_expans_gradients(0, theta, (int)X.n_cols, (int)X.n_rows, X.memptr(), QW.memptr(), g8), "gpss_expans_gradients");
  for (int i = 0; i < 8; i++) g[i] = g8[i];
}

// ---------------------------------------------------------------------------------------------------
// Kern_Exponential, Kern_RBF: host-matrix computeK through the device (gpss_compute_K with the kernel kind, bias 0, unit noise).
// The members' getGradients with a HOST QW are not needed by GP_utils (GradLL runs device-resident, gpss_nlml_grad) and are
// not provided for these two kernels.
// ---------------------------------------------------------------------------------------------------
namespace {
void iso_compute(int kind, const double* par, int npar, const mat& X1, const mat& X2, mat& K, mat& D2, const char* who)
{
  require_3d(X1, who);
  require_3d(X2, who);
  if (X1.n_cols != X2.n_cols) fatal(string(who) + ": X1 and X2 must have the same number of columns");
  double theta[GPSS_NPAR];
  for (int i = 0; i < GPSS_NPAR; i++) theta[i] = 0.0;
  for (int i = 0; i < npar; i++) theta[i] = par[i];
  theta[npar] = 0.0;          // Sigma_Bias
  theta[npar + 1] = 1.0;      // sn2 (unused by computeK)
  K.set_size(X1.n_rows, X2.n_rows);
  D2.set_size(X1.n_rows, X2.n_rows);
  check(gpss_compute_K(0, kind, theta, (int)X1.n_cols, (int)X1.n_rows, X1.memptr(), (int)X2.n_rows, X2.memptr(), K.memptr(), D2.memptr()),
        "gpss_compute_K");
}
}  // namespace

void Kern_Exponential::_init()
{
  nParams = 2;
  setKerName("Exp");
  setParamName("Hayper_Euc_Exp", 0);
  setParamName("Sigma_Exp", 1);
  setInitPars();
}
void Kern_Exponential::setParam(double val, unsigned int paramNo)
{
  if (paramNo >= 2) fatal("Requested parameter doesn't exist.");
  par[paramNo] = val;
}
double Kern_Exponential::getParam(unsigned int paramNo) const
{
  if (paramNo >= 2) fatal("Requested parameter doesn't exist.");
  return par[paramNo];
}
void Kern_Exponential::computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const
{
  iso_compute(GPSS_KERNEL_EXP, par, 2, X1, X2, K, D2, "Kern_Exponential::computeK");
}
void Kern_Exponential::getGradients(mat&, const mat&, const mat&, const mat&, const mat&) const
{
  fatal("Kern_Exponential::getGradients with host matrices is not provided: GP_utils::GradLL computes the gradient on the device");
}

void Kern_RBF::_init()
{
  nParams = 3;
  setKerName("RBF");
  setParamName("Hayper_Euc_RBF", 0);
  setParamName("inverseWidth_RBF", 1);
  setParamName("Sigma_RBF", 2);
  setInitPars();
}
void Kern_RBF::setParam(double val, unsigned int paramNo)
{
  if (paramNo >= 3) fatal("Requested parameter doesn't exist.");
  par[paramNo] = val;
}
double Kern_RBF::getParam(unsigned int paramNo) const
{
  if (paramNo >= 3) fatal("Requested parameter doesn't exist.");
  return par[paramNo];
}
void Kern_RBF::computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const
{
  iso_compute(GPSS_KERNEL_RBF, par, 3, X1, X2, K, D2, "Kern_RBF::computeK");
}
void Kern_RBF::getGradients(mat&, const mat&, const mat&, const mat&, const mat&) const
{
  fatal("Kern_RBF::getGradients with host matrices is not provided: GP_utils::GradLL computes the gradient on the device");
}

// ---------------------------------------------------------------------------------------------------
// model-file helpers (Kernel.cpp:1281-1307)
// ---------------------------------------------------------------------------------------------------
void WriteKernelPas(const Kernels& kern, std::ostream& out) { kern.StrmOut(out); }

Kernels* ReadKerFromFile(std::istream& in)
{
  string line;
  std::getline(in, line);                       // raw getline: '#' lines are NOT skipped here, as in the reference
  const string name = line.substr(line.find("=") + 1);
  Kernels* k = 0;
  if (name == "Bias") k = new Kern_Bias();
  else if (name == "ExpAns") k = new Kern_ExpAnisotropic();
  else if (name == "Hyb") k = new HybKerns();
  else if (name == "Exp") k = new Kern_Exponential();
  else if (name == "RBF") k = new Kern_RBF();
  else if (name == "white") k = new Kern_White();          // [quirk] the WRITER emits "White Noise" (getKerName), which lands in the branch below
  else fatal("Unknown kernel type ");
  k->FromFile_GP_Params(in);
  return k;
}
