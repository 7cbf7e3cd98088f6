// Covariance-function classes of the B200 exact-GP path, mirroring the reference's Kernel.h (/root/reference):
// abstract `Kernels`, the additive container `mainKernel` / `HybKerns`, the anisotropic exponential kernel
// `Kern_ExpAnisotropic` (3 rotation angles + 3 inverse widths + Sigma [+ the 4-th column width]) and `Kern_Bias`.
// They own parameters, names and the text (de)serialisation of the train_model file.  The matrix-valued members
// computeK / getGradients are COMPATIBILITY wrappers: they move host matrices through the C ABI
// (gpss_compute_K / gpss_expans_gradients); GP_utils never uses them on the hot path -- it hands the parameter
// vector to the device-resident handle instead.
// Additive members without a distance (Kern_Bias, Kern_White) may appear any number of times next to at most ONE distance-based
// member (ExpAns | Exp | RBF); sums of two distance-based members are outside this build (GP_utils::check_supported says so).
#ifndef GPSS_HOST_KERNEL_H
#define GPSS_HOST_KERNEL_H

#include <armadillo>
#include <iostream>
#include <string>
#include <vector>

#include "ModelInf.h"

using arma::mat;

class Kernels : public StreamIntfce {
 public:
  Kernels() : nParams(0), inputDim(0) {}
  explicit Kernels(const mat&) : nParams(0), inputDim(0) {}
  explicit Kernels(unsigned int) : nParams(0), inputDim(0) {}
  virtual ~Kernels() {}

  virtual Kernels* clone() const = 0;
  virtual void setInitPars() = 0;
  virtual double Diag_Kernel(const mat& X, unsigned int index) const = 0;
  virtual void diag_Compute(mat& d, const mat& X) const
  {
    for (unsigned int i = 0; i < X.n_rows; i++) d(i) = Diag_Kernel(X, i);
  }
  virtual void setParam(double, unsigned int) = 0;
  virtual double getParam(unsigned int) const = 0;
  // K (n1 x n2) and the squared Mahalanobis distance D2 between the rows of X1 and X2
  virtual void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const = 0;
  // g (1 x nParams) = sum_ij QW_ij dK_ij/dparam in the reference's own (non-textbook) form
  virtual void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const = 0;
  virtual unsigned int addNewKernel(const Kernels*)
  {
    std::cerr << "Error in adding new kernel." << std::endl;
    return 0;
  }

  void setParams(const mat& v) { for (unsigned int i = 0; i < nParams; i++) setParam(v(i), i); }
  void getParams(mat& v) const { for (unsigned int i = 0; i < nParams; i++) v(i) = getParam(i); }
  std::string getKerName() const { return kernName; }
  void setKerName(const std::string name) { kernName = name; }
  void setInputDim(unsigned int dim) { inputDim = dim; }
  unsigned getInputDim() const { return inputDim; }
  unsigned int getNPars() const { return nParams; }
  void setNPars(unsigned int np) { nParams = np; }
  void setParamName(const std::string name, unsigned int index)
  {
    if (paramNames.size() <= index) paramNames.resize(index + 1, "no name");
    paramNames[index] = name;
  }
  virtual std::string getParamName(unsigned int index) const { return paramNames[index]; }

  virtual void ToFile_GP_Params(std::ostream& out) const;
  virtual void FromFile_GP_Params(std::istream& in);
  virtual std::ostream& ShowKernelPars(std::ostream& os) const;
  void GetGrads(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const { getGradients(g, X, X2, D2, QW); }

 protected:
  unsigned int nParams;
  std::string kernName;
  std::vector<std::string> paramNames;

 private:
  unsigned int inputDim;
};

// container that concatenates the parameter vectors of its members
class mainKernel : public Kernels {
 public:
  mainKernel() : Kernels() {}
  explicit mainKernel(unsigned int inDim) : Kernels(inDim) {}
  explicit mainKernel(const mat& X) : Kernels(X) {}
  virtual unsigned int addNewKernel(const Kernels* kern)
  {
    MainKEl.push_back(kern->clone());
    nParams += kern->getNPars();
    return MainKEl.size() - 1;
  }
  virtual void setParam(double val, unsigned int paramNo);
  virtual double getParam(unsigned int paramNo) const;
  virtual std::string getParamName(unsigned int paramNo) const;
  virtual void FromFile_GP_Params(std::istream& in);
  virtual void ToFile_GP_Params(std::ostream& out) const;
  virtual unsigned int getNumKerns() const { return MainKEl.size(); }
  const Kernels* getKern(unsigned int i) const { return MainKEl[i]; }

 protected:
  // member index and local parameter index of global parameter `paramNo`; false when out of range
  bool locate(unsigned int paramNo, size_t& member, unsigned int& local) const;
  std::vector<Kernels*> MainKEl;
};

// additive ("hybrid") kernel: K = sum of the members' K
class HybKerns : public mainKernel {
 public:
  HybKerns();
  explicit HybKerns(unsigned int inDim);
  explicit HybKerns(const mat& X);
  HybKerns(const HybKerns&);          // deep copy (the reference's copy constructor appends to the list it iterates)
  ~HybKerns();
  HybKerns* clone() const { return new HybKerns(*this); }

  void setInitPars() {}
  double Diag_Kernel(const mat& X, unsigned int index) const;
  void diag_Compute(mat& d, const mat& X) const;
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const;

 private:
  HybKerns& operator=(const HybKerns&);
  void _init();
};

// constant kernel K_ij = Sigma_Bias on EVERY element (reference Kernel.cpp:285-377)
class Kern_Bias : public Kernels {
 public:
  Kern_Bias() : Kernels() { _init(); }
  explicit Kern_Bias(unsigned int inDim) : Kernels(inDim) { _init(); setInputDim(inDim); }
  explicit Kern_Bias(const mat& X) : Kernels(X) { _init(); setInputDim(X.n_cols); }
  Kern_Bias* clone() const { return new Kern_Bias(*this); }

  void setInitPars() { Sigma_Bias = 0.2; }
  double Diag_Kernel(const mat&, unsigned int) const { return Sigma_Bias; }
  void diag_Compute(mat& d, const mat&) const { d.fill(Sigma_Bias); }
  void setParam(double val, unsigned int paramNo);
  double getParam(unsigned int paramNo) const;
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const;

 private:
  void _init();
  double Sigma_Bias;
};

// white-noise kernel (reference Kernel.cpp:180-270, Kernel.h:256-282): K = Sigma_White on the diagonal, and only when computeK is
// handed the same point set twice -- the reference tests `X1(0) == X2(0) && X1.n_rows == X2.n_rows` (:261-262).
// [reference defect, reproduced as far as it can be] Kern_White does not override getGradients, and the base-class default calls
// ITSELF (Kernel.h:56-59): any optimiser run with a White member overflows the stack (the compiled reference: SIGSEGV right after
// "Log likelihood" of the initial model).  Objective and prediction work and are matched; for the gradient this build returns what
// the class's own getGradParam returns, 0 (Kernel.cpp:265-269), so an optimiser leaves Sigma_White where it started.
class Kern_White : public Kernels {
 public:
  Kern_White() : Kernels() { _init(); }
  explicit Kern_White(unsigned int inDim) : Kernels(inDim) { _init(); setInputDim(inDim); }
  explicit Kern_White(const mat& X) : Kernels(X) { _init(); setInputDim(X.n_cols); }
  Kern_White* clone() const { return new Kern_White(*this); }

  void setInitPars() { Sigma_White = 0.10; }                       // Kernel.cpp:217-220
  double Diag_Kernel(const mat&, unsigned int) const { return Sigma_White; }
  void diag_Compute(mat& d, const mat&) const { d.fill(Sigma_White); }
  void setParam(double val, unsigned int paramNo);
  double getParam(unsigned int paramNo) const;
  double getGradParam(unsigned int, const mat&, const mat&, const mat&, const mat&) const { return 0.0; }
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat&, const mat&, const mat&, const mat&) const { g[0] = 0.0; }

 private:
  void _init();
  double Sigma_White;
};

// K_ij = Sigma^2 exp(-sqrt(D2_ij)),  D2 = (x - x')' S^2 (x - x'),  S = Rot(AngleX, AngleY, AngleZ) diag(iWx, iWy, iWz) Rot'
// (reference Kernel.cpp:700-1263, 1370-1435)
class Kern_ExpAnisotropic : public Kernels {
 public:
  Kern_ExpAnisotropic() : Kernels() { _init(); }
  explicit Kern_ExpAnisotropic(unsigned int inDim) : Kernels(inDim) { _init(); setInputDim(inDim); }
  explicit Kern_ExpAnisotropic(const mat& X) : Kernels(X) { _init(); setInputDim(X.n_cols); }
  Kern_ExpAnisotropic* clone() const { return new Kern_ExpAnisotropic(*this); }

  void setAngleX(double v) { par[0] = v; }
  double getAngleX() const { return par[0]; }
  void setInverseWidthx(double v) { par[1] = v; }
  double getInverseWidthx() const { return par[1]; }
  void setAngleY(double v) { par[2] = v; }
  double getAngleY() const { return par[2]; }
  void setInverseWidthy(double v) { par[3] = v; }
  double getInverseWidthy() const { return par[3]; }
  void setAngleZ(double v) { par[4] = v; }
  double getAngleZ() const { return par[4]; }
  void setInverseWidthz(double v) { par[5] = v; }
  double getInverseWidthz() const { return par[5]; }

  void setInitPars();
  double Diag_Kernel(const mat&, unsigned int) const { return par[6] * par[6]; }
  void diag_Compute(mat& d, const mat&) const { d.fill(par[6] * par[6]); }
  void setParam(double val, unsigned int paramNo);
  double getParam(unsigned int paramNo) const;
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const;

 private:
  void _init();
  // AngleX, iWx, AngleY, iWy, AngleZ, iWz, Sigma, iWR -- the order of the parameter vector (Kernel.cpp:803-838)
  double par[8];
};

// Isotropic members of the family.  Distance: EuclDist, D2 = |x - x'|^2 / Hayper_Euc^2
// (reference Kernel.cpp:1343-1368, 1437-1441).  On the device they run through the ExpAns pair-distance code with
// sigInv = (1/Hayper_Euc) I (gpss_set_kernel, include/gpss.h).
// K_ij = Sigma_Exp^2 exp(-sqrt(D2_ij))   (reference Kernel.cpp:544-695)
class Kern_Exponential : public Kernels {
 public:
  Kern_Exponential() : Kernels() { _init(); }
  explicit Kern_Exponential(unsigned int inDim) : Kernels(inDim) { _init(); setInputDim(inDim); }
  explicit Kern_Exponential(const mat& X) : Kernels(X) { _init(); setInputDim(X.n_cols); }
  Kern_Exponential* clone() const { return new Kern_Exponential(*this); }

  void setInitPars() { par[0] = 0.5; par[1] = 0.9; }              // Hayper_Euc_Exp, Sigma_Exp (Kernel.cpp:585-589)
  double Diag_Kernel(const mat&, unsigned int) const { return par[1] * par[1]; }
  void diag_Compute(mat& d, const mat&) const { d.fill(par[1] * par[1]); }
  void setParam(double val, unsigned int paramNo);
  double getParam(unsigned int paramNo) const;
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const;

 private:
  void _init();
  double par[2];
};

// K_ij = exp(-0.5 inverseWidth_RBF D2_ij) Sigma_RBF^2   (reference Kernel.cpp:380-541)
class Kern_RBF : public Kernels {
 public:
  Kern_RBF() : Kernels() { _init(); }
  explicit Kern_RBF(unsigned int inDim) : Kernels(inDim) { _init(); setInputDim(inDim); }
  explicit Kern_RBF(const mat& X) : Kernels(X) { _init(); setInputDim(X.n_cols); }
  Kern_RBF* clone() const { return new Kern_RBF(*this); }

  void setInitPars() { par[0] = 0.5; par[1] = 0.9; par[2] = 0.5; }   // Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF (Kernel.cpp:425-431)
  double Diag_Kernel(const mat&, unsigned int) const { return par[2] * par[2]; }
  void diag_Compute(mat& d, const mat&) const { d.fill(par[2] * par[2]); }
  void setParam(double val, unsigned int paramNo);
  double getParam(unsigned int paramNo) const;
  void computeK(const mat& X1, const mat& X2, mat& K, mat& D2) const;
  void getGradients(mat& g, const mat& X, const mat& X2, const mat& D2, const mat& QW) const;

 private:
  void _init();
  double par[3];
};

void WriteKernelPas(const Kernels& kern, std::ostream& out);
Kernels* ReadKerFromFile(std::istream& in);

#endif
