// GP_utils: the exact-GP model object of GP_SS_AK with the reference's public surface (GP_Utils.h in /root/reference:
// same class name, bases, enums, virtuals, public data members Xinp / yTarg / Alpha / yhat / hyperlf / KerenlW and
// free functions writeGPFile / readGpFromFile), but every O(n^2)-and-up operation runs on a B200 behind the C ABI of
// include/gpss.h.  The n x n members of the reference (K, D2, Lchol, Q, QW, ...) have no host copy here: they live in
// HBM inside the gpss handle (K -> L in place, U = L^-T, B^-1), see DESIGN.md.
//
// Protocol (unchanged): set_GP_Pars(x) is cheap and only invalidates; ObjVal()/logLikelihood() and
// Grad_Values(g)/GradLL(g) return the NEGATIVE log marginal likelihood (NaN when the Cholesky factorisation fails);
// an objective asked again at the same parameters costs nothing.
//
// Supported configuration = the scope table: HybKerns{Kern_ExpAnisotropic, Kern_Bias}, 3-column inputs, one output,
// Gaussian likelihood, zero mean.  Anything else stops with a message (there is no CPU fallback).
#ifndef GPSS_HOST_GP_UTILS_H
#define GPSS_HOST_GP_UTILS_H

#include <armadillo>
#include <iostream>
#include <string>

#include "../../include/gpss.h"
#include "Kernel.h"
#include "ModelInf.h"
#include "Opt_pars.h"

class GP_utils : public Modeling, public Main_Opt_Algs, public StreamIntfce {
 public:
  enum likelihoodType { likeL_Gaussian, likeL_WarpGauss };
  enum InferenceType { inf_laplace, inf_EP };
  enum MeanType { mean_zero, mean_sum };

  GP_utils();
  GP_utils(Kernels* kernel, mat Xin, mat Yin, int Inf_type = inf_laplace, int likeLtype = likeL_Gaussian, int mean_type = mean_zero,
           unsigned int numhyper = 1, unsigned int numlik_par = 1, unsigned int numMF_par = 0, int verbos = 2);
  ~GP_utils();

  // (re)binds the device state to Xinp / yTarg of the current getNumData(); call after assigning them (gp_ss_ak.cpp:389-395)
  void initialize_vars();

  // predictive mean [and variance] of the rows of inData, exactly as the reference returns them (GP_Utils.cpp:159-178)
  void Calc_Out(mat& yPred, const mat& inData) const;
  void Calc_Out(mat& yPred, mat& yVar, const mat& inData) const;
  void Calc_Out(mat& yPred, mat& yVar, mat& probPred, const mat& inData) const;
  void posteriorMeanVar(mat& mu, mat& varSigma, const mat& X) const;
  void posteriorMean(mat& mu, const mat& X) const;

  void updateAlpha() const;                       // brings Alpha and yhat (host copies) up to date with the parameters
  virtual double logLikelihood() const;
  virtual double GradLL(mat& g) const;

  void OptimisePars(unsigned int iters = 1000);
  void ShowKernelPars(std::ostream& os) const;

  virtual unsigned int getNumPars() const;
  virtual void get_GP_Pars(mat& param) const;
  virtual void set_GP_Pars(mat& param) const;
  void FromFile_GP_Params(std::istream& in);
  void ToFile_GP_Params(std::ostream& out) const;

  double getHypermfVal(unsigned int i) const { return hypermf(i); }
  void setHypermfVal(double v, unsigned int i) { hypermf(i) = v; }
  double getHyperlfVal(unsigned int i) const { return hyperlf(i); }
  void setHyperlfVal(double v, unsigned int i) { hyperlf(i) = v; dirty = true; }
  void setHyperlf(const mat& v) { hyperlf = v; dirty = true; }

  int getLikelihoodType() const { return likelihoodType_; }
  std::string getLiklihoodStr() const;
  void setLikelihoodType(const int val);
  int getInferenceType() const { return InferenceType_; }
  std::string getInferenceStr() const;
  void setInferenceType(const int val) { InferenceType_ = val; }
  int getMeanType() const { return MeanType_; }
  std::string getMeanTypeStr() const;
  void setMeanType(const int val) { MeanType_ = val; }
  const Kernels* getKernel() const { return KerenlW; }

  // device selection for this model (default: GPSS_DEVICE environment variable, else 0)
  void setDevice(int dev) { device = dev; }
  // device time of the last objective / prediction call in ms (CUDA events), for the CLI's -v 3 report
  double lastDeviceMs() const;

  // ---- public data, as in the reference (GP_Utils.h:306-358) ----
  mat Xinp;                 // standardised inputs, n x 3
  mat yTarg;                // standardised targets, n x 1
  mutable mat Alpha;        // (K + sn2 I)^-1 y, n x 1
  mutable mat yhat;         // K * Alpha, n x 1
  mutable mat L;            // 1 x 1: the last objective value
  mutable mat hypermf, hyperlf;
  mutable mat g_hyperlf, g_hypermf, g_param;
  Kernels* KerenlW;         // not owned when passed to the constructor (reference behaviour)

 private:
  void _init();
  void check_supported() const;
  void theta_now(double theta[GPSS_NPAR]) const;
  double white_now() const;                        // sum of the White members' Sigma_White (gpss_set_white)
  mutable int kind2_dev;                           // second distance-based member last sent to the device (-1: none)
  mutable double theta2_dev[8];
  mutable int white_cross;                         // Kern_White's cross-covariance condition for the next prediction
  void sync_device() const;                        // create the handle / push data and parameters if stale
  [[noreturn]] void device_failure(const char* what) const;

  int likelihoodType_, InferenceType_, MeanType_;
  int device;
  mutable gpss_handle handle;
  mutable int handle_n;
  mutable int kind_dev;                            // kernel kind last sent to the device (gpss_set_kernel)
  int kernel_kind() const;                         // GPSS_KERNEL_* of the Hyb kernel's first member, -1 if unsupported
  mutable bool data_stale;                         // Xinp / yTarg changed since the last upload
  mutable bool dirty;                              // parameters changed since the last gpss_set_theta
  mutable double theta_dev[GPSS_NPAR];
  mutable bool Chol_fail;
};

void writeGpToStream(const GP_utils& model, std::ostream& out);
void writeGPFile(const GP_utils& model, const std::string modelFileName, const std::string comment = "");
GP_utils* readGpFromStream(std::istream& in);
GP_utils* readGpFromFile(const std::string modelfileName, int verbosity = 2);

#endif
