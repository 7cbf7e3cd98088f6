// Host-side optimiser of the B200 exact-GP path: the reference's callback surface (Opt_pars.h:20-253 in
// /root/reference) with the bound-aware L-BFGS driver and the Potra-Shi line search that decide WHICH theta the
// device path is evaluated at.  Same class / virtual / setter names and argument meaning as the reference so that
// GP_utils (and anything else deriving from Opt_Algs) drops in; the arithmetic follows the reference operation by
// operation (quirks included, each one cited in Opt_pars.cpp) because a 1e-12 difference can flip a line-search
// branch (hard part 4).  All algebra here is 10-dimensional and stays on the host.
#ifndef GPSS_HOST_OPT_PARS_H
#define GPSS_HOST_OPT_PARS_H

#include <armadillo>
#include <cstdlib>
#include <iostream>
#include <string>

class Opt_Algs {
 public:
  enum { SCG, BFGS, LBFGS };

  Opt_Algs();
  virtual ~Opt_Algs() {}

  // ---- the five callbacks the model implements (reference Opt_pars.h:51-55) ----
  virtual unsigned int getNumPars() const = 0;
  virtual void get_GP_Pars(arma::mat& param) const = 0;
  virtual void set_GP_Pars(arma::mat& param) const = 0;
  virtual double Grad_Values(arma::mat& g) const = 0;     // returns the objective, fills the 1 x nPars gradient
  virtual double ObjVal() const = 0;

  virtual void setVerbose(int val) const { verbose = val; }
  virtual int getVerbose() const { return verbose; }

  void setMaxIters(unsigned int val) { maxIters = val; }
  unsigned int getMaxIters() const { return maxIters; }
  void setTolObjVal(double val) { ToObj = val; }
  double getTolObjVal() const { return ToObj; }
  void setTolPars(double val) { TolPars = val; }
  double getTolPars() const { return TolPars; }
  void setOptimiser(int val) const { DefOpt = val; }
  int getOptimiser() const { return DefOpt; }
  void setOptimiserStr(std::string val);
  std::string getDefaultOptimiserStr() const;

  // dispatch on the selected optimiser (reference Opt_pars.h:176-195)
  void Optimise();
  void LBFGSOptimise();
  void BFGSOptimize();
  void scgOptimise();

  // pieces of the L-BFGS iteration, public as in the reference (Opt_pars.h:57-80)
  void cauchy_point(const arma::mat g, const arma::mat X, const arma::mat Wk, const arma::mat Mk, arma::mat& C, arma::mat& xcp,
                    arma::mat& index_r, const double theta, const double mnc);
  void Primal_Conjugate_grad(const arma::mat index_r, const arma::mat xcp, const arma::mat X, const arma::mat Wk, const arma::mat Mk,
                             const arma::mat C, const arma::mat g, const double theta, arma::mat& direction);
  void Efficient_line_search(const double fxk, const arma::mat X, const arma::mat gk, arma::mat& sk, double& steplength);

  // bound helpers (reference Opt_pars.h:92-108).  ChkBnd maps entries ABOVE ub to lb -- as the reference does.
  void ChkBnd(arma::mat& A, const arma::mat lb, const arma::mat ub);
  bool ChkBndStat(arma::mat& A, const arma::mat lb, const arma::mat ub);

 private:
  // X + s*d pulled back into [lb, ub] by dividing s by `div` (Opt_pars.cpp:253-265 and its eight siblings)
  void pull_step_inside(const arma::mat& X, const arma::mat& d, double& s, arma::mat& Xnew, double div);

  double ToObj, TolPars;
  arma::mat lb, ub;
  unsigned int maxIters;
  mutable int verbose;
  mutable int DefOpt;
  // The reference never initialises this flag (Opt_pars.h:218, read at Opt_pars.cpp:577); it starts false here, which
  // is what the zero-filled storage used to record the reference fixtures gives (tests/golden/make_ref_golden.py).
  bool fail_pre_bfgs;
};

// maps the two generic callbacks to the GP's names (reference Opt_pars.h:236-253)
class Main_Opt_Algs : public Opt_Algs {
 public:
  Main_Opt_Algs() : Opt_Algs() {}
  virtual double logLikelihood() const = 0;
  virtual double GradLL(arma::mat& g) const = 0;
  virtual double Grad_Values(arma::mat& g) const { return GradLL(g); }
  virtual double ObjVal() const { return logLikelihood(); }
};

#endif
