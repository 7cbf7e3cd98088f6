// Bound-aware L-BFGS (Byrd-Lu-Nocedal compact form) and the Potra-Shi line search, behaviour-identical to the
// reference's Opt_Algs (file:line citations are into /root/reference/Opt_pars.cpp unless another file is named).
// The reference's deviations from the textbook algorithms are kept on purpose and marked [quirk].
#include "Opt_pars.h"

#include <cmath>
#include <limits>

using namespace arma;
using std::cout;
using std::endl;

namespace {
const double kEps = std::numeric_limits<double>::epsilon();

inline double dotp(const mat& a, const mat& b) { return accu(a % b); }

// the compact-representation middle matrix  Mk = inv([[-D, L'], [L, theta S'S]])  with L = trimatl(S'Y)
// [quirk] trimatl keeps the diagonal (the published algorithm uses the strictly lower part), :222-228, 301-308, 316-321
mat middle_matrix(const mat& Dk, const mat& Sk, const mat& Yk, double theta)
{
  const uword nc = Dk.n_rows;
  const mat Lk = trimatl(Sk.t() * Yk);
  mat Mk(2 * nc, 2 * nc);
  Mk.zeros();
  Mk.submat(0, 0, nc - 1, nc - 1) = -Dk;
  Mk.submat(0, nc, nc - 1, 2 * nc - 1) = Lk.t();
  Mk.submat(nc, 0, 2 * nc - 1, nc - 1) = Lk;
  Mk.submat(nc, nc, 2 * nc - 1, 2 * nc - 1) = theta * Sk.t() * Sk;
  return inv(Mk);
}
}  // namespace

Opt_Algs::Opt_Algs()
{
  setVerbose(0);
  setOptimiser(LBFGS);
  setMaxIters(100);
  setTolObjVal(1e-0);
  setTolPars(1e0);
  fail_pre_bfgs = false;
}

void Opt_Algs::setOptimiserStr(std::string val)
{
  if (val == "SCG") DefOpt = SCG;
  else if (val == "BFGS") DefOpt = BFGS;
  else if (val == "LBFGS") DefOpt = LBFGS;
  else { cout << "Unknown optimisation. \n"; exit(1); }
}

std::string Opt_Algs::getDefaultOptimiserStr() const
{
  switch (DefOpt) {
    case SCG: return "SCG";
    case BFGS: return "quasinew";
    case LBFGS: return "LBFGS";
    default: cout << "Unknown optimisation. \n"; exit(1);
  }
}

void Opt_Algs::Optimise()
{
  switch (DefOpt) {
    case SCG: scgOptimise(); break;
    case BFGS: BFGSOptimize(); break;
    case LBFGS: LBFGSOptimise(); break;
    default: cout << "Unknown optimisation.\n"; exit(1);
  }
}

// ---------------------------------------------------------------------------------------------------
// BFGS (:451-538) -- the reference's DEFAULT optimiser (gp_ss_ak.cpp:91).  Restated operation by operation; what differs
// from the textbook method is kept and marked [quirk].
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::BFGSOptimize()
{
  const int D = getNumPars();
  lb.ones(1, D);
  lb = lb * 1e-4;
  ub.ones(1, D);
  ub = 6 * ub;
  mat Hes(D, D);
  Hes.eye(D, D);

  mat X0(1, D), g(1, D);
  get_GP_Pars(X0);
  set_GP_Pars(X0);
  mat X = X0;
  double fx = Grad_Values(g);
  const int Maxit = getMaxIters();
  int iter = 0;
  mat gnew = g, Xnew = X;
  Hes = eye<mat>(D, D) / accu(g * X.t());               // [quirk] initial inverse Hessian = I / (g0 . x0) (:481-482)
  double final_steplength = 1;
  while (true) {
    iter++;
    const mat gold = gnew;
    const mat Xold = Xnew;
    // [quirk] the direction is always built from g, the INITIAL gradient: g is never refreshed inside the loop (:491)
    mat search_direction = -g * Hes;
    Efficient_line_search(fx, X0, gold, search_direction, final_steplength);
    pull_step_inside(X0, search_direction, final_steplength, Xnew, 1.2);
    set_GP_Pars(Xnew);
    const double fnew = Grad_Values(gnew);
    if (fnew < fx) {
      X0 = Xnew;
      fx = fnew;
    }
    const mat yk = gnew - gold;                          // gradient of the last TRIED point against the previous tried point
    const mat sk = Xnew - Xold;
    if (iter == 1) {
      Hes = eye<mat>(D, D) * accu(sk * yk.t()) / accu(yk * yk.t());
    } else {
      const double rho = 1.0 / accu(yk * sk.t());
      Hes = (eye<mat>(D, D) - rho * sk.t() * yk) * Hes * (eye<mat>(D, D) - rho * yk.t() * sk) + rho * sk.t() * sk;
    }
    if (iter >= Maxit) break;
    if (getVerbose() > 0) cout << "Iteration: " << iter << " -logL: " << fx << endl;
  }
  set_GP_Pars(X0);
}

// ---------------------------------------------------------------------------------------------------
// Moller's scaled conjugate gradient as the reference writes it (:979-1124).
// [deviation] the reference never initialises its scale `lambda` (:1004) and reads it at :1036: undefined behaviour.  The
// compiled reference prints "Scale: 6.9e-310" (a stale pointer seen as a denormal), which is 0 in every operation it
// enters; lambda starts at 0.0 here, which reproduces the compiled reference's probe trace (tests/test_host_cpu.py).
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::scgOptimise()
{
  if (getVerbose() > 2) cout << "Scaled Conjugate Gradient Optimisation." << endl;
  const int D = getNumPars();
  lb.ones(1, D);
  lb = lb * 1e-4;
  ub.ones(1, D);
  ub = 6 * ub;

  mat w(1, D), rk(1, D), pk(1, D);
  double sigmak;
  const double sigma = 1e-4;
  double lambdaBar = 0.0, fw = 0.0, muk = 0.0, alphak = 0.0, betak, Deltak;
  double lambda = 0.0;
  bool success = true;

  get_GP_Pars(w);
  mat Xnew = w;
  fw = Grad_Values(rk);
  double fnew = fw;
  mat gnew = rk, rk_old = rk, g = gnew, gold = gnew, sk = gnew;
  const double fw_old = fw;                              // [quirk] never updated: the stop test compares with the START (:1105)
  rk = -1 * rk;
  pk = rk;
  double deltak = 0.0;
  int iter = 0;
  if (getVerbose() > 0) cout << "Iteration: " << iter << " -logL: " << fw << " Scale: " << lambda << endl;
  while (true) {
    iter++;
    if (success) {                                       // step 2: second-order information by a finite difference
      const double div = std::sqrt(accu(pk * pk.t()));
      sigmak = sigma / div;
      Xnew = w + pk * sigmak;
      set_GP_Pars(Xnew);
      fnew = Grad_Values(gnew);
      (void)fnew;
      sk = (gnew - gold) / sigma + lambda * pk;          // [quirk] divides by sigma, not sigmak; gold is the last PROBE's gradient
      deltak = accu(pk * sk.t());
      gold = gnew;
    }
    const double norm2p = accu(pk * pk.t());             // step 3
    deltak += (lambda - lambdaBar) * norm2p;
    sk += (lambda - lambdaBar) * pk;
    if (deltak <= 0.0) {                                 // step 4: make the Hessian estimate positive definite
      lambdaBar = 2.0 * (lambda - deltak / norm2p);
      deltak -= lambda * norm2p;
      lambda = lambdaBar;
    }
    muk = accu(pk * rk.t());                             // step 5
    alphak = muk / deltak;
    pull_step_inside(w, pk, alphak, Xnew, 1.2);          // step 6 (:1062-1075)
    set_GP_Pars(Xnew);
    const double falpha = ObjVal();
    Deltak = 2.0 * deltak * (fw - falpha) / std::pow(muk, 2.0);
    if (Deltak >= 0.0) {                                 // step 7
      w = Xnew;
      fw = falpha;
      Grad_Values(g);
      rk = -g;
      lambdaBar = 0;
      success = true;
      if (iter % D == 0) {
        pk = rk;
      } else {
        betak = (accu(rk * rk.t()) - accu(rk * rk_old.t())) / muk;
        pk = rk + betak * pk;
      }
      if (Deltak >= 0.75) lambda *= 0.25;
    } else {
      set_GP_Pars(w);
      lambdaBar = lambda;
      success = false;
    }
    rk_old = rk;
    if (std::fabs(fw - fw_old) < getTolObjVal() && (((iter % 3) == 0) & (iter > 10))) return;
    if (Deltak < 0.25) lambda += (deltak * (1 - Deltak) / accu(pk * pk.t()));       // step 8
    bool all_zero = true;                                // step 9
    for (uword i = 0; i < rk.n_elem; i++) if (rk[i] != 0) all_zero = false;
    if (all_zero) break;
    if (iter >= (int)getMaxIters()) break;
    if (getVerbose() > 0) cout << "Iteration: " << iter << " -logL: " << fw << " Scale: " << lambda << endl;
  }
  set_GP_Pars(w);
}

void Opt_Algs::ChkBnd(mat& A, const mat lo, const mat hi)
{
  for (uword i = 0; i < A.n_elem; i++) {
    if (A[i] < lo[i]) A[i] = lo[i];
    if (A[i] > hi[i]) A[i] = lo[i];          // [quirk] upper violations are mapped to the LOWER bound (Opt_pars.h:96-97)
  }
}

bool Opt_Algs::ChkBndStat(mat& A, const mat lo, const mat hi)
{
  for (uword i = 0; i < A.n_elem; i++)
    if (A[i] < lo[i] || A[i] > hi[i]) return true;
  return false;
}

void Opt_Algs::pull_step_inside(const mat& X, const mat& d, double& s, mat& Xnew, double div)
{
  Xnew = X + s * d;
  bool violate = ChkBndStat(Xnew, lb, ub);
  while (violate) {
    s /= div;
    Xnew = X + s * d;
    violate = ChkBndStat(Xnew, lb, ub);
    if (s < kEps) {      // checked after the re-test, and it ends the loop either way (:260-264)
      Xnew = X;
      s = 0.0;
      break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// generalised Cauchy point (:11-105)
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::cauchy_point(const mat g, const mat X, const mat Wk, const mat Mk, mat& C, mat& xcp, mat& index_r, const double theta,
                            const double mnc)
{
  const double tiny = 1e-100;
  const int D = getNumPars();
  mat c((uword)(2 * mnc), 1);
  c.zeros();
  index_r.resize(1, 1);
  xcp = X;
  mat t(D, 1);
  mat d = -g.t();
  for (int j = 0; j < D; j++) {
    if (g[j] < 0) t[j] = (X[j] - ub[j]) / g[j];
    else if (g[j] > 0) t[j] = (X[j] - lb[j]) / g[j];
    else t[j] = std::numeric_limits<double>::max();
    if (t[j] > -tiny && t[j] < tiny) d[j] = 0.0;
  }
  // break points still ahead of us; [quirk] from here on `b` is a position in this SHRINKING list, yet it is used to
  // index the variables (:45-51, 56-80)
  mat F = t.elem(find(t > 0.0));
  mat p = Wk.t() * d;
  double fprime = accu(-d.t() * d);
  double fsec = accu(-theta * fprime - p.t() * Mk * p);
  double dt_min = -fprime / fsec;
  double t_old = 0.0;
  double mt = F.min();
  uword b = F.index_min();
  F.shed_row(b);
  index_r[0] = (double)b;
  double dt = mt - t_old;
  int rowN = 1;
  while (dt_min >= dt && F.n_elem > 0) {
    if (d[b] > 0) xcp[b] = ub[b];
    else if (d[b] < 0) xcp[b] = lb[b];
    const double zb = xcp[b] - X[b];
    c += dt * p;
    const mat wb = Wk.row(b);
    fprime += dt * fsec + std::pow(g[b], 2) + theta * g[b] * zb - g[b] * accu(wb * (Mk * c));
    fsec += -theta * std::pow(g[b], 2) - 2.0 * g[b] * accu(wb * Mk * p) - std::pow(g[b], 2) * accu(wb * Mk * wb.t());
    p += g[b] * wb.t();
    d[b] = 0.0;
    dt_min = -fprime / fsec;
    t_old = mt;
    mt = F.min();                    // aborts like Armadillo when the list has just been emptied
    b = F.index_min();
    F.shed_row(b);
    rowN++;
    mat bb(1, 1);
    bb[0] = (double)b;
    index_r.insert_rows(rowN - 1, bb);
    dt = mt - t_old;
  }
  dt_min = std::max(dt_min, 0.0);
  t_old += dt_min;
  for (int i = 0; i < D; i++)
    if (t[i] >= mt) xcp[i] = X[i] + t_old * d[i];
  for (uword i = 0; i < F.n_rows; i++) {     // F.n_rows is re-read every pass, as in the reference (:93-103)
    if (t[i] == mt) {
      F.shed_row(i);
      rowN++;
      mat bb(1, 1);
      bb[0] = (double)i;
      index_r.insert_rows(rowN - 1, bb);
    }
  }
  C = c + dt_min * p;
}

// ---------------------------------------------------------------------------------------------------
// subspace minimisation by conjugate gradients over the free variables (:108-174)
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::Primal_Conjugate_grad(const mat index_r, const mat xcp, const mat X, const mat Wk, const mat Mk, const mat C, const mat g,
                                     const double theta, mat& direction)
{
  const int maxit = 50;
  direction.zeros();
  const int D = getNumPars();
  if (D - (int)index_r.n_elem == 0) {
    direction = xcp - X;
    return;
  }
  mat Zk(D, D);
  Zk.eye();
  for (uword j = 0; j < index_r.n_elem; j++) {
    const double v = index_r[j];
    for (int i = 0; i < D; i++)
      if (v == i) Zk(i, i) = 0.0;
  }
  const mat rc = Zk.t() * ((g + theta * (xcp - X)).t() - Wk * Mk * C);
  mat r = rc;
  mat p = -r;
  double rho2 = accu(r.t() * r), rho1 = 0.0;
  int it = 0;
  while (norm(r) >= std::min(0.1, std::sqrt(norm(rc))) * norm(rc)) {
    if (it > maxit) break;
    it++;
    double alpha1 = -std::numeric_limits<double>::infinity();
    for (int i = 0; i < D; i++) {            // [quirk] the LARGEST bound ratio is kept (:150-156)
      if (p[i] < 0) alpha1 = std::max(alpha1, (lb[i] - xcp[i] - direction[i]) / p[i]);
      else if (p[i] > 0) alpha1 = std::max(alpha1, (ub[i] - xcp[i] - direction[i]) / p[i]);
    }
    const mat Bk = theta * eye<mat>(D, D) - Wk * Mk * Wk.t();
    const mat q = Bk * p;
    const double alpha2 = rho2 / accu(p.t() * q);
    if (alpha2 > alpha1) {
      direction += alpha1 * p.t();
      break;
    }
    direction += alpha2 * p.t();
    r += alpha2 * q;
    rho1 = rho2;
    rho2 = accu(r.t() * r);
    const double beta = rho2 / rho1;
    p = -r + beta * p;
  }
}

// ---------------------------------------------------------------------------------------------------
// the driver (:179-332).  No convergence test: it always runs getMaxIters() iterations.
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::LBFGSOptimise()
{
  const int D = getNumPars();
  lb.ones(1, D);
  lb = lb * 1e-4;
  ub.ones(1, D);
  ub = 6 * ub;
  int nc = 1;
  const int mnc = 6;
  double theta = 0.9;

  mat X0(1, D), g(1, D);
  get_GP_Pars(X0);
  set_GP_Pars(X0);
  double fx = Grad_Values(g);          // g keeps the INITIAL gradient for the whole run (used again at :313)
  mat xcp = X0;
  mat index_r(1, 1);
  mat C(2 * mnc, 1);
  C.zeros();
  mat search_direction(1, D);

  // the first "correction pair" is (S, Y) = (x0, g0) [quirk] (:216-221)
  mat Dk(1, 1), Yk(D, 1), Sk(D, 1), Wk(D, 2);
  Dk(0, 0) = accu(X0 * g.t());
  Yk.col(0) = g.t();
  Sk.col(0) = X0.t();
  Wk.col(0) = g.t();
  Wk.col(1) = theta * X0.t();
  mat Mk = middle_matrix(Dk, Sk, Yk, theta);

  const int Maxit = getMaxIters();
  int iter = 0;
  mat gnew = g, Xnew = X0;
  double final_steplength = 1;          // in/out of the line search: persists across iterations (:237, 250)
  while (true) {
    iter++;
    const mat gold = gnew;              // [quirk] gradient of the last TRIED point, paired below with the best point X0
    const mat Xold = Xnew;
    cauchy_point(gold, X0, Wk, Mk, C, xcp, index_r, theta, nc);
    Primal_Conjugate_grad(index_r, xcp, X0, Wk, Mk, C, gold, theta, search_direction);
    Efficient_line_search(fx, X0, gold, search_direction, final_steplength);
    pull_step_inside(X0, search_direction, final_steplength, Xnew, 1.2);
    set_GP_Pars(Xnew);
    const double fnew = Grad_Values(gnew);
    if (fnew < fx) {                    // NaN (Cholesky failure) compares false: the step is rejected
      X0 = Xnew;
      fx = fnew;
    }
    const mat yk = gnew - gold;
    const mat sk = Xnew - Xold;
    if (accu(sk.t() * yk) <= kEps * accu(yk.t() * yk)) {      // n x n outer products summed, as written (:281)
      if (getVerbose() > 0) cout << "Iteration: " << iter << " -logL: " << fx << endl;
      if (iter >= Maxit) break;
      continue;
    }
    if (nc < mnc) {
      nc++;
      mat diagonal = Dk.diag();
      diagonal.insert_rows(diagonal.n_rows, sk * yk.t());
      Dk.zeros(nc, nc);
      for (int i = 0; i < nc; i++) Dk(i, i) = diagonal[i];
      Yk.insert_cols(nc - 1, yk.t());
      Sk.insert_cols(nc - 1, sk.t());
      Wk.resize(D, 2 * nc);
      Wk.submat(0, 0, D - 1, nc - 1) = Yk;
      Wk.submat(0, nc, D - 1, 2 * nc - 1) = theta * Sk;
      Mk = middle_matrix(Dk, Sk, Yk, theta);
    } else {
      // [quirk] a full memory is not rotated: slot 0 is overwritten, and Wk's columns 0 and mnc are refreshed with
      // the INITIAL gradient and the current best point (:310-321)
      Dk(0, 0) = accu(sk * yk.t());
      Yk.col(0) = yk.t();
      Sk.col(0) = sk.t();
      Wk.col(0) = g.t();
      Wk.col(mnc) = theta * X0.t();
      Mk = middle_matrix(Dk, Sk, Yk, theta);
    }
    theta = accu(yk * yk.t()) / accu(yk * sk.t());
    if (iter >= Maxit) break;
    if (getVerbose() > 0) cout << "Iteration: " << iter << " -logL: " << fx << endl;
  }
  set_GP_Pars(X0);
}

// ---------------------------------------------------------------------------------------------------
// Potra-Shi "efficient line search" (:543-974)
// ---------------------------------------------------------------------------------------------------
void Opt_Algs::Efficient_line_search(const double fxk, const mat X, const mat gk, mat& sk, double& final_steplength)
{
  const double rho = 1e-14, sig = 0.99, J = 2.0, tau3 = 2.1;
  double tau1 = 1e-14, tau2 = 0.49;          // shrunk while probing near the bounds; locals, so reset per call
  const int maxls = 4;

  double steplength = 1.0;
  double a = 0.0, b = steplength;
  bool done = false;
  const double f0 = fxk;
  double best = f0;                          // lowest objective seen; final_steplength follows it
  const double fprim0 = accu(gk.t() * sk);
  mat gnew = gk;
  mat Xnew;
  auto track = [&](double f, double s) { if (f < best) { final_steplength = s; best = f; } };
  auto finish = [&](bool strict) { fail_pre_bfgs = strict ? !(best < f0) : !(best <= f0); };

  if (fail_pre_bfgs) steplength = -1.0;      // the previous search failed: first trial goes backwards (:577-578)
  pull_step_inside(X, sk, steplength, Xnew, 1.2);
  set_GP_Pars(Xnew);
  const double f1 = Grad_Values(gnew);
  track(f1, 1.0);                            // [quirk] records 1.0, not the trial step (:597-601)

  double fa, fb;
  if (f1 > f0 + rho * fprim0) {
    // step 1: the unit step is too long -> bracket [0, steplength]
    a = 0.0;
    b = steplength;
    Xnew = X + a * sk;
    set_GP_Pars(Xnew);
    fa = ObjVal();
    track(fa, a);
    pull_step_inside(X, sk, b, Xnew, 1.2);
    set_GP_Pars(Xnew);
    fb = ObjVal();
    track(fb, b);
  } else {
    if (sig > 0.5) {
      if (f1 >= f0 + sig * fprim0) { final_steplength = 1.0; done = true; }
    } else {
      const double fprim1 = accu(gnew.t() * sk);
      if (fprim1 >= sig * fprim0) { final_steplength = 1.0; done = true; }
    }
    if (done) { finish(false); return; }

    // step 2: extrapolate [an, bn] = [1, J], [J, J^2], ...
    double an = 1.0, bn = J;
    pull_step_inside(X, sk, an, Xnew, 1.2);
    set_GP_Pars(Xnew);
    fa = ObjVal();
    track(fa, an);
    pull_step_inside(X, sk, bn, Xnew, 1.2);
    set_GP_Pars(Xnew);
    fb = ObjVal();
    track(fb, bn);
    while (true) {
      if (fb > fa + (bn - an) * rho * fprim0) { a = an; b = bn; break; }                            // 2a
      if (fb >= fa + (bn - an) * sig * fprim0) { final_steplength = bn; done = true; break; }       // 2b
      an = bn;                                                                                      // 2c
      bn = J * bn;
      {
        Xnew = X + an * sk;
        bool violate = ChkBndStat(Xnew, lb, ub);
        while (violate) {                 // [quirk] this copy of the pull-back loop divides twice per pass (:760-775)
          an /= 1.2;
          Xnew = X + an * sk;
          violate = ChkBndStat(Xnew, lb, ub);
          an /= 2.0;
          if (an < kEps) { Xnew = X; an = 0.0; break; }
        }
      }
      if (fa != fa || fb != fb) { done = true; break; }
      set_GP_Pars(Xnew);
      fa = ObjVal();
      track(fa, an);
      pull_step_inside(X, sk, bn, Xnew, 1.2);
      set_GP_Pars(Xnew);
      fb = ObjVal();
      track(fb, bn);
    }
  }
  if (done) { finish(false); return; }

  // step 3: interpolation inside [an, bn]
  double an = a, bn = b, cn = an, deltan = 0.0;
  double it = 0;
  while (it < maxls) {
    it++;
    double lowv = an + tau1 * (bn - an);
    double highv = an + tau2 * (bn - an);
    {
      Xnew = X + lowv * sk;
      bool violate = ChkBndStat(Xnew, lb, ub);
      while (violate) {
        tau1 /= 1.2;
        lowv = an + tau1 * (bn - an);
        Xnew = X + lowv * sk;
        violate = ChkBndStat(Xnew, lb, ub);
        if (lowv < kEps) { Xnew = X; lowv = 0.0; break; }
      }
    }
    set_GP_Pars(Xnew);
    mat glow = gk, ghigh = gk;
    const double flow = Grad_Values(glow);
    track(flow, lowv);
    {
      Xnew = X + highv * sk;
      bool violate = ChkBndStat(Xnew, lb, ub);
      while (violate) {
        tau2 /= 1.1;
        highv = an + tau2 * (bn - an);
        Xnew = X + highv * sk;
        violate = ChkBndStat(Xnew, lb, ub);
        if (tau2 >= tau1) break;           // [quirk] true on the first pass, so at most one shrink happens (:866-867)
        if (highv < kEps) { Xnew = X; highv = 0.0; break; }
      }
    }
    set_GP_Pars(Xnew);
    const double fhigh = Grad_Values(ghigh);
    track(fhigh, highv);
    const double fprimlow = accu(glow.t() * sk);
    const double fprimhigh = accu(ghigh.t() * sk);
    // two-point Hermite-like interpolant sampled at 1/4, 1/2, 3/4 of (lowv + highv) [quirk: of the SUM] (:888-905)
    auto model = [&](double x) {
      return (flow + (x - lowv) * fprimlow) * (highv - x) / (highv - lowv) + (fhigh + (x - highv) * fprimhigh) * (x - lowv) / (highv - lowv);
    };
    const double x0 = 0.25 * (lowv + highv), x1 = 0.5 * (lowv + highv), x2 = 0.75 * (lowv + highv);
    const double y0 = model(x0), y1 = model(x1), y2 = model(x2);
    double minf = std::min(y0, y1);
    minf = std::min(minf, y2);
    if (minf == y0) cn = x0;
    else if (minf == y1) cn = x1;
    else if (minf == y2) cn = x2;
    pull_step_inside(X, sk, cn, Xnew, 1.1);
    set_GP_Pars(Xnew);
    const double fcn = ObjVal();
    track(fcn, cn);
    if (it == 1) deltan = std::fabs(((fb - fcn) / (bn - cn) - (fcn - fa) / (cn - an)) / (bn - an));
    // 3b
    if (fcn <= fa + (cn - an) * rho * fprim0 && fcn >= fa + (cn - an) * sig * fprim0) {
      final_steplength = cn;
      done = true;
      break;
    }
    deltan = std::fabs(((fb - fcn) / (bn - cn) - (fcn - fa) / (cn - an)) / (bn - an));      // 3c
    if (fcn <= fa + (cn - an) * rho * fprim0) {                                             // 3d
      if ((rho - sig) * fprim0 >= tau3 * (bn - an) * deltan) {
        steplength = cn;
      } else {
        an = cn;
        Xnew = X + an * sk;
        ChkBnd(Xnew, lb, ub);
        set_GP_Pars(Xnew);
        fa = ObjVal();
        track(fa, an);
      }
    } else {                                                                                // 3e
      if ((rho - sig) * fprim0 >= tau3 * (bn - an) * deltan && an > 0) {
        final_steplength = an;
        done = true;
        break;
      }
      bn = cn;
      Xnew = X + bn * sk;
      ChkBnd(Xnew, lb, ub);
      set_GP_Pars(Xnew);
      fb = ObjVal();
      if (fcn < best) {                    // [quirk] tests fcn but stores fb (:957-961)
        final_steplength = bn;
        best = fb;
      }
    }
  }
  finish(true);
  if (!done) final_steplength = steplength;
}
