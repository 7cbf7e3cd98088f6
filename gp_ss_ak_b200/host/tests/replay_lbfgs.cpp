// CPU test harness: runs the host L-BFGS driver (gp_ss_ak_b200/host/Opt_pars.cpp) against a RECORDED objective -- the
// probe trace of the unmodified reference (tests/golden/ref_n300.npz, written out by tests/test_host_cpu.py).  Every
// ObjVal / Grad_Values call must arrive in the recorded order, of the recorded kind, at the recorded theta; it is then
// answered with the recorded f (and g).  Any deviation of the optimiser's decision logic shows up as a mismatch.
//   replay_lbfgs trace.txt iters tol [LBFGS|BFGS|SCG] [npar]
// trace line:  kind(0=ObjVal,1=Grad_Values) theta[npar] f g[npar]      (npar = 10: Hyb{ExpAns, Bias}; 4 / 5: Hyb{Exp | RBF, Bias})
#include "../Opt_pars.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <vector>

using namespace arma;

constexpr int MAXP = 10;
struct Probe { int kind; double th[MAXP]; double f; double g[MAXP]; };

class ReplayModel : public Opt_Algs {
 public:
  std::vector<Probe> trace;
  mutable size_t next = 0;
  mutable double cur[MAXP];
  mutable double worst = 0.0;
  double tol = 1e-9;
  int npar = MAXP;

  unsigned int getNumPars() const { return (unsigned)npar; }
  void get_GP_Pars(mat& p) const { for (int i = 0; i < npar; i++) p[i] = cur[i]; }
  void set_GP_Pars(mat& p) const { for (int i = 0; i < npar; i++) cur[i] = p[i]; }
  const Probe& take(int kind) const
  {
    if (next >= trace.size()) { printf("MISMATCH: the optimiser asks for probe %zu but the trace has %zu\n", next, trace.size()); exit(3); }
    const Probe& p = trace[next];
    if (p.kind != kind) { printf("MISMATCH at probe %zu: kind %d requested, %d recorded\n", next, kind, p.kind); exit(3); }
    for (int i = 0; i < npar; i++) {
      const double d = std::fabs(p.th[i] - cur[i]);
      if (d > worst) worst = d;
      if (!(d <= tol * std::max(1.0, std::fabs(p.th[i])))) {
        printf("MISMATCH at probe %zu: theta[%d] = %.17g, recorded %.17g\n", next, i, cur[i], p.th[i]);
        exit(3);
      }
    }
    next++;
    return p;
  }
  double ObjVal() const { return take(0).f; }
  double Grad_Values(mat& g) const
  {
    const Probe& p = take(1);
    for (int i = 0; i < npar; i++) g[i] = p.g[i];
    return p.f;
  }
};

int main(int argc, char** argv)
{
  if (argc < 4) { printf("usage: replay_lbfgs trace.txt iters tol [LBFGS|BFGS|SCG] [npar]\n"); return 2; }
  ReplayModel m;
  if (argc > 5) m.npar = atoi(argv[5]);
  if (m.npar < 1 || m.npar > MAXP) { printf("npar must be 1..%d\n", MAXP); return 2; }
  std::ifstream in(argv[1]);
  while (true) {
    Probe p;
    if (!(in >> p.kind)) break;
    for (int i = 0; i < m.npar; i++) in >> p.th[i];
    in >> p.f;
    for (int i = 0; i < m.npar; i++) in >> p.g[i];
    m.trace.push_back(p);
  }
  m.tol = atof(argv[3]);
  for (int i = 0; i < m.npar; i++) m.cur[i] = m.trace[0].th[i];
  m.setOptimiserStr(argc > 4 ? argv[4] : "LBFGS");      // LBFGS | BFGS | SCG
  m.setMaxIters(atoi(argv[2]));
  m.Optimise();
  printf("REPLAY OK probes %zu of %zu worst_theta_diff %.3e final", m.next, m.trace.size(), m.worst);
  for (int i = 0; i < m.npar; i++) printf(" %.17g", m.cur[i]);
  printf("\n");
  return m.next == m.trace.size() ? 0 : 4;
}
