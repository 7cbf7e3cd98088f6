// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference classes (GP_utils, HybKerns, Kern_ExpAnisotropic, Kern_Bias, Control, Opt_Algs;
// compiled from /root/reference by oracle/Makefile, never copied) exactly the way gp_ss_ak.cpp:train()/test() do
// (gp_ss_ak.cpp:134-190, 230, 288-306, 384-409) and dumps what the CLI only prints to 6 digits -- the objective,
// the 10-vector gradient, Alpha, samples of K and D2, predictions, and the full LBFGS probe trace -- with 17
// significant digits, so tests/golden/make_ref_golden.py can turn them into fixtures that pin oracle/gpss_oracle.py
// and the CUDA path to the reference's own numbers.
//
//   ref_driver <train.txt> <test.txt> <thetas.txt> <lbfgs_iters> <workdir> > dump.txt
//   ref_driver --time <train.txt> <workdir> <warmup> <steps>     one line per timed set_GP_Pars + Grad_Values (bench.py)
//   ref_driver --trace <SCG|BFGS|LBFGS> <train.txt> <iters> <workdir>      probe trace of one optimiser run
#include "gp_ss_ak.h"
#include <cstdio>
#include <new>

static void dump(const char* key, const mat& m)
{
  printf("%s %llu %llu", key, (unsigned long long)m.n_rows, (unsigned long long)m.n_cols);
  for (uword i = 0; i < m.n_elem; i++) printf(" %.17g", m[i]);
  printf("\n");
}
static void dump(const char* key, double v) { printf("%s 1 1 %.17g\n", key, v); }

// logs every probe the optimiser makes through the reference's own callback surface (Opt_pars.h:51-55, 236-253)
class TraceGP : public GP_utils {
 public:
  TraceGP(Kernels* k, mat X, mat y) : GP_utils(k, X, y, inf_laplace, likeL_Gaussian, mean_zero, 8, 1, 0, 0), on(false), count(0) {}
  virtual double ObjVal() const
  {
    const double f = GP_utils::ObjVal();
    if (on) { mat th(1, getNumPars()); get_GP_Pars(th); printf("probe %d O", count++); for (uword i = 0; i < th.n_elem; i++) printf(" %.17g", th[i]); printf(" f %.17g\n", f); }
    return f;
  }
  virtual double Grad_Values(mat& g) const
  {
    const double f = GP_utils::Grad_Values(g);
    if (on) {
      mat th(1, getNumPars()); get_GP_Pars(th);
      printf("probe %d G", count++); for (uword i = 0; i < th.n_elem; i++) printf(" %.17g", th[i]);
      printf(" f %.17g g", f); for (uword i = 0; i < g.n_elem; i++) printf(" %.17g", g[i]); printf("\n");
    }
    return f;
  }
  mutable bool on;
  mutable int count;
};

// The members of the Hyb kernel, in order: GPSS_REF_KERNEL = a '+'-separated list of ExpAns (default) | Exp | RBF | Bias | White
// (one -k each, gp_ss_ak.cpp:146-170); GPSS_REF_BIAS=0 drops the trailing Bias member (-kn 0, gp_ss_ak.cpp:179-184).
// GPSS_REF_NOGRAD=1 skips Grad_Values in the evaluation loop: with a White member the reference's Kernels::getGradients
// default calls itself (Kernel.h:56-59) and the process dies of stack overflow.
static void add_kernels(HybKerns& K, const mat& X)
{
  const char* e = getenv("GPSS_REF_KERNEL");
  string list = e ? e : "ExpAns";
  while (!list.empty()) {
    const size_t plus = list.find('+');
    const string name = list.substr(0, plus);
    list = (plus == string::npos) ? string() : list.substr(plus + 1);
    if (name == "Exp") K.addNewKernel(new Kern_Exponential(X));
    else if (name == "RBF") K.addNewKernel(new Kern_RBF(X));
    else if (name == "Bias") K.addNewKernel(new Kern_Bias(X));
    else if (name == "White") K.addNewKernel(new Kern_White(X));
    else K.addNewKernel(new Kern_ExpAnisotropic(X));
  }
  const char* b = getenv("GPSS_REF_BIAS");
  if (!b || atoi(b) != 0) K.addNewKernel(new Kern_Bias(X));
}

#include <chrono>
// bench.py --impl reference: W untimed + K timed LML+gradient evaluations (set_GP_Pars(theta_k); Grad_Values(g)) by the
// unmodified reference, a different theta every step so nothing is cached (GP_Utils.cpp:130-133)
static int time_mode(int argc, char** argv)
{
  if (argc < 6) { fprintf(stderr, "usage: ref_driver --time train.txt workdir warmup steps\n"); return 2; }
  const string trainFile = argv[2];
  const string model = string(argv[3]) + "/ref_time_model";
  const int warm = atoi(argv[4]), steps = atoi(argv[5]);
  char* fake[] = {argv[0], 0};
  int Data_mode = 0;
  bool yscale = true;
  Control ctl(1, fake);
  ctl.setMode("train");
  ctl.setprepM(1);
  int* sz = ctl.readDataSize(trainFile);
  mat X(sz[0], sz[1]), y(sz[0], 1);
  ctl.readDataFile(X, y, sz, trainFile);
  ctl.prepareData(X, y, Data_mode, yscale, model);
  HybKerns Kerns(X);
  add_kernels(Kerns, X);
  GP_utils gp(&Kerns, X, y, GP_utils::inf_laplace, GP_utils::likeL_Gaussian, GP_utils::mean_zero, 8, 1, 0, 0);
  const unsigned np = gp.getNumPars();
  mat th0(1, np), g(1, np);
  gp.get_GP_Pars(th0);
  for (int k = 0; k < warm + steps; k++) {
    mat th = th0 * (1.0 + 0.01 * ((k % 7) - 3));
    const auto t0 = std::chrono::steady_clock::now();
    gp.set_GP_Pars(th);
    const double L = gp.Grad_Values(g);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("eval %d %s seconds %.6f nlml %.12g\n", k, k < warm ? "warmup" : "timed", sec, L);
    fflush(stdout);
  }
  return 0;
}

// ref_driver --trace <SCG|BFGS|LBFGS> train.txt iters workdir: the probe trace of one optimiser run of the unmodified
// reference (Opt_pars.cpp:451-538 BFGS, 979-1124 SCG, 179-332 LBFGS), for the host drivers' replay tests.
static int trace_mode(int argc, char** argv)
{
  if (argc < 6) { fprintf(stderr, "usage: ref_driver --trace OPT train.txt iters workdir\n"); return 2; }
  const string opt = argv[2], trainFile = argv[3];
  const int iters = atoi(argv[4]);
  const string model = string(argv[5]) + "/ref_trace_model";
  char* fake[] = {argv[0], 0};
  int Data_mode = 0;
  bool yscale = true;
  Control ctl(1, fake);
  ctl.setMode("train");
  ctl.setprepM(1);
  int* sz = ctl.readDataSize(trainFile);
  mat X(sz[0], sz[1]), y(sz[0], 1);
  ctl.readDataFile(X, y, sz, trainFile);
  ctl.prepareData(X, y, Data_mode, yscale, model);
  HybKerns Kerns(X);
  add_kernels(Kerns, X);
  void* raw = calloc(1, sizeof(TraceGP));                     // zero-filled storage: see the LBFGS block in main()
  TraceGP& gp = *new (raw) TraceGP(&Kerns, X, y);
  gp.setOptimiser(opt == "SCG" ? GP_utils::SCG : opt == "BFGS" ? GP_utils::BFGS : GP_utils::LBFGS);
  gp.setMaxIters(iters);
  gp.setVerbose(1);                                           // the per-iteration lines (SCG prints its scale there)
  const unsigned np = gp.getNumPars();
  mat th0(1, np);
  gp.get_GP_Pars(th0);
  dump("theta_start", th0);
  gp.on = true;
  gp.Optimise();
  gp.on = false;
  mat th(1, np);
  gp.get_GP_Pars(th);
  dump("theta_fit", th);
  dump("nlml_fit", gp.logLikelihood());
  dump("n_probes", (double)gp.count);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc > 1 && string(argv[1]) == "--time") return time_mode(argc, argv);
  if (argc > 1 && string(argv[1]) == "--trace") return trace_mode(argc, argv);
  if (argc < 6) { fprintf(stderr, "usage: ref_driver train.txt test.txt thetas.txt lbfgs_iters workdir\n"); return 2; }
  const string trainFile = argv[1], testFile = argv[2], thetaFile = argv[3];
  const int iters = atoi(argv[4]);
  const string model = string(argv[5]) + "/ref_model";
  char* fake[] = {argv[0], 0};
  int Data_mode = 0;
  bool yscale = true;

  // ---- train-mode data path (gp_ss_ak.cpp:137-142) ----
  Control ctl(1, fake);
  ctl.setMode("train");
  ctl.setprepM(1);
  int* sz = ctl.readDataSize(trainFile);
  mat X(sz[0], sz[1]), y(sz[0], 1);
  ctl.readDataFile(X, y, sz, trainFile);
  dump("X_raw", X);
  dump("y_raw", y);
  ctl.prepareData(X, y, Data_mode, yscale, model);
  dump("Xs", X);
  dump("ys", y);
  dump("params", ctl.params);

  // ---- kernels and model (gp_ss_ak.cpp:146-190, 230) ----
  HybKerns Kerns(X);
  add_kernels(Kerns, X);
  TraceGP gp(&Kerns, X, y);
  const unsigned np = gp.getNumPars();

  // ---- test-mode data path (gp_ss_ak.cpp:379-388) ----
  Control ctl2(1, fake);
  ctl2.setMode("test");
  ctl2.setprepM(1);
  int* szt = ctl2.readDataSize(testFile);
  mat Xt(szt[0], szt[1]), yt(szt[0], 1);
  ctl2.readDataFile(Xt, yt, szt, testFile);
  int Data_mode_t = 1;
  ctl2.prepareData(Xt, yt, Data_mode_t, yscale, model);
  dump("Xt", Xt);

  // ---- evaluations at the given thetas ----
  std::ifstream tf(thetaFile.c_str());
  int nth = 0;
  while (true) {
    mat th(1, np);
    bool ok = true;
    for (unsigned i = 0; i < np; i++) { double v; if (!(tf >> v)) { ok = false; break; } th[i] = v; }
    if (!ok) break;
    char key[64];
    gp.set_GP_Pars(th);
    const double L = gp.logLikelihood();
    mat g(1, np);
    g.zeros();
    gp.set_GP_Pars(th);                    // the optimiser always calls set_GP_Pars before Grad_Values (Opt_pars.cpp:266-267)
    const bool nograd = getenv("GPSS_REF_NOGRAD") != 0 && atoi(getenv("GPSS_REF_NOGRAD")) != 0;
    const double L2 = nograd ? gp.logLikelihood() : gp.Grad_Values(g);
    snprintf(key, sizeof key, "theta_%d", nth); dump(key, th);
    snprintf(key, sizeof key, "nlml_%d", nth); dump(key, L);
    snprintf(key, sizeof key, "nlml_grad_%d", nth); dump(key, L2);
    snprintf(key, sizeof key, "g_%d", nth); dump(key, g);
    snprintf(key, sizeof key, "alpha_%d", nth); dump(key, gp.Alpha);
    snprintf(key, sizeof key, "K_diag_%d", nth); dump(key, mat(gp.K.diag()));
    snprintf(key, sizeof key, "D2_diag_%d", nth); dump(key, mat(gp.D2.diag()));
    snprintf(key, sizeof key, "K_col0_%d", nth); dump(key, mat(gp.K.col(0)));
    snprintf(key, sizeof key, "K_col17_%d", nth); dump(key, mat(gp.K.col(17)));
    mat mu(Xt.n_rows, 1), var(Xt.n_rows, 1);
    gp.Calc_Out(mu, var, Xt);
    snprintf(key, sizeof key, "mu_%d", nth); dump(key, mu);
    snprintf(key, sizeof key, "var_%d", nth); dump(key, var);
    // back-transformed outputs as the CLI reports them (gp_ss_ak.cpp:410-412, Control.cpp:197-255)
    mat Xc = Xt, muc = mu, varc = var;
    ctl2.postData(Xc, muc, yscale, model);
    ctl2.postData_var(varc, yscale, model);
    snprintf(key, sizeof key, "yhat_raw_%d", nth); dump(key, muc);
    snprintf(key, sizeof key, "std_raw_%d", nth); dump(key, varc);        // postData_var already takes the square root (Control.cpp:253-254)
    nth++;
  }
  dump("n_theta", (double)nth);

  // ---- LBFGS fit with the probe trace (gp_ss_ak.cpp:288-296; Opt_pars.cpp:179-332) ----
  if (iters > 0) {
    HybKerns Kerns2(X);
    add_kernels(Kerns2, X);
    // Opt_Algs never initialises fail_pre_bfgs (Opt_pars.h:218, read at Opt_pars.cpp:577): construct the object in
    // zero-filled storage so the flag starts false, which is what a fresh heap page gives the CLI's `new GP_utils`
    void* raw = calloc(1, sizeof(TraceGP));
    TraceGP& gp2 = *new (raw) TraceGP(&Kerns2, X, y);
    gp2.setOptimiser(GP_utils::LBFGS);
    gp2.setMaxIters(iters);
    gp2.on = true;
    gp2.Optimise();
    gp2.on = false;
    mat th(1, np);
    gp2.get_GP_Pars(th);
    dump("theta_fit", th);
    dump("nlml_fit", gp2.logLikelihood());
    dump("n_probes", (double)gp2.count);
    writeGPFile(gp2, model, "# GP_SS_AK Model File ");
  }
  return 0;
}
