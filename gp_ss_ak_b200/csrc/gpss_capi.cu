// C ABI (include/gpss.h) and host-side drivers of the sm_100a exact-GP path: blocked potrf, trsv, trtri,
// lauum, gradient reductions and batched prediction, each a fixed sequence of launches of the kernels in
// gpss_kernels.cuh / gpss_gemm.cuh on one stream.  No cuBLAS / cuSOLVER, no CPU fallback.
// File:line citations are into /root/reference.
#include "../../include/gpss.h"
#include "gpss_gemm.cuh"
#include "gpss_kernels.cuh"
#include "gpss_params.h"

#include <dlfcn.h>
#include <nccl.h>       // types and prototypes only: the library is dlopen'ed when a communicator is first needed

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <limits>

using namespace gpss;

static thread_local std::string g_last_error;

static int fail_cuda(cudaError_t e, const char* what, int line)
{
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error '%s' in %s (gpss_capi.cu:%d)", cudaGetErrorString(e), what, line);
  g_last_error = buf;
  return GPSS_ERR_CUDA;
}
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail_cuda(e__, #x, __LINE__); } while (0)
#define RET(x) do { int r__ = (x); if (r__ < 0) return r__; } while (0)

static int fail_arg(const char* msg) { g_last_error = msg; return GPSS_ERR_ARG; }

// ---------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (libnccl.so.2: the copy torch has already loaded, else the system one), so that the
// single-GPU library has no hard dependency on it.
// ---------------------------------------------------------------------------------------------------
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static int nccl_load()
{
  if (g_nccl.ok) return GPSS_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_last_error = std::string("cannot load libnccl.so.2: ") + dlerror(); return GPSS_ERR_NCCL; }
#define NCCL_SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(h, name); if (!g_nccl.field) { g_last_error = "libnccl.so.2 lacks " name; return GPSS_ERR_NCCL; }
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(Broadcast, "ncclBroadcast")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(AllGather, "ncclAllGather")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
  g_nccl.ok = true;
  return GPSS_OK;
}
static int fail_nccl(ncclResult_t r, const char* what, int line)
{
  char buf[512];
  snprintf(buf, sizeof buf, "NCCL error '%s' in %s (gpss_capi.cu:%d)", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what, line);
  g_last_error = buf;
  return GPSS_ERR_NCCL;
}
#define NC(x) do { ncclResult_t r__ = (x); if (r__ != ncclSuccess) return fail_nccl(r__, #x, __LINE__); } while (0)

constexpr int NBO = 512;      // outer block (k-depth of the big trailing updates)
constexpr int PRED_BATCH = 8192;

enum QState { Q_NONE = 0, Q_IS_BINV = 1, Q_IS_W = 2 };

struct gpss_ctx {
  int device = 0;
  int n = 0, n_pad = 0, nblk = 0;
  int d = 3;                                                   // input columns: 3, or 4 with the rock-type column
  int kind = 0;                                                // main kernel: GPSS_KERNEL_EXPANS | _EXP | _RBF
  cudaStream_t st = nullptr;                                  // main stream (highest priority): critical-path kernels
  cudaStream_t st2 = nullptr;                                 // look-ahead stream (lowest priority): bulk trailing updates
  cudaStream_t st3 = nullptr;                                 // second look-ahead stream: consecutive bulk updates alternate so
                                                              // the tail wave of one is filled by the head of the next
  cudaStream_t st4 = nullptr;                                 // communication stream of the pipelined panel broadcast (highest priority)
  std::vector<cudaEvent_t> ev_pipe;                           // per 128-column sub-panel: [2 i] factored on the owner, [2 i + 1] received
  cudaEvent_t ev_main = nullptr, ev_side = nullptr;           // cross-stream dependencies of the look-ahead
  std::vector<cudaEvent_t> ev_pool;                           // per-panel events of the look-ahead Cholesky / inverse
  // data
  double *xs = nullptr, *y = nullptr, *zs = nullptr;           // NX x n_pad, n_pad, NZ x n_pad
  double *Lm = nullptr, *Um = nullptr, *Qm = nullptr;          // n_pad^2 each (Um, Qm lazily)
  double *Winv = nullptr;                                      // nblk x 128 x 128
  double *logdet_parts = nullptr;                              // nblk
  double *rvec = nullptr, *zvec = nullptr, *alpha = nullptr, *fvec = nullptr;   // n_pad each
  double *Tpanel = nullptr, *Wjj = nullptr;                    // n_pad x NBO, NBO x NBO (lazily)
  double *partial = nullptr; long partial_blocks = 0;          // gradient partial sums
  double *red = nullptr;                                       // 32 doubles of reduced scalars
  DevParams* dP = nullptr;                                     // [0] training, [1] prediction
  int* dflag = nullptr;
  // prediction scratch (lazily)
  double *xt = nullptr, *zt = nullptr, *zsp = nullptr, *Bm = nullptr, *Vm = nullptr, *mu_part = nullptr, *dmu = nullptr, *dvar = nullptr;
  int pred_cap = 0;
  // distributed evaluation (one process per GPU; rank/world = 0/1 when not initialised)
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  double* stage = nullptr; size_t stage_count = 0;            // contiguous staging for strided sub-matrices
  double* Tsplit = nullptr; size_t Tsplit_cap = 0;             // split-k partial products of the row-sliced inverse
  double* pgather = nullptr;                                   // partitioned inverse: my piece of an L row strip + the all-gathered pieces
  bool partitioned = false;                                    // Lm holds only my block columns, packed (n too large to replicate)
  int nq = 0; long lcols = 0;                                  //   number of own block columns / local column count
  int urow0 = 0, urow1 = 0;                                    // my rows of U = L^-T
  int qrow0 = 0, qrow1 = 0;                                    // my rows of B^-1
  // host state
  double theta[GPSS_NPAR];
  double sums_train[4];
  bool have_factor = false, have_alpha = false, have_U = false;
  int qstate = Q_NONE;
  int chol_fail = 0;
  double nlml = std::numeric_limits<double>::quiet_NaN();
  double s3 = 0.0;
  // CUDA graphs of the launch-bound small-n evaluation: [0] K build + Cholesky + solves + objective terms, [1] inverse + gradient pass
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  long graph_launches[2] = {0, 0};
  bool graph_failed = false;
  // instrumentation
  bool profiling = false;
  double phase_ms[16];
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t ev_call[2] = {nullptr, nullptr};   // bracket the device work of the last objective / predict call
  double last_call_ms = 0.0;
  long launches = 0;
};

// ---------------------------------------------------------------------------------------------------
static int configure_kernels()
{
  CU(cudaFuncSetAttribute(gemm_nt_ws_kernel<GemmTileWideWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTileWideWS::SMEM_BYTES));
  CU(cudaFuncSetAttribute(gemm_nt_kernel<GemmTileWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTileWide::SMEM_BYTES));
  CU(cudaFuncSetAttribute(potrf_diag_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  return GPSS_OK;
}

// Every O(n^3) product of the path goes through the warp-specialised 128x64 DMMA kernel (gemm_nt_ws_kernel).
static int gemm_ws_on(gpss_ctx* c, const GemmArgs& g, cudaStream_t stream)
{
  using T = GemmTileWideWS;
  if (g.M <= 0 || g.N <= 0) return GPSS_OK;
  if (g.M % T::BM || g.N % T::BN || g.K % T::BK) return fail_arg("gemm: dimensions not tile multiples");
  GemmArgs ga = g;
  ga.mt = g.M / T::BM;
  ga.nt = g.N / T::BN;
  const int parts = ga.ksplit > 1 ? ga.ksplit : 1;
  gemm_nt_ws_kernel<T><<<(unsigned)(ga.mt * ga.nt * parts), T::THREADS, T::SMEM_BYTES, stream>>>(ga);
  c->launches++;
  CU(cudaGetLastError());
  return GPSS_OK;
}

// legacy cp.async kernel (kept as the A/B baseline of bench_micro/gemm_bench.cu and for the tile=1 test hook)
static int gemm_legacy_on(gpss_ctx* c, const GemmArgs& g, cudaStream_t stream)
{
  using T = GemmTileWide;
  if (g.M <= 0 || g.N <= 0) return GPSS_OK;
  if (g.M % T::BM || g.N % T::BN || g.K % T::BK) return fail_arg("gemm: dimensions not tile multiples");
  GemmArgs ga = g;
  ga.mt = g.M / T::BM;
  ga.nt = g.N / T::BN;
  gemm_nt_kernel<T><<<(unsigned)(ga.mt * ga.nt), T::THREADS, T::SMEM_BYTES, stream>>>(ga);
  c->launches++;
  CU(cudaGetLastError());
  return GPSS_OK;
}

static int gemm(gpss_ctx* c, const GemmArgs& g) { return gemm_ws_on(c, g, c->st); }
static GemmArgs gemm_args(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int M, int N, int K)
{
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  return g;
}

// CUDA-event bracket of a whole C-ABI call on the handle's stream (always on; read with gpss_get_last_call_ms)
struct CallTimer {
  gpss_ctx* c;
  explicit CallTimer(gpss_ctx* c_) : c(c_) { cudaEventRecord(c->ev_call[0], c->st); }
  ~CallTimer()
  {
    cudaEventRecord(c->ev_call[1], c->st);
    cudaEventSynchronize(c->ev_call[1]);
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_call[0], c->ev_call[1]);
    c->last_call_ms = ms;
  }
};

struct PhaseTimer {
  gpss_ctx* c; int idx;
  PhaseTimer(gpss_ctx* c_, int idx_) : c(c_), idx(idx_) { if (c->profiling) cudaEventRecord(c->ev[0], c->st); }
  ~PhaseTimer()
  {
    if (c->profiling) {
      cudaEventRecord(c->ev[1], c->st);
      cudaEventSynchronize(c->ev[1]);
      float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
      c->phase_ms[idx] += ms;
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// parameters -> device
// ---------------------------------------------------------------------------------------------------
static void fill_params(const double theta[GPSS_NPAR], const double* centre, DevParams& P, int dim = 3, int kind = 0)
{
  memset(&P, 0, sizeof P);
  for (int j = 0; j < dim; j++) P.c[j] = centre[j];
  P.dim = dim;
  P.kind = kind;
  if (kind == 0) {
    sig_inv(theta, P.S);
    P.lr = theta[7];                                   // InversewidthR: sigInv(3,3) of the 4-column branch (Kernel.cpp:1411-1424)
  } else {
    // EuclDist (Kernel.cpp:1343-1368): D2 = |x - x'|^2 / hyp^2 -> the same pair-distance code with sigInv = (1/hyp) I
    const double ih = 1.0 / theta[0];
    P.S[0] = P.S[4] = P.S[8] = ih;
    P.lr = ih;
    if (kind == 2) P.rbf_c = -0.5 * theta[1];
  }
  const double sig = theta_sigma(kind, theta), sn2 = theta_sn2(kind, theta);
  P.var2 = sig * sig;
  P.bias = theta_bias(kind, theta);
  P.sn2 = sn2;
  P.inv_sn2 = 1 / sn2;
  P.sw = std::sqrt(P.inv_sn2);
  P.sww = P.sw * P.sw;
  P.lp_const = std::log(2.0 * M_PI * sn2) / 2;
}

// ---------------------------------------------------------------------------------------------------
// blocked right-looking Cholesky, two-level (outer NBO = 512 for deep-k trailing updates, inner 128)
// A: n_pad x n_pad lower, in place.  Replaces arma::chol -> dpotrf (GP_Utils.cpp:881,903).
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// distributed evaluation helpers (world > 1): staging buffer, balanced row partitions
// ---------------------------------------------------------------------------------------------------
static int ensure_stage(gpss_ctx* c, size_t count)
{
  if (c->stage_count >= count) return GPSS_OK;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  c->stage_count = 0;
  CU(cudaMalloc(&c->stage, count * sizeof(double)));
  c->stage_count = count;
  return GPSS_OK;
}

// Row boundaries (multiples of 128) that give every rank the same share of work:
//   kind 0: rows of U = L^-T in the block-column inverse, cost(row i) ~ (n - i)^2 / 2   -> (n - r_k)^3 = n^3 (1 - k/P)
//   kind 1: rows of B^-1 = U U^T (lower),                 cost(row i) ~ (i + 1)(n - i)  -> n x^2/2 - x^3/3 = (k/P) n^3/6
static void balanced_rows(int n_pad, int world, int kind, std::vector<int>& bounds)
{
  bounds.assign(world + 1, 0);
  bounds[world] = n_pad;
  const double n = n_pad;
  for (int k = 1; k < world; k++) {
    const double f = (double)k / world;
    double x;
    if (kind == 0) {
      x = n * (1.0 - std::cbrt(1.0 - f));
    } else {
      double lo = 0.0, hi = n;
      const double target = f * n * n * n / 6.0;
      for (int it = 0; it < 100; it++) {
        const double mid = 0.5 * (lo + hi);
        if (n * mid * mid / 2.0 - mid * mid * mid / 3.0 < target) lo = mid; else hi = mid;
      }
      x = 0.5 * (lo + hi);
    }
    int b = (int)std::lround(x / NB) * NB;
    if (b < bounds[k - 1]) b = bounds[k - 1];
    if (b > n_pad) b = n_pad;
    bounds[k] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// The distributed Cholesky as a per-rank list of operations (pure host logic; gpss_dist_potrf_schedule exposes it so the
// CPU tests can replay all ranks and check that every block column sees every earlier panel exactly once, in an order
// the broadcasts make possible).  Block column j (owner j % P) receives, all on ONE low-priority side stream (they
// update the same tiles, so they serialise anyway):
//     chunk A(j):   panels 0 .. j-P        one long-k GEMM, issued as soon as the owner has factored its previous column
//     single(j,t):  panel t, j-P < t < j-1 (k = NBO), issued when panel t arrives
// and on the main stream U2(j) = panel j-1, the panel factorisation and the broadcast.  A rank therefore always has
// about P panel periods of bulk work queued behind the critical path instead of one.
// ---------------------------------------------------------------------------------------------------
enum { DIST_WAIT_SIDE = 0, DIST_UPDATE_MAIN = 1, DIST_FACTOR = 2, DIST_BCAST = 3, DIST_UPDATE_SIDE = 4 };
struct DistOp { int kind, col, pbeg, pcnt, root, stream; };   // update ops apply panels pbeg .. pbeg+pcnt-1 to block column col
static void dist_potrf_schedule(int nblk_o, int P, int me, std::vector<DistOp>& ops)
{
  ops.clear();
  for (int t = 0; t < nblk_o; t++) {
    const bool mine = (t % P) == me;
    if (mine) {
      if (t >= 2) ops.push_back({DIST_WAIT_SIDE, t, 0, 0, 0, 0});
      if (t >= 1) ops.push_back({DIST_UPDATE_MAIN, t, P == 1 ? 0 : t - 1, P == 1 ? t : 1, 0, 0});   // alone: plain left-looking
      ops.push_back({DIST_FACTOR, t, 0, 0, 0, 0});
    }
    ops.push_back({DIST_BCAST, t, 0, 0, t % P, 0});
    int j = t + ((me - t) % P + P) % P;        // my next block column after t
    if (j == t) j = t + P;
    if (j >= nblk_o || t >= j - 1) continue;   // panel j-1 is U2(j)
    const int stream = (j / P) & 1;
    if (mine) ops.push_back({DIST_UPDATE_SIDE, j, 0, t + 1, 0, stream});     // chunk A
    else ops.push_back({DIST_UPDATE_SIDE, j, t, 1, 0, stream});              // one panel
  }
}

// One outer panel: factor the NBO-wide block column starting at K0 (all rows below), 128 columns at a time.
template <class StepDone>
static int potrf_panel(gpss_ctx* c, double* A, long ld, int n_pad, int K0, int nbk, double* Winv, double* logdet_parts, int* dflag,
                       StepDone step_done)
{
  for (int k = K0; k < K0 + nbk; k += NB) {
    double* Akk = A + (long)k * ld + k;
    double* Wk = Winv + (long)(k / NB) * NB * NB;
    potrf_diag_inv_kernel<<<1, DIAG_THREADS, DIAG_SMEM, c->st>>>(Akk, ld, Wk, logdet_parts + k / NB, dflag);
    c->launches++;
    CU(cudaGetLastError());
    const int m = n_pad - k - NB;
    if (m <= 0) { RET(step_done(k)); continue; }
    double* A21 = A + (long)k * ld + (k + NB);
    {  // panel solve, in place: A21 <- A21 * inv(L11)^T, with the 128x64 tile (it shares an SM with a resident
       // trailing-update CTA, which the 128x128 tile cannot).  Columns 64..127 first: they read all 128 input
       // columns; columns 0..63 then need only inputs 0..63 because inv(L11) is lower triangular.
      GemmArgs g1 = gemm_args(A21, ld, Wk + 64, NB, A21 + 64 * ld, ld, m, 64, NB);
      RET(gemm(c, g1));
      GemmArgs g2 = gemm_args(A21, ld, Wk, NB, A21, ld, m, 64, 64);
      RET(gemm(c, g2));
    }
    RET(step_done(k));            // columns k .. k+127 of the factor are final from here on
    const int ncols = K0 + nbk - (k + NB);
    if (ncols > 0) {  // update of the remaining columns of the outer panel
      double* A22 = A + (long)(k + NB) * ld + (k + NB);
      GemmArgs g = gemm_args(A21, ld, A21, ld, A22, ld, m, ncols, NB);
      g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = k + NB; g.gcol0 = k + NB;
      RET(gemm(c, g));
    }
  }
  return GPSS_OK;
}

static int potrf_panel(gpss_ctx* c, double* A, long ld, int n_pad, int K0, int nbk, double* Winv, double* logdet_parts, int* dflag)
{
  return potrf_panel(c, A, ld, n_pad, K0, nbk, Winv, logdet_parts, dflag, [](int) { return (int)GPSS_OK; });
}

// LEFT-looking blocked Cholesky with look-ahead.  Block column T (width NBO) receives
//     U1(T):  A[T:, T] -= L[T:, 0:T-1] L[T, 0:T-1]^T     (panels 0..T-2: one long-k DMMA GEMM, side stream)
//     U2(T):  A[T:, T] -= L[T:, T-1]   L[T, T-1]^T       (panel T-1, k = NBO, main stream)
// and is then factored by potrf_panel on the main stream.  U1(T+1) only needs panels <= T-1, so it runs on the
// side stream WHILE the main stream does U2(T) and the latency-bound panel T: the DMMA pipe never waits for a
// panel, every output tile is written once per update instead of once per outer step (the right-looking form
// re-read and re-wrote the whole trailing matrix n/NBO times), and nearly all flops run in long-k GEMMs.
static int potrf_blocked(gpss_ctx* c, double* A, long ld, int n_pad, double* Winv, double* logdet_parts, int* dflag)
{
  const int P = c->world, me = c->rank;
  const bool la = c->st2 != nullptr && !getenv("GPSS_NO_LOOKAHEAD");
  if (P > 1 && !la) return fail_arg("the distributed factorisation needs the look-ahead streams");
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  if (la && (int)c->ev_pool.size() < 2 * nblk_o + 2) {
    const size_t want = 2 * nblk_o + 2;
    while (c->ev_pool.size() < want) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->ev_pool.push_back(e);
    }
  }
  // staging layout of one broadcast: [panel rows T0.. x nbT | the panel's 128x128 diagonal inverses | their log-dets]
  const size_t stage_need = (size_t)n_pad * NBO + (size_t)(NBO / NB) * NB * NB + NBO / NB;
  if (P > 1) RET(ensure_stage(c, stage_need));
  auto update = [&](int T0, int nbT, int kbeg, int klen, cudaStream_t stream) -> int {
    // A[T0:, T0:T0+nbT] -= L[T0:, kbeg:kbeg+klen] L[T0:T0+nbT, kbeg:kbeg+klen]^T
    const double* Lp = A + (long)kbeg * ld + T0;
    GemmArgs g = gemm_args(Lp, ld, Lp, ld, A + (long)T0 * ld + T0, ld, n_pad - T0, nbT, klen);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = T0; g.gcol0 = T0;
    return gemm_ws_on(c, g, stream);
  };
  if (P > 1) {
    std::vector<DistOp> ops;
    dist_potrf_schedule(nblk_o, P, me, ops);
    int kchunk = 1 << 30;
    if (const char* e = getenv("GPSS_DIST_KCHUNK")) { const int v = atoi(e); if (v >= NBO) kchunk = (v / NBO) * NBO; }
    // GPSS_DIST_TRACE: timing events around every main-stream step, summed per kind after the factorisation (diagnostic)
    const bool trace = getenv("GPSS_DIST_TRACE") != nullptr;
    std::vector<std::pair<int, cudaEvent_t>> marks;
    auto mark = [&](int what) {
      if (!trace) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->st);
      marks.push_back({what, e});
    };
    mark(-1);
    // GPSS_DIST_PIPE=1 (experimental): the panel travels in its four 128-column sub-panels, each broadcast -- on a separate
    // communication stream -- as soon as the owner's step has finalised it, and the next owner applies U2 in four k = 128
    // pieces as they arrive: U2 and 3/4 of the broadcast overlap the factorisation instead of following it.
    const bool pipe = getenv("GPSS_DIST_PIPE") != nullptr && atoi(getenv("GPSS_DIST_PIPE")) != 0;
    const int nsub_all = n_pad / NB;
    if (pipe) {
      if (!c->st4) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->st4, cudaStreamNonBlocking, hi));
      }
      while ((int)c->ev_pipe.size() < 2 * nsub_all + 2) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_pipe.push_back(e);
      }
      CU(cudaEventRecord(c->ev_main, c->st));                             // the K build precedes everything on the comm stream too
      CU(cudaStreamWaitEvent(c->st4, c->ev_main, 0));
    }
    for (const DistOp& op : ops) {
      const int T0 = op.col * NBO;
      const int nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
      if (pipe && (op.kind == DIST_UPDATE_MAIN || op.kind == DIST_FACTOR || op.kind == DIST_BCAST)) {
        if (op.kind == DIST_UPDATE_MAIN) {
          // U2 in k = 128 pieces, each as soon as its sub-panel has been received (op.pcnt == 1 whenever P > 1)
          const int Kp = op.pbeg * NBO;
          const int nbK = (n_pad - Kp < NBO) ? (n_pad - Kp) : NBO;
          for (int k0 = 0; k0 < nbK; k0 += NB) {
            CU(cudaStreamWaitEvent(c->st, c->ev_pipe[2 * ((Kp + k0) / NB) + 1], 0));
            RET(update(T0, nbT, Kp + k0, NB, c->st));
          }
        } else if (op.kind == DIST_FACTOR) {
          RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag, [&](int k) -> int {
            CU(cudaEventRecord(c->ev_pipe[2 * (k / NB)], c->st));           // sub-panel k is final on the owner
            return GPSS_OK;
          }));
        } else {
          const bool mine = op.root == me;
          for (int k = T0; k < T0 + nbT; k += NB) {
            const long rows = n_pad - k;
            const size_t n_panel = (size_t)rows * NB, n_w = (size_t)NB * NB;
            double* Wk = Winv + (size_t)(k / NB) * NB * NB;
            if (mine) {
              CU(cudaStreamWaitEvent(c->st4, c->ev_pipe[2 * (k / NB)], 0));
              pack_kernel<<<592, 256, 0, c->st4>>>(c->stage, A + (long)k * ld + k, ld, rows, NB);
              CU(cudaMemcpyAsync(c->stage + n_panel, Wk, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              CU(cudaMemcpyAsync(c->stage + n_panel + n_w, logdet_parts + k / NB, sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              c->launches++;
            }
            NC(g_nccl.Broadcast(c->stage, c->stage, n_panel + n_w + 1, ncclDouble, op.root, c->comm, c->st4));
            if (!mine) {
              unpack_kernel<<<592, 256, 0, c->st4>>>(A + (long)k * ld + k, ld, c->stage, rows, NB);
              CU(cudaMemcpyAsync(Wk, c->stage + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              CU(cudaMemcpyAsync(logdet_parts + k / NB, c->stage + n_panel + n_w, sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              c->launches++;
            }
            CU(cudaEventRecord(c->ev_pipe[2 * (k / NB) + 1], c->st4));      // sub-panel k is complete on this rank
          }
          CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st4));              // panel op.col is complete on this rank
        }
        continue;
      }
      switch (op.kind) {
        case DIST_WAIT_SIDE:                                               // every side-stream update of my column
          CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * op.col + 1], 0));
          mark(0);
          break;
        case DIST_UPDATE_MAIN:                                             // U2: the panel just received, on the critical path
          RET(update(T0, nbT, op.pbeg * NBO, op.pcnt * NBO, c->st));
          mark(1);
          break;
        case DIST_FACTOR:
          RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
          mark(2);
          break;
        case DIST_BCAST: {
          // the owner's finished block column (+ its diagonal inverses and log-dets) goes to every rank: after the loop L,
          // Winv and logdet_parts are replicated.  One NCCL broadcast per panel (<= 205 MB at n = 50k), on the main stream.
          const bool mine = op.root == me;
          const long rows = n_pad - T0;
          const size_t n_panel = (size_t)rows * nbT, n_w = (size_t)(nbT / NB) * NB * NB, n_l = nbT / NB;
          double* Wt = Winv + (size_t)(T0 / NB) * NB * NB;
          if (mine) {
            pack_kernel<<<592, 256, 0, c->st>>>(c->stage, A + (long)T0 * ld + T0, ld, rows, nbT);
            CU(cudaMemcpyAsync(c->stage + n_panel, Wt, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CU(cudaMemcpyAsync(c->stage + n_panel + n_w, logdet_parts + T0 / NB, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            c->launches++;
          }
          if (mine) mark(3);
          NC(g_nccl.Broadcast(c->stage, c->stage, n_panel + n_w + n_l, ncclDouble, op.root, c->comm, c->st));
          mark(mine ? 4 : 5);
          if (!mine) {
            unpack_kernel<<<592, 256, 0, c->st>>>(A + (long)T0 * ld + T0, ld, c->stage, rows, nbT);
            CU(cudaMemcpyAsync(Wt, c->stage + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CU(cudaMemcpyAsync(logdet_parts + T0 / NB, c->stage + n_panel + n_w, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            c->launches++;
          }
          CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st));            // panel op.col is complete on this rank
          if (!mine) mark(6);
          break;
        }
        case DIST_UPDATE_SIDE: {                                           // look-ahead: panels pbeg .. pbeg+pcnt-1 -> my column
          cudaStream_t side = op.stream ? c->st2 : c->st3;
          CU(cudaStreamWaitEvent(side, c->ev_pool[2 * (op.pbeg + op.pcnt - 1)], 0));
          int klen = op.pcnt * NBO;
          if (op.pbeg * NBO + klen > n_pad) klen = n_pad - op.pbeg * NBO;
          // Optional cut of the long-k chunk into launches of <= kchunk (GPSS_DIST_KCHUNK).  Measured at 8 GPUs, n = 50k:
          // potrf 219 / 219 / 223 / 226 ms for kchunk = inf / 8192 / 4096 / 2048 (profiles/r01_dist_kchunk_sweep_8gpu.log),
          // i.e. the critical path is NOT waiting for CTA slots held by long-lived bulk CTAs; default: one launch.
          for (int k0 = 0; k0 < klen; k0 += kchunk)
            RET(update(T0, nbT, op.pbeg * NBO + k0, (klen - k0 < kchunk) ? (klen - k0) : kchunk, side));
          CU(cudaEventRecord(c->ev_pool[2 * op.col + 1], side));
          break;
        }
      }
    }
    if (pipe) CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (nblk_o - 1)], 0));   // the last panel has arrived on the comm stream
    if (trace) {
      CU(cudaStreamSynchronize(c->st));
      static const char* names[7] = {"wait for look-ahead updates", "U2 (panel j-1 -> my column)", "panel factorisation", "pack", "broadcast (as root)",
                                     "broadcast (as receiver, incl. waiting for the owner)", "unpack"};
      double sum[7] = {0, 0, 0, 0, 0, 0, 0};
      int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      for (size_t i = 1; i < marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
        sum[marks[i].first] += ms;
        cnt[marks[i].first]++;
      }
      for (auto& m : marks) cudaEventDestroy(m.second);
      fprintf(stderr, "[gpss dist trace] rank %d of %d, n_pad %d:", me, P, n_pad);
      for (int k = 0; k < 7; k++) fprintf(stderr, " %s: %.1f ms / %d;", names[k], sum[k], cnt[k]);
      fprintf(stderr, "\n");
    }
  } else
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO;
    const int nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    if (!la) {
      if (t >= 1) RET(update(T0, nbT, 0, T0, c->st));
      RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
      continue;
    }
    cudaEvent_t evP = c->ev_pool[2 * t], evU = c->ev_pool[2 * t + 1];
    if (t >= 2) CU(cudaStreamWaitEvent(c->st, evU, 0));              // U1(t) was issued on the side stream below
    if (t >= 1) RET(update(T0, nbT, T0 - NBO, NBO, c->st));          // U2(t)
    RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
    CU(cudaEventRecord(evP, c->st));
    // issue U1(t+1) = panels 0..t-1 applied to block column t+1; needs panel t-1 (complete: main stream order) --
    // here, right after panel t was ENQUEUED, the side stream must only wait for panel t-1.
    if (t + 1 < nblk_o && t >= 1) {
      const int T1 = T0 + NBO;
      const int nb1 = (n_pad - T1 < NBO) ? (n_pad - T1) : NBO;
      cudaStream_t side = (t & 1) ? c->st2 : c->st3;
      CU(cudaStreamWaitEvent(side, c->ev_pool[2 * (t - 1)], 0));
      RET(update(T1, nb1, 0, T0, side));
      CU(cudaEventRecord(c->ev_pool[2 * (t + 1) + 1], side));
    }
  }
  if (P > 1) {   // a failed pivot anywhere must be seen everywhere
    NC(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMax, c->comm, c->st));
  }
  return GPSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// PARTITIONED storage (n_pad^2 too large to replicate, e.g. n = 200 000: 320 GB): every rank keeps only the block columns
// it owns (j % P == rank), packed side by side (40 GB per rank at n = 200 000, P = 8).  RIGHT-looking factorisation: the
// owner factors block column t and broadcasts it; the broadcast buffer itself is the GEMM operand with which every rank
// updates its own remaining block columns (one k = 512 DMMA launch per panel over all of them, lower-triangle tiles only
// through the cyclic column map of gemm_nt_ws_kernel).  Look-ahead of one panel: the owner of t+1 updates that single
// column on the high-priority stream, factors and broadcasts it while the bulk update with panel t is still running.
// ---------------------------------------------------------------------------------------------------
static int potrf_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ld = n_pad;
  double* A = c->Lm;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  while ((int)c->ev_pool.size() < 2 * nblk_o + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  const size_t per = (size_t)n_pad * NBO + (size_t)(NBO / NB) * NB * NB + NBO / NB;
  RET(ensure_stage(c, 3 * per));                                         // panels t, t-1, t-2 stay live (see the look-ahead below)
  // update of my local block columns [q0, q0 + cnt) with panel t, which lies in its broadcast buffer (rows T0.., ld = rows)
  auto update = [&](int t, int q0, int cnt, cudaStream_t stream) -> int {
    if (cnt <= 0) return GPSS_OK;
    const double* pan = c->stage + (size_t)(t % 3) * per;
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    const long rows = n_pad - T0;
    const int Rb = (q0 * P + me) * NBO;                                  // first global row (= first global column) touched
    long ncols = (long)cnt * NBO;
    if ((long)q0 * NBO + ncols > c->lcols) ncols = c->lcols - (long)q0 * NBO;   // ragged last block column
    GemmArgs g = gemm_args(pan + (Rb - T0), rows, pan, rows, A + (long)q0 * NBO * ld + Rb, ld, n_pad - Rb, (int)ncols, nbT);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = Rb;
    g.cyc_P = P; g.cyc_me = me; g.cyc_w = NBO; g.cyc_lcol0 = q0 * NBO; g.cyc_boff = T0;
    return gemm_ws_on(c, g, stream);
  };
  // Look-ahead: the bulk update with panel s (side stream) covers my block columns j >= s + 3 only; column j receives
  // panels j-2 and j-1 on the MAIN stream when panel j-1 arrives.  The critical path (two k = 512 updates of one column,
  // the panel factorisation, the broadcast) therefore waits for the bulk update that finished a whole panel period
  // earlier (s = j - 3), never for the one in flight.
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    const bool mine = (t % P) == me;
    double* buf = c->stage + (size_t)(t % 3) * per;
    const long rows = n_pad - T0;
    const size_t n_panel = (size_t)rows * nbT, n_w = (size_t)(nbT / NB) * NB * NB, n_l = nbT / NB;
    double* Wt = c->Winv + (size_t)(T0 / NB) * NB * NB;
    // bulk update t-3 read this buffer and was the last side-stream launch to write block column t
    if (t >= 3) CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (t - 3) + 1], 0));
    if (mine) {
      const int q = t / P;
      if (t >= 2) RET(update(t - 2, q, 1, c->st));
      if (t >= 1) RET(update(t - 1, q, 1, c->st));
      double* Acol = A + (long)q * NBO * ld;                              // my packed copy of global block column t
      RET(potrf_panel(c, Acol - (long)T0 * ld, ld, n_pad, T0, nbT, c->Winv, c->logdet_parts, c->dflag));   // indexes by global column
      pack_kernel<<<592, 256, 0, c->st>>>(buf, Acol + T0, ld, rows, nbT);
      CU(cudaMemcpyAsync(buf + n_panel, Wt, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      CU(cudaMemcpyAsync(buf + n_panel + n_w, c->logdet_parts + T0 / NB, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      c->launches++;
    }
    NC(g_nccl.Broadcast(buf, buf, n_panel + n_w + n_l, ncclDouble, t % P, c->comm, c->st));
    if (!mine) {
      CU(cudaMemcpyAsync(Wt, buf + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      CU(cudaMemcpyAsync(c->logdet_parts + T0 / NB, buf + n_panel + n_w, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    }
    CU(cudaEventRecord(c->ev_pool[2 * t], c->st));
    int q0 = 0;                                                          // my first block column j >= t + 3
    while (q0 < c->nq && q0 * P + me < t + 3) q0++;
    CU(cudaStreamWaitEvent(c->st2, c->ev_pool[2 * t], 0));
    RET(update(t, q0, c->nq - q0, c->st2));
    CU(cudaEventRecord(c->ev_pool[2 * t + 1], c->st2));
  }
  CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (nblk_o - 1) + 1], 0));
  NC(g_nccl.AllReduce(c->dflag, c->dflag, 1, ncclInt, ncclMax, c->comm, c->st));
  return GPSS_OK;
}

// alpha = L^-T L^-1 rhs with the partitioned factor.  Forward: the owner of block column t runs its four 128-steps and
// broadcasts the updated tail of the right-hand side and the finished piece of z.  Backward: every rank keeps the
// right-hand side current at the columns it owns and updates them with each new x_k; the owner of tile k-1 produces
// x_{k-1}, broadcast 128 doubles at a time.  rhs in c->rvec (destroyed), result in c->alpha (replicated).
static int potrs_vec_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad, nblk = c->nblk;
  const long ld = n_pad;
  const int w = NBO / NB;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  if (me == 0) { trsv_fwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv, c->rvec, c->zvec); c->launches++; }
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    if ((t % P) == me) {
      const double* Lg = c->Lm + (long)(t / P) * NBO * ld - (long)T0 * ld;     // indexed by global column inside my block column
      for (int k = T0 / NB; k < (T0 + nbT) / NB && k + 1 < nblk; k++) {
        trsv_fwd_step_kernel<<<nblk - 1 - k, TRSV_THREADS, 0, c->st>>>(Lg, ld, c->Winv, c->rvec, c->zvec, k * NB);
        c->launches++;
      }
    }
    const int zend = (T0 + nbT + NB <= n_pad) ? T0 + nbT + NB : n_pad;          // z of this block column and of the next tile
    NC(g_nccl.Broadcast(c->zvec + T0, c->zvec + T0, (size_t)(zend - T0), ncclDouble, t % P, c->comm, c->st));
    if (T0 + nbT < n_pad)
      NC(g_nccl.Broadcast(c->rvec + T0 + nbT, c->rvec + T0 + nbT, (size_t)(n_pad - T0 - nbT), ncclDouble, t % P, c->comm, c->st));
  }
  CU(cudaGetLastError());
  const int own_last = ((nblk - 1) / w) % P;
  if (me == own_last) {
    trsv_bwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv + (long)(nblk - 1) * NB * NB, c->zvec, c->alpha, (nblk - 1) * NB);
    c->launches++;
  }
  NC(g_nccl.Broadcast(c->alpha + (long)(nblk - 1) * NB, c->alpha + (long)(nblk - 1) * NB, NB, ncclDouble, own_last, c->comm, c->st));
  const int ltiles = (int)(c->lcols / NB);
  for (int k = nblk - 1; k >= 1; k--) {
    trsv_bwd_step_part_kernel<<<ltiles, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->zvec, c->alpha, k * NB, P, me, w);
    c->launches++;
    const int owner = ((k - 1) / w) % P;
    NC(g_nccl.Broadcast(c->alpha + (long)(k - 1) * NB, c->alpha + (long)(k - 1) * NB, NB, ncclDouble, owner, c->comm, c->st));
  }
  CU(cudaGetLastError());
  return GPSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// PARTITIONED storage, gradient: B^-1 = U U^T with U = L^-T held as block ROWS owned cyclically (rank r keeps rows
// I = q P + r, packed: local row block q), 40 GB per rank at n = 200 000 like L itself.
//   inverse (left-looking over block columns J of U, as trtri_upper):
//       every rank contributes its blocks of the ROW strip L[J, 0:J] (ncclAllGather, then laid out in global column order);
//       the owner of J inverts the diagonal block and broadcasts W_JJ = inv(L_JJ);
//       T = U_loc[rows < J, 0:J0] L[J, 0:J0]^T  (k from each row's own start: cyclic row map of the GEMM kernel, split-k for
//       short slices),  U_loc[rows < J, J] = -T W_JJ^T.
//   B^-1 and the gradient, one block column J at a time: the owner broadcasts the row strip U[J, J0:], every rank forms
//       Q[I >= J, J] = U_loc[I, J0:] U[J, J0:]^T for its rows and feeds the 512-wide strip straight into the fused gradient
//       reductions -- B^-1 is never stored.
// Everything on the main stream; results (13 sums) all-reduced at the end.
// ---------------------------------------------------------------------------------------------------
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj);
static int pick_ksplit(int tiles, int klen, long part_doubles, size_t cap_doubles);
static int ensure_lazy(double** p, size_t count);

static int part_buffers(gpss_ctx* c)
{
  const size_t ldu = (size_t)(c->nq > 0 ? c->nq : 1) * NBO;
  RET(ensure_lazy(&c->Um, ldu * c->n_pad));
  RET(ensure_lazy(&c->Tpanel, ldu * NBO));                       // T, later the Q strip
  RET(ensure_lazy(&c->Wjj, (size_t)2 * NBO * NBO));              // W_JJ and the owner's U_JJ scratch
  const int nblk_o = (c->n_pad + NBO - 1) / NBO;
  const size_t cmax = (size_t)(nblk_o + c->world - 1) / c->world;
  RET(ensure_lazy(&c->pgather, (size_t)(c->world + 1) * cmax * NBO * NBO));   // [my piece | P gathered pieces]
  if (!c->Tsplit) {
    const size_t cap = (size_t)24576 * NBO;
    CU(cudaMalloc(&c->Tsplit, cap * sizeof(double)));
    c->Tsplit_cap = cap;
  }
  const long nblocks = (long)(ldu / NB) * c->nblk;               // every (local row tile, global column tile)
  if (c->partial_blocks < nblocks || !c->partial) {
    if (c->partial) cudaFree(c->partial);
    c->partial = nullptr;
    CU(cudaMalloc(&c->partial, sizeof(double) * (nblocks > 0 ? nblocks : 1) * NGRAD));
    c->partial_blocks = nblocks;
  }
  return GPSS_OK;
}

static int trtri_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ld = n_pad;
  const long ldu = (long)(c->nq > 0 ? c->nq : 1) * NBO;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  const int cmax = (nblk_o + P - 1) / P;
  const size_t blk = (size_t)NBO * NBO;
  double* piece = c->pgather;                                     // my blocks of the row strip
  double* gathered = c->pgather + (size_t)cmax * blk;             // P segments of cmax blocks
  double* Lrow = c->stage;                                        // the strip in global column order (the panel buffers are free now)
  double* Wjj = c->Wjj;
  double* Ujj = c->Wjj + blk;
  CU(cudaMemsetAsync(c->Um, 0, sizeof(double) * (size_t)ldu * n_pad, c->st));
  for (int J = 0; J < nblk_o; J++) {
    const int J0 = J * NBO, nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    const int owner = J % P;
    if (owner == me) {
      // U_JJ from my packed copy of block column J (global addressing through shifted base pointers), W_JJ = U_JJ^T
      const double* Lg = c->Lm + (long)(J / P) * NBO * ld - (long)J0 * ld;
      double* Ug = Ujj - ((long)J0 * NBO + J0);
      CU(cudaMemsetAsync(Ujj, 0, sizeof(double) * blk, c->st));
      RET(trtri_diag_block(c, Ug, NBO, Lg, ld, J0, nbj));
      transpose_kernel<<<dim3(nbj / 32, nbj / 32), 256, 0, c->st>>>(Wjj, NBO, Ujj, NBO, 1);
      c->launches++;
      CU(cudaGetLastError());
    }
    NC(g_nccl.Broadcast(Wjj, Wjj, blk, ncclDouble, owner, c->comm, c->st));
    int cnt = 0;                                                  // my row blocks above J = my column blocks left of J
    while (cnt < c->nq && cnt * P + me < J) cnt++;
    if (J > 0) {
      const int jc = (J + P - 1) / P;                             // blocks per segment needed for this J (<= cmax)
      if (cnt > 0) {
        pack_rowstrip_kernel<<<592, 256, 0, c->st>>>(piece, c->Lm, ld, J0, nbj, NBO, cnt);
        c->launches++;
      }
      NC(g_nccl.AllGather(piece, gathered, (size_t)jc * blk, ncclDouble, c->comm, c->st));
      order_rowstrip_kernel<<<592, 256, 0, c->st>>>(Lrow, gathered, nbj, NBO, J, P, (long)jc * (long)blk);
      c->launches++;
      CU(cudaGetLastError());
    }
    if (cnt > 0 && J > 0) {
      const int rows = cnt * NBO;
      // T = U_loc[0:rows, 0:J0] * Lrow^T, k from each row's own global start
      GemmArgs g = gemm_args(c->Um, ldu, Lrow, nbj, c->Tpanel, ldu, rows, nbj, J0);
      g.kbeg_row = 1; g.rcyc_P = P; g.rcyc_me = me; g.rcyc_w = NBO; g.rcyc_l0 = 0; g.rcyc_koff = 0;
      const int S = pick_ksplit(rows / GemmTileWideWS::BM * (nbj / GemmTileWideWS::BN), J0, (long)rows * nbj, c->Tsplit_cap);
      if (S > 1) {
        g.C = c->Tsplit; g.ldc = rows; g.ksplit = S; g.csplit = (long)rows * nbj;
        RET(gemm(c, g));
        split_sum_kernel<<<296, 256, 0, c->st>>>(c->Tpanel, ldu, c->Tsplit, rows, nbj, S);
        c->launches++;
        CU(cudaGetLastError());
      } else {
        RET(gemm(c, g));
      }
      GemmArgs g2 = gemm_args(c->Tpanel, ldu, Wjj, NBO, c->Um + (long)J0 * ldu, ldu, rows, nbj, nbj);
      g2.negate_out = 1; g2.kend_col = 1;
      RET(gemm(c, g2));
    }
    if (owner == me) {                                            // my diagonal block
      copy2d_kernel<<<64, 256, 0, c->st>>>(c->Um + (long)J0 * ldu + (long)(J / P) * NBO, ldu, Ujj, NBO, nbj, nbj);
      c->launches++;
      CU(cudaGetLastError());
    }
  }
  return GPSS_OK;
}

static int gradient_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ldu = (long)(c->nq > 0 ? c->nq : 1) * NBO;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  const int ltiles = (int)(ldu / NB);                             // my local row tiles (128 high)
  const int w = NBO / NB;
  double* strip = c->stage;                                       // U[J, J0:] as nbj x (n_pad - J0), contiguous
  double* Qs = c->Tpanel;                                         // Q[my rows >= J, J], ld = ldu
  const long nblocks = (long)ltiles * c->nblk;
  CU(cudaMemsetAsync(c->partial, 0, sizeof(double) * nblocks * NGRAD, c->st));
  for (int J = 0; J < nblk_o; J++) {
    const int J0 = J * NBO, nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    const int owner = J % P;
    const long cols = n_pad - J0;
    if (owner == me) {
      pack_kernel<<<592, 256, 0, c->st>>>(strip, c->Um + (long)J0 * ldu + (long)(J / P) * NBO, ldu, nbj, cols);
      c->launches++;
    }
    NC(g_nccl.Broadcast(strip, strip, (size_t)nbj * cols, ncclDouble, owner, c->comm, c->st));
    int q0 = 0;                                                   // my first row block I >= J
    while (q0 < c->nq && q0 * P + me < J) q0++;
    const int rows = (c->nq - q0) * NBO - ((q0 < c->nq && (c->nq - 1) * P + me == nblk_o - 1) ? (NBO - (n_pad - (nblk_o - 1) * NBO)) : 0);
    if (rows <= 0) continue;
    // Q strip = U_loc[q0 rows.., J0:] * strip^T, k from each row's own start (relative to J0)
    GemmArgs g = gemm_args(c->Um + (long)J0 * ldu + (long)q0 * NBO, ldu, strip, nbj, Qs, ldu, rows, nbj, (int)cols);
    g.kbeg_row = 1; g.rcyc_P = P; g.rcyc_me = me; g.rcyc_w = NBO; g.rcyc_l0 = q0 * NBO; g.rcyc_koff = J0;
    RET(gemm(c, g));
    // fused gradient reductions over the strip: local row tiles q0*w .., global column tiles J*w ..
    const int ntm = rows / NB, ntn = nbj / NB;
    grad_pass_kernel<<<dim3(ntm, ntn), 256, 0, c->st>>>(Qs, ldu, c->zs, n_pad, c->xs, n_pad, c->alpha, c->n, c->dP,
                                                       c->partial + (long)(J * w) * ltiles * NGRAD, q0 * w, J * w, P, me, w);
    c->launches++;
    CU(cudaGetLastError());
  }
  sum_partials_kernel<NGRAD><<<1, 256, 0, c->st>>>(c->partial, nblocks, c->red + 8);
  c->launches++;
  CU(cudaGetLastError());
  NC(g_nccl.AllReduce(c->red + 8, c->red + 8, NGRAD, ncclDouble, ncclSum, c->comm, c->st));
  return GPSS_OK;
}

static int create_streams(gpss_ctx* c)
{
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (largest number), hi = greatest
  CU(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, hi));
  CU(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, lo));
  CU(cudaStreamCreateWithPriority(&c->st3, cudaStreamNonBlocking, lo));
  CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
  return GPSS_OK;
}

static void destroy_streams(gpss_ctx* c)
{
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  c->ev_pool.clear();
  for (auto e : c->ev_pipe) cudaEventDestroy(e);
  c->ev_pipe.clear();
  if (c->st4) cudaStreamDestroy(c->st4);
  c->st4 = nullptr;
  if (c->st2) cudaStreamDestroy(c->st2);
  if (c->st3) cudaStreamDestroy(c->st3);
  if (c->st) cudaStreamDestroy(c->st);
  c->ev_main = c->ev_side = nullptr;
  c->st = c->st2 = c->st3 = nullptr;
}

// Number of k-parts for a GEMM of `tiles` output tiles on 2 x 148 CTA slots: the smallest S whose CTA count fills
// whole waves best, subject to parts of >= 2048 in k and to the capacity of the partial-product buffer.
static int pick_ksplit(int tiles, int klen, long part_doubles, size_t cap_doubles)
{
  const int slots = 296;
  if (tiles <= 0 || tiles >= 4 * slots) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= 8; S++) {
    if (S > 1 && (klen / S < 2048 || (size_t)part_doubles * S > cap_doubles)) break;
    const int ctas = tiles * S;
    const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
    if (eff > best_eff + 0.03) { best_eff = eff; best = S; }
  }
  return best;
}

// U = L^-T (upper), block-column left-looking; only NT products (see gpss_gemm.cuh header):
//     U[J,J]   = inv(L[J,J])^T                       built from the stored 128x128 inverses      (main stream)
//     U[0:J,J] = -(U[0:J,0:J] L[J,0:J]^T) U[J,J]     one long-k GEMM + one k = NBO GEMM           (side stream)
// The diagonal blocks depend only on L, so the main stream produces them (latency-bound small launches) ahead of
// the side stream, which runs the bulk GEMMs back to back.
// The diagonal block U[J0:J0+nbj, J0:J0+nbj] = inv(L[J0.., J0..])^T in 128-steps from the stored 128 x 128 inverses (main stream).
// U and L are addressed by GLOBAL row / column (callers with packed storage pass suitably shifted base pointers).
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj)
{
  for (int i0 = J0; i0 < J0 + nbj; i0 += NB) {
    const double* Wi = c->Winv + (long)(i0 / NB) * NB * NB;
    put_transposed_block_kernel<<<dim3(NB / 32, NB / 32), 256, 0, c->st>>>(U + (long)i0 * ldu + i0, ldu, Wi);
    c->launches++;
    CU(cudaGetLastError());
    const int mr = i0 - J0;
    if (mr > 0) {
      double* Uc = U + (long)i0 * ldu + J0;                 // U[J0:i0, i0:i0+128]
      GemmArgs g = gemm_args(U + (long)J0 * ldu + J0, ldu, L + (long)J0 * ldl + i0, ldl, Uc, ldu, mr, NB, mr);
      g.kbeg_row = 1;
      RET(gemm(c, g));
      // Uc <- -Uc Wi^T in place: columns 64..127 first (all 128 inputs), then 0..63 (inputs 0..63 only)
      GemmArgs g1 = gemm_args(Uc, ldu, Wi + 64, NB, Uc + 64 * ldu, ldu, mr, 64, NB);
      g1.negate_out = 1;
      RET(gemm(c, g1));
      GemmArgs g2 = gemm_args(Uc, ldu, Wi, NB, Uc, ldu, mr, 64, 64);
      g2.negate_out = 1;
      RET(gemm(c, g2));
    }
  }
  return GPSS_OK;
}

static int trtri_upper(gpss_ctx* c)
{
  const long ld = c->n_pad;
  const int n_pad = c->n_pad;
  double *L = c->Lm, *U = c->Um;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  while ((int)c->ev_pool.size() < 2 * nblk_o + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  // Distributed: every row of U depends only on L and on the SAME row of earlier block columns, so a rank computes
  // the rows [urow0, urow1) of its balanced slice with no communication (the 128-step diagonal blocks, which every
  // rank needs as right factors, are cheap and computed redundantly).
  const int R0 = c->urow0, R1 = c->urow1;
  // the side stream must not start before the factor is complete on the main stream
  CU(cudaEventRecord(c->ev_main, c->st));
  CU(cudaStreamWaitEvent(c->st2, c->ev_main, 0));
  for (int t = 0; t < nblk_o; t++) {
    const int J0 = t * NBO;
    const int nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    double* Wjj = c->Wjj + (size_t)t * NBO * NBO;
    // (1) the diagonal NBO-block of U in 128-steps
    RET(trtri_diag_block(c, U, ld, L, ld, J0, nbj));
    if (t == 0) continue;
    // (2) W_JJ = U_JJ^T into this block's scratch
    transpose_kernel<<<dim3(nbj / 32, nbj / 32), 256, 0, c->st>>>(Wjj, NBO, U + (long)J0 * ld + J0, ld, 1);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_pool[2 * t], c->st));
    CU(cudaStreamWaitEvent(c->st2, c->ev_pool[2 * t], 0));
    const int ra = R0, rb = (R1 < J0) ? R1 : J0;             // my rows above this block column
    if (rb <= ra) continue;
    // (3) T[ra:rb] = U[ra:rb, 0:J0] * L[Jblk, 0:J0]^T      (k starts at each tile's own row: U is upper triangular)
    GemmArgs g = gemm_args(U + ra, ld, L + J0, ld, c->Tpanel + ra, ld, rb - ra, nbj, J0);
    g.kbeg_row = 1; g.krow_off = ra;
    // A row slice has few tiles per step (rank 0 of 8 at n = 50k: 17 x 8 = 136 for 296 CTA slots) and the steps are
    // sequential, so a distributed rank cuts the long k-range of every tile into S parts (one CTA each), sized to
    // fill whole waves; the parts are summed in a fixed order by split_sum_kernel.
    int S = 1;
    if (c->world > 1) S = pick_ksplit((rb - ra) / GemmTileWideWS::BM * (nbj / GemmTileWideWS::BN), J0 - ra, (long)(rb - ra) * nbj, c->Tsplit_cap);
    if (S > 1) {
      const int rows = rb - ra;
      g.C = c->Tsplit; g.ldc = rows; g.ksplit = S; g.csplit = (long)rows * nbj;
      RET(gemm_ws_on(c, g, c->st2));
      split_sum_kernel<<<296, 256, 0, c->st2>>>(c->Tpanel + ra, ld, c->Tsplit, rows, nbj, S);
      c->launches++;
      CU(cudaGetLastError());
    } else {
      RET(gemm_ws_on(c, g, c->st2));
    }
    // (4) U[ra:rb, Jblk] = -T * W_JJ^T
    GemmArgs g2 = gemm_args(c->Tpanel + ra, ld, Wjj, NBO, U + (long)J0 * ld + ra, ld, rb - ra, nbj, nbj);
    g2.negate_out = 1; g2.kend_col = 1;
    RET(gemm_ws_on(c, g2, c->st2));
  }
  CU(cudaEventRecord(c->ev_side, c->st2));
  CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
  return GPSS_OK;
}

// Distributed: every rank has computed its rows of U; B^-1 = U U^T and W = U^T need all of them.  Each rank's slice
// (rows [b_k, b_k+1) x columns b_k..n, a strided region of the column-major buffer) is packed, broadcast and unpacked.
static int allgather_U(gpss_ctx* c)
{
  if (c->world == 1) return GPSS_OK;
  const long ld = c->n_pad;
  std::vector<int> b;
  balanced_rows(c->n_pad, c->world, 0, b);
  size_t need = 0;
  for (int k = 0; k < c->world; k++) need = std::max(need, (size_t)(b[k + 1] - b[k]) * (size_t)(c->n_pad - b[k]));
  RET(ensure_stage(c, need));
  for (int k = 0; k < c->world; k++) {
    const long rows = b[k + 1] - b[k], cols = c->n_pad - b[k];
    if (rows <= 0) continue;
    double* slice = c->Um + (long)b[k] * ld + b[k];
    if (k == c->rank) { pack_kernel<<<1184, 256, 0, c->st>>>(c->stage, slice, ld, rows, cols); c->launches++; }
    NC(g_nccl.Broadcast(c->stage, c->stage, (size_t)rows * cols, ncclDouble, k, c->comm, c->st));
    if (k != c->rank) { unpack_kernel<<<1184, 256, 0, c->st>>>(slice, ld, c->stage, rows, cols); c->launches++; }
  }
  CU(cudaGetLastError());
  return GPSS_OK;
}

// Q (lower) = U U^T = B^-1; a rank computes the rows [qrow0, qrow1) of its balanced slice (all rows when alone)
static int lauum_lower(gpss_ctx* c)
{
  const long ld = c->n_pad;
  const int q0 = c->qrow0, q1 = c->qrow1;
  if (q1 <= q0) return GPSS_OK;
  GemmArgs g = gemm_args(c->Um + q0, ld, c->Um, ld, c->Qm + q0, ld, q1 - q0, q1, c->n_pad);
  g.lower_only = 1; g.kbeg_row = 1; g.krow_off = q0; g.grow0 = q0; g.gcol0 = 0;
  return gemm(c, g);
}

// x = L^-T L^-1 rhs through the stored diagonal inverses; rhs in c->rvec (destroyed), result in c->alpha
static int potrs_vec(gpss_ctx* c)
{
  const long ld = c->n_pad;
  const int nblk = c->nblk;
  trsv_fwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv, c->rvec, c->zvec);
  c->launches++;
  for (int k = 0; k + 1 < nblk; k++) {
    trsv_fwd_step_kernel<<<nblk - 1 - k, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->rvec, c->zvec, k * NB);
    c->launches++;
  }
  CU(cudaGetLastError());
  trsv_bwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv + (long)(nblk - 1) * NB * NB, c->zvec, c->alpha, (nblk - 1) * NB);
  c->launches++;
  for (int k = nblk - 1; k >= 1; k--) {
    trsv_bwd_step_kernel<<<k, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->zvec, c->alpha, k * NB);
    c->launches++;
  }
  CU(cudaGetLastError());
  return GPSS_OK;
}

// dst = src / sn2 (the factor comes from the device parameters, so the launch carries no theta-dependent argument and can sit in a graph)
__global__ void scale_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, const DevParams* __restrict__ P, int n, int n_pad)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (i < n) ? src[i] * P->inv_sn2 : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// objective pieces
// ---------------------------------------------------------------------------------------------------
static int upload_params(gpss_ctx* c, int slot, const double* centre)
{
  DevParams P;
  fill_params(c->theta, centre, P, c->d, c->kind);
  CU(cudaMemcpyAsync(c->dP + slot, &P, sizeof P, cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));   // P is a stack object
  return GPSS_OK;
}

static int upload_train_params(gpss_ctx* c)
{
  double centre[4];
  maha_centre(c->n, c->sums_train, c->n, c->sums_train, centre, c->d);
  return upload_params(c, 0, centre);
}

static int enqueue_factor(gpss_ctx* c);
static int ensure_factor(gpss_ctx* c)
{
  if (c->have_factor) return GPSS_OK;
  RET(upload_train_params(c));
  return enqueue_factor(c);
}

// stream work only (no host synchronisation, no allocation): may be captured into a graph
static int enqueue_factor(gpss_ctx* c)
{
  const int n_pad = c->n_pad;
  const long ld = n_pad;
  CU(cudaMemsetAsync(c->dflag, 0, sizeof(int), c->st));
  {
    PhaseTimer t(c, 0);
    transform_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->xs, ld, c->zs, ld, c->n, n_pad, c->dP);
    if (c->partitioned)
      kbuild_lower_kernel<<<dim3(c->nblk, (unsigned)(c->lcols / NB)), 256, 0, c->st>>>(c->Lm, ld, c->zs, ld, c->n, c->dP, 0, c->world, c->rank,
                                                                                      -(NBO / NB));
    else
      kbuild_lower_kernel<<<dim3(c->nblk, c->nblk), 256, 0, c->st>>>(c->Lm, ld, c->zs, ld, c->n, c->dP, 0, c->world, c->rank, NBO / NB);
    c->launches += 2;
    CU(cudaGetLastError());
  }
  {
    PhaseTimer t(c, 1);
    if (c->partitioned) RET(potrf_partitioned(c));
    else RET(potrf_blocked(c, c->Lm, ld, n_pad, c->Winv, c->logdet_parts, c->dflag));
  }
  c->have_factor = true;
  c->have_alpha = false;
  c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

// alpha, f = K alpha, and the scalar terms; sets c->nlml (GP_Utils.cpp:1138-1162)
static int enqueue_solves(gpss_ctx* c)
{
  const int n_pad = c->n_pad;
  {
    PhaseTimer t(c, 2);
    // rhs = y / sn2: the IRLS fixed point alpha = B^-1 (y/sn2) = (K + sn2 I)^-1 y (GP_Utils.cpp:214-223)
    scale_copy_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->rvec, c->y, c->dP, c->n, n_pad);
    c->launches++;
    if (c->partitioned) RET(potrs_vec_partitioned(c));
    else RET(potrs_vec(c));
    kmatvec_kernel<<<(c->n + 63) / 64, 256, 0, c->st>>>(c->zs, n_pad, c->alpha, c->fvec, c->n, c->dP);
    lml_terms_kernel<<<1, 256, 0, c->st>>>(c->y, c->alpha, c->fvec, c->n, c->dP, c->logdet_parts, c->nblk, c->red);
    c->launches += 2;
    CU(cudaGetLastError());
  }
  return GPSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// CUDA graphs.  For small n an evaluation is a chain of a few hundred short, dependent launches on three streams
// (n = 2000: ~3.6 ms for ~0.3 ms of arithmetic).  The launch sequence depends only on n -- theta reaches the kernels
// through the DevParams block in device memory -- so it is captured ONCE per handle (stream capture follows the
// cross-stream events of the look-ahead) and replayed for every later evaluation.  Single-GPU handles, n_pad <=
// GPSS_GRAPH_MAX_N (default 8192), not while profiling; GPSS_NO_GRAPH=1 disables it.  Any capture error falls back to
// plain launches for the rest of the handle's life.
// ---------------------------------------------------------------------------------------------------
static bool graphs_enabled(const gpss_ctx* c)
{
  if (c->world != 1 || c->partitioned || c->profiling || c->graph_failed || !c->st2) return false;
  if (getenv("GPSS_NO_GRAPH")) return false;
  int max_n = 8192;
  if (const char* e = getenv("GPSS_GRAPH_MAX_N")) max_n = atoi(e);
  return c->n_pad <= max_n;
}

static int ensure_event_pool(gpss_ctx* c, size_t want)
{
  while (c->ev_pool.size() < want) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  return GPSS_OK;
}

static int enqueue_gradient(gpss_ctx* c);
static int ensure_gradient_buffers(gpss_ctx* c);
// run graph `which` (capturing it first if needed); returns 1 if the caller must fall back to plain launches
static int run_graph(gpss_ctx* c, int which)
{
  if (!c->graph[which]) {
    const long l0 = c->launches;
    if (cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_failed = true; return 1; }
    int rc = (which == 0) ? enqueue_factor(c) : enqueue_gradient(c);
    if (rc >= 0 && which == 0) rc = enqueue_solves(c);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->st, &g);
    c->graph_launches[which] = c->launches - l0;
    c->launches = l0;
    if (rc < 0 || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      c->graph_failed = true;
      return 1;
    }
    const cudaError_t ei = cudaGraphInstantiate(&c->graph[which], g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { cudaGetLastError(); c->graph[which] = nullptr; c->graph_failed = true; return 1; }
  }
  CU(cudaGraphLaunch(c->graph[which], c->st));
  c->launches += c->graph_launches[which];
  return GPSS_OK;
}

static int ensure_objective(gpss_ctx* c)
{
  if (!c->have_factor && graphs_enabled(c)) {
    const int nblk_o = (c->n_pad + NBO - 1) / NBO;
    RET(ensure_event_pool(c, 2 * nblk_o + 2));
    RET(upload_train_params(c));
    const int r = run_graph(c, 0);
    if (r < 0) return r;
    if (r == 0) {
      c->have_factor = true;
      c->have_U = false;
      c->qstate = Q_NONE;
    } else {
      RET(enqueue_factor(c));
      RET(enqueue_solves(c));
    }
  } else {
    RET(ensure_factor(c));
    if (c->have_alpha) return GPSS_OK;
    RET(enqueue_solves(c));
  }
  double red[4];
  int flag = 0;
  CU(cudaMemcpyAsync(red, c->red, sizeof red, cudaMemcpyDeviceToHost, c->st));
  CU(cudaMemcpyAsync(&flag, c->dflag, sizeof flag, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  c->chol_fail = flag;
  if (flag) {
    c->nlml = std::numeric_limits<double>::quiet_NaN();
  } else {
    // L = Alpha' * ydif - accu(lp) + Lchol_db2 (GP_Utils.cpp:1159)
    c->nlml = red[0] - red[1] + red[3];
    c->s3 = red[2];
  }
  c->have_alpha = true;
  return GPSS_OK;
}

static int ensure_lazy(double** p, size_t count)
{
  if (*p) return GPSS_OK;
  CU(cudaMalloc(p, count * sizeof(double)));
  return GPSS_OK;
}

static int ensure_U(gpss_ctx* c)
{
  if (c->have_U) return GPSS_OK;
  const size_t nn = (size_t)c->n_pad * c->n_pad;
  RET(ensure_lazy(&c->Um, nn));
  RET(ensure_lazy(&c->Tpanel, (size_t)c->n_pad * NBO));
  RET(ensure_lazy(&c->Wjj, (size_t)NBO * NBO * ((c->n_pad + NBO - 1) / NBO)));
  if (c->world > 1 && !c->Tsplit) {
    const size_t cap = (size_t)24576 * NBO;                    // S * rows <= 24k rows of a 512-wide block column
    CU(cudaMalloc(&c->Tsplit, cap * sizeof(double)));
    c->Tsplit_cap = cap;
  }
  {
    PhaseTimer t(c, 3);
    RET(trtri_upper(c));
  }
  {
    PhaseTimer t(c, 8);
    RET(allgather_U(c));
  }
  c->have_U = true;
  return GPSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int gpss_version(void) { return 100; }

const char* gpss_last_error(void) { return g_last_error.c_str(); }

int gpss_device_count(int* count)
{
  if (!count) return fail_arg("gpss_device_count: null");
  CU(cudaGetDeviceCount(count));
  return GPSS_OK;
}

int gpss_destroy(gpss_handle c)
{
  if (!c) return GPSS_OK;
  cudaSetDevice(c->device);
  double** bufs[] = {&c->xs, &c->y, &c->zs, &c->Lm, &c->Um, &c->Qm, &c->Winv, &c->logdet_parts, &c->rvec, &c->zvec, &c->alpha,
                     &c->fvec, &c->Tpanel, &c->Wjj, &c->partial, &c->red, &c->xt, &c->zt, &c->zsp, &c->Bm, &c->Vm, &c->mu_part,
                     &c->dmu, &c->dvar};
  for (auto b : bufs) if (*b) cudaFree(*b);
  if (c->dP) cudaFree(c->dP);
  if (c->dflag) cudaFree(c->dflag);
  if (c->graph[0]) cudaGraphExecDestroy(c->graph[0]);
  if (c->graph[1]) cudaGraphExecDestroy(c->graph[1]);
  if (c->stage) cudaFree(c->stage);
  if (c->Tsplit) cudaFree(c->Tsplit);
  if (c->pgather) cudaFree(c->pgather);
  if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  if (c->ev_call[0]) cudaEventDestroy(c->ev_call[0]);
  if (c->ev_call[1]) cudaEventDestroy(c->ev_call[1]);
  destroy_streams(c);
  delete c;
  return GPSS_OK;
}

int gpss_set_data(gpss_handle c, const double* X, const double* y)
{
  if (!c || !X || !y) return fail_arg("gpss_set_data: null argument");
  CU(cudaSetDevice(c->device));
  const int n = c->n, n_pad = c->n_pad;
  seq_colsums(X, n, c->sums_train, c->d);
  CU(cudaMemsetAsync(c->xs, 0, sizeof(double) * NX * n_pad, c->st));
  for (int j = 0; j < c->d; j++)
    CU(cudaMemcpyAsync(c->xs + (long)j * n_pad, X + (long)j * n, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->y, 0, sizeof(double) * n_pad, c->st));
  CU(cudaMemcpyAsync(c->y, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));
  c->have_factor = c->have_alpha = c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

static int create_impl(int device, int n, int d, const double* X, const double* y, int part_world, int part_rank, const void* id128,
                       gpss_handle* out)
{
  if (!out || !X || !y) return fail_arg("gpss_create: null argument");
  if (d != 3 && d != 4) return fail_arg("gpss_create: d must be 3, or 4 with a rock-type column (Kernel.cpp:872-878)");
  if (n < 2) return fail_arg("gpss_create: n must be >= 2");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail_arg("gpss_create: no such CUDA device");
  CU(cudaSetDevice(device));
  RET(configure_kernels());
  gpss_ctx* c = new gpss_ctx();
  c->device = device;
  c->n = n;
  c->d = d;
  c->n_pad = ((n + NB - 1) / NB) * NB;
  c->nblk = c->n_pad / NB;
  c->urow0 = c->qrow0 = 0;
  c->urow1 = c->qrow1 = c->n_pad;
  memset(c->phase_ms, 0, sizeof c->phase_ms);
  const size_t np = c->n_pad;
  size_t lm_cols = np;
  if (part_world > 1) {
    // partitioned storage: only my block columns j = q P + rank, packed; the last global block column may be narrower
    const int nblk_o = (c->n_pad + NBO - 1) / NBO;
    c->partitioned = true;
    c->world = part_world;
    c->rank = part_rank;
    c->nq = 0;
    c->lcols = 0;
    for (int j = part_rank; j < nblk_o; j += part_world) {
      c->nq++;
      c->lcols += (c->n_pad - j * NBO < NBO) ? (c->n_pad - j * NBO) : NBO;
    }
    lm_cols = c->lcols > 0 ? (size_t)c->lcols : 1;
  }
  auto fail = [&](int code) { gpss_destroy(c); return code; };
#define CUF(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail(fail_cuda(e__, #x, __LINE__)); } while (0)
  { int r__ = create_streams(c); if (r__ != GPSS_OK) return fail(r__); }
  CUF(cudaEventCreate(&c->ev[0]));
  CUF(cudaEventCreate(&c->ev[1]));
  CUF(cudaEventCreate(&c->ev_call[0]));
  CUF(cudaEventCreate(&c->ev_call[1]));
  CUF(cudaMalloc(&c->xs, sizeof(double) * NX * np));
  CUF(cudaMalloc(&c->y, sizeof(double) * np));
  CUF(cudaMalloc(&c->zs, sizeof(double) * NZ * np));
  CUF(cudaMalloc(&c->Lm, sizeof(double) * np * lm_cols));
  CUF(cudaMalloc(&c->Winv, sizeof(double) * (size_t)c->nblk * NB * NB));
  CUF(cudaMalloc(&c->logdet_parts, sizeof(double) * c->nblk));
  CUF(cudaMalloc(&c->rvec, sizeof(double) * np));
  CUF(cudaMalloc(&c->zvec, sizeof(double) * np));
  CUF(cudaMalloc(&c->alpha, sizeof(double) * np));
  CUF(cudaMalloc(&c->fvec, sizeof(double) * np));
  CUF(cudaMalloc(&c->red, sizeof(double) * 32));
  CUF(cudaMalloc(&c->dP, sizeof(DevParams) * 2));
  CUF(cudaMalloc(&c->dflag, sizeof(int)));
  CUF(cudaMemsetAsync(c->alpha, 0, sizeof(double) * np, c->st));
  CUF(cudaMemsetAsync(c->fvec, 0, sizeof(double) * np, c->st));
#undef CUF
  for (int i = 0; i < GPSS_NPAR; i++) c->theta[i] = 0;
  int r = gpss_set_data(c, X, y);
  if (r != GPSS_OK) return fail(r);
  if (part_world > 1) {
    r = nccl_load();
    if (r != GPSS_OK) return fail(r);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t nr = g_nccl.CommInitRank(&c->comm, part_world, id, part_rank);
    if (nr != ncclSuccess) return fail(fail_nccl(nr, "ncclCommInitRank", __LINE__));
  }
  *out = c;
  return GPSS_OK;
}

int gpss_create(int device, int n, int d, const double* X, const double* y, gpss_handle* out)
{
  return create_impl(device, n, d, X, y, 1, 0, nullptr, out);
}

int gpss_create_partitioned(int device, int rank, int world, const void* id128, int n, int d, const double* X, const double* y,
                            gpss_handle* out)
{
  if (world < 2 || rank < 0 || rank >= world || !id128) return fail_arg("gpss_create_partitioned: needs world >= 2, 0 <= rank < world and the NCCL id");
  return create_impl(device, n, d, X, y, world, rank, id128, out);
}

int gpss_set_theta(gpss_handle c, const double theta[GPSS_NPAR])
{
  if (!c || !theta) return fail_arg("gpss_set_theta: null argument");
  memcpy(c->theta, theta, sizeof c->theta);
  c->have_factor = c->have_alpha = c->have_U = false;   // setKUpdateStat(false) (GP_Utils.cpp:132)
  c->qstate = Q_NONE;
  return GPSS_OK;
}

// Main kernel of the Hyb{main, Bias} covariance (HybKerns, Kernel.cpp:140-169): the reference's -k choice.
int gpss_set_kernel(gpss_handle c, int kind)
{
  if (!c || kind < 0 || kind > 2) return fail_arg("gpss_set_kernel: kind must be GPSS_KERNEL_EXPANS, _EXP or _RBF");
  c->kind = kind;
  c->have_factor = c->have_alpha = c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

int gpss_get_theta(gpss_handle c, double theta[GPSS_NPAR])
{
  if (!c || !theta) return fail_arg("gpss_get_theta: null argument");
  memcpy(theta, c->theta, sizeof c->theta);
  return GPSS_OK;
}

int gpss_nlml(gpss_handle c, double* nlml)
{
  if (!c || !nlml) return fail_arg("gpss_nlml: null argument");
  CU(cudaSetDevice(c->device));
  if (c->profiling) memset(c->phase_ms, 0, sizeof c->phase_ms);
  CallTimer ct(c);
  RET(ensure_objective(c));
  *nlml = c->nlml;
  return c->chol_fail ? GPSS_NOT_POSDEF : GPSS_OK;
}

int gpss_nlml_grad(gpss_handle c, double* nlml, double g[GPSS_NPAR])
{
  if (!c || !nlml || !g) return fail_arg("gpss_nlml_grad: null argument");
  CU(cudaSetDevice(c->device));
  if (c->profiling) memset(c->phase_ms, 0, sizeof c->phase_ms);
  CallTimer ct(c);
  RET(ensure_objective(c));
  *nlml = c->nlml;
  if (c->chol_fail) {
    for (int i = 0; i < GPSS_NPAR; i++) g[i] = std::numeric_limits<double>::quiet_NaN();
    return GPSS_NOT_POSDEF;
  }
  if (c->partitioned) {
    RET(part_buffers(c));
    if (!c->have_U) {
      PhaseTimer t(c, 3);
      RET(trtri_partitioned(c));
      c->have_U = true;
    }
    {
      PhaseTimer t(c, 4);
      RET(gradient_partitioned(c));
    }
    double redp[NGRAD];
    CU(cudaMemcpyAsync(redp, c->red + 8, sizeof redp, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (c->kind == 0) combine_gradient(c->theta, redp, c->s3, g, c->d, c->n);
    else { for (int i = 0; i < GPSS_NPAR; i++) g[i] = 0.0; combine_gradient_iso(c->kind, c->theta, redp, c->s3, g); }
    return GPSS_OK;
  }
  RET(ensure_gradient_buffers(c));
  bool replayed = false;
  if (graphs_enabled(c) && !c->have_U && c->qstate != Q_IS_BINV) {
    const int nblk_o = (c->n_pad + NBO - 1) / NBO;
    RET(ensure_event_pool(c, 2 * nblk_o + 2));
    const int r = run_graph(c, 1);
    if (r < 0) return r;
    if (r == 0) {
      c->have_U = true;
      c->qstate = Q_IS_BINV;
      replayed = true;
    }
  }
  if (!replayed) RET(enqueue_gradient(c));
  double red[NGRAD];
  CU(cudaMemcpyAsync(red, c->red + 8, sizeof red, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (c->kind == 0) combine_gradient(c->theta, red, c->s3, g, c->d, c->n);
  else { for (int i = 0; i < GPSS_NPAR; i++) g[i] = 0.0; combine_gradient_iso(c->kind, c->theta, red, c->s3, g); }
  return GPSS_OK;
}

}  // extern "C"

// every buffer the inverse and the gradient pass need (allocation is not allowed while a graph is being captured)
static int ensure_gradient_buffers(gpss_ctx* c)
{
  const size_t nn = (size_t)c->n_pad * c->n_pad;
  RET(ensure_lazy(&c->Um, nn));
  RET(ensure_lazy(&c->Tpanel, (size_t)c->n_pad * NBO));
  RET(ensure_lazy(&c->Wjj, (size_t)NBO * NBO * ((c->n_pad + NBO - 1) / NBO)));
  if (c->world > 1 && !c->Tsplit) {
    const size_t cap = (size_t)24576 * NBO;                    // S * rows <= 24k rows of a 512-wide block column
    CU(cudaMalloc(&c->Tsplit, cap * sizeof(double)));
    c->Tsplit_cap = cap;
  }
  RET(ensure_lazy(&c->Qm, nn));
  const long nblocks = (long)((c->qrow1 - c->qrow0) / NB) * c->nblk;
  if (c->partial_blocks < nblocks || !c->partial) {
    if (c->partial) cudaFree(c->partial);
    c->partial = nullptr;
    CU(cudaMalloc(&c->partial, sizeof(double) * (nblocks > 0 ? nblocks : 1) * NGRAD));
    c->partial_blocks = nblocks;
  }
  return GPSS_OK;
}

// U = L^-T, B^-1 = U U^T and the fused gradient reductions into c->red[8..]: stream work only
static int enqueue_gradient(gpss_ctx* c)
{
  RET(ensure_U(c));
  if (c->qstate != Q_IS_BINV) {
    PhaseTimer t(c, 4);
    RET(lauum_lower(c));
    c->qstate = Q_IS_BINV;
  }
  const int tm0 = c->qrow0 / NB, ntm = (c->qrow1 - c->qrow0) / NB;
  const long nblocks = (long)ntm * c->nblk;
  {
    PhaseTimer t(c, 5);
    if (ntm > 0) {
      grad_pass_kernel<<<dim3(ntm, c->nblk), 256, 0, c->st>>>(c->Qm + (long)tm0 * NB, c->n_pad, c->zs, c->n_pad, c->xs, c->n_pad, c->alpha, c->n,
                                                             c->dP, c->partial, tm0, 0, 0, 0, 1);
      c->launches++;
    }
    sum_partials_kernel<NGRAD><<<1, 256, 0, c->st>>>(c->partial, nblocks, c->red + 8);
    c->launches++;
    CU(cudaGetLastError());
    if (c->world > 1) NC(g_nccl.AllReduce(c->red + 8, c->red + 8, NGRAD, ncclDouble, ncclSum, c->comm, c->st));
  }
  return GPSS_OK;
}

extern "C" {

int gpss_get_alpha(gpss_handle c, double* alpha)
{
  if (!c || !alpha) return fail_arg("gpss_get_alpha: null argument");
  CU(cudaSetDevice(c->device));
  RET(ensure_objective(c));
  CU(cudaMemcpyAsync(alpha, c->alpha, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return c->chol_fail ? GPSS_NOT_POSDEF : GPSS_OK;
}

int gpss_get_yhat(gpss_handle c, double* yhat)
{
  if (!c || !yhat) return fail_arg("gpss_get_yhat: null argument");
  CU(cudaSetDevice(c->device));
  RET(ensure_objective(c));
  CU(cudaMemcpyAsync(yhat, c->fvec, sizeof(double) * c->n, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return c->chol_fail ? GPSS_NOT_POSDEF : GPSS_OK;
}

// --- distributed evaluation: one process per GPU, NCCL over NVLink ----------------------------------------------
int gpss_nccl_unique_id(void* id128)
{
  if (!id128) return fail_arg("gpss_nccl_unique_id: null");
  RET(nccl_load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return GPSS_OK;
}

int gpss_dist_init(gpss_handle c, int rank, int world, const void* id128)
{
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail_arg("gpss_dist_init: bad argument");
  if (c->comm || c->partitioned) return fail_arg("gpss_dist_init: already initialised");
  CU(cudaSetDevice(c->device));
  if (world == 1) return GPSS_OK;
  RET(nccl_load());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC(g_nccl.CommInitRank(&c->comm, world, id, rank));
  c->rank = rank;
  c->world = world;
  std::vector<int> b;
  balanced_rows(c->n_pad, world, 0, b);
  c->urow0 = b[rank]; c->urow1 = b[rank + 1];
  balanced_rows(c->n_pad, world, 1, b);
  c->qrow0 = b[rank]; c->qrow1 = b[rank + 1];
  c->have_factor = c->have_alpha = c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

int gpss_dist_potrf_schedule(int nblk, int world, int rank, int* ops6, int cap, int* count)
{
  if (!count || nblk < 1 || world < 1 || rank < 0 || rank >= world || (cap > 0 && !ops6)) return fail_arg("gpss_dist_potrf_schedule: bad argument");
  std::vector<DistOp> ops;
  dist_potrf_schedule(nblk, world, rank, ops);
  *count = (int)ops.size();
  for (int i = 0; i < (int)ops.size() && i < cap; i++) {
    const DistOp& o = ops[i];
    const int v[6] = {o.kind, o.col, o.pbeg, o.pcnt, o.root, o.stream};
    memcpy(ops6 + 6 * i, v, sizeof v);
  }
  return GPSS_OK;
}

int gpss_dist_partition(int n_pad, int world, int kind, int* bounds)
{
  if (!bounds || world < 1 || n_pad < NB || n_pad % NB || (kind != 0 && kind != 1)) return fail_arg("gpss_dist_partition: bad argument");
  std::vector<int> b;
  balanced_rows(n_pad, world, kind, b);
  for (int k = 0; k <= world; k++) bounds[k] = b[k];
  return GPSS_OK;
}

// --- prediction ---------------------------------------------------------------------------------
static int ensure_W(gpss_ctx* c)
{
  RET(ensure_U(c));
  if (c->qstate == Q_IS_W) return GPSS_OK;
  const size_t nn = (size_t)c->n_pad * c->n_pad;
  RET(ensure_lazy(&c->Qm, nn));
  PhaseTimer t(c, 3);
  transpose_kernel<<<dim3(c->n_pad / 32, c->n_pad / 32), 256, 0, c->st>>>(c->Qm, c->n_pad, c->Um, c->n_pad, 1);
  c->launches++;
  CU(cudaGetLastError());
  c->qstate = Q_IS_W;
  return GPSS_OK;
}

// mean and RAW variance (kD - k*' A k*, no post-processing) of one shard
static int predict_core(gpss_ctx* c, long m_total, const double* sums_total, long count, const double* Xs, long ldx, double* mu, double* var)
{
  if (!c || !sums_total || (count > 0 && (!Xs || !mu))) return fail_arg("gpss_predict_shard: null argument");
  if (m_total < 1 || count < 0) return fail_arg("gpss_predict_shard: bad sizes");
  if (c->partitioned && var) { g_last_error = "gpss_predict: the predictive variance needs L^-1, which partitioned storage does not hold yet (pass var = NULL for the mean)"; return GPSS_ERR_STATE; }
  CU(cudaSetDevice(c->device));
  if (c->profiling) memset(c->phase_ms, 0, sizeof c->phase_ms);
  CallTimer ct(c);
  RET(ensure_objective(c));          // _postMean -> updateAlpha (GP_Utils.cpp:961)
  if (c->chol_fail) return GPSS_NOT_POSDEF;
  if (var) RET(ensure_W(c));
  const int n_pad = c->n_pad;
  const int cap = PRED_BATCH;
  if (!c->pred_cap) {
    RET(ensure_lazy(&c->xt, (size_t)NX * cap));
    RET(ensure_lazy(&c->zt, (size_t)NZ * cap));
    RET(ensure_lazy(&c->zsp, (size_t)NZ * n_pad));
    RET(ensure_lazy(&c->mu_part, (size_t)c->nblk * cap));
    RET(ensure_lazy(&c->dmu, cap));
    RET(ensure_lazy(&c->dvar, cap));
    c->pred_cap = cap;
  }
  if (var) {
    RET(ensure_lazy(&c->Bm, (size_t)cap * n_pad));
    RET(ensure_lazy(&c->Vm, (size_t)cap * n_pad));
  }
  // centre over the training set and ALL test points (Kernel.cpp:1391-1392)
  double centre[4];
  maha_centre(c->n, c->sums_train, m_total, sums_total, centre, c->d);
  RET(upload_params(c, 1, centre));
  transform_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->xs, n_pad, c->zsp, n_pad, c->n, n_pad, c->dP + 1);
  c->launches++;
  const double sig_k = theta_sigma(c->kind, c->theta);
  const double kD = sig_k * sig_k + theta_bias(c->kind, c->theta);   // diag_Compute (Kernel.cpp:782, 449, 594, 331, 127-136)
  for (long off = 0; off < count; off += cap) {
    const int mb = (int)((count - off < cap) ? (count - off) : cap);
    const int m_pad = ((mb + NB - 1) / NB) * NB;
    CU(cudaMemsetAsync(c->xt, 0, sizeof(double) * NX * cap, c->st));
    for (int j = 0; j < c->d; j++)
      CU(cudaMemcpyAsync(c->xt + (long)j * cap, Xs + (long)j * ldx + off, sizeof(double) * mb, cudaMemcpyHostToDevice, c->st));
    transform_kernel<<<(m_pad + 255) / 256, 256, 0, c->st>>>(c->xt, cap, c->zt, cap, mb, m_pad, c->dP + 1);
    {
      PhaseTimer t(c, 6);
      cross_build_kernel<<<dim3(m_pad / NB, c->nblk), 256, 0, c->st>>>(c->Bm, m_pad, c->zt, cap, c->zsp, n_pad, c->alpha, mb, c->n,
                                                                    c->dP + 1, c->mu_part, cap, var ? 1 : 0);
      mean_finish_kernel<<<(mb + 255) / 256, 256, 0, c->st>>>(c->mu_part, cap, c->nblk, mb, c->dmu);
      c->launches += 3;
      CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(mu + off, c->dmu, sizeof(double) * mb, cudaMemcpyDeviceToHost, c->st));
    if (var) {
      PhaseTimer t(c, 7);
      // V = L^-1 (Sw o kX): A = W (lower), B = Bm (test index contiguous)
      GemmArgs g = gemm_args(c->Qm, n_pad, c->Bm, m_pad, c->Vm, n_pad, n_pad, m_pad, n_pad);
      g.kend_row = 1; g.rev_order = 1;
      RET(gemm(c, g));
      var_finish_kernel<<<(mb + 7) / 8, 256, 0, c->st>>>(c->Vm, n_pad, n_pad, mb, kD, c->dvar);
      c->launches++;
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(var + off, c->dvar, sizeof(double) * mb, cudaMemcpyDeviceToHost, c->st));
    }
    CU(cudaStreamSynchronize(c->st));
  }
  CU(cudaStreamSynchronize(c->st));
  return GPSS_OK;
}

int gpss_predict_shard(gpss_handle c, long m_total, const double* sums_total, long count, const double* Xs, double* mu, double* var)
{
  RET(predict_core(c, m_total, sums_total, count, Xs, count, mu, var));
  return GPSS_OK;
}

// The reference's treatment of the variance vector, literally (GP_Utils.cpp:1001-1003, 1033-1040):
//     uvec ind = varSigma < 0;  varSigma.elem(ind) = zeros(...);      ind holds 0/1 FLAGS but is used as INDICES, so
// element 0 is zeroed when any entry is non-negative, element 1 when any entry is negative, and negative entries
// themselves are NOT clamped; then sn2 is added to every element unless sn2 == 1.0.  Reproduced as is (verified
// against the compiled reference: tests/golden/ref_*.npz has var[0] == sn2 for every theta).
int gpss_var_postprocess(long m, double sn2, double* var)
{
  if (!var || m < 1) return fail_arg("gpss_var_postprocess: bad argument");
  bool any_neg = false, any_nonneg = false;
  for (long i = 0; i < m; i++) { if (var[i] < 0) any_neg = true; else any_nonneg = true; }
  if (any_neg && m < 2) { g_last_error = "gpss_var_postprocess: the reference aborts here (Mat::elem(): index out of bounds)"; return GPSS_ERR_STATE; }
  if (any_nonneg) var[0] = 0.0;
  if (any_neg) var[1] = 0.0;
  if (sn2 != 1.0) for (long i = 0; i < m; i++) var[i] += sn2;
  return GPSS_OK;
}

int gpss_predict(gpss_handle c, long m, const double* Xs, double* mu, double* var)
{
  if (!c || !Xs || !mu) return fail_arg("gpss_predict: null argument");
  if (m < 1) return fail_arg("gpss_predict: m must be >= 1");
  double sums[4];
  seq_colsums(Xs, m, sums, c->d);
  if (c->world == 1) {
    RET(predict_core(c, m, sums, m, Xs, m, mu, var));
  } else {
    // Distributed handle (collective call, same Xs on every rank): the test points are split over the ranks -- L and alpha
    // are replicated -- every rank predicts rows [m r / P, m (r+1) / P) with the GLOBAL Mahalanobis centre, and the slices
    // are exchanged with one ncclBroadcast per rank and output vector, so every rank returns the full vectors.
    const long P = c->world;
    const long lo = m * c->rank / P, hi = m * (c->rank + 1) / P;
    RET(predict_core(c, m, sums, hi - lo, Xs + lo, m, mu + lo, var ? var + lo : nullptr));
    const int nvec = var ? 2 : 1;
    RET(ensure_stage(c, (size_t)nvec * m));
    if (hi > lo) {
      CU(cudaMemcpyAsync(c->stage + lo, mu + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, c->st));
      if (var) CU(cudaMemcpyAsync(c->stage + m + lo, var + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, c->st));
    }
    for (long k = 0; k < P; k++) {
      const long a = m * k / P, b = m * (k + 1) / P;
      if (b <= a) continue;
      for (int v = 0; v < nvec; v++)
        NC(g_nccl.Broadcast(c->stage + v * m + a, c->stage + v * m + a, (size_t)(b - a), ncclDouble, (int)k, c->comm, c->st));
    }
    CU(cudaMemcpyAsync(mu, c->stage, sizeof(double) * m, cudaMemcpyDeviceToHost, c->st));
    if (var) CU(cudaMemcpyAsync(var, c->stage + m, sizeof(double) * m, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
  }
  if (var) return gpss_var_postprocess(m, theta_sn2(c->kind, c->theta), var);
  return GPSS_OK;
}

// --- Kernels::computeK compatibility ----------------------------------------------------------
__global__ void __launch_bounds__(256) full_K_kernel(double* __restrict__ Km, double* __restrict__ D2m, long ld, const double* __restrict__ z1,
                                                     long ld1, const double* __restrict__ z2, long ld2, int n1, int n2,
                                                     const DevParams* __restrict__ Pp)
{
  const DevParams P = *Pp;
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int jb = blockIdx.y * 64 + (threadIdx.x >> 6) * 16;
  if (i >= n1) return;
  const double a0 = z1[i], a1 = z1[ld1 + i], a2 = z1[2 * ld1 + i], aa = z1[3 * ld1 + i], a3 = z1[4 * ld1 + i];
  for (int j = jb; j < jb + 16 && j < n2; j++) {
    const double d2 = pair_d2(a0, a1, a2, aa, z2[j], z2[ld2 + j], z2[2 * ld2 + j], z2[3 * ld2 + j], a3, z2[4 * ld2 + j]);
    if (Km) Km[(long)j * ld + i] = kern_val(d2, P);
    if (D2m) D2m[(long)j * ld + i] = d2;
  }
}

int gpss_compute_K(int device, int kind, const double theta[GPSS_NPAR], int d, int n1, const double* X1, int n2, const double* X2, double* K, double* D2)
{
  if (!theta || !X1 || !X2 || n1 < 1 || n2 < 1 || (d != 3 && d != 4) || kind < 0 || kind > 2) return fail_arg("gpss_compute_K: bad argument");
  CU(cudaSetDevice(device));
  double s1[4], s2[4], centre[4];
  seq_colsums(X1, n1, s1, d);
  seq_colsums(X2, n2, s2, d);
  maha_centre(n1, s1, n2, s2, centre, d);
  DevParams P;
  fill_params(theta, centre, P, d, kind);
  double *dx1 = nullptr, *dx2 = nullptr, *dz1 = nullptr, *dz2 = nullptr, *dK = nullptr, *dD = nullptr;
  DevParams* dP = nullptr;
  int rc = GPSS_OK;
  auto cleanup = [&]() { cudaFree(dx1); cudaFree(dx2); cudaFree(dz1); cudaFree(dz2); cudaFree(dK); cudaFree(dD); cudaFree(dP); };
#define CUK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = fail_cuda(e__, #x, __LINE__); cleanup(); return rc; } } while (0)
  CUK(cudaMalloc(&dx1, sizeof(double) * NX * n1));
  CUK(cudaMalloc(&dx2, sizeof(double) * NX * n2));
  CUK(cudaMalloc(&dz1, sizeof(double) * NZ * n1));
  CUK(cudaMalloc(&dz2, sizeof(double) * NZ * n2));
  CUK(cudaMemset(dx1, 0, sizeof(double) * NX * n1));
  CUK(cudaMemset(dx2, 0, sizeof(double) * NX * n2));
  CUK(cudaMalloc(&dP, sizeof(DevParams)));
  if (K) CUK(cudaMalloc(&dK, sizeof(double) * (size_t)n1 * n2));
  if (D2) CUK(cudaMalloc(&dD, sizeof(double) * (size_t)n1 * n2));
  CUK(cudaMemcpy(dx1, X1, sizeof(double) * d * n1, cudaMemcpyHostToDevice));
  CUK(cudaMemcpy(dx2, X2, sizeof(double) * d * n2, cudaMemcpyHostToDevice));
  CUK(cudaMemcpy(dP, &P, sizeof P, cudaMemcpyHostToDevice));
  transform_kernel<<<(n1 + 255) / 256, 256>>>(dx1, n1, dz1, n1, n1, n1, dP);
  transform_kernel<<<(n2 + 255) / 256, 256>>>(dx2, n2, dz2, n2, n2, n2, dP);
  full_K_kernel<<<dim3((n1 + 63) / 64, (n2 + 63) / 64), 256>>>(dK, dD, n1, dz1, n1, dz2, n2, n1, n2, dP);
  CUK(cudaGetLastError());
  if (K) CUK(cudaMemcpy(K, dK, sizeof(double) * (size_t)n1 * n2, cudaMemcpyDeviceToHost));
  if (D2) CUK(cudaMemcpy(D2, dD, sizeof(double) * (size_t)n1 * n2, cudaMemcpyDeviceToHost));
#undef CUK
  cleanup();
  return GPSS_OK;
}

// --- Kernels::getGradients compatibility -------------------------------------------------------------
// Kern_ExpAnisotropic::getGradients with a HOST n x n matrix QW (Kernel.cpp:886-1263), for callers of the Kernels
// interface that do not go through the device-resident GradLL.  One pass over ALL (i, j) pairs (QW need not be
// symmetric) accumulating  T = X' w X (3x3),  V1_k = sum w x_ik^2,  V2_k = sum w x_jk^2,  G6 = sum QW e^{-s}  with
// w_ij = Sigma^2 QW_ij e^{-s_ij} (-0.5 / s_ij), zero on the diagonal and where s_ij == 0 (Kernel.cpp:1176-1185).
constexpr int NGFULL = 17;
__global__ void __launch_bounds__(256) grad_full_kernel(const double* __restrict__ QW, long ldq, const double* __restrict__ z, long ldz,
                                                        const double* __restrict__ x, long ldx, int n, const DevParams* __restrict__ Pp,
                                                        double* __restrict__ partial)
{
  const DevParams P = *Pp;
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int jb = blockIdx.y * 64 + (threadIdx.x >> 6) * 16;
  double v[NGFULL];
#pragma unroll
  for (int q = 0; q < NGFULL; q++) v[q] = 0.0;
  if (i < n) {
    const double zi0 = z[i], zi1 = z[ldz + i], zi2 = z[2 * ldz + i], ai = z[3 * ldz + i], zi3 = z[4 * ldz + i];
    const double xi[3] = {x[i], x[ldx + i], x[2 * ldx + i]};
    const double xr = x[3 * ldx + i];
    for (int j = jb; j < jb + 16 && j < n; j++) {
      const double d2 = pair_d2(zi0, zi1, zi2, ai, z[j], z[ldz + j], z[2 * ldz + j], z[3 * ldz + j], zi3, z[4 * ldz + j]);
      const double s = sqrt(d2), es = exp(-s);
      const double q = QW[(long)j * ldq + i];
      v[15] += q * es;
      const double dr = xr - x[3 * ldx + j];
      v[16] = fma(es, dr * dr, v[16]);            // 4-column branch: sum_ij exp(-s) (x_i3 - x_j3)^2 (Kernel.cpp:1246-1255, no QW)
      if (i != j && s != 0.0) {
        const double w = (P.var2 * q) * (es * (-0.5 / s));
        const double xj[3] = {x[j], x[ldx + j], x[2 * ldx + j]};
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
          for (int l = 0; l < 3; l++) v[k * 3 + l] = fma(w * xi[k], xj[l], v[k * 3 + l]);
          v[9 + k] = fma(w, xi[k] * xi[k], v[9 + k]);
          v[12 + k] = fma(w, xj[k] * xj[k], v[12 + k]);
        }
      }
    }
  }
  block_reduce_store<NGFULL>(v, partial + ((long)blockIdx.y * gridDim.x + blockIdx.x) * NGFULL);
}

int gpss_expans_gradients(int device, const double theta[GPSS_NPAR], int d, int n, const double* X, const double* QW, double g8[8])
{
  if (!theta || !X || !QW || !g8 || n < 1 || (d != 3 && d != 4)) return fail_arg("gpss_expans_gradients: bad argument");
  CU(cudaSetDevice(device));
  double s1[4], centre[4];
  seq_colsums(X, n, s1, d);
  maha_centre(n, s1, n, s1, centre, d);
  DevParams P;
  fill_params(theta, centre, P, d);
  const dim3 grid((n + 63) / 64, (n + 63) / 64);
  const long nblocks = (long)grid.x * grid.y;
  double *dx = nullptr, *dz = nullptr, *dQ = nullptr, *dpart = nullptr, *dred = nullptr;
  DevParams* dP = nullptr;
  int rc = GPSS_OK;
  auto cleanup = [&]() { cudaFree(dx); cudaFree(dz); cudaFree(dQ); cudaFree(dpart); cudaFree(dred); cudaFree(dP); };
#define CUG(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = fail_cuda(e__, #x, __LINE__); cleanup(); return rc; } } while (0)
  CUG(cudaMalloc(&dx, sizeof(double) * NX * n));
  CUG(cudaMalloc(&dz, sizeof(double) * NZ * n));
  CUG(cudaMemset(dx, 0, sizeof(double) * NX * n));
  CUG(cudaMalloc(&dQ, sizeof(double) * (size_t)n * n));
  CUG(cudaMalloc(&dpart, sizeof(double) * nblocks * NGFULL));
  CUG(cudaMalloc(&dred, sizeof(double) * NGFULL));
  CUG(cudaMalloc(&dP, sizeof(DevParams)));
  CUG(cudaMemcpy(dx, X, sizeof(double) * d * n, cudaMemcpyHostToDevice));
  CUG(cudaMemcpy(dQ, QW, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
  CUG(cudaMemcpy(dP, &P, sizeof P, cudaMemcpyHostToDevice));
  transform_kernel<<<(n + 255) / 256, 256>>>(dx, n, dz, n, n, n, dP);
  grad_full_kernel<<<grid, 256>>>(dQ, n, dz, n, dx, n, n, dP, dpart);
  sum_partials_kernel<NGFULL><<<1, 256>>>(dpart, nblocks, dred);
  CUG(cudaGetLastError());
  double red[NGFULL];
  CUG(cudaMemcpy(red, dred, sizeof red, cudaMemcpyDeviceToHost));
#undef CUG
  cleanup();
  double M[6][9];
  grad_M_matrices(theta, M);
  for (int p = 0; p < 6; p++) {
    double qv = 0.0, mt = 0.0;
    for (int k = 0; k < 3; k++) {
      const double rho = (M[p][k * 3 + 0] + M[p][k * 3 + 1]) + M[p][k * 3 + 2];
      qv += rho * (red[9 + k] + red[12 + k]);
      for (int l = 0; l < 3; l++) mt += M[p][k * 3 + l] * red[k * 3 + l];
    }
    g8[p] = 2.0 * qv - 4.0 * mt;               // sum_ij w_ij (2 q_p(x_i) + 2 q_p(x_j) - 4 x_i' M_p x_j), Kernel.cpp:1192-1233
  }
  g8[6] = 2.0 * red[15] * theta[6];            // Kernel.cpp:1239-1242
  g8[7] = (d == 4) ? (-2.0 * (2.0 * red[16])) / (double)n : 0.0;   // Kernel.cpp:1246-1257 (see combine_gradient)
  return GPSS_OK;
}

// --- instrumentation ------------------------------------------------------------------------------
int gpss_set_profiling(gpss_handle c, int on)
{
  if (!c) return fail_arg("gpss_set_profiling: null");
  c->profiling = on != 0;
  return GPSS_OK;
}

int gpss_get_phase_ms(gpss_handle c, double ms[16])
{
  if (!c || !ms) return fail_arg("gpss_get_phase_ms: null");
  memcpy(ms, c->phase_ms, sizeof c->phase_ms);
  return GPSS_OK;
}

int gpss_get_last_call_ms(gpss_handle c, double* ms)
{
  if (!c || !ms) return fail_arg("gpss_get_last_call_ms: null");
  *ms = c->last_call_ms;
  return GPSS_OK;
}

// Register-only DMMA.8x8x4 loop: the FP64 tensor-pipe peak used as the roofline denominator
// (MEASURED_PEAKS.json carries no FP64 figure).  Same loop as bench_micro/fp64_peak.cu.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b)
{
  double c0[32], c1[32];
#pragma unroll
  for (int i = 0; i < 32; i++) { c0[i] = 0; c1[i] = 0; }
  const double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 32; i++) dmma884(c0[i], c1[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int gpss_measure_fp64_peak(int device, double* tflops)
{
  if (!tflops) return fail_arg("gpss_measure_fp64_peak: null");
  CU(cudaSetDevice(device));
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  const int grid = p.multiProcessorCount * 2, iters = 20000;
  double* out;
  CU(cudaMalloc(&out, sizeof(double) * grid * 256));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(e0, 0));
    dmma_peak_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-7);
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * grid * 8.0 * 256 * 32 * (double)iters / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaFree(out);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return GPSS_OK;
}

int gpss_get_launch_count(gpss_handle c, long* launches)
{
  if (!c || !launches) return fail_arg("gpss_get_launch_count: null");
  *launches = c->launches;
  return GPSS_OK;
}

int gpss_padded_n(gpss_handle c, int* n_pad)
{
  if (!c || !n_pad) return fail_arg("gpss_padded_n: null");
  *n_pad = c->n_pad;
  return GPSS_OK;
}

// which: 0 = L factor, 1 = U, 2 = Q/W buffer, 3 = zs (4 x n_pad)
int gpss_debug_fetch(gpss_handle c, int which, double* host_out, long count)
{
  if (!c || !host_out) return fail_arg("gpss_debug_fetch: null");
  CU(cudaSetDevice(c->device));
  const double* src = which == 0 ? c->Lm : which == 1 ? c->Um : which == 2 ? c->Qm : c->zs;
  if (!src) return fail_arg("gpss_debug_fetch: buffer not allocated");
  if (c->partitioned && which != 3) return fail_arg("gpss_debug_fetch: a partitioned handle holds packed block columns / rows, not n_pad x n_pad buffers");
  if (count < 0 || (size_t)count > (which == 3 ? (size_t)NZ * c->n_pad : (size_t)c->n_pad * c->n_pad)) return fail_arg("gpss_debug_fetch: count exceeds the buffer");
  CU(cudaStreamSynchronize(c->st));
  CU(cudaMemcpy(host_out, src, sizeof(double) * count, cudaMemcpyDeviceToHost));
  return GPSS_OK;
}

// --- kernel-level test hooks --------------------------------------------------------------------------
int gpss_test_gemm_nt(int device, int tile, int M, int N, int K, const double* A, const double* B, double* C, int subtract_from_C,
                      double* ms_out)
{
  if (!A || !B || !C) return fail_arg("gpss_test_gemm_nt: null");
  CU(cudaSetDevice(device));
  RET(configure_kernels());
  gpss_ctx tmp;
  tmp.st = nullptr;
  double *dA, *dB, *dC;
  CU(cudaMalloc(&dA, sizeof(double) * (size_t)M * K));
  CU(cudaMalloc(&dB, sizeof(double) * (size_t)N * K));
  CU(cudaMalloc(&dC, sizeof(double) * (size_t)M * N));
  CU(cudaMemcpy(dA, A, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dB, B, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dC, C, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
  GemmArgs g = gemm_args(dA, M, dB, N, dC, M, M, N, K);
  if (subtract_from_C) { g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  int rc;
  CU(cudaEventRecord(e0, 0));
  double* dparts = nullptr;
  if (tile == 0) rc = gemm_ws_on(&tmp, g, tmp.st);
  else if (tile == 1) rc = gemm_legacy_on(&tmp, g, tmp.st);
  else {
    // tile = S in 2..8: the split-k form used by the distributed triangular inverse (S partial products + fixed-order sum)
    if (tile > 8 || subtract_from_C) return fail_arg("gpss_test_gemm_nt: split-k hook takes 2 <= tile <= 8 and no accumulate");
    CU(cudaMalloc(&dparts, sizeof(double) * (size_t)M * N * tile));
    g.C = dparts; g.ldc = M; g.ksplit = tile; g.csplit = (long)M * N;
    rc = gemm_ws_on(&tmp, g, tmp.st);
    split_sum_kernel<<<296, 256, 0, tmp.st>>>(dC, M, dparts, M, N, tile);
  }
  CU(cudaEventRecord(e1, 0));
  if (rc < 0) return rc;
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  CU(cudaMemcpy(C, dC, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dparts);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return GPSS_OK;
}

int gpss_test_potrf(int device, int n, double* A, double* logdet_half, double* ms_out)
{
  if (!A || n < 1) return fail_arg("gpss_test_potrf: bad argument");
  CU(cudaSetDevice(device));
  RET(configure_kernels());
  gpss_ctx tmp;
  RET(create_streams(&tmp));
  const int n_pad = ((n + NB - 1) / NB) * NB;
  const int nblk = n_pad / NB;
  double *dA, *dW, *dl;
  int* dflag;
  CU(cudaMalloc(&dA, sizeof(double) * (size_t)n_pad * n_pad));
  CU(cudaMalloc(&dW, sizeof(double) * (size_t)nblk * NB * NB));
  CU(cudaMalloc(&dl, sizeof(double) * nblk));
  CU(cudaMalloc(&dflag, sizeof(int)));
  CU(cudaMemset(dflag, 0, sizeof(int)));
  // identity padding
  std::vector<double> pad((size_t)n_pad * n_pad, 0.0);
  for (int j = 0; j < n_pad; j++)
    for (int i = 0; i < n_pad; i++) pad[(size_t)j * n_pad + i] = (i < n && j < n) ? A[(size_t)j * n + i] : (i == j ? 1.0 : 0.0);
  CU(cudaMemcpy(dA, pad.data(), sizeof(double) * pad.size(), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(e0, tmp.st));
  int rc = potrf_blocked(&tmp, dA, n_pad, n_pad, dW, dl, dflag);
  CU(cudaEventRecord(e1, tmp.st));
  if (rc < 0) return rc;
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  CU(cudaMemcpy(pad.data(), dA, sizeof(double) * pad.size(), cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) A[(size_t)j * n + i] = (i >= j) ? pad[(size_t)j * n_pad + i] : 0.0;
  std::vector<double> ld(nblk);
  int flag = 0;
  CU(cudaMemcpy(ld.data(), dl, sizeof(double) * nblk, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost));
  double s = 0;
  for (double v : ld) s += v;
  if (logdet_half) *logdet_half = s;
  cudaFree(dA); cudaFree(dW); cudaFree(dl); cudaFree(dflag);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  destroy_streams(&tmp);
  return flag ? GPSS_NOT_POSDEF : GPSS_OK;
}

}  // extern "C"
