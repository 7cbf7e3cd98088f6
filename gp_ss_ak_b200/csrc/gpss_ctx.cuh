// gpss_ctx.cuh -- part of the single translation unit gpss_capi.cu (not a standalone header): error plumbing, the run-time
// NCCL binding, the handle (gpss_ctx), GEMM launch helpers, timers and the theta -> device-parameter mapping.
#pragma once
#include "../../include/gpss.h"
#include "gpss_gemm.cuh"
#include "gpss_kernels.cuh"
#include "gpss_params.h"
#include "gpss_ozaki.cuh"

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

static const char* base_name(const char* path) { const char* b = std::strrchr(path, '/'); return b ? b + 1 : path; }
static int fail_cuda(cudaError_t e, const char* what, const char* file, int line)
{
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e), what, base_name(file), line);
  g_last_error = buf;
  return GPSS_ERR_CUDA;
}
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail_cuda(e__, #x, __FILE__, __LINE__); } while (0)
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
static int fail_nccl(ncclResult_t r, const char* what, const char* file, int line)
{
  char buf[512];
  snprintf(buf, sizeof buf, "NCCL error '%s' in %s (%s:%d)", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what, base_name(file), line);
  g_last_error = buf;
  return GPSS_ERR_NCCL;
}
#define NC(x) do { ncclResult_t r__ = (x); if (r__ != ncclSuccess) return fail_nccl(r__, #x, __FILE__, __LINE__); } while (0)

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
  cudaStream_t st5 = nullptr;                                 // distributed Cholesky: the full-height part of U2 runs here (highest priority) WHILE the main
                                                              // stream factors the 512 x 512 diagonal block of the same block column
  cudaStream_t st6 = nullptr;                                 // distributed Cholesky: digit planes of a received panel (they gate bulk updates only)
  cudaEvent_t ev_u2 = nullptr, ev_unpacked = nullptr;         // st5 / st6 dependencies of the above
  cudaStream_t st8 = nullptr, st9 = nullptr;                  // inverse issued DURING the distributed Cholesky: diagonal blocks (highest priority) / bulk (lowest)
  bool want_trtri_interleaved = false;                        // set by the overlapped gpss_nlml_grad before the factorisation is enqueued
  bool trtri_inflight = false;                                // potrf_blocked issued every step of the inverse: trtri_upper only joins the streams
  cudaStream_t st7 = nullptr;                                 // gpss_nlml_grad on a fresh theta: the vector solves for alpha run here beside the inverse
  cudaEvent_t ev_factored = nullptr, ev_solved = nullptr;
  bool solve_pending = false;                                 // the solves were enqueued on st7: the gradient pass must wait for ev_solved, and the
                                                              // objective's scalars are read together with the gradient sums
  double* Wpan = nullptr;                                     // 2 x NBO x NBO: inverse of the current diagonal block (transposed scratch | lower, column-major)
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
  double *Tpanel2 = nullptr;                                   // second product buffer of the two-stream inverse (int8 path)
  double *partial = nullptr; long partial_blocks = 0;          // gradient partial sums
  double *red = nullptr;                                       // 32 doubles of reduced scalars
  DevParams* dP = nullptr;                                     // [0] training, [1] prediction
  int* dflag = nullptr;                                        // [0] Cholesky failure, [1] an int8 digit-plane operand exceeded its a-priori bound
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
  int urow_kind = 0;                                           // which partition of balanced_rows they come from (0: flops, 2: int8 wave model)
  int qrow0 = 0, qrow1 = 0;                                    // my rows of B^-1
  // int8 tensor-core path (the default above n_pad = 8192; GPSS_OZAKI, gpss_ozaki.cuh): signed digit planes of L and of U = L^-T,
  // [oz_s][n_pad rows][n_pad bytes of k] each, and their TMA descriptors ([0] 128-row box = A operand, [1] 64-row box = B operand)
  int oz_s = 0;
  int oz_s_grad = 0;                                           // GPSS_OZAKI_GRAD=6 (measured, NOT adopted: -10 % time, gradient 1e-8): fewer slices for the inverse / B^-1
                                                               // products (they feed the gradient only), read from the TOP planes of the same tensors
  int oz_bits = 7;                                             // digit width: 8 with the size rule's 7 slices (default), 7 with a forced GPSS_OZAKI unless GPSS_OZAKI_BITS=8
  bool oz_blocked = false;                                     // this theta stays on the DMMA path: Sigma_Bias < 0 or sn2 <= 0 (K not PSD, so |L^-1| <= 1 is not
                                                               // guaranteed), or dflag[1] was raised by the previous attempt at this theta
  bool dmma_coresident = false;                                // set while an int8 bulk phase is in flight: main-stream DMMA GEMMs use the 2-stage ring (gemm_ws_on)
  long oz_fallbacks = 0;                                       // evaluations repeated on the DMMA path because dflag[1] was raised
  bool oz_auto = false;                                        // chosen by the size rule, not by GPSS_OZAKI: falls back to DMMA if the planes do not fit
  int8_t *ozL = nullptr, *ozU = nullptr;
  bool ozL_valid = false, ozU_valid = false;                   // the planes hold the CURRENT factor / inverse (set by the drivers that cut them)
  CUtensorMap oz_tmL[2], oz_tmU[2];
  // the prediction GEMM on the int8 kernel too (default; GPSS_OZAKI_PREDICT=0: DMMA): V = W (Sw o k*)^T on the same kernel -- planes of W = L^-1
  // (cut once per factor) and of the cross-covariance batch (PRED_BATCH rows, cut per batch)
  // the int8 path on replicated-layout multi-GPU handles (gpss_dist_init; default, GPSS_OZAKI_DIST=0: DMMA)
  bool oz_dist = false;
  bool oz_predict = false, ozW_valid = false, oz_w_fresh = false;
  int8_t *ozW = nullptr, *ozB = nullptr;
  CUtensorMap oz_tmW[2], oz_tmB[2];
  // a SECOND distance-based member of the Hyb sum (gpss_set_kernel2 / gpss_set_theta2; kind2 < 0: none).  Its transformed coordinates
  // and its parameter blocks (dP[2] training, dP[3] prediction) sit beside the first member's; the kernels that evaluate K take both.
  int kind2 = -1;
  double theta2[8] = {0, 0, 0, 0, 0, 0, 0, 0};                 // the member's own parameters, in its own order (Kernel.h)
  double g2[8] = {0, 0, 0, 0, 0, 0, 0, 0};                     // its gradient entries from the last gpss_nlml_grad
  double *zs2 = nullptr, *zsp2 = nullptr, *zt2 = nullptr, *partial2 = nullptr;
  // host state
  double white = 0.0;                                          // sum of the White members' Sigma_White (gpss_set_white); 0 without one
  int white_cross = 0;                                         // the next gpss_predict adds it on the cross-covariance diagonal (Kernel.cpp:261-262)
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

// streams and events of one pass of the triangular inverse (gpss_inverse.cuh: trtri_step)
struct TrtriRun {
  cudaStream_t sm, ss;
  std::vector<cudaEvent_t>* evs;     // [2 t]: diagonal block t ready (sm -> bulk stream); [2 t + 1]: digit planes of block column t cut
  bool ozk;
  cudaStream_t ss2;                  // second bulk stream (nullptr: none): consecutive block columns alternate between ss and ss2
  double* T2;                        // second n_pad x NBO product buffer for the steps on ss2
};

// ---------------------------------------------------------------------------------------------------
static int configure_kernels()
{
  CU(cudaFuncSetAttribute(gemm_nt_ws_kernel<GemmTileWideWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTileWideWS::SMEM_BYTES));
  CU(cudaFuncSetAttribute(gemm_nt_ws_kernel<GemmTileWideWS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTileWideWS2::SMEM_BYTES));
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
  // While oz_gemm_kernel CTAs hold the SMs (one per SM, 169 KB of shared memory with 7 digit planes) the panel work of the main
  // stream -- U2, the panel solves, the rank-128 updates, the diagonal blocks of the inverse -- would wait for whole SMs to drain.
  // The same kernel with a 2-stage ring (51 KB, same registers, bitwise the same sums) fits NEXT TO a resident int8 CTA, on the
  // FP64 pipe that CTA leaves idle.  (8 planes of 7 bits take 193 KB: no room, the hardware then simply queues these CTAs.)
  if (c->dmma_coresident && (stream == c->st || (c->st5 && stream == c->st5) || (c->st8 && stream == c->st8)) && parts == 1) {
    using T2 = GemmTileWideWS2;
    gemm_nt_ws_kernel<T2><<<(unsigned)(ga.mt * ga.nt), T2::THREADS, T2::SMEM_BYTES, stream>>>(ga);
    c->launches++;
    CU(cudaGetLastError());
    return GPSS_OK;
  }
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

// ---------------------------------------------------------------------------------------------------
// int8 (Ozaki) path: active on replicated-storage handles with the look-ahead streams -- single-GPU, or multi-GPU with GPSS_OZAKI_DIST=1
// ---------------------------------------------------------------------------------------------------
static int oz_active(const gpss_ctx* c)
{
  return (c->oz_s > 0 && !c->oz_blocked && !c->partitioned && c->st2 && (c->world == 1 || c->oz_dist)) ? c->oz_s : 0;
}

static int oz_configure()
{
  CU(oz::configure<5>());
  CU(oz::configure<6>());
  CU(oz::configure<7>());
  CU(oz::configure<8>());
  return GPSS_OK;
}

// digit planes of one n_pad x n_pad operand + its two TMA descriptors (allocation: never while a graph is being captured)
static int oz_ensure_planes(gpss_ctx* c, int8_t** planes, CUtensorMap* tm)
{
  if (*planes) return GPSS_OK;
  const size_t bytes = (size_t)c->oz_s * c->n_pad * c->n_pad;
  if (c->oz_auto) {                                            // not asked for explicitly: no room means the DMMA path, not an error
    if (cudaMalloc(planes, bytes) != cudaSuccess) { cudaGetLastError(); *planes = nullptr; return GPSS_OK; }
  } else {
    CU(cudaMalloc(planes, bytes));
  }
  CU(cudaMemsetAsync(*planes, 0, bytes, c->st));
  if (oz::make_plane_map(&tm[0], *planes, (long)c->oz_s * c->n_pad, c->n_pad, oz::BM) != 0 ||
      oz::make_plane_map(&tm[1], *planes, (long)c->oz_s * c->n_pad, c->n_pad, oz::BN) != 0)
    return fail_arg("GPSS_OZAKI: cuTensorMapEncodeTiled failed");
  return GPSS_OK;
}

// slices = 0: the handle's slice count; otherwise the product reads only the top `slices` planes of the (oz_s-plane) tensors -- the
// leading digits of a signed-digit expansion are the expansion of the value rounded to fewer digits
static int oz_gemm_on(gpss_ctx* c, const CUtensorMap& ta, const CUtensorMap& tb, oz::Args a, cudaStream_t st, int slices = 0)
{
  const int s_use = (slices > 0 && slices < c->oz_s) ? slices : c->oz_s;
  if (a.m <= 0 || a.n <= 0) return GPSS_OK;
  if (a.m % oz::BM || a.n % oz::BN || a.k0 % oz::BK || a.k1 % oz::BK) return fail_arg("oz_gemm: dimensions not tile multiples");
  if (!a.a_rows) a.a_rows = c->n_pad;
  if (!a.b_rows) a.b_rows = c->n_pad;
  if (!a.dP) a.dP = c->dP;
  a.digit_bits = c->oz_bits;
  // 8-bit digits: an int32 accumulation holds at most kseg bytes of k exactly; the kernel drains its accumulators into C at every absolute
  // multiple of kseg and restarts them (Args::kseg), all in ONE launch.  (Until round 2 this was one launch per segment: every tile paid
  // its prologue and epilogue up to three times and each launch had its own triangle of short-k tiles -- 13 % of B^-1 = U U^T at n = 50 000.)
  const int seg = oz::kseg(s_use, c->oz_bits);
  a.kseg = seg >= (1 << 30) ? 0 : seg;
  switch (s_use) {
    case 5: oz::launch<5>(ta, tb, a, st); break;
    case 6: oz::launch<6>(ta, tb, a, st); break;
    case 7: oz::launch<7>(ta, tb, a, st); break;
    case 8: oz::launch<8>(ta, tb, a, st); break;
    default: return fail_arg("GPSS_OZAKI must be 6, 7 or 8");
  }
  c->launches++;
  CU(cudaGetLastError());
  return GPSS_OK;
}

// X addressed by global (row, k) -> planes[p][row][k] for the block [row0, row0 + rows) x [k0, k0 + kcnt)
static int oz_slice_on(gpss_ctx* c, const double* X, long ldx, int row0, int rows, int k0, int kcnt, int kind, int mask, int8_t* planes,
                       cudaStream_t st)
{
  if (rows <= 0 || kcnt <= 0) return GPSS_OK;
  switch (c->oz_s) {
    case 6: oz::slice<6>(X, ldx, row0, rows, k0, kcnt, kind, mask, c->dP, planes, c->n_pad, c->n_pad, st, c->oz_bits, c->dflag ? c->dflag + 1 : nullptr); break;
    case 7: oz::slice<7>(X, ldx, row0, rows, k0, kcnt, kind, mask, c->dP, planes, c->n_pad, c->n_pad, st, c->oz_bits, c->dflag ? c->dflag + 1 : nullptr); break;
    case 8: oz::slice<8>(X, ldx, row0, rows, k0, kcnt, kind, mask, c->dP, planes, c->n_pad, c->n_pad, st, c->oz_bits, c->dflag ? c->dflag + 1 : nullptr); break;
    default: return fail_arg("GPSS_OZAKI must be 6, 7 or 8");
  }
  c->launches++;
  CU(cudaGetLastError());
  return GPSS_OK;
}
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
static void fill_params(const double theta[GPSS_NPAR], const double* centre, DevParams& P, int dim = 3, int kind = 0, double white = 0.0,
                        double var2x = 0.0)
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
  P.white = white;
  P.var2x = var2x;
  P.sn2 = sn2;
  P.inv_sn2 = 1 / sn2;
  P.sw = std::sqrt(P.inv_sn2);
  P.sww = P.sw * P.sw;
  P.lp_const = std::log(2.0 * M_PI * sn2) / 2;
}

