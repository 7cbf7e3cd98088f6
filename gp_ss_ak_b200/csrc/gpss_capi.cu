// C ABI (include/gpss.h) and host-side drivers of the sm_100a exact-GP path: blocked potrf, trsv, trtri,
// lauum, gradient reductions and batched prediction, each a fixed sequence of launches of the kernels in
// gpss_kernels.cuh / gpss_gemm.cuh on one stream.  No cuBLAS / cuSOLVER, no CPU fallback.
// File:line citations are into /root/reference.
#include "gpss_ctx.cuh"
#include "gpss_potrf.cuh"
#include "gpss_inverse.cuh"

__global__ void scale_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, const DevParams* __restrict__ P, int n, int n_pad)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (i < n) ? src[i] * P->inv_sn2 : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// objective pieces
// ---------------------------------------------------------------------------------------------------
// the second member's parameters in the 10-slot layout of ITS kind: [own parameters ..., Sigma_Bias = 0, sn2]
static void member2_theta(const gpss_ctx* c, double t2[GPSS_NPAR])
{
  for (int i = 0; i < GPSS_NPAR; i++) t2[i] = 0.0;
  const int np = kernel_npar(c->kind2);
  for (int i = 0; i < np - 2; i++) t2[i] = c->theta2[i];
  t2[np - 2] = 0.0;
  t2[np - 1] = theta_sn2(c->kind, c->theta);
}

static int upload_params(gpss_ctx* c, int slot, const double* centre)
{
  DevParams P, P2;
  double var2x = 0.0;
  if (c->kind2 >= 0) {
    // the second member's block: its own parameters in the 10-slot layout of its kind (bias 0), the same sn2
    double t2[GPSS_NPAR];
    member2_theta(c, t2);
    fill_params(t2, centre, P2, c->d, c->kind2, 0.0);
    var2x = P2.var2;
    CU(cudaMemcpyAsync(c->dP + 2 + slot, &P2, sizeof P2, cudaMemcpyHostToDevice, c->st));
  }
  fill_params(c->theta, centre, P, c->d, c->kind, c->white, var2x);
  CU(cudaMemcpyAsync(c->dP + slot, &P, sizeof P, cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));   // P, P2 are stack objects
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
  if (c->trtri_inflight) {                             // an inverse issued beside the previous factorisation was never collected: L is about to change
    c->trtri_inflight = false;
    CU(cudaEventRecord(c->ev_side, c->st9));
    CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
    CU(cudaEventRecord(c->ev_side, c->st8));
    CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
  }
  CU(cudaMemsetAsync(c->dflag, 0, 2 * sizeof(int), c->st));
  {
    PhaseTimer t(c, 0);
    transform_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->xs, ld, c->zs, ld, c->n, n_pad, c->dP);
    if (c->kind2 >= 0) {
      transform_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->xs, ld, c->zs2, ld, c->n, n_pad, c->dP + 2);
      kbuild_lower_kernel<true><<<dim3(c->nblk, c->nblk), 256, 0, c->st>>>(c->Lm, ld, c->zs, ld, c->n, c->dP, 0, c->world, c->rank, NBO / NB,
                                                                          c->zs2, c->dP + 2);
      c->launches++;
    } else if (c->partitioned)
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
    if (c->kind2 >= 0) kmatvec_kernel<true><<<(c->n + 63) / 64, 256, 0, c->st>>>(c->zs, n_pad, c->alpha, c->fvec, c->n, c->dP, c->zs2, c->dP + 2);
    else kmatvec_kernel<<<(c->n + 63) / 64, 256, 0, c->st>>>(c->zs, n_pad, c->alpha, c->fvec, c->n, c->dP);
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
  if (c->oz_s > 0 && c->oz_blocked) return false;               // the captured graphs hold the int8 launches
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
  if (c->solve_pending) {                              // an overlapped evaluation that ended early (error path): order the main stream after its solves
    CU(cudaStreamWaitEvent(c->st, c->ev_solved, 0));
    c->solve_pending = false;
  }
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

// gpss_nlml_grad on a theta that has no factor yet (every optimiser probe, every bench step): the vector solves for alpha are a chain of
// 2 x n/128 dependent HBM-bound step launches (24 ms at n = 50 000, the same on every rank of a multi-GPU handle) that need only L, like
// the inverse.  They are enqueued on their own stream and run BESIDE the tensor-bound inverse; the gradient pass waits for both, and the
// objective's scalars travel to the host with the gradient sums (one synchronisation per evaluation instead of two).
// Not for graph-replayed sizes, partitioned storage or while phase timers are on; GPSS_NO_SOLVE_OVERLAP=1 restores the serial order.
static bool solve_overlap_enabled(const gpss_ctx* c)
{
  if (c->have_factor || c->partitioned || c->profiling || !c->st2 || graphs_enabled(c)) return false;
  return getenv("GPSS_NO_SOLVE_OVERLAP") == nullptr;
}

static int enqueue_objective_overlapped(gpss_ctx* c)
{
  if (!c->st7) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&c->st7, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&c->ev_factored, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_solved, cudaEventDisableTiming));
  }
  RET(upload_train_params(c));
  if (c->world > 1 && oz_active(c)) {                 // multi-GPU int8 handle: the inverse is issued inside the factorisation (potrf_blocked)
    RET(ensure_gradient_buffers(c));
    c->want_trtri_interleaved = true;
  }
  const int rc_f = enqueue_factor(c);
  c->want_trtri_interleaved = false;
  RET(rc_f);
  CU(cudaEventRecord(c->ev_factored, c->st));
  CU(cudaStreamWaitEvent(c->st7, c->ev_factored, 0));
  cudaStream_t main_stream = c->st;
  c->st = c->st7;                                     // enqueue_solves launches on the handle's stream
  const int rc = enqueue_solves(c);
  c->st = main_stream;
  RET(rc);
  CU(cudaEventRecord(c->ev_solved, c->st7));
  c->solve_pending = true;
  return GPSS_OK;
}

// after the synchronisation that follows an overlapped evaluation: what ensure_objective does with the scalars
static void finish_objective(gpss_ctx* c, const double red[4], int flag)
{
  c->chol_fail = flag;
  if (flag) {
    c->nlml = std::numeric_limits<double>::quiet_NaN();
  } else {
    c->nlml = red[0] - red[1] + red[3];
    c->s3 = red[2];
  }
  c->have_alpha = true;
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
  if (oz_active(c)) RET(oz_ensure_planes(c, &c->ozU, c->oz_tmU));
  RET(ensure_lazy(&c->Tpanel, (size_t)c->n_pad * NBO));
  if (oz_active(c)) RET(ensure_lazy(&c->Tpanel2, (size_t)c->n_pad * NBO));
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
                     &c->dmu, &c->dvar, &c->zs2, &c->zsp2, &c->zt2, &c->partial2, &c->Wpan, &c->Tpanel2};
  for (auto b : bufs) if (*b) cudaFree(*b);
  if (c->dP) cudaFree(c->dP);
  if (c->dflag) cudaFree(c->dflag);
  if (c->graph[0]) cudaGraphExecDestroy(c->graph[0]);
  if (c->graph[1]) cudaGraphExecDestroy(c->graph[1]);
  if (c->stage) cudaFree(c->stage);
  if (c->Tsplit) cudaFree(c->Tsplit);
  if (c->pgather) cudaFree(c->pgather);
  if (c->ozL) cudaFree(c->ozL);
  if (c->ozU) cudaFree(c->ozU);
  if (c->ozW) cudaFree(c->ozW);
  if (c->ozB) cudaFree(c->ozB);
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
#define CUF(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail(fail_cuda(e__, #x, __FILE__, __LINE__)); } while (0)
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
  CUF(cudaMalloc(&c->red, sizeof(double) * 64));
  CUF(cudaMalloc(&c->dP, sizeof(DevParams) * 4));
  CUF(cudaMalloc(&c->dflag, 2 * sizeof(int)));
  CUF(cudaMemsetAsync(c->dflag, 0, 2 * sizeof(int), c->st));
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
    if (nr != ncclSuccess) return fail(fail_nccl(nr, "ncclCommInitRank", __FILE__, __LINE__));
  }
  // int8 tensor-core path of the three long-k contractions (gpss_ozaki.cuh).
  //   GPSS_OZAKI unset : 7 slices of 8 bits (base-256 signed digits: operands carried to 2^-55 of their a-priori bound, at or below the
  //                      rounding of the DMMA path; 28 int8 products per FP64 product) for 8192 < n_pad <= 57 344, i.e. where an
  //                      evaluation is GEMM-bound and the two extra 7 n_pad^2-byte plane buffers still fit beside L, U and B^-1 in
  //                      180 GB; the DMMA path otherwise.  Measured at n = 50 000 against the DMMA path: nlml 1e-13, g 8e-12, alpha
  //                      4e-11 (profiles/r02_ozaki_digits_n50k.log) -- closer than 8 slices of 7 bits (36 products), and 12 % faster.
  //   GPSS_OZAKI=0     : always the FP64 DMMA path
  //   GPSS_OZAKI=6|7|8 : that many slices for every n; 7-bit digits unless GPSS_OZAKI_BITS=8 (then 6 or 7 slices)
  //   GPSS_OZAKI_BITS=7 without GPSS_OZAKI: the size rule with 8 slices of 7 bits (the round-1 default)
  if (part_world <= 1) {
    int v = (c->n_pad > 8192 && c->n_pad <= 57344) ? 7 : 0;
    int bits_default = 8;
    c->oz_auto = true;
    if (const char* eb = getenv("GPSS_OZAKI_BITS")) { if (atoi(eb) == 7 && v) { v = 8; bits_default = 7; } }
    if (const char* e = getenv("GPSS_OZAKI")) {
      c->oz_auto = false;
      bits_default = 7;
      v = atoi(e);
      if (v != 0 && (v < 6 || v > 8)) return fail(fail_arg("GPSS_OZAKI must be 0 (FP64 DMMA path), 6, 7 or 8 (slices per operand)"));
      // |G_g| <= S n_pad 64^2 must stay below 2^31 (S = 8: n_pad < 65 536)
      const char* eb = getenv("GPSS_OZAKI_BITS");
      if (v != 0 && !(eb && atoi(eb) == 8) && (long long)v * c->n_pad * 4096 >= (1ll << 31)) return fail(fail_arg("GPSS_OZAKI: n too large for exact int32 accumulation with this many slices"));
    }
    c->oz_bits = bits_default;
    if (v != 0) {
      c->oz_s = v;
      c->oz_predict = true;                                   // the variance GEMM V = W (Sw o k*) on the same pipe (41.1 k vs 14.2 k points/s at
      if (const char* e = getenv("GPSS_OZAKI_PREDICT")) c->oz_predict = atoi(e) != 0;   // n = 50 000, var within 2e-12: profiles/r02_int8_predict.log)
      if (const char* e = getenv("GPSS_OZAKI_GRAD")) { const int sg = atoi(e); if (sg >= 5 && sg < v) c->oz_s_grad = sg; }
      if (const char* e = getenv("GPSS_OZAKI_BITS")) {
        if (atoi(e) != 7 && atoi(e) != 8) return fail(fail_arg("GPSS_OZAKI_BITS must be 7 or 8"));
        if (atoi(e) == 8 && v == 8) return fail(fail_arg("GPSS_OZAKI_BITS=8 takes GPSS_OZAKI=6 or 7 (7 x 8 bits already exceed FP64)"));
        c->oz_bits = atoi(e);
      }
      r = oz_configure();
      if (r == GPSS_OK) r = oz_ensure_planes(c, &c->ozL, c->oz_tmL);
      if (r != GPSS_OK) return fail(r);
      if (!c->ozL) c->oz_s = 0;                              // size rule only: the planes did not fit, stay on the DMMA path
    }
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
  // the int8 path scales U = L^-T and W = L^-1 by the a-priori bound |L^-1_ij| <= 1, which needs B = I + K / sn2 >= I: every kernel
  // of the path is positive semi-definite for any widths / angles, the bias term only for Sigma_Bias >= 0 (Kern_Bias uses it raw,
  // Kernel.cpp:362-367, and no optimiser constrains it).  Outside that region this theta is evaluated on the DMMA path.
  c->oz_blocked = (theta_bias(c->kind, theta) + (c->white < 0.0 ? c->white : 0.0) < 0.0 || !(theta_sn2(c->kind, theta) > 0.0)) && !getenv("GPSS_OZAKI_TRUST_THETA");   // (test hook: leave it to the device flag)
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

// White members of the Hyb covariance (Kern_White, Kernel.cpp:180-270): K_ii += sigma_white for the training covariance and the prior
// variance; its gradient entry is 0 (getGradParam, Kernel.cpp:266-270), so the host never asks the device for it.
int gpss_set_white(gpss_handle c, double sigma_white, int cross_diagonal)
{
  if (!c) return fail_arg("gpss_set_white: null");
  if (c->partitioned && sigma_white != 0.0) return fail_arg("gpss_set_white: not supported on partitioned handles");
  if (c->white != sigma_white) {
    c->white = sigma_white;
    c->have_factor = c->have_alpha = c->have_U = false;
    c->qstate = Q_NONE;
  }
  c->white_cross = cross_diagonal != 0;
  return GPSS_OK;
}

// A SECOND distance-based member of the additive covariance (HybKerns with two of ExpAns | Exp | RBF, gp_ss_ak.cpp:146-175;
// HybKerns::computeK / getGradients, Kernel.cpp:140-169).  kind2 < 0 removes it.
int gpss_set_kernel2(gpss_handle c, int kind2)
{
  if (!c || kind2 > 2) return fail_arg("gpss_set_kernel2: kind must be -1 (none), GPSS_KERNEL_EXPANS, _EXP or _RBF");
  if (kind2 < 0) kind2 = -1;
  if (kind2 >= 0 && c->partitioned) return fail_arg("gpss_set_kernel2: not supported on partitioned handles");
  if (kind2 == c->kind2) return GPSS_OK;
  CU(cudaSetDevice(c->device));
  if (kind2 >= 0) RET(ensure_lazy(&c->zs2, (size_t)NZ * c->n_pad));
  c->kind2 = kind2;
  // captured graphs hold the kernel instantiations and pointer arguments of the old member set
  CU(cudaStreamSynchronize(c->st));
  for (int w = 0; w < 2; w++) if (c->graph[w]) { cudaGraphExecDestroy(c->graph[w]); c->graph[w] = nullptr; }
  c->have_factor = c->have_alpha = c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

int gpss_set_theta2(gpss_handle c, const double* theta2)
{
  if (!c || !theta2) return fail_arg("gpss_set_theta2: null argument");
  memcpy(c->theta2, theta2, sizeof c->theta2);
  c->have_factor = c->have_alpha = c->have_U = false;
  c->qstate = Q_NONE;
  return GPSS_OK;
}

int gpss_get_grad2(gpss_handle c, double g2[8])
{
  if (!c || !g2) return fail_arg("gpss_get_grad2: null argument");
  memcpy(g2, c->g2, sizeof c->g2);
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
  const bool overlapped = solve_overlap_enabled(c);
  if (overlapped) {
    RET(enqueue_objective_overlapped(c));            // no host synchronisation: nlml and the failure flag are read below, with the gradient sums
  } else {
    RET(ensure_objective(c));
    *nlml = c->nlml;
    if (c->chol_fail) {
      for (int i = 0; i < GPSS_NPAR; i++) g[i] = std::numeric_limits<double>::quiet_NaN();
      return GPSS_NOT_POSDEF;
    }
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
  for (int attempt = 0;; attempt++) {
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
    double red[2 * NGRAD];
    int viol = 0;
    CU(cudaMemcpyAsync(red, c->red + 8, sizeof(double) * (c->kind2 >= 0 ? 2 : 1) * NGRAD, cudaMemcpyDeviceToHost, c->st));
    if (oz_active(c)) CU(cudaMemcpyAsync(&viol, c->dflag + 1, sizeof viol, cudaMemcpyDeviceToHost, c->st));
    double red_obj[4];
    int flag_obj = 0;
    if (overlapped && attempt == 0) {
      CU(cudaMemcpyAsync(red_obj, c->red, sizeof red_obj, cudaMemcpyDeviceToHost, c->st));
      CU(cudaMemcpyAsync(&flag_obj, c->dflag, sizeof flag_obj, cudaMemcpyDeviceToHost, c->st));
    }
    CU(cudaStreamSynchronize(c->st));
    if (overlapped && attempt == 0) {
      finish_objective(c, red_obj, flag_obj);
      *nlml = c->nlml;
      if (c->chol_fail) {                              // the inverse above ran on a failed factor: nothing of it is kept
        c->have_U = false;
        c->qstate = Q_NONE;
        for (int i = 0; i < GPSS_NPAR; i++) g[i] = std::numeric_limits<double>::quiet_NaN();
        return GPSS_NOT_POSDEF;
      }
    }
    if (viol && attempt == 0) {
      // an operand of the int8 products left its a-priori bound (oz_slice_kernel): U, B^-1 and the sums above are not trustworthy.
      // Repeat this theta on the FP64 DMMA path (oz_active() is 0 while oz_blocked is set; gpss_set_theta clears it).
      c->oz_blocked = true;
      c->oz_fallbacks++;
      c->have_factor = c->have_alpha = c->have_U = false;
      c->qstate = Q_NONE;
      RET(ensure_objective(c));
      *nlml = c->nlml;
      if (c->chol_fail) {
        for (int i = 0; i < GPSS_NPAR; i++) g[i] = std::numeric_limits<double>::quiet_NaN();
        return GPSS_NOT_POSDEF;
      }
      continue;
    }
    if (c->kind == 0) combine_gradient(c->theta, red, c->s3, g, c->d, c->n);
    else { for (int i = 0; i < GPSS_NPAR; i++) g[i] = 0.0; combine_gradient_iso(c->kind, c->theta, red, c->s3, g); }
    if (c->kind2 >= 0) {
      // the second member's entries from ITS pass (its own 10-slot layout; the bias / sn2 slots of that result are meaningless: tr QW and
      // sum Q o K were accumulated in the first pass only)
      double t2[GPSS_NPAR], gm[GPSS_NPAR];
      member2_theta(c, t2);
      for (int i = 0; i < GPSS_NPAR; i++) gm[i] = 0.0;
      if (c->kind2 == 0) combine_gradient(t2, red + NGRAD, c->s3, gm, c->d, c->n);
      else combine_gradient_iso(c->kind2, t2, red + NGRAD, c->s3, gm);
      for (int i = 0; i < 8; i++) c->g2[i] = (i < kernel_npar(c->kind2) - 2) ? gm[i] : 0.0;
    }
    return GPSS_OK;
  }
}

}  // extern "C"

// every buffer the inverse and the gradient pass need (allocation is not allowed while a graph is being captured)
static int ensure_gradient_buffers(gpss_ctx* c)
{
  const size_t nn = (size_t)c->n_pad * c->n_pad;
  RET(ensure_lazy(&c->Um, nn));
  if (oz_active(c)) RET(oz_ensure_planes(c, &c->ozU, c->oz_tmU));
  RET(ensure_lazy(&c->Tpanel, (size_t)c->n_pad * NBO));
  if (oz_active(c)) RET(ensure_lazy(&c->Tpanel2, (size_t)c->n_pad * NBO));
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
    if (c->partial2) { cudaFree(c->partial2); c->partial2 = nullptr; }
    c->partial_blocks = nblocks;
  }
  if (c->kind2 >= 0 && !c->partial2) CU(cudaMalloc(&c->partial2, sizeof(double) * (nblocks > 0 ? nblocks : 1) * NGRAD));
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
  if (c->solve_pending) {                              // alpha comes from the solve stream (enqueue_objective_overlapped)
    CU(cudaStreamWaitEvent(c->st, c->ev_solved, 0));
    c->solve_pending = false;
  }
  {
    PhaseTimer t(c, 5);
    if (ntm > 0 && c->kind2 >= 0) {
      // one pass per distance-based member (grad_pass_kernel<true>): the first also carries tr QW and sum Q o K
      grad_pass_kernel<true><<<dim3(ntm, c->nblk), 256, 0, c->st>>>(c->Qm + (long)tm0 * NB, c->n_pad, c->zs, c->n_pad, c->xs, c->n_pad, c->alpha,
                                                                   c->n, c->dP, c->partial, tm0, 0, 0, 0, 1, 1, c->zs2, c->dP + 2);
      grad_pass_kernel<true><<<dim3(ntm, c->nblk), 256, 0, c->st>>>(c->Qm + (long)tm0 * NB, c->n_pad, c->zs2, c->n_pad, c->xs, c->n_pad, c->alpha,
                                                                   c->n, c->dP + 2, c->partial2, tm0, 0, 0, 0, 1, 0, c->zs, c->dP);
      c->launches += 2;
    } else if (ntm > 0) {
      grad_pass_kernel<<<dim3(ntm, c->nblk), 256, 0, c->st>>>(c->Qm + (long)tm0 * NB, c->n_pad, c->zs, c->n_pad, c->xs, c->n_pad, c->alpha, c->n,
                                                             c->dP, c->partial, tm0, 0, 0, 0, 1);
      c->launches++;
    }
    sum_partials_kernel<NGRAD><<<1, 256, 0, c->st>>>(c->partial, nblocks, c->red + 8);
    c->launches++;
    if (c->kind2 >= 0) {
      sum_partials_kernel<NGRAD><<<1, 256, 0, c->st>>>(c->partial2, nblocks, c->red + 8 + NGRAD);
      c->launches++;
    }
    CU(cudaGetLastError());
    if (c->world > 1) NC(g_nccl.AllReduce(c->red + 8, c->red + 8, (c->kind2 >= 0 ? 2 : 1) * NGRAD, ncclDouble, ncclSum, c->comm, c->st));
    if (c->world > 1 && oz_active(c)) NC(g_nccl.AllReduce(c->dflag + 1, c->dflag + 1, 1, ncclInt, ncclMax, c->comm, c->st));   // all ranks take the same branch
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
  c->oz_dist = true;                                     // the int8 pipe stays on (measured: profiles/r02_dist_int8_2gpu.log); GPSS_OZAKI_DIST=0: DMMA
  if (const char* e = getenv("GPSS_OZAKI_DIST")) c->oz_dist = atoi(e) != 0;
  if (!c->oz_dist) {                                     // asked for the DMMA pipe on this multi-GPU handle: give the planes back
    c->oz_s = 0;
    if (c->ozL) { cudaFree(c->ozL); c->ozL = nullptr; }
    if (c->ozU) { cudaFree(c->ozU); c->ozU = nullptr; }
  }
  std::vector<int> b;
  // rows of the inverse: flop-balanced (kind 0).  The wave-aware partition (kind 2, GPSS_UROW_KIND=2) is faster when the inverse runs AFTER the
  // factorisation (8 GPUs, n = 50 000: 326 -> 310.5 ms) but not when it is issued inside it, where other work fills the partial waves
  // (300 ms with kind 0, 306 ms with kind 2: profiles/r02_dist_row_partition_8gpu.log) -- and that is the default from 6 ranks up.
  c->urow_kind = 0;
  if (const char* e = getenv("GPSS_UROW_KIND")) c->urow_kind = (atoi(e) == 2) ? 2 : 0;
  balanced_rows(c->n_pad, world, c->urow_kind, b);
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

// The packed layout of one rank's slice of U = L^-T in the exchange of allgather_U (host logic only, for the CPU tests): offsets[j - r0] =
// position of column j's first kept entry, lens[j - r0] = how many rows of it travel, *count = doubles allocated for the slice.
int gpss_dist_uslice_layout(int n_pad, int r0, int rows, long* offsets, int* lens, long* count)
{
  if (!offsets || !lens || !count || n_pad < NB || n_pad % NB || r0 < 0 || rows < 0 || r0 + rows > n_pad) return fail_arg("gpss_dist_uslice_layout: bad argument");
  for (int j = r0; j < n_pad; j++) {
    int len = (j / NBO) * NBO - r0;
    len = len < 0 ? 0 : (len > rows ? rows : len);
    lens[j - r0] = len;
    offsets[j - r0] = uslice_offset(j, r0, rows, NBO);
  }
  *count = uslice_count(n_pad, r0, rows);
  return GPSS_OK;
}

int gpss_dist_partition(int n_pad, int world, int kind, int* bounds)
{
  if (!bounds || world < 1 || n_pad < NB || n_pad % NB || (kind != 0 && kind != 1 && kind != 2)) return fail_arg("gpss_dist_partition: bad argument");
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
  c->ozW_valid = false;                                        // digit planes of W are cut on first use (predict_core)
  return GPSS_OK;
}

// mean and RAW variance (kD - k*' A k*, no post-processing) of one shard
static int predict_core(gpss_ctx* c, long m_total, const double* sums_total, long count, const double* Xs, long ldx, double* mu, double* var,
                        long shard_off = 0)
{
  if (!c || !sums_total || (count > 0 && (!Xs || !mu))) return fail_arg("gpss_predict_shard: null argument");
  if (m_total < 1 || count < 0) return fail_arg("gpss_predict_shard: bad sizes");
  if (c->partitioned && var && count != m_total) {
    g_last_error = "gpss_predict_shard: on a partitioned handle the variance is a collective over the SAME test points on every rank (use gpss_predict)";
    return GPSS_ERR_STATE;
  }
  CU(cudaSetDevice(c->device));
  if (c->profiling) memset(c->phase_ms, 0, sizeof c->phase_ms);
  CallTimer ct(c);
  RET(ensure_objective(c));          // _postMean -> updateAlpha (GP_Utils.cpp:961)
  if (c->chol_fail) return GPSS_NOT_POSDEF;
  if (var && c->partitioned) {                         // U = L^-T as cyclic block rows (variance_partitioned streams its strips)
    RET(part_buffers(c));
    if (!c->have_U) {
      PhaseTimer t(c, 3);
      RET(trtri_partitioned(c));
      c->have_U = true;
    }
  } else if (var) RET(ensure_W(c));
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
  if (c->kind2 >= 0) {
    RET(ensure_lazy(&c->zsp2, (size_t)NZ * n_pad));
    RET(ensure_lazy(&c->zt2, (size_t)NZ * cap));
    transform_kernel<<<(n_pad + 255) / 256, 256, 0, c->st>>>(c->xs, n_pad, c->zsp2, n_pad, c->n, n_pad, c->dP + 3);
    c->launches++;
  }
  const double sig_k = theta_sigma(c->kind, c->theta);
  double kD = sig_k * sig_k + theta_bias(c->kind, c->theta) + c->white;   // diag_Compute (Kernel.cpp:782, 449, 594, 331, 222-225, 127-136)
  if (c->kind2 >= 0) { double t2[GPSS_NPAR]; member2_theta(c, t2); const double s2 = theta_sigma(c->kind2, t2); kD += s2 * s2; }
  const double white_x = (c->white_cross && m_total == c->n) ? c->white : 0.0;   // Kern_White::computeK on (X_train, X_test), Kernel.cpp:257-264
  for (long off = 0; off < count; off += cap) {
    const int mb = (int)((count - off < cap) ? (count - off) : cap);
    const int m_pad = ((mb + NB - 1) / NB) * NB;
    CU(cudaMemsetAsync(c->xt, 0, sizeof(double) * NX * cap, c->st));
    for (int j = 0; j < c->d; j++)
      CU(cudaMemcpyAsync(c->xt + (long)j * cap, Xs + (long)j * ldx + off, sizeof(double) * mb, cudaMemcpyHostToDevice, c->st));
    transform_kernel<<<(m_pad + 255) / 256, 256, 0, c->st>>>(c->xt, cap, c->zt, cap, mb, m_pad, c->dP + 1);
    if (c->kind2 >= 0) transform_kernel<<<(m_pad + 255) / 256, 256, 0, c->st>>>(c->xt, cap, c->zt2, cap, mb, m_pad, c->dP + 3);
    {
      PhaseTimer t(c, 6);
      if (c->kind2 >= 0)
        cross_build_kernel<true><<<dim3(m_pad / NB, c->nblk), 256, 0, c->st>>>(c->Bm, m_pad, c->zt, cap, c->zsp, n_pad, c->alpha, mb, c->n,
                                                                            c->dP + 1, c->mu_part, cap, var ? 1 : 0, white_x, shard_off + off,
                                                                            c->zt2, c->zsp2, c->dP + 3);
      else
      cross_build_kernel<<<dim3(m_pad / NB, c->nblk), 256, 0, c->st>>>(c->Bm, m_pad, c->zt, cap, c->zsp, n_pad, c->alpha, mb, c->n,
                                                                    c->dP + 1, c->mu_part, cap, var ? 1 : 0, white_x, shard_off + off);
      mean_finish_kernel<<<(mb + 255) / 256, 256, 0, c->st>>>(c->mu_part, cap, c->nblk, mb, c->dmu);
      c->launches += 3;
      CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(mu + off, c->dmu, sizeof(double) * mb, cudaMemcpyDeviceToHost, c->st));
    if (var && c->partitioned) {
      PhaseTimer t(c, 7);
      RET(variance_partitioned(c, mb, m_pad));
      CU(cudaMemcpyAsync(var + off, c->dvar, sizeof(double) * mb, cudaMemcpyDeviceToHost, c->st));
      CU(cudaStreamSynchronize(c->st));
      for (int j = 0; j < mb; j++) var[off + j] = kD - var[off + j];      // raw variance, as var_finish_kernel
    } else if (var) {
      PhaseTimer t(c, 7);
      // V = L^-1 (Sw o kX): A = W (lower), B = Bm (test index contiguous)
      if (oz_active(c) && c->oz_predict && !c->ozW) {
        // the planes of W = L^-1 and of one batch: if they do not fit beside L, U, W and the planes of L and U, predict on the DMMA pipe
        const size_t wb = (size_t)c->oz_s * n_pad * n_pad, bb = (size_t)c->oz_s * cap * n_pad;
        if (cudaMalloc(&c->ozW, wb) != cudaSuccess) { cudaGetLastError(); c->ozW = nullptr; c->oz_predict = false; }
        else if (cudaMalloc(&c->ozB, bb) != cudaSuccess) { cudaGetLastError(); cudaFree(c->ozW); c->ozW = nullptr; c->ozB = nullptr; c->oz_predict = false; }
        else c->oz_w_fresh = true;
      }
      if (oz_active(c) && c->oz_predict) {
        // the same product on the int8 tensor cores -- planes of W once per factor, planes of the batch per batch
        if (c->oz_w_fresh) {
          const size_t bb = (size_t)c->oz_s * cap * n_pad;
          c->oz_w_fresh = false;
          CU(cudaMemsetAsync(c->ozB, 0, bb, c->st));
          if (oz::make_plane_map(&c->oz_tmW[0], c->ozW, (long)c->oz_s * n_pad, n_pad, oz::BM) != 0 ||
              oz::make_plane_map(&c->oz_tmB[1], c->ozB, (long)c->oz_s * cap, n_pad, oz::BN) != 0)
            return fail_arg("GPSS_OZAKI_PREDICT: cuTensorMapEncodeTiled failed");
          c->ozW_valid = false;
        }
        if (!c->ozW_valid) {
          RET(oz_slice_on(c, c->Qm, n_pad, 0, n_pad, 0, n_pad, oz::SCALE_UNIT, oz::MASK_LOWER, c->ozW, c->st));
          c->ozW_valid = true;
        }
        // batch planes: rows = test index (plane_rows = cap), k = training index; Bm(j, k) at Bm[j + k * m_pad]
        switch (c->oz_s) {
          case 6: oz::slice<6>(c->Bm, m_pad, 0, m_pad, 0, n_pad, oz::SCALE_CROSS, oz::MASK_NONE, c->dP + 1, c->ozB, cap, n_pad, c->st, c->oz_bits); break;
          case 7: oz::slice<7>(c->Bm, m_pad, 0, m_pad, 0, n_pad, oz::SCALE_CROSS, oz::MASK_NONE, c->dP + 1, c->ozB, cap, n_pad, c->st, c->oz_bits); break;
          default: oz::slice<8>(c->Bm, m_pad, 0, m_pad, 0, n_pad, oz::SCALE_CROSS, oz::MASK_NONE, c->dP + 1, c->ozB, cap, n_pad, c->st, c->oz_bits); break;
        }
        c->launches++;
        CU(cudaGetLastError());
        oz::Args a;
        memset(&a, 0, sizeof a);
        a.C = c->Vm; a.ldc = n_pad; a.m = n_pad; a.n = m_pad;
        a.a_rows = n_pad; a.b_rows = cap; a.k0 = 0; a.k1 = n_pad; a.kend_row = 1; a.rev_order = 1;
        a.accumulate = 0; a.sign = 1.0; a.a_kind = oz::SCALE_UNIT; a.b_kind = oz::SCALE_CROSS;
        a.dP = c->dP + 1;
        RET(oz_gemm_on(c, c->oz_tmW[0], c->oz_tmB[1], a, c->st));
      } else {
        GemmArgs g = gemm_args(c->Qm, n_pad, c->Bm, m_pad, c->Vm, n_pad, n_pad, m_pad, n_pad);
        g.kend_row = 1; g.rev_order = 1;
        RET(gemm(c, g));
      }
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
  if (c->world == 1 || (c->partitioned && var)) {
    // (partitioned storage with the variance: the factor is spread over the ranks, so every rank works on ALL test points -- the call is
    //  a collective inside predict_core and returns the same vectors everywhere)
    RET(predict_core(c, m, sums, m, Xs, m, mu, var));
  } else {
    // Distributed handle (collective call, same Xs on every rank): the test points are split over the ranks -- L and alpha
    // are replicated -- every rank predicts rows [m r / P, m (r+1) / P) with the GLOBAL Mahalanobis centre, and the slices
    // are exchanged with one ncclBroadcast per rank and output vector, so every rank returns the full vectors.
    const long P = c->world;
    const long lo = m * c->rank / P, hi = m * (c->rank + 1) / P;
    RET(predict_core(c, m, sums, hi - lo, Xs + lo, m, mu + lo, var ? var + lo : nullptr, lo));
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
#define CUK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = fail_cuda(e__, #x, __FILE__, __LINE__); cleanup(); return rc; } } while (0)
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
#define CUG(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = fail_cuda(e__, #x, __FILE__, __LINE__); cleanup(); return rc; } } while (0)
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

int gpss_measure_int8_peak(int device, double* tops_burst, double* tops_sustained)
{
  if (!tops_burst || !tops_sustained) return fail_arg("gpss_measure_int8_peak: null");
  CU(cudaSetDevice(device));
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  const int sms = p.multiProcessorCount, smem = 24 * 1024 + 1024 + 64, iters = 20000;
  CU(cudaFuncSetAttribute(oz::int8_peak_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long* cyc;
  CU(cudaMalloc(&cyc, 8));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  // one launch = iters x 2 accumulators x 2 k-steps instructions of 128 x 256 x 32 MACs per SM
  const double ops = 2.0 * 128 * 256 * 32 * ((double)iters * 2 * 2) * sms;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {                         // burst: best single launch (~6 ms) after a warm-up launch
    CU(cudaEventRecord(e0, 0));
    oz::int8_peak_kernel<256><<<sms, 128, smem>>>(iters, cyc);
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double t = ops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && t > best) best = t;
  }
  const int reps = 160;                                       // sustained: ~1 s back to back, the power cap has pulled the clocks down
  CU(cudaEventRecord(e0, 0));
  for (int r = 0; r < reps; r++) oz::int8_peak_kernel<256><<<sms, 128, smem>>>(iters, cyc);
  CU(cudaEventRecord(e1, 0));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  CU(cudaGetLastError());
  cudaFree(cyc);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tops_burst = best;
  *tops_sustained = ops * reps / (ms * 1e-3) * 1e-12;
  return GPSS_OK;
}

int gpss_get_launch_count(gpss_handle c, long* launches)
{
  if (!c || !launches) return fail_arg("gpss_get_launch_count: null");
  *launches = c->launches;
  return GPSS_OK;
}

int gpss_get_ozaki(gpss_handle c, int* slices)
{
  if (!c || !slices) return fail_arg("gpss_get_ozaki: null");
  *slices = oz_active(c);
  return GPSS_OK;
}

int gpss_get_ozaki_bits(gpss_handle c, int* bits)
{
  if (!c || !bits) return fail_arg("gpss_get_ozaki_bits: null");
  *bits = c->oz_bits;
  return GPSS_OK;
}

int gpss_get_ozaki_fallbacks(gpss_handle c, long* count)
{
  if (!c || !count) return fail_arg("gpss_get_ozaki_fallbacks: null");
  *count = c->oz_fallbacks;
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
  else if (tile == 9) { tmp.dmma_coresident = true; rc = gemm_ws_on(&tmp, g, tmp.st); tmp.dmma_coresident = false; }   // the 2-stage ring (51 KB)
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

// C <- C - A B^T (subtract_from_C) or C <- A B^T through the int8 tensor-core kernel with `slices` 7-bit digits per operand; host
// buffers, column-major, |A|, |B| <= 1 (unit scale), M % 128 == 0, N % 64 == 0, K % 64 == 0.  The result is reproduced to the last
// bit by the numpy restatement the tests hold (exact integer accumulation, fixed-order FP64 recombination).
int gpss_test_oz_gemm(int device, int slices, int M, int N, int K, const double* A, const double* B, double* C, int subtract_from_C,
                      double* ms_out)
{
  if (!A || !B || !C) return fail_arg("gpss_test_oz_gemm: null");
  if (slices < 6 || slices > 8) return fail_arg("gpss_test_oz_gemm: slices must be 6, 7 or 8");
  if (M <= 0 || N <= 0 || K <= 0 || M % oz::BM || N % oz::BN || K % oz::BK) return fail_arg("gpss_test_oz_gemm: M % 128, N % 64, K % 64 must be 0");
  CU(cudaSetDevice(device));
  RET(oz_configure());
  gpss_ctx tmp;
  tmp.st = nullptr;
  tmp.oz_s = slices;
  if (const char* e = getenv("GPSS_OZAKI_BITS")) tmp.oz_bits = atoi(e) == 8 ? 8 : 7;
  tmp.n_pad = K;                                  // plane geometry of the helpers below: kpad = plane_rows = n_pad
  double *dA, *dB, *dC;
  int8_t *pa, *pb;
  const long R = (M > N ? M : N);
  if (R > K) return fail_arg("gpss_test_oz_gemm: M, N <= K (the planes are K x K)");
  CU(cudaMalloc(&dA, sizeof(double) * (size_t)M * K));
  CU(cudaMalloc(&dB, sizeof(double) * (size_t)N * K));
  CU(cudaMalloc(&dC, sizeof(double) * (size_t)M * N));
  CU(cudaMalloc(&pa, (size_t)slices * K * K));
  CU(cudaMalloc(&pb, (size_t)slices * K * K));
  CU(cudaMemcpy(dA, A, sizeof(double) * (size_t)M * K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dB, B, sizeof(double) * (size_t)N * K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dC, C, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  if (oz::make_plane_map(&ta, pa, (long)slices * K, K, oz::BM) != 0 || oz::make_plane_map(&tb, pb, (long)slices * K, K, oz::BN) != 0)
    return fail_arg("gpss_test_oz_gemm: cuTensorMapEncodeTiled failed");
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  RET(oz_slice_on(&tmp, dA, M, 0, M, 0, K, oz::SCALE_UNIT, oz::MASK_NONE, pa, tmp.st));
  RET(oz_slice_on(&tmp, dB, N, 0, N, 0, K, oz::SCALE_UNIT, oz::MASK_NONE, pb, tmp.st));
  oz::Args a;
  memset(&a, 0, sizeof a);
  a.C = dC; a.ldc = M; a.m = M; a.n = N; a.k0 = 0; a.k1 = K;
  a.accumulate = subtract_from_C ? 1 : 0; a.sign = subtract_from_C ? -1.0 : 1.0;
  CU(cudaEventRecord(e0, 0));
  RET(oz_gemm_on(&tmp, ta, tb, a, tmp.st));
  CU(cudaEventRecord(e1, 0));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  CU(cudaMemcpy(C, dC, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(pa); cudaFree(pb);
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

