/* gpss.h -- C ABI of the B200-native exact-GP hot path of GP_SS_AK.
 *
 * This is the ONLY boundary between host code (the C++ classes that mirror the reference's
 * Kernels / Opt_Algs / GP_utils surface, the ctypes test harness, bench.py) and the CUDA
 * implementation in gp_ss_ak_b200/csrc.  Plain pointers and sizes, no C++/torch types.
 *
 * The reference has no FFI of its own (single C++ process, Armadillo); each entry point below
 * replaces the body of one reference member function -- the file:line it stands in for is cited.
 * INTEGRATION.md shows the few lines a maintainer of the reference adds to route GP_utils through
 * this library.
 *
 * Conventions
 *   - matrices are column-major like arma::mat: X(i,d) at X[i + d*n]
 *   - theta[10] = {AngleX, iWx, AngleY, iWy, AngleZ, iWz, Sigma, iWR, Sigma_Bias, sn2}: the order
 *     GP_utils::get_GP_Pars produces (GP_Utils.cpp:101-128, Kernel.cpp:803-838, 350-360)
 *   - every function returns an int status: GPSS_OK, GPSS_NOT_POSDEF (the reference's Chol_fail ->
 *     quiet-NaN path, GP_Utils.cpp:881-888,1145-1146) or a negative error; gpss_last_error() gives text
 *   - a handle is bound to one CUDA device and is not thread-safe (the reference is single-threaded)
 *   - there is NO CPU fallback: if no CUDA device / kernel image is available every call fails loudly
 */
#ifndef GPSS_H
#define GPSS_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPSS_OK 0
#define GPSS_NOT_POSDEF 1
#define GPSS_ERR_CUDA (-1)
#define GPSS_ERR_ARG (-2)
#define GPSS_ERR_NCCL (-3)
#define GPSS_ERR_STATE (-4)

#define GPSS_NPAR 10

typedef struct gpss_ctx* gpss_handle;

/* library / device ------------------------------------------------------------------------------ */
int gpss_version(void);
int gpss_device_count(int* count);
const char* gpss_last_error(void);

/* model state ------------------------------------------------------------------------------------ */
/* Replaces GP_utils::GP_utils + initialize_vars (GP_Utils.cpp:9-86): takes the standardised training
 * inputs X (n x d) and targets y (n) and makes them device resident.  d == 3, or d == 4 for the reference's rock-type
 * branch of the ExpAns kernel: the 4th column is scaled by theta[7] = InversewidthR and g[7] is no longer zero
 * (Kernel.cpp:872-878, 1411-1424, 1246-1255). */
int gpss_create(int device, int n, int d, const double* X_colmajor, const double* y, gpss_handle* out);
int gpss_destroy(gpss_handle h);
/* Re-upload training data of the same shape (test() assigns Xinp / yTarg, gp_ss_ak.cpp:389-395). */
int gpss_set_data(gpss_handle h, const double* X_colmajor, const double* y);
/* GP_utils::set_GP_Pars (GP_Utils.cpp:130-157): stores theta and invalidates the caches; no GPU work. */
int gpss_set_theta(gpss_handle h, const double theta[GPSS_NPAR]);
/* Kern_White members of the additive covariance (reference Kernel.cpp:180-270): sigma_white (the sum over the White members) is
 * added to K_ii of the training covariance (computeK fills the diagonal when both arguments are the training set, :257-264) and to
 * the prior variance of a test point (Diag_Kernel, :222-225).  cross_diagonal != 0: the reference's condition `X1(0) == X2(0) and
 * equal row counts` held for (X_train, X_test) of the NEXT gpss_predict, so element (i, i) of that cross-covariance carries it too
 * (gpss_predict only; gpss_predict_shard callers own the offset bookkeeping and do not get this quirk).  The gradient entry of a
 * White member is 0 (getGradParam, :266-270).  Invalidates the factorisation when the value changes. */
int gpss_set_white(gpss_handle h, double sigma_white, int cross_diagonal);
/* A SECOND distance-based member of the additive covariance: `gp_ss_ak train -k ExpAns -k RBF ...` builds Hyb{ExpAns, RBF[, Bias]}
 * (gp_ss_ak.cpp:146-175), HybKerns::computeK adds the members' K AND their D2 (Kernel.cpp:140-154), HybKerns::getGradients hands
 * every member the SUMMED D2 (:156-169) -- which Exp / RBF use as their distance (Kernel.cpp:646-695, 491-541) while ExpAns
 * recomputes its own (:925).  All of that is reproduced.  kind2 = -1 (none) | GPSS_KERNEL_EXPANS | _EXP | _RBF; theta2 = the member's
 * own parameters in its own order (8 slots read: ExpAns 8, Exp 2, RBF 3); its gradient entries come back through gpss_get_grad2
 * after gpss_nlml_grad.  The first member, Sigma_Bias and sn2 stay in gpss_set_kernel / gpss_set_theta.  Not on partitioned handles. */
int gpss_set_kernel2(gpss_handle h, int kind2);
int gpss_set_theta2(gpss_handle h, const double* theta2);
int gpss_get_grad2(gpss_handle h, double g2[8]);
int gpss_get_theta(gpss_handle h, double theta[GPSS_NPAR]);
/* The main kernel of the Hyb{main, Bias} covariance -- the reference's `-k` choice (gp_ss_ak.cpp:146-170; HybKerns,
 * Kernel.cpp:140-169).  theta and g keep their 10-slot arrays; the slots in use are
 *   GPSS_KERNEL_EXPANS  {AngleX, iWx, AngleY, iWy, AngleZ, iWz, Sigma, iWR, Sigma_Bias, sn2}      (default; Kernel.cpp:856-1263)
 *   GPSS_KERNEL_EXP     {Hayper_Euc_Exp, Sigma_Exp, Sigma_Bias, sn2}                              (Kernel.cpp:636-695)
 *   GPSS_KERNEL_RBF     {Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF, Sigma_Bias, sn2}            (Kernel.cpp:482-541)
 * (a covariance without the Bias member is the same call with Sigma_Bias = 0; its gradient slot is then ignored). */
#define GPSS_KERNEL_EXPANS 0
#define GPSS_KERNEL_EXP 1
#define GPSS_KERNEL_RBF 2
int gpss_set_kernel(gpss_handle h, int kind);

/* objective ---------------------------------------------------------------------------------------- */
/* Opt_Algs::ObjVal -> GP_utils::logLikelihood (Opt_pars.h:248-251, GP_Utils.cpp:1138-1162):
 * negative log marginal likelihood.  On GPSS_NOT_POSDEF *nlml is NaN. */
int gpss_nlml(gpss_handle h, double* nlml);
/* Opt_Algs::Grad_Values -> GP_utils::GradLL (Opt_pars.h:242-246, GP_Utils.cpp:1171-1262 +
 * Kernel.cpp:886-1263, 370-377): value and the reference's 10-vector g (quirks included). */
int gpss_nlml_grad(gpss_handle h, double* nlml, double g[GPSS_NPAR]);
/* GP_utils::Alpha after updateAlpha (GP_Utils.cpp:383-393): alpha = (K + sn2 I)^-1 y, n doubles. */
int gpss_get_alpha(gpss_handle h, double* alpha);
/* yhat = K*alpha as logLikelihood leaves it (GP_Utils.cpp:1147-1148), n doubles. */
int gpss_get_yhat(gpss_handle h, double* yhat);

/* distributed evaluation ---------------------------------------------------------------------------- */
/* One process per GPU.  After gpss_dist_init every rank holds a handle over the SAME data and gpss_set_theta /
 * gpss_nlml / gpss_nlml_grad / gpss_predict* become COLLECTIVE calls: all ranks must make them in the same order with
 * the same theta, and all receive the same results.  Layout: 512-wide block columns of the
 * Cholesky factor are owned round-robin, the owner factors its panel and broadcasts it (ncclBroadcast over NVLink),
 * every rank updates only its own block columns; L^-T is computed in balanced row slices without communication and
 * gathered once; B^-1 and the gradient reductions are split by rows and summed with one small ncclAllReduce.
 * The pipe of the long-k products is the handle's (gpss_get_ozaki): the int8 tensor-core pipe chosen at gpss_create stays on after
 * gpss_dist_init (its digit planes are replicated like the factor); GPSS_OZAKI_DIST=0 switches a multi-GPU handle to the DMMA pipe.
 * The reference has no counterpart (single process, single thread). */
int gpss_nccl_unique_id(void* id128);                           /* rank 0 creates it, the caller distributes the 128 bytes */
int gpss_dist_init(gpss_handle h, int rank, int world, const void* id128);
/* PARTITIONED storage for n too large to replicate (n = 200 000: n^2 doubles = 320 GB; 8 x B200 hold 40 GB of block columns
 * each): collective constructor, every rank passes the same data.  The handle supports gpss_set_theta, gpss_nlml,
 * gpss_nlml_grad, gpss_get_alpha, gpss_get_yhat and gpss_predict (mean, and mean + variance: the variance is a collective over the
 * SAME test points on every rank -- strips of U = L^-T are broadcast and every rank accumulates || U^T b ||^2 over the block columns
 * it owns; gpss_predict_shard with a variance on a slice of the points returns GPSS_ERR_STATE).  Right-looking block-column-cyclic Cholesky: the owner's NCCL panel broadcast buffer is
 * itself the DMMA operand of every rank's trailing update; the triangular solves pass the right-hand side around; for the gradient
 * U = L^-T is held as block rows owned cyclically and B^-1 is produced one 512-wide strip at a time and consumed by the fused
 * gradient reductions without ever being stored. */
int gpss_create_partitioned(int device, int rank, int world, const void* id128, int n, int d, const double* X_colmajor,
                            const double* y, gpss_handle* out);
/* The balanced row partitions used above (kind 0: rows of L^-T by flops, kind 1: rows of B^-1, kind 2: rows of L^-T for the int8 pipe --
 * narrow slices are whole waves of one CTA per SM, so their cost is their longest k-range): bounds[0..world], multiples of 128. */
int gpss_dist_partition(int n_pad, int world, int kind, int* bounds);
/* Host logic of the exchange of the L^-T row slices between ranks (no counterpart in the reference): the packed position and length of
 * every column j = r0 .. n_pad-1 of the slice [r0, r0 + rows), and the buffer size -- entries inside and below the 512-wide diagonal
 * blocks do not travel (every rank computes the diagonal blocks itself; below them U is zero). */
int gpss_dist_uslice_layout(int n_pad, int r0, int rows, long* offsets, int* lens, long* count);
/* The operation list one rank executes for a factor of `nblk` 512-wide block columns (pure host logic, no device needed):
 * 6 ints per operation {kind, column, first panel, panel count, broadcast root, side stream}; kind 0 = main stream waits
 * for the column's look-ahead updates, 1 = update on the main stream, 2 = factor the column, 3 = broadcast it from root,
 * 4 = look-ahead update on a side stream.  Writes min(*count, cap) operations. */
int gpss_dist_potrf_schedule(int nblk, int world, int rank, int* ops6, int cap, int* count);

/* prediction --------------------------------------------------------------------------------------- */
/* GP_utils::Calc_Out / posteriorMeanVar (GP_Utils.cpp:159-178, 1016-1041): predictive mean and variance
 * of m standardised test points EXACTLY as the reference returns them, i.e. including its post-processing of the
 * variance vector (gpss_var_postprocess below); var may be NULL (posteriorMean, :1005-1015).
 * The Mahalanobis centre uses the training set and ALL m test points as MahaDist does (Kernel.cpp:1391).
 * On a distributed handle this is a COLLECTIVE call (same Xs on every rank): the test points are split over the ranks,
 * L and alpha being replicated, and the slices are exchanged over NCCL so that every rank returns the full vectors. */
int gpss_predict(gpss_handle h, long m, const double* Xs_colmajor, double* mu, double* var);
/* One shard [first, first+count) of a test set of m_total points whose column sums are sums_total[3]: lets ranks
 * split the test points (L and alpha replicated) yet use the global centre.  var receives the RAW variance
 * kD - k*' (K + sn2 I)^-1 k* (GP_Utils.cpp:997-998) with no post-processing: the reference's next step is index
 * arithmetic over the whole vector, so the caller gathers the shards and calls gpss_var_postprocess once. */
int gpss_predict_shard(gpss_handle h, long m_total, const double* sums_total /* [d] */, long count,
                       const double* Xs_shard_colmajor, double* mu, double* var);
/* GP_Utils.cpp:1001-1003 + 1033-1040, literally: `uvec ind = varSigma < 0; varSigma.elem(ind) = 0` uses the 0/1
 * comparison flags AS INDICES (element 0 is zeroed if any entry is >= 0, element 1 if any entry is < 0, negative
 * entries are left as they are), then sn2 is added unless sn2 == 1.0.  Host-side, O(m). */
int gpss_var_postprocess(long m, double sn2, double* var_inout);

/* Kernels::computeK compatibility (host matrices; Kernel.cpp:140-154, 856-882, 362-367) ---------- */
/* K and D2 are n1 x n2 column-major host buffers (either may be NULL); X1, X2 have d (3 or 4) columns. */
int gpss_compute_K(int device, int kind, const double theta[GPSS_NPAR], int d, int n1, const double* X1, int n2,
                   const double* X2, double* K, double* D2);

/* Kern_ExpAnisotropic::getGradients compatibility (Kernel.cpp:886-1263): the 8 kernel-parameter entries of the
 * reference's gradient for a HOST n x n matrix QW (column-major, need not be symmetric) and X1 == X2 == X (n x d). */
int gpss_expans_gradients(int device, const double theta[GPSS_NPAR], int d, int n, const double* X_colmajor,
                          const double* QW_colmajor, double g8[8]);

/* instrumentation ---------------------------------------------------------------------------------- */
/* Device time in ms of the phases of the last gpss_nlml / gpss_nlml_grad / gpss_predict on this handle:
 * [0] K build, [1] potrf, [2] solves+objective terms, [3] trtri, [4] lauum, [5] gradient pass,
 * [6] cross-covariance+mean, [7] variance GEMM, [8..15] reserved.  Requires gpss_set_profiling(h,1). */
int gpss_set_profiling(gpss_handle h, int on);
int gpss_get_phase_ms(gpss_handle h, double ms[16]);
/* Device time (CUDA events on the handle's stream) of the last gpss_nlml / gpss_nlml_grad / gpss_predict*. */
int gpss_get_last_call_ms(gpss_handle h, double* ms);
/* FP64 tensor-pipe (DMMA.8x8x4) micro-peak in TFLOP/s, measured live: the roofline denominator. */
int gpss_measure_fp64_peak(int device, double* tflops);
/* int8 tensor-pipe micro-peak in TOP/s, measured live: tcgen05.mma kind::i8 (M 128, N 256, K 32) from operands resident in
 * shared memory on every SM -- burst (best ~6 ms launch) and sustained (~1 s back to back, under the power cap).  The roofline
 * denominator of oz_gemm_kernel (MEASURED_PEAKS.json carries no int8 figure). */
int gpss_measure_int8_peak(int device, double* tops_burst, double* tops_sustained);
/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
int gpss_get_launch_count(gpss_handle h, long* launches);
/* Raw device pointer to the n_pad x n_pad factor / inverse and the padded size (tests only). */
int gpss_debug_fetch(gpss_handle h, int which, double* host_out, long count);
int gpss_padded_n(gpss_handle h, int* n_pad);
/* Which pipe runs the three long-k contractions of this handle (potrf look-ahead update, bulk product of the triangular
 * inverse, B^-1 = U U^T): 0 = FP64 DMMA (gemm_nt_ws_kernel), 6 | 7 | 8 = int8 tensor cores with that many 7-bit Ozaki slices
 * (oz_gemm_kernel, csrc/gpss_ozaki.cuh).  Chosen at gpss_create from n and the GPSS_OZAKI environment variable.
 * The answer is per theta: the int8 scaling rests on |L^-1_ij| <= 1, i.e. on B = I + K / sn2 >= I, which the reference's
 * unconstrained parameters do not guarantee (Kern_Bias adds Sigma_Bias raw, Kernel.cpp:362-367).  A theta with Sigma_Bias < 0 or
 * sn2 <= 0 is evaluated on the DMMA pipe (0 is returned after gpss_set_theta), and if an operand still leaves its bound the
 * device flags it and the evaluation is repeated on the DMMA pipe: gpss_get_ozaki_fallbacks counts those repeats. */
int gpss_get_ozaki(gpss_handle h, int* slices);
int gpss_get_ozaki_fallbacks(gpss_handle h, long* count);
/* Width of the signed digits the int8 pipe cuts operands into: 7 (base 128, digits in [-64, 64]) or 8 (base 256, [-128, 127]). */
int gpss_get_ozaki_bits(gpss_handle h, int* bits);

/* kernel-level test hooks (tests/ only) ------------------------------------------------------------ */
/* C(MxN) = A(MxK) * B(NxK)^T with host buffers, through the DMMA kernel; tile: 0 = the warp-specialised
 * bulk-copy/mbarrier kernel every product of the path uses, 1 = the legacy cp.async kernel (A/B baseline),
 * 2..8 = the same kernel in its split-k form (that many partial products summed in a fixed order), which the
 * distributed triangular inverse uses to fill the machine from a narrow row slice, 9 = tile 0 with a 2-stage ring (the 51 KB
 * variant that shares an SM with a resident int8 CTA). */
int gpss_test_gemm_nt(int device, int tile, int M, int N, int K, const double* A, const double* B, double* C,
                      int subtract_from_C, double* ms_out);
/* C(MxN) = A B^T, or C - A B^T, with host buffers through the int8 tensor-core kernel (csrc/gpss_ozaki.cuh) with `slices` in
 * {6,7,8} 7-bit digits per operand; |A|, |B| <= 1; M % 128 == N % 64 == K % 64 == 0, M, N <= K.  Bit-exact against the numpy
 * restatement oracle/ozaki_oracle.py::oz_gemm_nt. */
int gpss_test_oz_gemm(int device, int slices, int M, int N, int K, const double* A, const double* B, double* C,
                      int subtract_from_C, double* ms_out);
/* In-place blocked Cholesky of a host n x n SPD matrix (lower), through the full potrf driver. */
int gpss_test_potrf(int device, int n, double* A_colmajor, double* logdet_half, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* GPSS_H */
