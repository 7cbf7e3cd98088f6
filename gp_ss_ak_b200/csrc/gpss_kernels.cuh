// Hand-written sm_100a kernels for the exact-GP hot path of GP_SS_AK (everything except the DMMA GEMM,
// which lives in gpss_gemm.cuh).  File:line citations are into /root/reference.
//
// Data layout in HBM (all FP64, column-major like arma::mat):
//   xs[4][n_pad]   standardised, UNcentred coordinates (the reference's Xinp), SoA; row 3 = rock-type column (zeros when d = 3)
//   zs[5][n_pad]   z = (x - c) * sigInv (rows 0-2), a = |z|^2 (row 3), z_3 = (x_3 - c_3) * InversewidthR (row 4; 0 when d = 3), SoA
//   Lm[n_pad^2]    B = I + (Sw Sw') o K  -> overwritten by its Cholesky factor L (lower)
//   Um[n_pad^2]    U = L^-T (upper)
//   Qm[n_pad^2]    Q = B^-1 (lower triangle)  (also W = L^-1 for prediction)
//   Winv[nblk][128*128]  inverses of the 128x128 diagonal blocks of L
// n_pad = n rounded up to 128; the padding rows/cols carry the identity so no kernel needs edge handling.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gpss {

constexpr int NB = 128;          // diagonal / tile block

constexpr int NZ = 5;            // rows of a transformed-coordinate array
constexpr int NX = 4;            // rows of a coordinate array

struct DevParams {
  double S[9];        // sigInv = Rot*diag(l)*Rot' (Kernel.cpp:1425), S[k*3+j]
  double c[4];        // MahaDist centre (Kernel.cpp:1391-1392)
  double lr;          // InversewidthR_ExpAns = sigInv(3,3) of the 4-column branch (Kernel.cpp:1411-1424)
  int dim;            // 3, or 4 with the rock-type column
  int kind;           // main kernel of Hyb{main, Bias}: 0 ExpAns, 1 Exp, 2 RBF (GPSS_KERNEL_*); the isotropic kernels use
                      // S = (1/hyp) I, lr = 1/hyp, i.e. EuclDist (Kernel.cpp:1343-1368) through the same pair-distance code
  double rbf_c;       // -0.5 * inverseWidth_RBF (Kernel.cpp:486)
  double var2;        // Sigma_ExpAns^2 (Kernel.cpp:861)
  double bias;        // Sigma_Bias (Kernel.cpp:366)
  double white;       // Sigma_White, summed over the White members: added to K_ii of the training covariance only (Kernel.cpp:257-264)
  double var2x;       // Sigma^2 of a SECOND distance-based member of the Hyb sum (0 without one; its own parameters live in a second
                      // DevParams block with bias = white = 0 and the same sn2 fields): enters the a-priori bound of the int8 scaling
  double sn2;         // hyperlf(0) (GP_Utils.cpp:406)
  double inv_sn2;     // 1/sn2 = d2lp (GP_Utils.cpp:412-413)
  double sw;          // Sw = sqrt(d2lp) (GP_Utils.cpp:897)
  double sww;         // fl(Sw*Sw), the element of Sw*Sw.t() (GP_Utils.cpp:900)
  double lp_const;    // log(2*pi*sn2)/2 (GP_Utils.cpp:810)
};

// ---------------------------------------------------------------------------------------------------
// defined-order Mahalanobis pieces (SURVEY.md section 7 hard part 1; mirrors oracle maha_dist_defined)
// ---------------------------------------------------------------------------------------------------
// The 4th coordinate (z_3, zero in the 3-column case) enters last, so with d = 3 every value below is bit-identical to
// the 3-term form: fma(0, 0, c) = c and a + 0 = a.
__device__ __forceinline__ double pair_d2(double zi0, double zi1, double zi2, double ai,
                                          double zj0, double zj1, double zj2, double aj, double zi3, double zj3)
{
  const double cij = fma(zi3, zj3, fma(zi2, zj2, fma(zi1, zj1, __dmul_rn(zi0, zj0))));
  const double d2 = __dadd_rn(__dadd_rn(ai, aj), __dmul_rn(-2.0, cij));
  return d2 < 0.0 ? 0.0 : d2;     // find(D2<0) -> 0 (Kernel.cpp:1433-1434)
}

// z = (x - c) * S, a = fl(fl(z0^2+z1^2)+z2^2)   (Kernel.cpp:1393-1397, 1426-1432)
__global__ void transform_kernel(const double* __restrict__ xs, long ldx, double* __restrict__ zs, long ldz,
                                 int n, int n_pad, const DevParams* __restrict__ P)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i >= n) { zs[i] = 0; zs[ldz + i] = 0; zs[2 * ldz + i] = 0; zs[3 * ldz + i] = 0; zs[4 * ldz + i] = 0; return; }
  const double d0 = __dsub_rn(xs[i], P->c[0]);
  const double d1 = __dsub_rn(xs[ldx + i], P->c[1]);
  const double d2 = __dsub_rn(xs[2 * ldx + i], P->c[2]);
  double z[3];
#pragma unroll
  for (int j = 0; j < 3; j++) z[j] = fma(d2, P->S[6 + j], fma(d1, P->S[3 + j], __dmul_rn(d0, P->S[j])));
  zs[i] = z[0]; zs[ldz + i] = z[1]; zs[2 * ldz + i] = z[2];
  const double z3 = (P->dim == 4) ? __dmul_rn(__dsub_rn(xs[3 * ldx + i], P->c[3]), P->lr) : 0.0;
  zs[4 * ldz + i] = z3;
  zs[3 * ldz + i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(z[0], z[0]), __dmul_rn(z[1], z[1])), __dmul_rn(z[2], z[2])), __dmul_rn(z3, z3));
}

// ---------------------------------------------------------------------------------------------------
// K1: lower-triangle tiles of B = I + (Sw Sw') o K,  K = var2*exp(-sqrt(D2)) + bias
//     (Kernel.cpp:881, 366, 140-154; GP_Utils.cpp:898-902).  One 128x128 tile per CTA, 256 threads,
//     each thread owns 2 consecutive rows (16-byte coalesced stores down a column) x 32 columns.
// ---------------------------------------------------------------------------------------------------
// one member's covariance without the Bias term
__device__ __forceinline__ double kern_main(double d2, const DevParams& P)
{
  if (P.kind == 2) return __dmul_rn(exp(__dmul_rn(P.rbf_c, d2)), P.var2);   // RBF (Kernel.cpp:486)
  return __dmul_rn(P.var2, exp(-sqrt(d2)));                                  // ExpAns, Exp (Kernel.cpp:881, 640)
}
// Hyb{member 1, member 2, Bias} (HybKerns::computeK, Kernel.cpp:140-154): the members' values added in order, then the bias
__device__ __forceinline__ double kern_val2(double d2a, const DevParams& Pa, double d2b, const DevParams& Pb)
{
  return __dadd_rn(__dadd_rn(kern_main(d2a, Pa), kern_main(d2b, Pb)), Pa.bias);
}
__device__ __forceinline__ double kern_val(double d2, const DevParams& P)
{
  if (P.kind == 2) return __dadd_rn(__dmul_rn(exp(__dmul_rn(P.rbf_c, d2)), P.var2), P.bias);   // RBF (Kernel.cpp:486)
  return __dadd_rn(__dmul_rn(P.var2, exp(-sqrt(d2))), P.bias);                                  // ExpAns, Exp (Kernel.cpp:881, 640)
}

template <bool TWO = false>
__global__ void __launch_bounds__(256) kbuild_lower_kernel(double* __restrict__ Bm, long ld, const double* __restrict__ zs, long ldz,
                                                           int n, const DevParams* __restrict__ Pp, int raw_K, int own_world, int own_rank,
                                                           int own_width, const double* __restrict__ zs2 = nullptr,
                                                           const DevParams* __restrict__ P2p = nullptr)
{
  // TWO: a second distance-based member (its transformed coordinates zs2, its parameters P2p) is added to every element
  // own_world > 1: only the tile columns of the block columns (own_width tiles wide) this rank owns in the distributed
  // Cholesky are built -- every other block column arrives already factored with the owner's broadcast.
  // own_width < 0 (partitioned storage): Bm holds ONLY this rank's block columns, packed; blockIdx.y is the local tile column.
  const int tm = blockIdx.x;
  int tn = blockIdx.y;
  const int tn_store = tn;
  if (own_width < 0) {
    const int w = -own_width;
    tn = ((tn / w) * own_world + own_rank) * w + tn % w;
    if (tn >= (int)gridDim.x) return;
  } else if (own_world > 1 && (tn / own_width) % own_world != own_rank) return;
  if (tn > tm) return;
  __shared__ double cz[NZ][NB];
  __shared__ double cz2[TWO ? NZ : 1][TWO ? NB : 1];
  __shared__ DevParams P, P2;
  const int tid = threadIdx.x;
  if (tid == 0) { P = *Pp; if constexpr (TWO) P2 = *P2p; }
  const int r0 = tm * NB, c0 = tn * NB;
  for (int idx = tid; idx < NZ * NB; idx += 256) {
    cz[idx / NB][idx % NB] = zs[(long)(idx / NB) * ldz + c0 + idx % NB];
    if constexpr (TWO) cz2[idx / NB][idx % NB] = zs2[(long)(idx / NB) * ldz + c0 + idx % NB];
  }
  const int tx = tid & 63, ty = tid >> 6;
  const int i0 = r0 + 2 * tx;
  double zi[2][NZ], zi2[2][TWO ? NZ : 1];
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int q = 0; q < NZ; q++) {
      zi[e][q] = zs[(long)q * ldz + i0 + e];
      if constexpr (TWO) zi2[e][q] = zs2[(long)q * ldz + i0 + e];
    }
  __syncthreads();
#pragma unroll 4
  for (int jj = ty; jj < NB; jj += 4) {
    const int j = c0 + jj;
    double2 out;
    double* o = &out.x;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = i0 + e;
      double v;
      if (i < n && j < n) {
        const double d2 = pair_d2(zi[e][0], zi[e][1], zi[e][2], zi[e][3], cz[0][jj], cz[1][jj], cz[2][jj], cz[3][jj], zi[e][4], cz[4][jj]);
        if constexpr (TWO) {
          const double d2b = pair_d2(zi2[e][0], zi2[e][1], zi2[e][2], zi2[e][3], cz2[0][jj], cz2[1][jj], cz2[2][jj], cz2[3][jj], zi2[e][4], cz2[4][jj]);
          v = kern_val2(d2, P, d2b, P2);
        } else {
          v = kern_val(d2, P);
        }
        if (i == j) v = __dadd_rn(v, P.white);                 // Kern_White: K.diag() += Sigma_White (0 without a White member: v unchanged)
        if (!raw_K) {
          v = __dmul_rn(P.sww, v);
          if (i == j) v = __dadd_rn(v, 1.0);
        }
      } else {
        v = (i == j) ? 1.0 : 0.0;
      }
      o[e] = v;
    }
    *reinterpret_cast<double2*>(Bm + (long)(tn_store * NB + jj) * ld + i0) = out;
  }
}

// ---------------------------------------------------------------------------------------------------
// K3: Cholesky of one 128x128 diagonal block in shared memory + its triangular inverse.
//     Replaces the diagonal-block dpotf2 inside arma::chol (GP_Utils.cpp:881,903) and supplies
//     inv(L11) so that every panel solve becomes a DMMA GEMM.  Also accumulates sum(log(diag))
//     (GP_Utils.cpp:913) and raises *flag when a pivot is not positive (chol() == false, :882-886).
//
//     One CTA of 256 threads.  The lower triangle lives in shared memory as ten packed 32x32 blocks
//     (column-major, leading dimension 33: row- and column-walks are bank-conflict free) plus the four
//     diagonal-block inverses: 119 KB and <= 144 registers/thread, so the kernel fits on an SM NEXT TO a
//     resident trailing-update CTA -- the look-ahead in potrf_blocked depends on that.
//     Blocked right-looking with 32-wide steps:
//       (a) warp 0 factors the 32x32 diagonal block entirely in registers (lane = row) with warp-shuffle
//           broadcasts of the pivot column, then inverts it (lane = column, forward substitution);
//       (b) panel blocks are multiplied by inv(L_kk)^T and (c) the rank-32 trailing update is applied, one
//           32x32 output block per 64-thread group (row-per-thread, the other operand broadcast from smem).
//     The 128x128 inverse is then assembled recursively:  W = [[W11,0],[-W22 L21 W11, W22]]
//     at block size 32 -> 64 -> 128.
// ---------------------------------------------------------------------------------------------------
constexpr int DIAG_THREADS = 256;
constexpr int BLD = 33;                      // leading dimension of a packed 32x32 block
constexpr int BSZ = 32 * BLD;                // doubles per block
constexpr size_t DIAG_SMEM = (size_t)(14 * BSZ + NB) * sizeof(double);

__device__ __forceinline__ int blk_index(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }   // bi >= bj

// out[q] += sum_k A(r,k) * B(k, 16g+q)   with A(r,k) at Ab[k*BLD + r] and
//   TRANSB = false: B(k,c) at Bb[c*BLD + k] ;  TRANSB = true: B(k,c) = Bt(c,k) at Bb[k*BLD + c]
template <bool TRANSB>
__device__ __forceinline__ void blk_acc(double (&out)[16], const double* __restrict__ Ab, const double* __restrict__ Bb, int r, int g)
{
  const double* bq = Bb + (TRANSB ? 16 * g : 16 * g * BLD);
#pragma unroll 4
  for (int k = 0; k < 32; k++) {
    const double a = Ab[k * BLD + r];
#pragma unroll
    for (int q = 0; q < 16; q++) out[q] = fma(a, TRANSB ? bq[k * BLD + q] : bq[q * BLD + k], out[q]);
  }
}

__device__ __forceinline__ void blk_zero(double (&out)[16])
{
#pragma unroll
  for (int q = 0; q < 16; q++) out[q] = 0.0;
}

// 32x32 Cholesky in registers: lane = row.  Lb = packed block (leading dimension BLD).
__device__ __forceinline__ void warp_chol32(double* Lb, double* invd, double& my_diag, int* flag, int lane)
{
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; c++) a[c] = Lb[c * BLD + lane];
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const double ajj = __shfl_sync(0xffffffffu, a[j], j);
    double d;
    if (!(ajj > 0.0)) { if (lane == 0) atomicExch(flag, 1); d = nan(""); }
    else d = sqrt(ajj);
    const double inv = 1.0 / d;
    const double l = a[j] * inv;            // column j of L for rows (lanes) > j
    if (lane == j) { a[j] = d; my_diag = d; invd[j] = inv; }
    else a[j] = l;
#pragma unroll
    for (int c = j + 1; c < 32; c++) {
      const double lc = __shfl_sync(0xffffffffu, l, c);
      a[c] = fma(-l, lc, a[c]);             // meaningful for rows >= c only
    }
  }
#pragma unroll
  for (int c = 0; c < 32; c++) Lb[c * BLD + lane] = (lane >= c) ? a[c] : 0.0;
}

// inverse of the 32x32 lower-triangular block Lb: lane = column of W, forward substitution with the
// matrix entries broadcast from shared memory.  W(r,c) -> Wb[c*BLD + r] (zeros above the diagonal).
__device__ __forceinline__ void warp_inv32(const double* Lb, const double* invd, double* Wb, int lane)
{
  double s[32];
#pragma unroll
  for (int r = 0; r < 32; r++) s[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < 32; k++) {
    const double wk = s[k] * invd[k];
    s[k] = wk;
#pragma unroll
    for (int r = k + 1; r < 32; r++) s[r] = fma(-Lb[k * BLD + r], wk, s[r]);
  }
#pragma unroll
  for (int r = 0; r < 32; r++) Wb[lane * BLD + r] = s[r];
}

__global__ void __maxnreg__(144)
potrf_diag_inv_kernel(double* __restrict__ A, long ld, double* __restrict__ Winv, double* __restrict__ logdet_part, int* __restrict__ flag)
{
  extern __shared__ double sm[];
  double* Lb = sm;                       // ten packed lower blocks, blk_index(bi,bj)
  double* Wd = sm + 10 * BSZ;            // four diagonal-block inverses
  double* invd = Wd + 4 * BSZ;           // reciprocal pivots
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid >> 6, r = tid & 31, g = (tid >> 5) & 1;      // 64-thread groups: row r, column half g
  {
    const int i = tid & (NB - 1), bi = i >> 5;
    for (int k = tid >> 7; k < NB; k += DIAG_THREADS / NB) {
      const int bk = k >> 5;
      if (bk <= bi) Lb[blk_index(bi, bk) * BSZ + (k & 31) * BLD + (i & 31)] = A[(long)k * ld + i];
    }
  }
  __syncthreads();
  double logsum = 0.0;
  for (int kb = 0; kb < 4; kb++) {
    if (warp == 0) {
      double my_diag = 1.0;
      double* Lkk = Lb + blk_index(kb, kb) * BSZ;
      warp_chol32(Lkk, invd + 32 * kb, my_diag, flag, lane);
      __syncwarp();
      warp_inv32(Lkk, invd + 32 * kb, Wd + kb * BSZ, lane);
      logsum += log(my_diag);
    }
    __syncthreads();
    if (kb == 3) break;
    // (b) panel blocks P_bi <- P_bi * inv(L_kk)^T, one block per group (in place: barrier between read and write)
    {
      const int bi = kb + 1 + grp;
      double out[16];
      blk_zero(out);
      double* Pb = Lb + blk_index(bi <= 3 ? bi : 3, kb) * BSZ;
      if (bi <= 3) blk_acc<true>(out, Pb, Wd + kb * BSZ, r, g);
      __syncthreads();
      if (bi <= 3) {
#pragma unroll
        for (int q = 0; q < 16; q++) Pb[(16 * g + q) * BLD + r] = out[q];
      }
      __syncthreads();
    }
    // (c) trailing blocks (bi >= bj > kb):  C -= P_bi P_bj^T
    {
      const int nrem = 3 - kb, nout = nrem * (nrem + 1) / 2;
      for (int o = grp; o < nout; o += 4) {
        int bi = kb + 1, bj = kb + 1, cnt = o;
        while (cnt > bi - (kb + 1)) { cnt -= bi - kb; bi++; }      // row-by-row enumeration of the lower blocks
        bj = kb + 1 + cnt;
        double out[16];
        blk_zero(out);
        blk_acc<true>(out, Lb + blk_index(bi, kb) * BSZ, Lb + blk_index(bj, kb) * BSZ, r, g);
        double* Cb = Lb + blk_index(bi, bj) * BSZ;
#pragma unroll
        for (int q = 0; q < 16; q++) Cb[(16 * g + q) * BLD + r] -= out[q];
      }
      __syncthreads();
    }
  }
  // factor back to global (the strict upper part of the tile is zeroed)
  {
    const int i = tid & (NB - 1), bi = i >> 5;
    for (int k = tid >> 7; k < NB; k += DIAG_THREADS / NB) {
      const int bk = k >> 5;
      A[(long)k * ld + i] = (bk <= bi) ? Lb[blk_index(bi, bk) * BSZ + (k & 31) * BLD + (i & 31)] : 0.0;
    }
  }
  if (warp == 0) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) logsum += __shfl_xor_sync(0xffffffffu, logsum, off);
    if (lane == 0) *logdet_part = logsum;
  }
  __syncthreads();
  // ---- inverse: W(b,b) = Wd[b]; off-diagonal blocks overwrite the packed L blocks ----
  {
    // level 1: X(1,0) = -W11 (L10 W00) [groups 0,1 idle in pairs: group 0 -> (1,0), group 1 -> (3,2)]
    const int b = 2 * grp;                                   // grp 0 -> b=0, grp 1 -> b=2
    const bool on = grp < 2;
    double* Xb = Lb + blk_index(on ? b + 1 : 1, on ? b : 0) * BSZ;
    double out[16];
    blk_zero(out);
    if (on) blk_acc<false>(out, Xb, Wd + b * BSZ, r, g);     // T = L10 * W00
    __syncthreads();
    if (on) {
#pragma unroll
      for (int q = 0; q < 16; q++) Xb[(16 * g + q) * BLD + r] = out[q];
    }
    __syncthreads();
    blk_zero(out);
    if (on) blk_acc<false>(out, Wd + (b + 1) * BSZ, Xb, r, g);   // W11 * T
    __syncthreads();
    if (on) {
#pragma unroll
      for (int q = 0; q < 16; q++) Xb[(16 * g + q) * BLD + r] = -out[q];
    }
    __syncthreads();
  }
  {
    // level 2: blocks (2+bi, bj), bi,bj in {0,1}, one per group:  X = -W22' (L21 W11')
    const int bi = grp >> 1, bj = grp & 1;
    double* Xb = Lb + blk_index(2 + bi, bj) * BSZ;
    double out[16];
    blk_zero(out);
    // T(bi,bj) = sum_{kb >= bj} L(2+bi, kb) W11'(kb, bj)
    blk_acc<false>(out, Lb + blk_index(2 + bi, bj) * BSZ, Wd + bj * BSZ, r, g);
    if (bj == 0) blk_acc<false>(out, Lb + blk_index(2 + bi, 1) * BSZ, Lb + blk_index(1, 0) * BSZ, r, g);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; q++) Xb[(16 * g + q) * BLD + r] = out[q];
    __syncthreads();
    // X(bi,bj) = - sum_{kb <= bi} W22'(bi, kb) T(kb, bj)
    blk_zero(out);
    blk_acc<false>(out, Wd + (2 + bi) * BSZ, Lb + blk_index(2 + bi, bj) * BSZ, r, g);
    if (bi == 1) blk_acc<false>(out, Lb + blk_index(3, 2) * BSZ, Lb + blk_index(2, bj) * BSZ, r, g);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; q++) Xb[(16 * g + q) * BLD + r] = -out[q];
    __syncthreads();
  }
  {
    const int i = tid & (NB - 1), bi = i >> 5;
    for (int k = tid >> 7; k < NB; k += DIAG_THREADS / NB) {
      const int bk = k >> 5;
      double v = 0.0;
      if (bk == bi) v = Wd[bi * BSZ + (k & 31) * BLD + (i & 31)];
      else if (bk < bi) v = Lb[blk_index(bi, bk) * BSZ + (k & 31) * BLD + (i & 31)];
      Winv[k * NB + i] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K4 (vector right-hand side): blocked triangular solves with the stored diagonal-block inverses.
//     Replaces solve_chol's two dtrtrs for alpha (GP_Utils.cpp:841-845 via :893).
//     One launch per 128-block step; a launch streams the 128-column panel of L exactly once (HBM-bound: the two
//     sweeps read 8 n^2 bytes in total).  Every CTA applies the current block solution to its 128 rows (columns);
//     the CTA that owns the NEXT diagonal block then produces that block's solution, so consecutive launches need
//     no extra kernel in between.
// ---------------------------------------------------------------------------------------------------
constexpr int TRSV_THREADS = 256;

// s[0..127] = Winv(128x128 lower, column-major ld 128) * v   (TRANS = false)   or   Winv^T * v   (TRANS = true); v, s in smem
template <bool TRANS>
__device__ __forceinline__ void block_tri_matvec(const double* __restrict__ Winv, const double* v, double* s_out, double* scratch)
{
  const int t = threadIdx.x, r = t & 127, h = t >> 7;
  double acc = 0.0;
  if (!TRANS) {
    // s[r] = sum_{c <= r} W(r,c) v[c]; the two halves of the CTA split the columns
    for (int c = h * 64; c < h * 64 + 64 && c <= r; c++) acc = fma(Winv[c * NB + r], v[c], acc);
  } else {
    // s[r] = sum_{q >= r} W(q,r) v[q]; column r is contiguous
    for (int q = max(r, h * 64); q < h * 64 + 64; q++) acc = fma(Winv[r * NB + q], v[q], acc);
  }
  scratch[h * NB + r] = acc;
  __syncthreads();
  if (t < NB) s_out[t] = scratch[t] + scratch[NB + t];
  __syncthreads();
}

// first block of the forward sweep: z_0 = Winv_0 r_0
__global__ void __launch_bounds__(TRSV_THREADS) trsv_fwd_first_kernel(const double* __restrict__ Winv, const double* __restrict__ r,
                                                                       double* __restrict__ z)
{
  __shared__ double v[NB], s[NB], scratch[2 * NB];
  if (threadIdx.x < NB) v[threadIdx.x] = r[threadIdx.x];
  __syncthreads();
  block_tri_matvec<false>(Winv, v, s, scratch);
  if (threadIdx.x < NB) z[threadIdx.x] = s[threadIdx.x];
}

// forward step k: rows below block k get  r[i] -= L[i, kblk] z_k ;  the CTA of block k+1 then stores z_{k+1} = Winv_{k+1} r_{k+1}.
// grid = number of 128-row tiles below block k.
__global__ void __launch_bounds__(TRSV_THREADS) trsv_fwd_step_kernel(const double* __restrict__ L, long ld, const double* __restrict__ Winv,
                                                                      double* __restrict__ r, double* __restrict__ z, int k0)
{
  __shared__ double zk[NB], rn[NB], s[NB], scratch[2 * NB];
  const int t = threadIdx.x, row = t & 127, h = t >> 7;
  if (t < NB) zk[t] = z[k0 + t];
  __syncthreads();
  const long i = (long)k0 + NB + (long)blockIdx.x * NB + row;
  const double* Lp = L + (long)(k0 + h * 64) * ld + i;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    double l[8];
#pragma unroll
    for (int q = 0; q < 8; q++) l[q] = Lp[(long)(c + q) * ld];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = fma(l[q], zk[h * 64 + c + q], a[q]);
  }
  scratch[h * NB + row] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (t < NB) {
    const double v = r[i] - (scratch[t] + scratch[NB + t]);
    r[i] = v;
    rn[t] = v;
  }
  if (blockIdx.x != 0) return;
  __syncthreads();
  block_tri_matvec<false>(Winv + (long)(k0 / NB + 1) * NB * NB, rn, s, scratch);
  if (t < NB) z[k0 + NB + t] = s[t];
}

// last block of the backward sweep's start: x_last = Winv_last^T r_last
__global__ void __launch_bounds__(TRSV_THREADS) trsv_bwd_first_kernel(const double* __restrict__ Winv_last, const double* __restrict__ r,
                                                                       double* __restrict__ x, int k0)
{
  __shared__ double v[NB], s[NB], scratch[2 * NB];
  if (threadIdx.x < NB) v[threadIdx.x] = r[k0 + threadIdx.x];
  __syncthreads();
  block_tri_matvec<true>(Winv_last, v, s, scratch);
  if (threadIdx.x < NB) x[k0 + threadIdx.x] = s[threadIdx.x];
}

// backward step k: column tiles left of block k get  r[j] -= L[kblk, j]^T x_k ;  the CTA of block k-1 then stores
// x_{k-1} = Winv_{k-1}^T r_{k-1}.  grid = k (number of 128-column tiles left of block k); tile b covers columns 128 b.
__global__ void __launch_bounds__(TRSV_THREADS) trsv_bwd_step_kernel(const double* __restrict__ L, long ld, const double* __restrict__ Winv,
                                                                      double* __restrict__ r, double* __restrict__ x, int k0)
{
  __shared__ double xk[NB], rn[NB], s[NB], scratch[2 * NB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < NB) xk[t] = x[k0 + t];
  __syncthreads();
  const int j0 = blockIdx.x * NB;
  // warp w owns columns j0 + 16 w .. + 15; lanes run down the 128 rows of the tile (4 each), then a shuffle reduction
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const double* Lp = L + (long)(j0 + warp * 16 + c) * ld + k0 + lane;
    acc[c] = fma(Lp[0], xk[lane], fma(Lp[32], xk[lane + 32], fma(Lp[64], xk[lane + 64], Lp[96] * xk[lane + 96])));
  }
#pragma unroll
  for (int c = 0; c < 16; c++) {
    double v = acc[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[c] = v;
  }
  if (lane < 16) {
    double mine = acc[0];
#pragma unroll
    for (int c = 1; c < 16; c++) if (lane == c) mine = acc[c];
    const int j = j0 + warp * 16 + lane;
    const double v = r[j] - mine;
    r[j] = v;
    rn[warp * 16 + lane] = v;
  }
  if ((int)blockIdx.x != k0 / NB - 1) return;
  __syncthreads();
  block_tri_matvec<true>(Winv + (long)(k0 / NB - 1) * NB * NB, rn, s, scratch);
  if (t < NB) x[k0 - NB + t] = s[t];
}

// Partitioned storage (L held as packed block columns, `w` 128-tiles per block column, owner = block column % P): the backward
// step k over THIS rank's column tiles.  blockIdx.x is a local tile; its global tile must lie left of block k.  r is valid on a
// rank only at the columns it owns (each rank keeps its own columns current); the owner of tile k-1 produces x_{k-1}, which
// the host then broadcasts.
__global__ void __launch_bounds__(TRSV_THREADS) trsv_bwd_step_part_kernel(const double* __restrict__ Lloc, long ld, const double* __restrict__ Winv,
                                                                           double* __restrict__ r, double* __restrict__ x, int k0, int P,
                                                                           int me, int w)
{
  const int lt = blockIdx.x;
  const int gt = ((lt / w) * P + me) * w + lt % w;            // global 128-column tile
  if (gt >= k0 / NB) return;
  __shared__ double xk[NB], rn[NB], s[NB], scratch[2 * NB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < NB) xk[t] = x[k0 + t];
  __syncthreads();
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const double* Lp = Lloc + (long)(lt * NB + warp * 16 + c) * ld + k0 + lane;
    acc[c] = fma(Lp[0], xk[lane], fma(Lp[32], xk[lane + 32], fma(Lp[64], xk[lane + 64], Lp[96] * xk[lane + 96])));
  }
#pragma unroll
  for (int c = 0; c < 16; c++) {
    double v = acc[c];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    acc[c] = v;
  }
  if (lane < 16) {
    double mine = acc[0];
#pragma unroll
    for (int c = 1; c < 16; c++) if (lane == c) mine = acc[c];
    const int j = gt * NB + warp * 16 + lane;
    const double v = r[j] - mine;
    r[j] = v;
    rn[warp * 16 + lane] = v;
  }
  if (gt != k0 / NB - 1) return;
  __syncthreads();
  block_tri_matvec<true>(Winv + (long)(k0 / NB - 1) * NB * NB, rn, s, scratch);
  if (t < NB) x[k0 - NB + t] = s[t];
}

// ---------------------------------------------------------------------------------------------------
// f = K * alpha with K regenerated from coordinates (mvmK_exact, GP_Utils.cpp:394-397, used at :1147).
// 64 rows per CTA, 256 threads: thread (row, part) strides over columns, partials combined in smem.
// ---------------------------------------------------------------------------------------------------
template <bool TWO = false>
__global__ void __launch_bounds__(256) kmatvec_kernel(const double* __restrict__ zs, long ldz, const double* __restrict__ alpha,
                                                      double* __restrict__ f, int n, const DevParams* __restrict__ Pp,
                                                      const double* __restrict__ zs2 = nullptr, const DevParams* __restrict__ P2p = nullptr)
{
  __shared__ double cz[6][256];
  __shared__ double cy[TWO ? 5 : 1][TWO ? 256 : 1];             // the second member's transformed coordinates of the column block
  __shared__ double part[4][64];
  __shared__ DevParams P, P2;
  const int tid = threadIdx.x, rr = tid & 63, pp = tid >> 6;
  if (tid == 0) { P = *Pp; if constexpr (TWO) P2 = *P2p; }
  const int i = blockIdx.x * 64 + rr;
  const bool vi = i < n;
  const int ic = vi ? i : 0;
  const double z0 = zs[ic], z1 = zs[ldz + ic], z2 = zs[2 * ldz + ic], ai = zs[3 * ldz + ic], z3 = zs[4 * ldz + ic];
  double y0 = 0, y1 = 0, y2 = 0, bi = 0, y3 = 0;
  if constexpr (TWO) { y0 = zs2[ic]; y1 = zs2[ldz + ic]; y2 = zs2[2 * ldz + ic]; bi = zs2[3 * ldz + ic]; y3 = zs2[4 * ldz + ic]; }
  double acc = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    const int j = j0 + tid;
    if (j < n) {
      cz[0][tid] = zs[j]; cz[1][tid] = zs[ldz + j]; cz[2][tid] = zs[2 * ldz + j]; cz[3][tid] = zs[3 * ldz + j];
      cz[4][tid] = alpha[j]; cz[5][tid] = zs[4 * ldz + j];
      if constexpr (TWO) { cy[0][tid] = zs2[j]; cy[1][tid] = zs2[ldz + j]; cy[2][tid] = zs2[2 * ldz + j]; cy[3][tid] = zs2[3 * ldz + j]; cy[4][tid] = zs2[4 * ldz + j]; }
    } else {
      cz[0][tid] = cz[1][tid] = cz[2][tid] = cz[3][tid] = 0; cz[4][tid] = 0; cz[5][tid] = 0;
      if constexpr (TWO) { cy[0][tid] = cy[1][tid] = cy[2][tid] = cy[3][tid] = cy[4][tid] = 0; }
    }
    __syncthreads();
#pragma unroll 4
    for (int q = pp; q < 256; q += 4) {
      const double d2 = pair_d2(z0, z1, z2, ai, cz[0][q], cz[1][q], cz[2][q], cz[3][q], z3, cz[5][q]);
      if constexpr (TWO) {
        const double d2b = pair_d2(y0, y1, y2, bi, cy[0][q], cy[1][q], cy[2][q], cy[3][q], y3, cy[4][q]);
        acc = fma(kern_val2(d2, P, d2b, P2), cz[4][q], acc);
      } else {
        acc = fma(kern_val(d2, P), cz[4][q], acc);
      }
    }
  }
  part[pp][rr] = acc;
  __syncthreads();
  if (pp == 0 && vi) {
    double fi = (part[0][rr] + part[1][rr]) + (part[2][rr] + part[3][rr]);
    if (P.white != 0.0) fi = fma(P.white, alpha[i], fi);        // the White members' diagonal
    f[i] = fi;
  }
}

// block-wide sum of NV values per thread -> out[blockIdx][NV]; 256 threads
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* out)
{
  __shared__ double red[8][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double s = v[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// scalar pieces of the objective (GP_Utils.cpp:1147-1159, 810, 858-862):
//   out[0] = sum alpha_i * 0.5 f_i ;  out[1] = sum lp_i ;  out[2] = sum ((y_i-f_i)^2/sn2 - 1)
// single CTA, fixed summation tree -> bitwise reproducible.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lml_terms_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                        const double* __restrict__ f, int n, const DevParams* __restrict__ Pp,
                                                        const double* __restrict__ logdet_parts, int nblk, double* __restrict__ out)
{
  const DevParams P = *Pp;
  double v[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += 256) {
    const double ymmu = y[i] - f[i];
    v[0] = fma(alpha[i], 0.5 * f[i], v[0]);
    v[1] += ymmu * ymmu * (-1.0 / (2.0 * P.sn2)) - P.lp_const;
    v[2] += P.inv_sn2 * (ymmu * ymmu) - 1.0;
  }
  for (int b = threadIdx.x; b < nblk; b += 256) v[3] += logdet_parts[b];
  block_reduce_store<4>(v, out);
}

// ---------------------------------------------------------------------------------------------------
// K7: one pass over the lower triangle of Q = B^-1 accumulating every reduction GradLL needs
//     (GP_Utils.cpp:1164-1169, 1206, 1222-1235; Kernel.cpp:1176-1242, 370-377):
//       QW_ij = Q_ij*d2lp - alpha_i alpha_j
//       w_ij  = var2*QW_ij*exp(-s_ij)*(-0.5/s_ij)   (0 on the diagonal and where s_ij == 0)
//       T_kl  = sum_ij w_ij x_ik x_jl ; V_k = sum_ij w_ij x_ik^2 ; G6 = sum_ij QW_ij exp(-s_ij)
//       TR    = sum_i QW_ii ; QK = sum_ij Q_ij K_ij
//       RK    = sum_{i>j} e^{-s_ij} (x_i3 - x_j3)^2   (4-column branch only: g[7], Kernel.cpp:1246-1255 -- no QW, see combine_gradient)
//     partial[block][13] = {T00,T01,T02,T11,T12,T22, V0,V1,V2, G6, TR, QK, RK}; combined on the host into g[0..9].
// ---------------------------------------------------------------------------------------------------
constexpr int NGRAD = 13;

// TWO: the Hyb sum holds a second distance-based member.  The pass is then run once per member (zs / Pp = the member whose entries are
// being accumulated, zs2 / P2p = the other one): the ExpAns sums use the member's OWN distance (getGradients recomputes DD2 itself,
// Kernel.cpp:925), the Exp / RBF sums use the D2 argument -- which HybKerns::computeK has SUMMED over the members (Kernel.cpp:140-154,
// 646-695, 491-541: reproduced).  common != 0 (first member's pass): also tr QW and sum Q o K with the full covariance.
template <bool TWO = false>
__global__ void __launch_bounds__(256) grad_pass_kernel(const double* __restrict__ Qm, long ld, const double* __restrict__ zs, long ldz,
                                                        const double* __restrict__ xs, long ldx, const double* __restrict__ alpha,
                                                        int n, const DevParams* __restrict__ Pp, double* __restrict__ partial, int tm0,
                                                        int tn0, int rmap_P, int rmap_me, int rmap_w, int common = 1,
                                                        const double* __restrict__ zs2 = nullptr, const DevParams* __restrict__ P2p = nullptr)
{
  // Qm points at the FIRST tile handed to this launch: tile (blockIdx.x, blockIdx.y) of the buffer is global tile (tm, tn) with
  //   tm = tm0 + blockIdx.x                      rows of a contiguous slice (single GPU, replicated layout), or, rmap_P > 1,
  //   tm = ((l / w) P + me) w + l % w, l = tm0 + blockIdx.x     packed block rows owned cyclically (partitioned storage)
  //   tn = tn0 + blockIdx.y
  int tm = tm0 + blockIdx.x;
  if (rmap_P > 1) tm = ((tm / rmap_w) * rmap_P + rmap_me) * rmap_w + tm % rmap_w;
  const int tn = tn0 + blockIdx.y;
  double* out = partial + ((long)blockIdx.y * gridDim.x + blockIdx.x) * NGRAD;
  if (tn > tm) { if (threadIdx.x < NGRAD) out[threadIdx.x] = 0.0; return; }
  __shared__ double cz[NZ][NB], cx[NX][NB], ca[NB];
  __shared__ double cy[TWO ? NZ : 1][TWO ? NB : 1];
  __shared__ DevParams P, P2;
  const int tid = threadIdx.x;
  if (tid == 0) { P = *Pp; if constexpr (TWO) P2 = *P2p; }
  const int r0 = tm * NB, c0 = tn * NB;
  for (int idx = tid; idx < NB; idx += 256) {
    const int j = c0 + idx;
#pragma unroll
    for (int q = 0; q < NZ; q++) {
      cz[q][idx] = zs[(long)q * ldz + j];
      if constexpr (TWO) cy[q][idx] = zs2[(long)q * ldz + j];
    }
#pragma unroll
    for (int q = 0; q < NX; q++) cx[q][idx] = xs[(long)q * ldx + j];
    ca[idx] = alpha[j];
  }
  const int tx = tid & 63, ty = tid >> 6;
  const int i0 = r0 + 2 * tx;
  double zi[2][NZ], xi[2][NX], al[2], yi[2][TWO ? NZ : 1];
#pragma unroll
  for (int e = 0; e < 2; e++) {
#pragma unroll
    for (int q = 0; q < NZ; q++) {
      zi[e][q] = zs[(long)q * ldz + i0 + e];
      if constexpr (TWO) yi[e][q] = zs2[(long)q * ldz + i0 + e];
    }
#pragma unroll
    for (int q = 0; q < NX; q++) xi[e][q] = xs[(long)q * ldx + i0 + e];
    al[e] = alpha[i0 + e];
  }
  __syncthreads();
  double om[2] = {0, 0}, u[2][3] = {{0, 0, 0}, {0, 0, 0}}, pq[2][3] = {{0, 0, 0}, {0, 0, 0}};
  double g6 = 0, tr = 0, qk = 0, rk = 0;
  for (int jj = ty; jj < NB; jj += 4) {
    const int j = c0 + jj;
    const double2 qv = *reinterpret_cast<const double2*>(Qm + (long)(blockIdx.y * NB + jj) * ld + (blockIdx.x * NB + 2 * tx));
    const double qe[2] = {qv.x, qv.y};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int i = i0 + e;
      if (i < n && j < n && i >= j) {
        const double d2own = pair_d2(zi[e][0], zi[e][1], zi[e][2], zi[e][3], cz[0][jj], cz[1][jj], cz[2][jj], cz[3][jj], zi[e][4], cz[4][jj]);
        double d2 = d2own, Ktot = 0.0;
        if constexpr (TWO) {
          const double d2o = pair_d2(yi[e][0], yi[e][1], yi[e][2], yi[e][3], cy[0][jj], cy[1][jj], cy[2][jj], cy[3][jj], yi[e][4], cy[4][jj]);
          if (P.kind != 0) d2 = __dadd_rn(d2own, d2o);                       // Exp / RBF: the D2 argument is the members' sum
          if (common) Ktot = kern_val2(d2own, P, d2o, P2);
        }
        const double QWij = qe[e] * P.inv_sn2 - al[e] * ca[jj];
        if (P.kind != 0) {
          // isotropic kernels: two pair sums each (slots G6 and RK), see combine_gradient_iso
          //   Exp (Kernel.cpp:646-695):  A = sum QW e^{-2s} (all pairs),  B = sum_{i != j} QW (e^{-s} (-0.5/s)) D2  (NaN for duplicates, as there)
          //   RBF (Kernel.cpp:491-541):  A = sum QW KD2 D2,               B = sum QW KD2,  KD2 = exp(-0.5 w D2)
          const double mult = (i == j) ? 1.0 : 2.0;
          double kd, a, b;
          if (P.kind == 1) {
            const double s = sqrt(d2);
            kd = exp(-s);
            a = QWij * (kd * kd);
            b = (i == j) ? 0.0 : QWij * ((kd * (-0.5 / s)) * d2);
          } else {
            kd = exp(__dmul_rn(P.rbf_c, d2));
            a = (QWij * kd) * d2;
            b = QWij * kd;
          }
          g6 += mult * a;
          rk += mult * b;
          if constexpr (TWO) {
            if (common) {
              if (i == j) tr += QWij;
              qk += mult * (qe[e] * ((i == j) ? __dadd_rn(Ktot, P.white) : Ktot));
            }
          } else {
            if (i == j) tr += QWij;
            qk += mult * (qe[e] * ((i == j) ? __dadd_rn(kern_val(d2, P), P.white) : kern_val(d2, P)));
          }
          continue;
        }
        const double s = sqrt(d2);
        const double es = exp(-s);
        const double Kij = TWO ? Ktot : __dadd_rn(__dmul_rn(P.var2, es), P.bias);
        const bool cm = !TWO || common;
        if (i == j) {
          g6 += QWij * es;
          if (cm) { tr += QWij; qk += qe[e] * __dadd_rn(Kij, P.white); }
        } else {
          g6 += 2.0 * (QWij * es);
          if (cm) qk += 2.0 * (qe[e] * Kij);
          const double dr = xi[e][3] - cx[3][jj];
          rk = fma(es, dr * dr, rk);
          if (s != 0.0) {
            const double w = (P.var2 * QWij) * (es * (-0.5 / s));
            om[e] += w;
#pragma unroll
            for (int q = 0; q < 3; q++) {
              u[e][q] = fma(w, cx[q][jj], u[e][q]);
              pq[e][q] = fma(w, cx[q][jj] * cx[q][jj], pq[e][q]);
            }
          }
        }
      }
    }
  }
  // ordered-pair sums from the unordered (i>j) traversal:
  //   V_k  = sum_{i>j} w (x_ik^2 + x_jk^2) ;  T_kl = sum_{i>j} w (x_ik x_jl + x_jk x_il)
  double v[NGRAD];
#pragma unroll
  for (int q = 0; q < NGRAD; q++) v[q] = 0;
#pragma unroll
  for (int e = 0; e < 2; e++) {
    v[0] += 2.0 * xi[e][0] * u[e][0];
    v[1] += xi[e][0] * u[e][1] + xi[e][1] * u[e][0];
    v[2] += xi[e][0] * u[e][2] + xi[e][2] * u[e][0];
    v[3] += 2.0 * xi[e][1] * u[e][1];
    v[4] += xi[e][1] * u[e][2] + xi[e][2] * u[e][1];
    v[5] += 2.0 * xi[e][2] * u[e][2];
#pragma unroll
    for (int q = 0; q < 3; q++) v[6 + q] += xi[e][q] * xi[e][q] * om[e] + pq[e][q];
  }
  v[9] = g6; v[10] = tr; v[11] = qk; v[12] = rk;
  block_reduce_store<NGRAD>(v, out);
}

// deterministic final sum of per-CTA partials: out[q] = sum_b partial[b][q]
template <int NV>
__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, long nblocks, double* __restrict__ out)
{
  double v[NV];
#pragma unroll
  for (int q = 0; q < NV; q++) v[q] = 0;
  for (long b = threadIdx.x; b < nblocks; b += 256)
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] += partial[b * NV + q];
  block_reduce_store<NV>(v, out);
}

// ---------------------------------------------------------------------------------------------------
// small data-movement helpers
// ---------------------------------------------------------------------------------------------------
// dst(128x128 tile, upper) = transpose of Winv block (lower); dst is column-major with leading dimension ld
__global__ void __launch_bounds__(256) put_transposed_block_kernel(double* __restrict__ dst, long ld, const double* __restrict__ Winv)
{
  __shared__ double t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) t[r][tx] = Winv[(long)(by + r) * NB + bx + tx];   // Winv(row bx+tx, col by+r)
  __syncthreads();
  for (int r = ty; r < 32; r += 8) dst[(long)(bx + r) * ld + by + tx] = t[tx][r];     // dst(row by+tx, col bx+r) = Winv(bx+r, by+tx)
}

// out(r x c region) = in^T : out(i,j) = in(j,i); generic tiled transpose, dims multiples of 32
__global__ void __launch_bounds__(256) transpose_kernel(double* __restrict__ out, long ldo, const double* __restrict__ in, long ldi,
                                                        int upper_src_only)
{
  __shared__ double t[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32;     // block of the SOURCE: rows bi.., cols bj..
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool zero = upper_src_only && (bi > bj + 31);       // strictly below the diagonal of an upper-triangular source
  for (int r = ty; r < 32; r += 8) t[r][tx] = zero ? 0.0 : in[(long)(bj + r) * ldi + bi + tx];   // in(bi+tx, bj+r)
  __syncthreads();
  for (int r = ty; r < 32; r += 8) out[(long)(bi + r) * ldo + bj + tx] = t[tx][r];               // out(bj+tx, bi+r) = in(bi+r, bj+tx)
}

// dst(i, j) = sum_s parts[s][i + j*rows]  (s ascending: a fixed order, so the result is reproducible run to run)
__global__ void __launch_bounds__(256) split_sum_kernel(double* __restrict__ dst, long ld, const double* __restrict__ parts, long rows,
                                                        long cols, int S)
{
  const long total = rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    double v = parts[idx];
    for (int s = 1; s < S; s++) v += parts[(long)s * total + idx];
    dst[(idx / rows) * ld + (idx % rows)] = v;
  }
}

// Partitioned storage: rows [r0, r0 + h) of this rank's first `cnt` packed block columns (w wide each) -> dst as cnt contiguous
// h x w blocks (the rank's piece of a ROW strip of L), and the inverse permutation that lays the all-gathered pieces
// (P segments of cmax blocks, owner-major) out in global column order: strip(:, k w + c) = seg[k % P][k / P](:, c).
__global__ void pack_rowstrip_kernel(double* __restrict__ dst, const double* __restrict__ Lloc, long ld, int r0, int h, int w, int cnt)
{
  const long total = (long)cnt * w * h;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % h);
    const long col = idx / h;                       // local column 0 .. cnt*w-1
    dst[idx] = Lloc[col * ld + r0 + r];
  }
}
__global__ void order_rowstrip_kernel(double* __restrict__ strip, const double* __restrict__ gathered, int h, int w, int nblocks, int P,
                                      long seg_stride)
{
  const long blk = (long)h * w, total = (long)nblocks * blk;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int k = (int)(idx / blk);
    const long within = idx % blk;
    strip[idx] = gathered[(long)(k % P) * seg_stride + (long)(k / P) * blk + within];
  }
}
// dst(i, j) = src(i, j) for an rows x cols block, both with their own leading dimension
__global__ void copy2d_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src, long lds, long rows, long cols)
{
  const long total = rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    dst[(idx / rows) * ldd + (idx % rows)] = src[(idx / rows) * lds + (idx % rows)];
}

// dst (rows x cols, contiguous) <- src (leading dimension ld), and back: staging of strided sub-matrices for NCCL
__global__ void pack_kernel(double* __restrict__ dst, const double* __restrict__ src, long ld, long rows, long cols)
{
  const long total = rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    dst[idx] = src[(idx / rows) * ld + (idx % rows)];
}
__global__ void unpack_kernel(double* __restrict__ dst, long ld, const double* __restrict__ src, long rows, long cols)
{
  const long total = rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    dst[(idx / rows) * ld + (idx % rows)] = src[idx];
}

// Row slice [r0, r0 + rows) of the upper-triangular U = L^-T, WITHOUT what every rank already holds: a rank needs from another rank's
// slice only the entries strictly above the w-wide diagonal blocks (the diagonal blocks themselves are computed redundantly everywhere,
// everything below them is zero).  Column j (global) contributes its rows r0 .. min(r0 + rows, w floor(j / w)) - 1, columns packed one
// after the other.  One CTA per column (grid-stride); dir = 0: pack (dst = contiguous buffer), 1: unpack (src = buffer).
__host__ __device__ __forceinline__ long uslice_offset(int j, int r0, int rows, int w)
{
  // sum over the columns before j of clamp(w floor(j' / w) - r0, 0, rows): whole block columns first, then the columns of j's own block
  long off = 0;
  const int J = j / w;
  for (int b = 0; b < J; b++) {
    int len = b * w - r0;
    len = len < 0 ? 0 : (len > rows ? rows : len);
    off += (long)len * w;
  }
  int len = J * w - r0;
  len = len < 0 ? 0 : (len > rows ? rows : len);
  return off + (long)len * (j - J * w);
}
__global__ void __launch_bounds__(256) uslice_copy_kernel(double* __restrict__ buf, double* __restrict__ U, long ld, int n_pad, int r0, int rows,
                                                          int w, int dir)
{
  for (int j = r0 + blockIdx.x; j < n_pad; j += gridDim.x) {
    int len = (j / w) * w - r0;
    len = len < 0 ? 0 : (len > rows ? rows : len);
    if (len == 0) continue;
    double* b = buf + uslice_offset(j, r0, rows, w);
    double* u = U + (long)j * ld + r0;
    if (dir == 0) for (int i = threadIdx.x; i < len; i += blockDim.x) b[i] = u[i];
    else          for (int i = threadIdx.x; i < len; i += blockDim.x) u[i] = b[i];
  }
}

__global__ void fill_kernel(double* __restrict__ p, long count, double v)
{
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

// zero a rectangular region of a column-major matrix
__global__ void zero_region_kernel(double* __restrict__ p, long ld, int rows, int cols)
{
  const long total = (long)rows * cols;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    p[(idx / rows) * ld + (idx % rows)] = 0.0;
}

// ---------------------------------------------------------------------------------------------------
// K2/K8: scaled cross-covariance tile  Bm(j,i) = Sw * K(x*_j, x_i)  (test index j contiguous) with the
//        predictive-mean partial sums fused in:  mu_part[tile_i][j] = sum_{i in tile} alpha_i K(x*_j,x_i)
//        (GP_Utils.cpp:943-949, 958-972, 985-990).  grid = (m_pad/128, n_pad/128), 256 threads.
// ---------------------------------------------------------------------------------------------------
template <bool TWO = false>
__global__ void __launch_bounds__(256) cross_build_kernel(double* __restrict__ Bm, long ldb, const double* __restrict__ zt, long ldzt,
                                                          const double* __restrict__ zs, long ldz, const double* __restrict__ alpha,
                                                          int m, int n, const DevParams* __restrict__ Pp,
                                                          double* __restrict__ mu_part, long ldmu, int write_B, double white_x, long goff,
                                                          const double* __restrict__ zt2 = nullptr, const double* __restrict__ zs2 = nullptr,
                                                          const DevParams* __restrict__ P2p = nullptr)
{
  // white_x != 0: Kern_White::computeK's condition held for the whole call (X1(0) == X2(0) and equally many rows, Kernel.cpp:261-262),
  // so element (i, i) of the n x m cross-covariance carries Sigma_White; goff = global test index of this batch's column 0.
  __shared__ double cz[NZ][NB], ca[NB];
  __shared__ double cy[TWO ? NZ : 1][TWO ? NB : 1];
  __shared__ double mred[4][NB];
  __shared__ DevParams P, P2;
  const int tid = threadIdx.x;
  if (tid == 0) { P = *Pp; if constexpr (TWO) P2 = *P2p; }
  const int j0t = blockIdx.x * NB, i0t = blockIdx.y * NB;
  for (int idx = tid; idx < NB; idx += 256) {
    const int i = i0t + idx;
#pragma unroll
    for (int q = 0; q < NZ; q++) {
      cz[q][idx] = zs[(long)q * ldz + i];
      if constexpr (TWO) cy[q][idx] = zs2[(long)q * ldz + i];
    }
    ca[idx] = (i < n) ? alpha[i] : 0.0;
  }
  const int tx = tid & 63, ty = tid >> 6;
  const int j0 = j0t + 2 * tx;
  double zj[2][NZ], yj[2][TWO ? NZ : 1];
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int q = 0; q < NZ; q++) {
      zj[e][q] = zt[(long)q * ldzt + j0 + e];
      if constexpr (TWO) yj[e][q] = zt2[(long)q * ldzt + j0 + e];
    }
  __syncthreads();
  double mu[2] = {0, 0};
  for (int ii = ty; ii < NB; ii += 4) {
    const int i = i0t + ii;
    double2 out;
    double* o = &out.x;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double v = 0.0;
      if (i < n && (j0 + e) < m) {
        // K(X_train, X_test)(i,j): first argument is the training point (GP_Utils.cpp:946-947)
        const double d2 = pair_d2(cz[0][ii], cz[1][ii], cz[2][ii], cz[3][ii], zj[e][0], zj[e][1], zj[e][2], zj[e][3], cz[4][ii], zj[e][4]);
        double k;
        if constexpr (TWO) {
          const double d2b = pair_d2(cy[0][ii], cy[1][ii], cy[2][ii], cy[3][ii], yj[e][0], yj[e][1], yj[e][2], yj[e][3], cy[4][ii], yj[e][4]);
          k = kern_val2(d2, P, d2b, P2);
        } else {
          k = kern_val(d2, P);
        }
        if (white_x != 0.0 && (long)i == goff + j0 + e) k = __dadd_rn(k, white_x);
        mu[e] = fma(ca[ii], k, mu[e]);
        v = __dmul_rn(k, P.sw);
      }
      o[e] = v;
    }
    if (write_B) *reinterpret_cast<double2*>(Bm + (long)i * ldb + j0) = out;
  }
  mred[ty][2 * tx] = mu[0];
  mred[ty][2 * tx + 1] = mu[1];
  __syncthreads();
  if (tid < NB) mu_part[(long)blockIdx.y * ldmu + j0t + tid] = (mred[0][tid] + mred[1][tid]) + (mred[2][tid] + mred[3][tid]);
}

// mu[j] = sum_t mu_part[t][j]  (fixed order)
__global__ void mean_finish_kernel(const double* __restrict__ mu_part, long ldmu, int ntiles, int m, double* __restrict__ mu)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s = 0;
  for (int t = 0; t < ntiles; t++) s += mu_part[(long)t * ldmu + j];
  mu[j] = s;
}

// raw predictive variance  var[j] = kD - sum_i V(i,j)^2  (GP_Utils.cpp:997-998); V is n_pad x m_pad, one warp per column.
// The reference's post-processing of this vector (GP_Utils.cpp:1001-1003 and 1033-1040) is index arithmetic over the
// WHOLE test set and is applied on the host (gpss_capi.cu: apply_reference_var_postprocessing).
__global__ void __launch_bounds__(256) var_finish_kernel(const double* __restrict__ V, long ldv, int n_pad, int m, double kD,
                                                         double* __restrict__ var)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 8 + warp;
  if (j >= m) return;
  const double* col = V + (long)j * ldv;
  double s0 = 0, s1 = 0;
  for (int i = lane * 2; i < n_pad; i += 64) {
    const double2 v = *reinterpret_cast<const double2*>(col + i);
    s0 = fma(v.x, v.x, s0);
    s1 = fma(v.y, v.y, s1);
  }
  double s = s0 + s1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) var[j] = kD - s;
}

// out[j] = sum_k V(j, k)^2 over `cols` columns of an m x cols column-major matrix (row index contiguous): one thread per row, fixed order
__global__ void __launch_bounds__(256) rowsumsq_kernel(const double* __restrict__ V, long ld, int m, long cols, double* __restrict__ out)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s0 = 0, s1 = 0;
  long k = 0;
  for (; k + 1 < cols; k += 2) {
    const double a = V[j + k * ld], b = V[j + (k + 1) * ld];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (k < cols) { const double a = V[j + k * ld]; s0 = fma(a, a, s0); }
  out[j] = s0 + s1;
}

}  // namespace gpss
