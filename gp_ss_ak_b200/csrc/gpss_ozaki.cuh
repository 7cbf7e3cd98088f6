// gpss_ozaki.cuh -- the DEFAULT pipe for 8192 < n_pad <= 57 344 (GPSS_OZAKI=0 restores DMMA everywhere, GPSS_OZAKI=6|7|8 forces
// it for every n; size rule in gpss_create): the three long-k FP64 contractions of the
// path (look-ahead update of the Cholesky, bulk product of the triangular inverse, B^-1 = U U^T) evaluated on the int8
// tensor cores of sm_100a -- tcgen05.mma kind::i8, operands by TMA, int32 accumulators in TMEM -- with the Ozaki splitting,
// instead of the 37 TFLOP/s DMMA pipe.  Measured on B200 (profiles/r01_ozaki_int8_gemm_microbench.txt, 16384^2 x 8192):
// 83.7 FP64-equivalent TFLOP/s with S = 7 slices, 69.0 with S = 8, against 35.9 for gemm_nt_ws_kernel; error of S = 8
// against a long-double reference 1.3e-15 (DMMA: 1.4e-14).  Numerics of the whole path: scripts/ozaki_numerics.py.
//
// Digits: 7 slices of 8 bits by default (base-256 digits in [-128, 127]: the 55 bits that take 8 slices of 7 bits, 28 products instead
// of 36; the k-range of one int32 accumulation shrinks to 18 688, see kseg).  The formulas below are written for 7-bit digits
// (GPSS_OZAKI_BITS=7); with b-bit digits read 2^b for 128 and 2^(b-1) for 64.
//   operand x -> t = x / 2^e (ONE a-priori exponent per operand kind, see oz_exponent), v = rint(t 2^(7S-1)),
//   signed base-128 digits d_0 .. d_{S-1} in [-64, 64]:  t = sum_p d_p 2^-(7p+6)          (oz_slice_kernel, K-major int8 planes)
//   A B^T = 2^(eA+eB-12) sum_g 2^(-7g) G_g,  G_g = sum_{i+j=g} A_i B_j^T  exact in int32 (|G_g| <= (g+1) k 2^12 < 2^31 for
//   k <= 65 536 at S = 8); pairs with i + j >= S are dropped (below the rounding of v).
//
// oz_gemm_kernel<S>: one CTA per 128 x 64 tile of C, 192 threads.
//   warp 4   : TMA producer -- per 64-byte k-chunk all S planes of the A tile (128 rows) and of the B tile (64 rows) as
//              SWIZZLE_64B boxes, one mbarrier per stage (2 stages at S = 7 / 8);
//   warp 5   : allocates TMEM (S x 64 columns), issues S (S + 1) / 2 x 2 MMAs (M 128, N 64, K 32) per chunk; group g
//              accumulates in TMEM columns [64 g, 64 g + 64); tcgen05.commit releases the stage;
//   warps 0-3: epilogue -- tcgen05.ld of the S group accumulators, Horner in FP64, C written / updated (TMEM lane = row
//              of C, so a warp touches 32 consecutive doubles of a column).
// Every plane of a chunk is loaded ONCE and feeds up to S products.
#pragma once
#include "gpss_gemm.cuh"          // mbarrier helpers
#include "gpss_kernels.cuh"       // DevParams
#include <cuda.h>
#include <cstdint>

namespace oz {

constexpr int BM = 128, BN = 64, BK = 64;        // BK: granularity of every k-range in bytes (= int8 elements); a stage holds BKB = 64 or 32 of them
constexpr int UMMA_K = 32;                       // kind::i8: 32 bytes of k per instruction
constexpr int DIGIT_BITS = 7;
constexpr int RASTER_W = 8;

// operand kinds: |x| <= 1 (U = L^-T and W = L^-1: B >= I), |x| <= sqrt(max B_ii) (L), |x| <= Sw (Sigma^2 + Sigma_Bias) (the scaled
// cross-covariance tile of the prediction)
enum { SCALE_UNIT = 0, SCALE_CHOL = 1, SCALE_CROSS = 2 };
enum { MASK_NONE = 0, MASK_LOWER = 1, MASK_UPPER = 2 };

// Kernel variants (oz::variant(), GPSS_OZ_VARIANT; A/B-measured in profiles/r02_oz_gemm_variants.txt):
//   0: 64-byte stages (SWIZZLE_64B), one MMA of N = 64 per slice pair                       -- the round-1 kernel
//   1: 64-byte stages, MERGED MMAs: the B planes of a stage are contiguous in shared memory (64 rows each) and group i + j sits in
//      TMEM columns [64 (i + j), +64), so A_i x [B_j .. B_j+c-1] is ONE instruction of N = 64 c <= 256 whose accumulator columns
//      are exactly the groups i + j .. i + j + c - 1: 12 instructions instead of 36 per 32 bytes of k at S = 8, and the A tile is
//      read from shared memory 12 times instead of 36 (the round-1 kernel ran at the 128 B/clk shared-memory limit:
//      36 x (4 KB + 2 KB) per 32 clk of tensor work)
//   2: variant 1 with 32-byte stages (SWIZZLE_32B): twice the stages in the same shared memory, finer refill granularity
//   3: variant 1 with PER-PLANE barriers: a stage is S units {A_i, B_(S-1-i)} of 12 KB; step i of a chunk (A_i against B_0 .. B_(S-1-i))
//      is the last reader of unit i, so the MMA warp hands each unit back right after its step and the producer refills it at once --
//      the refill of a 2-stage ring starts (S - i) / S of a chunk earlier than with one barrier per stage.  With 7 planes a chunk is
//      1792 clk of tensor work against ~2300 clk to land an 84 KB stage: the whole-stage ring left the pipe waiting (0.82 of peak).
enum { VAR_PAIR64 = 0, VAR_MERGE64 = 1, VAR_MERGE32 = 2, VAR_UNIT64 = 3, VAR_DEFAULT = VAR_MERGE64 };

template <int S, int BKB = BK>
struct Cfg {
  static constexpr int PAIRS = S * (S + 1) / 2;
  static constexpr int A_BYTES = BM * BKB, B_BYTES = BN * BKB;
  static constexpr int STAGE_BYTES = S * (A_BYTES + B_BYTES);
  static constexpr int STAGES_FIT = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int TMEM_COLS = S * BN <= 64 ? 64 : S * BN <= 128 ? 128 : S * BN <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 1024 /* barriers: 2 x STAGES x S + 1 */;
  static_assert((2 * STAGES * S + 3) * 8 <= 1024, "barrier area");
  static_assert(S >= 1 && S * BN <= 512, "S group accumulators of BN columns must fit the 512 TMEM columns");
  static_assert(STAGES >= 2, "need at least two stages");
};

struct Args {
  double* C; long ldc;           // C(i,j) at C[i + j*ldc]
  int m, n;                      // tile grid: m % 128 == 0, n % 64 == 0
  int a_rows, b_rows;            // rows per plane of the plane tensors (plane p, row r -> TMA row p * rows + r)
  int a_row0, b_row0;            // plane row of C(0,.) in the A planes / of C(.,0) in the B planes
  int k0, k1;                    // k range [k0, k1), multiples of 64 (plane byte columns)
  int kbeg_row;                  // 1: k starts at the tile's first A row (A upper-triangular in (row, k)); a_row0 is then a k coordinate too
  int kend_row;                  // 1: k ends after the tile's last A row (A lower-triangular in (row, k))
  int rev_order;                 // 1: last tile row first (kend_row products: longest k-ranges first)
  int lower_only;                // 1: skip tiles entirely above the diagonal of the global matrix (C(0,0) sits at (grow0, gcol0))
  int grow0, gcol0;
  int accumulate;                // 0: C = sign * A B^T, 1: C += sign * A B^T
  double sign;
  int a_kind, b_kind;            // SCALE_*: which power of two each operand was divided by before slicing
  int digit_bits;                // 0 or 7: base-128 digits in [-64, 64]; 8: base-256 digits in [-128, 127] (the default of the size rule)
  int kseg;                      // 0: the whole k-range is ONE int32 accumulation; > 0 (multiple of 64): the accumulators are drained into C at
                                 // every absolute multiple of kseg (8-bit digits: 18 688 at S = 7) and restarted -- exact int32 sums per
                                 // segment, FP64 sums across segments, all inside one launch
  const gpss::DevParams* dP;     // device parameters (theta-dependent scale: never a launch argument, so launches can sit in a graph)
  int32_t* dbg;                  // test hook: raw int32 group accumulators, [S][m][n] row-major (else nullptr)
};

// The exponent e with |x| < 2^e for every element of an operand of the given kind.  L: |L_ij| <= sqrt(B_ii),
// B_ii = 1 + Sw^2 (Sigma^2 + Sigma_Bias) for all i (exp(0) = 1 on the diagonal of every kernel kind).
// 8-bit digits: the signed base-256 expansion of v needs |v| < 127.5 x 256^(S-1) for its top digit to fit int8, so the bound is
// widened by 128 / 127.5 before its exponent is taken (one bit less for the unit-bound operands, rarely one for the others).
__device__ __forceinline__ int oz_exponent(int kind, const gpss::DevParams* P, int bits = DIGIT_BITS)
{
  if (bits != 8 && (kind == SCALE_UNIT || P == nullptr)) return 0;
  double bound = 1.0;
  if (kind == SCALE_CROSS && P) bound = P->sw * (P->var2 + P->var2x + P->bias + P->white);
  else if (kind == SCALE_CHOL && P) bound = sqrt(1.0 + P->sww * (P->var2 + P->var2x + P->bias + P->white));
  if (bits == 8) bound *= 128.0 / 127.5;
  int e;
  frexp(bound, &e);                                           // bound = f 2^e, f in [0.5, 1)
  return e;
}

// ------------------------------------------------------------------ PTX helpers (sm_100a)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               :: "r"(gpss::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(gpss::smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)       // arrives on bar when all MMAs issued so far have completed
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(gpss::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major operand tile in shared memory, rows of BKB = 64 / 32 bytes, SWIZZLE_64B / SWIZZLE_32B (cute::UMMA::SmemDescriptor,
// mma_sm100_desc.hpp: start >> 4 in [0,14), LBO in [16,30) (ignored for swizzled K-major: 1), SBO = 8 rows x BKB bytes >> 4 in
// [32,46), version 1 in [46,48), layout SWIZZLE_64B = 4 / SWIZZLE_32B = 6 in [61,64)).  The tile base is 1024-byte aligned.
template <int BKB>
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t saddr)
{
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(8 * BKB / 16) << 32) | (1ull << 46) | ((BKB == 64 ? 4ull : 6ull) << 61);
}
// cute::UMMA::InstrDescriptor: c_format S32 = 2 [4,6), a/b_format INT8 = 1 [7,10) / [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t idesc_i8(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }
constexpr uint32_t IDESC_I8 = idesc_i8(BN);

// Registers: capped at OZ_MAXNREG so that the critical-path DMMA kernel (gemm_nt_ws_kernel, 2-stage ring: 168 x 160 threads, 51 KB) fits on
// an SM NEXT TO a resident CTA of this kernel (6 warps x 168 x 32 = 32 256 of 65 536 registers, 175 of 227 KB).  Uncapped, ptxas gave the
// epilogue's batched loads 222 registers: 43 008 + 26 880 > 65 536, i.e. every panel GEMM of the Cholesky waited for an SM to drain a bulk
// CTA (~1 ms long) -- 30-45 us per dependent launch in the GPSS_DIST_TRACE logs of round 2.
#ifndef OZ_MAXNREG
#define OZ_MAXNREG 168
#endif
template <int S, int BKB = BK, bool MERGE = false, bool UNIT = false>
__global__ void __maxnreg__(OZ_MAXNREG)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Args g)
{
  using T = Cfg<S, BKB>;
  static_assert(!UNIT || (MERGE && BKB == 64), "per-plane barriers come with the merged 64-byte kernel");
  constexpr int NBAR = UNIT ? T::STAGES * S : T::STAGES;        // full / empty barriers: per stage, or per stage and plane unit
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T::STAGES * T::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NBAR;
  uint64_t* acc_bar = empty_bar + NBAR;
  uint64_t* free_bar = acc_bar + 1;                               // epilogue -> MMA warp: the accumulators of a k-segment have been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rasterisation as in gemm_nt_ws_kernel: super-columns of RASTER_W tile columns, tile column fastest, so that one wave of
  // 148 CTAs shares ~19 A row-tiles x 8 B row-tiles through L2 instead of streaming every A plane once per tile column
  const int mt = g.m / BM, nt = g.n / BN;
  const int grp = blockIdx.x / (RASTER_W * mt), within = blockIdx.x % (RASTER_W * mt);
  const int gcols = min(RASTER_W, nt - grp * RASTER_W);
  const int tile_n = grp * RASTER_W + within % gcols, tile_m = g.rev_order ? mt - 1 - within / gcols : within / gcols;
  if (g.lower_only && g.gcol0 + tile_n * BN > g.grow0 + tile_m * BM + BM - 1) return;      // whole CTA, before any barrier
  int kb = g.k0;
  if (g.kbeg_row) { const int kr = (g.a_row0 + tile_m * BM) & ~(BK - 1); if (kr > kb) kb = kr; }
  int ke = g.k1;
  if (g.kend_row) { const int kr = g.a_row0 + tile_m * BM + BM; if (kr < ke) ke = kr; }
  const int nk = ke > kb ? (ke - kb) / BKB : 0;
  if (nk == 0 && g.accumulate) return;                                                      // nothing to add
  // k-segments (absolute multiples of kseg, so that every tile cuts at the same k whatever its own range): chunk kc is the last of its
  // segment when the next chunk starts on a boundary
  const int kseg = g.kseg > 0 ? g.kseg : (1 << 30);
  const int nseg = nk > 0 ? ((ke - 1) / kseg - kb / kseg + 1) : 1;
  auto seg_ends_after = [&](int kc) { return kc == nk - 1 || (kb + (kc + 1) * BKB) % kseg == 0; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NBAR; s++) { gpss::mbar_init(full_bar + s, 1); gpss::mbar_init(empty_bar + s, 1); }
    gpss::mbar_init(acc_bar, 1);
    gpss::mbar_init(free_bar, 4);                                 // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(gpss::smem_u32(tmem_slot)), "r"((uint32_t)T::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kc = 0; kc < nk; kc++) {
        const int st = kc % T::STAGES;
        const uint32_t ph = (uint32_t)(kc / T::STAGES) & 1u;
        uint8_t* sa = smem + st * T::STAGE_BYTES;
        uint8_t* sb = sa + S * T::A_BYTES;
        const int kq = kb + kc * BKB;
        if constexpr (UNIT) {
          // unit u = {A_u, B_(S-1-u)}: freed by step u of the chunk that used this stage last, refilled here in the same order
#pragma unroll
          for (int u = 0; u < S; u++) {
            uint64_t* fb = full_bar + st * S + u;
            gpss::mbar_wait(empty_bar + st * S + u, ph ^ 1u);
            gpss::mbar_arrive_expect_tx(fb, (uint32_t)(T::A_BYTES + T::B_BYTES));
            tma_load_2d(sa + u * T::A_BYTES, &tmA, fb, kq, u * g.a_rows + g.a_row0 + tile_m * BM);
            tma_load_2d(sb + (S - 1 - u) * T::B_BYTES, &tmB, fb, kq, (S - 1 - u) * g.b_rows + g.b_row0 + tile_n * BN);
          }
        } else {
          gpss::mbar_wait(empty_bar + st, ph ^ 1u);
          gpss::mbar_arrive_expect_tx(full_bar + st, (uint32_t)T::STAGE_BYTES);
#pragma unroll
          for (int p = 0; p < S; p++) {
            tma_load_2d(sa + p * T::A_BYTES, &tmA, full_bar + st, kq, p * g.a_rows + g.a_row0 + tile_m * BM);
            tma_load_2d(sb + p * T::B_BYTES, &tmB, full_bar + st, kq, p * g.b_rows + g.b_row0 + tile_n * BN);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && nk > 0) {
      bool seg_first = true;                 // the next chunk opens a k-segment: its first products overwrite the accumulators
      int seg = 0;
      for (int kc = 0; kc < nk; kc++) {
        const int st = kc % T::STAGES;
        const uint32_t ph = (uint32_t)(kc / T::STAGES) & 1u;
        const uint32_t sa = gpss::smem_u32(smem + st * T::STAGE_BYTES);
        const uint32_t sb = sa + S * T::A_BYTES;
        const bool opens = seg_first, closes = seg_ends_after(kc);
        if (opens && seg > 0) {              // the epilogue warps must have drained the previous segment out of TMEM
          gpss::mbar_wait(free_bar, (uint32_t)(seg - 1) & 1u);
          tc_fence_after();
        }
        seg_first = closes;
        if constexpr (UNIT) {
          // step 0 reads every B plane, i.e. every unit of the stage: all of them must have landed; step i is the last reader of unit i
#pragma unroll
          for (int u = 0; u < S; u++) gpss::mbar_wait(full_bar + st * S + u, ph);
          tc_fence_after();
#pragma unroll
          for (int i = 0; i < S; i++) {
#pragma unroll
            for (int ks = 0; ks < BKB / UMMA_K; ks++) {
              const uint64_t ad = smem_desc_k<BKB>(sa + i * T::A_BYTES + ks * UMMA_K);
              const uint32_t acc = (!opens || ks > 0 || i > 0) ? 1u : 0u;
              constexpr int MAXP = 256 / BN;
              const int cnt = S - i, nm = (cnt + MAXP - 1) / MAXP;
              int j0 = 0;
#pragma unroll
              for (int q = 0; q < nm; q++) {
                const int len = (cnt - j0 + (nm - q) - 1) / (nm - q);
                const uint64_t bd = smem_desc_k<BKB>(sb + j0 * T::B_BYTES + ks * UMMA_K);
                mma_i8(tmem_base + (uint32_t)((i + j0) * BN), ad, bd, idesc_i8(len * BN), acc);
                j0 += len;
              }
            }
            tc_commit(empty_bar + st * S + i);         // A_i and B_(S-1-i) have no reader left in this chunk
          }
          if (closes) { tc_commit(acc_bar); seg++; }   // the segment's accumulators are complete
          continue;
        }
        gpss::mbar_wait(full_bar + st, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < BKB / UMMA_K; ks++) {
#pragma unroll
          for (int i = 0; i < S; i++) {
            const uint64_t ad = smem_desc_k<BKB>(sa + i * T::A_BYTES + ks * UMMA_K);
            // group i + j: the first product that reaches it (i == 0 of the first k-step of a segment) overwrites, the rest accumulate
            const uint32_t acc = (!opens || ks > 0 || i > 0) ? 1u : 0u;
            if constexpr (MERGE) {
              // B planes 0 .. S-1-i are 64 (S - i) consecutive rows of shared memory and their groups i .. S-1 are 64 (S - i)
              // consecutive TMEM columns: ceil((S - i) / 4) instructions of N <= 256, planes split as evenly as possible
              constexpr int MAXP = 256 / BN;
              const int cnt = S - i, nm = (cnt + MAXP - 1) / MAXP;
              int j0 = 0;
#pragma unroll
              for (int q = 0; q < nm; q++) {
                const int len = (cnt - j0 + (nm - q) - 1) / (nm - q);
                const uint64_t bd = smem_desc_k<BKB>(sb + j0 * T::B_BYTES + ks * UMMA_K);
                mma_i8(tmem_base + (uint32_t)((i + j0) * BN), ad, bd, idesc_i8(len * BN), acc);
                j0 += len;
              }
            } else {
#pragma unroll
              for (int j = 0; j < S - i; j++) {
                const uint64_t bd = smem_desc_k<BKB>(sb + j * T::B_BYTES + ks * UMMA_K);
                mma_i8(tmem_base + (uint32_t)((i + j) * BN), ad, bd, IDESC_I8, acc);
              }
            }
          }
        }
        tc_commit(empty_bar + st);         // stage reusable once these MMAs have read it
        if (closes) { tc_commit(acc_bar); seg++; }     // the segment's accumulators are complete
      }
    }
  } else {
    // ------------------------------------------------ epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows of the tile
    const int row = tile_m * BM + warp * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int bits = g.digit_bits ? g.digit_bits : DIGIT_BITS;
    const double w = ldexp(1.0, -bits);
    const double scale = g.sign * ldexp(1.0, oz_exponent(g.a_kind, g.dP, bits) + oz_exponent(g.b_kind, g.dP, bits) - 2 * (bits - 1));
    double* const crow = g.C + row + (size_t)(tile_n * BN) * g.ldc;
    for (int seg = 0; seg < nseg; seg++) {
      if (nk > 0) {
        gpss::mbar_wait(acc_bar, (uint32_t)seg & 1u);
        tc_fence_after();
      }
      const bool rmw = g.accumulate || seg > 0;        // later segments add to what the first one wrote
      // C is read-modify-written in place: the 8 loads of a column group are issued together, BEFORE the TMEM read-back and the Horner
      // sums, and the next group's loads before this group's stores (one round trip to L2 / HBM per group instead of one per column --
      // the first version's 64 dependent load -> FMA -> store chains cost ~30 us per tile, profiles/r02_oz_gemm_variants.txt)
      double cin[8], cnx[8];
#pragma unroll
      for (int c = 0; c < 8; c++) cnx[c] = rmw ? __ldcg(crow + (size_t)c * g.ldc) : 0.0;
      for (int c0 = 0; c0 < BN; c0 += 8) {
        uint32_t v[S][8];
#pragma unroll
        for (int c = 0; c < 8; c++) cin[c] = cnx[c];
        if (rmw && c0 + 8 < BN) {
#pragma unroll
          for (int c = 0; c < 8; c++) cnx[c] = __ldcg(crow + (size_t)(c0 + 8 + c) * g.ldc);
        }
        if (nk > 0) {
#pragma unroll
          for (int gi = 0; gi < S; gi++) tmem_ld8(lane_base + (uint32_t)(gi * BN + c0), v[gi]);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int gi = 0; gi < S; gi++)
#pragma unroll
            for (int c = 0; c < 8; c++) v[gi][c] = 0u;
        }
        if (g.dbg) {
#pragma unroll
          for (int gi = 0; gi < S; gi++)
#pragma unroll
            for (int c = 0; c < 8; c++) g.dbg[((size_t)gi * g.m + row) * g.n + tile_n * BN + c0 + c] = (int32_t)v[gi][c];
        }
        double outv[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          double acc = 0.0;
#pragma unroll
          for (int gi = S - 1; gi >= 0; gi--) acc = acc * w + (double)(int32_t)v[gi][c];     // sum_g 2^(-7g) G_g, smallest first
          outv[c] = rmw ? (cin[c] + scale * acc) : (scale * acc);
        }
#pragma unroll
        for (int c = 0; c < 8; c++) crow[(size_t)(c0 + c) * g.ldc] = outv[c];
      }
      if (seg + 1 < nseg) {                            // hand the accumulators back: every lane's TMEM loads have completed (wait::ld above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) gpss::mbar_arrive(free_bar);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"((uint32_t)T::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ int8 tensor-pipe micro-peak (roofline denominator)
// tcgen05.mma kind::i8 (M 128, K 32, N = 64 | 128 | 256) issued back to back from operands that stay in shared memory -- no TMA, no
// epilogue -- one CTA per SM: what the pipe delivers for the instruction shapes oz_gemm_kernel uses (bench_micro/int8_peak.cu prints
// the table, gpss_measure_int8_peak times the N = 256 shape for bench.py; MEASURED_PEAKS.json has no int8 entry).
// shared memory: A tile 128 rows x 64 B (SWIZZLE_64B layout, 8 KB) + B tile 256 rows x 64 B (16 KB), filled with pseudo-random bytes
// (all-zero operands would flatter the clocks: tensor power is data dependent)
template <int N>
__global__ void __launch_bounds__(128, 1) int8_peak_kernel(int iters, unsigned long long* cycles)
{
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  unsigned s = 1234567u + blockIdx.x * 7919u + threadIdx.x;
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) { s = s * 1664525u + 1013904223u; smem[i] = (uint8_t)((int)((s >> 16) % 129u) - 64); }
  if (threadIdx.x == 0) { gpss::mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(gpss::smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // generic-proxy writes of the operands -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = gpss::smem_u32(smem), sb = sa + 8 * 1024;
    const long long t0 = clock64();
    constexpr int ACC = 512 / N;                  // independent accumulators of N columns (N = 192: 2)
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int q = 0; q < ACC; q++) {
#pragma unroll
        for (int ks = 0; ks < 2; ks++)
          mma_i8(tmem + (uint32_t)(q * N), smem_desc_k<64>(sa + ks * 32), smem_desc_k<64>(sb + ks * 32), idesc_i8(N), it > 0 || ks > 0);
      }
    }
    tc_commit(bar);
    gpss::mbar_wait(bar, 0);
    if (blockIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ slicing: FP64 column-major -> S K-major int8 planes
// X is addressed by GLOBAL (row, k): X[row + k * ldx].  The block [row0, row0 + rows) x [k0, k0 + kcnt) goes to
// planes[p][row * kpad + k] (plane p starts at p * plane_rows * kpad).  mask: MASK_LOWER keeps k <= row (L), MASK_UPPER keeps
// k >= row (U); everything else is written as 0, so the GEMM may run its k-ranges in whole 64-byte chunks over the
// triangle's edge.  32 x 32 tiles through shared memory: coalesced FP64 reads along rows, 32-byte row segments written.
template <int S>
__global__ void oz_slice_kernel(const double* __restrict__ X, long ldx, int row0, int rows, int k0, int kcnt, int kind, int mask,
                                const gpss::DevParams* dP, int8_t* __restrict__ planes, long plane_rows, long kpad, int bits, int* viol)
{
  __shared__ int8_t tile[S][32][33];
  const int r = row0 + blockIdx.x * 32 + threadIdx.x;
  const double lim = (double)(1ll << (bits * S - 1));
  const long long half = 1ll << (bits - 1), dmask = (1ll << bits) - 1;
  // base 256: [127, 127, ..., 127] is the largest value whose digits fit int8 (S <= 7); never reached given oz_exponent's margin
  const long long vmax = bits == 8 ? 127ll * (((1ll << (8 * (S < 8 ? S : 7))) - 1) / 255) : (1ll << (DIGIT_BITS * S - 1));
  const double mul = ldexp(lim, -oz_exponent(kind, dP, bits));         // exact power of two
  for (int ky = threadIdx.y; ky < 32; ky += blockDim.y) {
    const int kq = k0 + blockIdx.y * 32 + ky;
    int d[S];
#pragma unroll
    for (int p = 0; p < S; p++) d[p] = 0;
    const bool keep = (mask == MASK_NONE) || (mask == MASK_LOWER && kq <= r) || (mask == MASK_UPPER && kq >= r);
    if (r < row0 + rows && kq < k0 + kcnt && keep) {
      double sc = X[r + (size_t)kq * ldx] * mul;
      // |x| <= 2^e holds when B = I + K / sn2 >= I, i.e. for a positive semi-definite K; nothing constrains theta (Sigma_Bias enters
      // raw), so an operand that exceeds its a-priori bound by more than rounding noise (2^-40 relative) raises the flag the host reads
      // at the end of the evaluation, which is then repeated on the DMMA path (gpss_capi.cu: oz_blocked).  NaN raises it too.
      if (viol && !(fabs(sc) <= (double)vmax * (1.0 + 0x1p-40))) *viol = 1;      // vmax: the largest value the digits can hold (= lim with 7-bit digits)
      sc = fmin(fmax(sc, -lim), lim);                                  // rounding excess only
      long long v = __double2ll_rn(sc);
      v = v > vmax ? vmax : (v < -vmax ? -vmax : v);
#pragma unroll
      for (int p = S - 1; p >= 1; p--) {
        const int dg = (int)(((v + half) & dmask) - half);            // [-2^(b-1), 2^(b-1) - 1], exact remainder
        v = (v - dg) >> bits;
        d[p] = dg;
      }
      d[0] = (int)v;                                                   // |v| <= 2^(b-1) (b = 8: <= 127 by the clamp)
    }
#pragma unroll
    for (int p = 0; p < S; p++) tile[p][threadIdx.x][ky] = (int8_t)d[p];
  }
  __syncthreads();
  for (int ry = threadIdx.y; ry < 32; ry += blockDim.y) {
    const int rr = row0 + blockIdx.x * 32 + ry, kq = k0 + blockIdx.y * 32 + threadIdx.x;
    if (rr < row0 + rows && kq < k0 + kcnt) {
#pragma unroll
      for (int p = 0; p < S; p++) planes[((size_t)p * plane_rows + rr) * kpad + kq] = tile[p][ry][threadIdx.x];
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static inline EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// which oz_gemm_kernel variant runs (see the list above Cfg): GPSS_OZ_VARIANT = 0 | 1 | 2, read once
static inline int variant()
{
  static int v = -1;
  if (v < 0) {
    v = VAR_DEFAULT;
    if (const char* e = getenv("GPSS_OZ_VARIANT")) { const int x = atoi(e); if (x >= VAR_PAIR64 && x <= VAR_UNIT64) v = x; }
  }
  return v;
}
static inline int stage_k() { return variant() == VAR_MERGE32 ? 32 : 64; }

// planes: [total_rows][kpad] bytes, K contiguous; box = stage_k() bytes of k x box_rows rows, SWIZZLE_64B / _32B.  Returns 0 on success.
static inline int make_plane_map(CUtensorMap* tm, const int8_t* planes, long total_rows, long kpad, int box_rows)
{
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  const int bkb = stage_k();
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad};
  cuuint32_t box[2] = {(cuuint32_t)bkb, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            bkb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

template <int S>
static inline cudaError_t configure()
{
  cudaError_t e = cudaFuncSetAttribute(oz_gemm_kernel<S, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<S, 64>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(oz_gemm_kernel<S, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<S, 64>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(oz_gemm_kernel<S, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<S, 32>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(oz_gemm_kernel<S, 64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<S, 64>::SMEM_BYTES);
  return e;
}

template <int S>
static inline void launch(const CUtensorMap& ta, const CUtensorMap& tb, const Args& g, cudaStream_t st)
{
  const unsigned grid = (unsigned)((g.m / BM) * (g.n / BN));
  switch (variant()) {
    case VAR_PAIR64: oz_gemm_kernel<S, 64, false><<<grid, 192, Cfg<S, 64>::SMEM_BYTES, st>>>(ta, tb, g); break;
    case VAR_MERGE64: oz_gemm_kernel<S, 64, true><<<grid, 192, Cfg<S, 64>::SMEM_BYTES, st>>>(ta, tb, g); break;
    case VAR_MERGE32: oz_gemm_kernel<S, 32, true><<<grid, 192, Cfg<S, 32>::SMEM_BYTES, st>>>(ta, tb, g); break;
    default: oz_gemm_kernel<S, 64, true, true><<<grid, 192, Cfg<S, 64>::SMEM_BYTES, st>>>(ta, tb, g); break;
  }
}

template <int S>
static inline void slice(const double* X, long ldx, int row0, int rows, int k0, int kcnt, int kind, int mask, const gpss::DevParams* dP,
                         int8_t* planes, long plane_rows, long kpad, cudaStream_t st, int bits = DIGIT_BITS, int* viol = nullptr)
{
  if (rows <= 0 || kcnt <= 0) return;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((kcnt + 31) / 32)), block(32, 8);
  oz_slice_kernel<S><<<grid, block, 0, st>>>(X, ldx, row0, rows, k0, kcnt, kind, mask, dP, planes, plane_rows, kpad, bits, viol);
}

// Longest k-range ONE int32 accumulation may cover: |G_g| <= S k 2^(2 (bits - 1)) < 2^31.  7-bit digits: 65 536 / S x 8 >= n_pad
// for every admitted size (one launch); 8-bit digits: 18 688 at S = 7 -- the caller cuts the k-range into launches of this length,
// the later ones accumulating into C in FP64 (a handful of roundings instead of none, against k of them on the DMMA pipe).
static inline int kseg(int S, int bits) { return bits == 8 ? ((1 << 17) / S - 1) / 64 * 64 : (1 << 30); }

}  // namespace oz
