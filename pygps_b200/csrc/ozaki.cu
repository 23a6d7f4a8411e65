// ozaki.cu - the fp64 trailing update C -= P·P' of the blocked Cholesky, computed on the int8 tensor cores.
//
// Why: B200's fp64 pipe peaks at 37 TFLOP/s (DMMA and DFMA share it; profiles/r1_peaks_first_cut.json), and the
// rank-(W*128) update is >90 % of gp.GPR.getPosterior() at N=16384.  `tcgen05.mma.kind::i8` runs at 4.5 POP/s with
// EXACT int32 accumulation in TMEM.  The update is therefore computed by an error-free split (Ozaki scheme):
//
//   row i of the panel:  P[i,:] = 2^e_i * sum_{t=1..S} d_t[i,:] * 2^(-8t)  + O(2^(e_i-8S-1)),   d_t int8 in [-128,127]
//   (P P')[i,j]          = 2^(e_i+e_j) * sum_{m=2..S+1} 2^(-8m) * G_m[i,j],   G_m = sum_{t+u=m} d_t[i,:]·d_u[j,:]
//
// Each G_m is an exact integer (|G_m| <= 7 * 1024 * 128^2 < 2^31), the S accumulators G_2..G_{S+1} live in TMEM
// (S x 64 columns), and only the final weighted sum is rounded, in fp64.  Products with t+u > S+1 are below 2^(-8(S+2)) relative to
// the row scales and are dropped.  With S=7 (radix 256) the split keeps 56 bits per entry relative to the row maximum, i.e.
// the result is at least as accurate as an fp64 DMMA accumulation of the same contraction.
//
// Kernels:
//   oz_slice_kernel  : row exponents, then the int8 slices written directly in the tensor core's canonical K-major "core matrix" order
//                      [k-step 32][row group 8][slice t][k chunk 2][row 8][16 B], so that a (rows x 32 k) stage of
//                      ALL slices of a tile is ONE contiguous block: a single cp.async.bulk (TMA) per operand per stage
//   oz_syrk_kernel   : 128x64 tiles of C (one per CTA by default, lower triangle only); warp 9 = TMA producer (barriers
//                      and the first four stages before the CTA barrier), warp 8 = MMA issuer, warps 0-7 = epilogue
//                      (software-pipelined TMEM -> fp64 -> C).  S(S+1)/2 MMAs of 128x64x32 per k-step.
//   generalisations  : launch_oz_ex - rectangular products of two operands stacked in one slice buffer (ti_min),
//                      trapezoid contraction (U U'), C -= / = / += ; used by the derivative, FITC, EP and predict paths.
// Short CTAs (not one persistent CTA per SM) on purpose: the high-priority panel stream of the look-ahead
// Cholesky needs SMs to free up every few microseconds.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gpk_internal.cuh"
#include "tc_common.cuh"

namespace gpk {

constexpr int OZ_BN = 64;        // tile columns
constexpr int OZ_ST = 4;         // pipeline stages
constexpr int OZ_EPI_WARPS = 8;  // epilogue warps 0..7, then the MMA issuer warp and the TMA producer warp
constexpr int OZ_THREADS = (OZ_EPI_WARPS + 2) * 32;

// Balanced digits in radix 2^RB, all slices signed: x (|x| <= 0.48) is rounded ONCE to S*RB fractional bits
// (integer q, exact in int64 since S*RB <= 56), then q is split from the least significant digit upwards with
// carries, d_t in [-2^(RB-1), 2^(RB-1)-1].  x = sum_t d_t 2^(-RB(t+1)) + r, |r| <= 2^(-S*RB-1), unbiased.
// Balanced digits keep the dropped cross terms (t+u > S-1) at random-walk size; unsigned digits would add them
// coherently (measured: 10x larger error).  RB=8,S=7 and RB=7,S=8 both carry 56 bits; the former needs 28 instead
// of 36 tensor-core products per k-step.
// Returns the S digits packed one per byte, slice t in byte S-1-t (bit patterns of int8).
template <int S, int RB>
__device__ __forceinline__ unsigned long long oz_digits(double x) {
  long long q = __double2ll_rn(x * (double)(1ll << (S * RB)));
  if constexpr (RB == 8) {
    // q + sum_t 128*256^t has the unsigned base-256 digits d_t + 128: no carry loop, the xor removes the offset
    unsigned long long B = 0;
#pragma unroll
    for (int t = 0; t < S; ++t) B |= 0x80ull << (8 * t);
    return ((unsigned long long)q + B) ^ B;
  } else {
    unsigned long long out = 0;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const int lo = (int)(q & ((1 << RB) - 1));
      const int dd = (lo >= (1 << (RB - 1))) ? lo - (1 << RB) : lo;
      out |= (unsigned long long)(dd & 0xff) << (8 * t);
      q = (q - dd) >> RB;
    }
    return out;
  }
}

// max(m, |v|) that does NOT lose NaN/Inf (fmax ignores NaN): a non-finite entry turns the row maximum into +Inf, which
// the slice kernel maps to a NaN row scale, so every product that touches the row comes out NaN - the NaN/Inf of a
// panel propagate into C like they do in fp64 arithmetic (include/gpk.h: "NaN/Inf propagate into outputs").
__device__ __forceinline__ double oz_absmax(double m, double v) {
  const double av = fabs(v);
  return (av <= 1.0e300) ? fmax(m, av) : __longlong_as_double(0x7ff0000000000000ll);
}

// Row maxima for the split-k slicing: grid (rows/32, ksplit), every CTA reduces its k-range and merges with an atomic
// max on the bit pattern (non-negative doubles order like unsigned integers).  mx must be zero on entry.
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ P, int64_t lda, int kw,
                                                         unsigned long long* __restrict__ mx) {
  __shared__ double red[8][32];
  const int r0 = blockIdx.x * 32;
  const int r = threadIdx.x & 31, kq = threadIdx.x >> 5;
  const int per = ((kw / 32 + gridDim.y - 1) / gridDim.y) * 32;
  const int kb = blockIdx.y * per, ke = min(kb + per, kw);
  const double* prow = P + r0 + r;
  double m = 0.0;
#pragma unroll 4
  for (int k = kb + kq; k < ke; k += 8) m = oz_absmax(m, prow[(int64_t)k * lda]);
  red[kq][r] = m;
  __syncthreads();
  if (kq == 0) {
#pragma unroll
    for (int q = 1; q < 8; ++q) m = fmax(m, red[q][r]);
    // a NaN/Inf entry made m = +Inf (oz_absmax): its bit pattern orders above every finite value
    atomicMax(mx + r0 + r, (unsigned long long)__double_as_longlong(m));
  }
}

// One CTA = 32 panel rows x (a k-range of) the contraction length.  Pass 1: row exponent (from the CTA's own
// reduction, or from mx when the k-range is split over gridDim.y CTAs); pass 2: digits, written in the tensor core's
// canonical order so that the CTA's output per k-step (4 row groups x S x 256 B) is contiguous.
// FIXED: the row scales are GIVEN (sc[row], set once per factorisation from the bound |L_ij| <= sqrt(A_ii), see
// oz_fixed_scale_kernel): no row-maximum pass, and one set of digits serves every update that reads the panel.  An entry
// that does not fit its row's scale (possible only if the matrix is not positive definite) turns the scale into NaN, so
// the overflow surfaces as NaN in the updated matrix and as info > 0 at the next pivot - never as silent garbage.
template <int S, int RB, bool FIXED = false>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ P, int64_t lda, int n, int kw,
                                                        double* __restrict__ sc, int8_t* __restrict__ sl, int row0,
                                                        const unsigned long long* __restrict__ mx) {
  // n: rows of the whole slice buffer (k-step stride); this launch fills rows [row0, row0 + 32*gridDim.x) from P
  __shared__ __align__(16) int8_t out[4][S][2][8][16];
  __shared__ double red[8][32];
  __shared__ int sh_e[32];
  const int r0 = blockIdx.x * 32;
  const int r = threadIdx.x & 31, kq = threadIdx.x >> 5;
  const double* prow = P + r0 + r;
  double m = 0.0;
  if (FIXED) {
    if (kq == 0) {
      const double scv = sc[row0 + r0 + r];
      sh_e[r] = (scv > 0.0 && scv < 1.0e300) ? ilogb(scv) + RB : 0;      // NaN scale (bad row): digits are irrelevant
    }
  } else if (mx == nullptr) {
#pragma unroll 4
    for (int k = kq; k < kw; k += 8) m = oz_absmax(m, prow[(int64_t)k * lda]);
    red[kq][r] = m;
    __syncthreads();
  }
  if (!FIXED && kq == 0) {
    if (mx == nullptr) {
#pragma unroll
      for (int q = 1; q < 8; ++q) m = fmax(m, red[q][r]);
    } else {
      m = __longlong_as_double((long long)mx[r0 + r]);
    }
    // |x| < 2^(ilogb+1)  ->  |x| * 2^-(ilogb+2) < 0.5 ; exponents clamped so that 2^(e-RB) stays a normal double
    int e = 0;
    const bool bad = !(m < 1.0e150);          // NaN/Inf in the row (oz_absmax), or a magnitude the digits cannot carry
    if (m > 0.0 && !bad) {
      e = ilogb(m) + 2;
      if (scalbn(m, -e) > 0.48) ++e;         // head room for the carry into the leading digit
    }
    if (e < -500) e = -500;
    sh_e[r] = e;
    // bad rows: NaN scale -> every entry of C in that row and column becomes NaN in the epilogue
    if (blockIdx.y == 0) sc[row0 + r0 + r] = bad ? __longlong_as_double(0x7ff8000000000000ll) : scalbn(1.0, e - RB);
  }
  __syncthreads();
  const int e = sh_e[r];
  const size_t kstride = (size_t)n * S * 32;
  int4* dst4 = reinterpret_cast<int4*>(sl + (size_t)(row0 + r0) * S * 32);
  const int4* src4 = reinterpret_cast<const int4*>(&out[0][0][0][0][0]);
  uint32_t* out32 = reinterpret_cast<uint32_t*>(&out[0][0][0][0][0]);
  // this thread: row r, the four k positions 4*kq .. 4*kq+3 of every k-step -> one 32-bit word per slice
  const int wbase = ((r >> 3) * S * 256 + ((4 * kq) >> 4) * 128 + (r & 7) * 16 + ((4 * kq) & 15)) >> 2;
  const int nsteps = kw / 32, per = (nsteps + gridDim.y - 1) / gridDim.y;
  const int ks_b = blockIdx.y * per, ks_e = min(ks_b + per, nsteps);
  bool ovf = false;
  for (int ks = ks_b; ks < ks_e; ++ks) {
    unsigned long long qq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double xs = scalbn(prow[(int64_t)(ks * 32 + 4 * kq + j) * lda], -e);            // scaling exact; |x| <= 0.48
      if (FIXED) ovf |= !(fabs(xs) <= 0.5);                                                  // (also catches NaN/Inf)
      qq[j] = oz_digits<S, RB>(xs);
    }
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const int sh = 8 * (S - 1 - t);
      const uint32_t w = (uint32_t)((qq[0] >> sh) & 0xff) | ((uint32_t)((qq[1] >> sh) & 0xff) << 8) |
                         ((uint32_t)((qq[2] >> sh) & 0xff) << 16) | ((uint32_t)((qq[3] >> sh) & 0xff) << 24);
      out32[wbase + t * 64] = w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 4 * S * 16; c += 256) dst4[ks * (kstride / 16) + c] = src4[c];
    __syncthreads();
  }
  if (FIXED && ovf) sc[row0 + r0 + r] = __longlong_as_double(0x7ff8000000000000ll);
}

// Fixed row scales of a Cholesky factorisation: |L_ij| <= sqrt(A_ii) for a positive definite A, so 2^e_i with
// sqrt(A_ii) * 2^-e_i <= 0.48 is a valid digit scale for row i in EVERY panel.  diag_const > 0: A_ii is that constant
// (stationary kernels: sf2/sn2 + 1) and A is not read; else the diagonal of A (pitch lda).  Non-positive or non-finite
// diagonal entries get a NaN scale.
template <int RB>
__global__ void oz_fixed_scale_kernel(const double* __restrict__ A, int64_t lda, int n, double diag_const,
                                      double* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = (diag_const > 0.0) ? diag_const : A[i + (int64_t)i * lda];
  double out = __longlong_as_double(0x7ff8000000000000ll);
  if (d > 0.0 && d < 1.0e300) {
    const double m = sqrt(d) * (1.0 + 1.0e-12);      // rounding head room of the computed factor
    int e = ilogb(m) + 2;
    if (scalbn(m, -e) > 0.48) ++e;
    if (e < -500) e = -500;
    out = scalbn(1.0, e - RB);
  }
  sc[i] = out;
}

struct OzArgs {
  const int8_t* sl;   // slices of the panel rows [kw/32][n/8][S][2][8][16]
  const double* sc;   // 2^(e_i - RB) per row
  double* C;          // trailing matrix origin (row 0 / col 0 of the sliced rows), column-major
  int64_t ldc;
  int n, kw;          // sliced rows, contraction length
  int jb0, jb1;       // 128-column blocks [jb0, jb1) of this launch
  int ntiles;         // lower-triangle 128x64 tiles of this launch
  int tpc;            // tiles per CTA
  int skip00;         // leave the first diagonal tile (rows/cols 0..127) alone: the panel stream updates it itself
  int ti_min;         // only row tiles >= ti_min (rectangular products: operand rows stacked [B; A], see launch_oz_gemm)
  int trap;           // contraction of row tile ti starts at k = 128*ti (U U' of an upper-triangular U)
  int cmode;          // 0: C -= P P' ; 1: C = +P P' (C is not read) ; 2: C += P P'   (1, 2: pipelined epilogue only)
  int rect_cols;      // > 0: block-cyclic columns (multi-GPU): the launch covers rect_cols LOCAL 128-column blocks c,
  int cs, cfirst;     //      global block jb = cfirst + c*cs (relative to the sliced rows); tiles with ti < jb are skipped;
                      //      C columns are the packed local ones (c*128 + ...), slices / scales / masks use jb
  long long* dbg;     // optional per-CTA clock stamps [5] (debug timing), else nullptr
  int srows;          // rows of the slice BUFFER (k-step stride = srows*S*32 bytes); 0: n.  > n when sl / sc point into a
                      // larger buffer shared by several updates (fixed row scales, see potrf_device)
};

// tile index -> (128-row tile ti, 64-column tile tj) over the lower-triangular tile set {jb0 <= jb < jb1, ti >= jb}.
// Rasterised in BANDS of OZ_BAND row tiles: inside a band the column blocks run left to right and the band's rows
// innermost, so the ~148 concurrently running CTAs share OZ_BAND A-operand tiles and a handful of B tiles.  The A
// tiles of a band stay in L2 for the whole band; every B tile is fetched from HBM once per band instead of once
// per column block (ncu: 6.6 GB of DRAM reads per update with the column-major order, against 0.9 GB of C).
constexpr int OZ_BAND = 16;
__device__ __forceinline__ void oz_decode(int idx, int nt, int jb0, int jb1, int ti_min, int& ti, int& tj) {
  int r_lo = max(jb0, ti_min);
  for (;;) {
    const int r_hi = min(r_lo + OZ_BAND, nt), rows = r_hi - r_lo;
    const int nfull = max(min(jb1, r_lo + 1) - jb0, 0);        // column blocks that see all rows of the band
    const int cnt_full = 2 * rows * nfull;
    const int t0 = max(jb0, r_lo + 1), t1 = min(jb1, r_hi);    // column blocks that cut the band diagonally
    const int nt_ = max(t1 - t0, 0);
    const int cnt_tri = 2 * nt_ * r_hi - (t0 + t1 - 1) * nt_;
    if (idx < cnt_full + cnt_tri || r_hi >= nt) {
      if (idx < cnt_full) {
        tj = 2 * jb0 + idx / rows;
        ti = r_lo + idx % rows;
      } else {
        idx -= cnt_full;
        int jb = t0;
        while (jb < t1 - 1 && idx >= 2 * (r_hi - jb)) { idx -= 2 * (r_hi - jb); ++jb; }
        const int c = r_hi - jb;
        tj = 2 * jb + idx / c;
        ti = jb + idx % c;
      }
      return;
    }
    idx -= cnt_full + cnt_tri;
    r_lo = r_hi;
  }
}

// rectangular enumeration over (row tile, local column block), bands of OZ_BAND row tiles innermost-rows like oz_decode
__device__ __forceinline__ void oz_decode_rect(int idx, int nt, int ncols, int& ti, int& tjv) {
  int r_lo = 0;
  for (;;) {
    const int rows = min(OZ_BAND, nt - r_lo), cnt = 2 * rows * ncols;
    if (idx < cnt || r_lo + rows >= nt) { tjv = idx / rows; ti = r_lo + idx % rows; return; }
    idx -= cnt;
    r_lo += rows;
  }
}

// tile index -> row tile ti, GLOBAL 64-column tile tjg (slices, scales, diagonal mask), LOCAL 64-column tile tjl (C address).
// Returns false for a tile that is not part of the launch (above the diagonal in block-cyclic mode).
template <int GEN>
__device__ __forceinline__ bool oz_tile(const OzArgs& a, int idx, int nt, int& ti, int& tjg, int& tjl) {
  if (GEN && a.rect_cols > 0) {
    int tjv;
    oz_decode_rect(idx, nt, a.rect_cols, ti, tjv);
    const int jb = a.cfirst + (tjv >> 1) * a.cs;
    tjg = 2 * jb + (tjv & 1);
    tjl = tjv;
    return ti >= jb;
  }
  oz_decode(idx, nt, a.jb0, a.jb1, a.ti_min, ti, tjg);
  tjl = tjg;
  return true;
}

// GEN = 0: the Cholesky's update (C -= P P', full contraction) with the flags folded away at compile time - the
// generalised form costs 38 registers and 5 % of the hot kernel; GEN = 1: trap / cmode honoured.
template <int S, int RB, int EPI, int GEN>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_syrk_kernel(OzArgs a) {
  const int a_trap = GEN ? a.trap : 0;
  const int a_cmode = GEN ? a.cmode : 0;
  constexpr uint32_t A_BYTES = 128 * S * 32, B_BYTES = OZ_BN * S * 32, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t TCOLS = 512;
  constexpr bool ATMEM = (S * OZ_BN + S * 8 <= 512);   // room for one k-step of the A slices behind the accumulators
  constexpr double HORNER = 1.0 / (double)(1 << RB);
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[OZ_ST], empty[OZ_ST], done, tfree;
  __shared__ double sjs[OZ_BN];
  __shared__ uint32_t tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = a.kw / 32, nt = a.n / 128;
  const size_t kstride = (size_t)(a.srows > 0 ? a.srows : a.n) * S * 32;
  const int tile0 = blockIdx.x * a.tpc;
  const int tile1 = min(tile0 + a.tpc, a.ntiles);
  if (a.skip00) {                                          // (host guarantees tpc == 1 with skip00)
    int ti0, tj0;
    oz_decode(tile0, nt, a.jb0, a.jb1, a.ti_min, ti0, tj0);
    if (ti0 == 0 && tj0 < 2) return;
  }
  if (GEN && a.rect_cols > 0) {                            // (host guarantees tpc == 1 in block-cyclic mode)
    int ti0, tg0, tl0;
    if (!oz_tile<GEN>(a, tile0, nt, ti0, tg0, tl0)) return;
  }

  if (warp == OZ_EPI_WARPS + 1) {
    // the producer initialises the barriers itself and requests the first OZ_ST stages (all of the first tile:
    // nk >= OZ_ST) BEFORE the CTA-wide barrier, so their latency overlaps the TMEM allocation
    if (elect_one()) {
      for (int s = 0; s < OZ_ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(&done, 1);
      mbar_init(&tfree, OZ_EPI_WARPS * 32);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      if (tile0 < tile1) {
        int ti, tj, tjl;
        oz_tile<GEN>(a, tile0, nt, ti, tj, tjl);
        const int8_t* gA = a.sl + (size_t)ti * 128 * S * 32;
        const int8_t* gB = a.sl + (size_t)tj * OZ_BN * S * 32;
        const int ks0 = a_trap ? 4 * ti : 0;             // (the last row tile still has nk - ks0 = 4 = OZ_ST k-steps)
#pragma unroll
        for (int ks = 0; ks < OZ_ST; ++ks) {
          mbar_expect_tx(&full[ks], STAGE);
          uint8_t* sdst = sm + (size_t)ks * STAGE;
          bulk_g2s(sdst, gA + (ks0 + ks) * kstride, A_BYTES, &full[ks]);
          bulk_g2s(sdst + A_BYTES, gB + (ks0 + ks) * kstride, B_BYTES, &full[ks]);
        }
      }
    }
    __syncwarp();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  long long* dbg = (a.dbg && blockIdx.x < 4096) ? a.dbg + 5 * blockIdx.x : nullptr;
  if (dbg && tid == 0) dbg[0] = clock64();

  if (warp == OZ_EPI_WARPS + 1) {
    if (elect_one()) {
      int it = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        int ti, tj, tjl;
        oz_tile<GEN>(a, tile, nt, ti, tj, tjl);
        const int8_t* gA = a.sl + (size_t)ti * 128 * S * 32;
        const int8_t* gB = a.sl + (size_t)tj * OZ_BN * S * 32;
        for (int ks = a_trap ? 4 * ti : 0; ks < nk; ++ks, ++it) {
          const int slot = it % OZ_ST;
          if (it < OZ_ST) continue;                      // requested before the CTA barrier (see above)
          mbar_wait(&empty[slot], ((it / OZ_ST) - 1) & 1);
          mbar_expect_tx(&full[slot], STAGE);
          uint8_t* s = sm + (size_t)slot * STAGE;
          bulk_g2s(s, gA + ks * kstride, A_BYTES, &full[slot]);
          bulk_g2s(s + A_BYTES, gB + ks * kstride, B_BYTES, &full[slot]);
        }
      }
    }
  } else if (warp == OZ_EPI_WARPS) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_i8(128, OZ_BN);
      // smem descriptors differ only in the 14-bit start-address field: slice t sits 256 B (= 16 units) further
      const uint64_t dbase = make_smem_desc(0, 128, S * 256);
      int it = 0, tcount = 0;
      for (int tile = tile0; tile < tile1; ++tile, ++tcount) {
        if (tcount > 0) { mbar_wait(&tfree, (tcount - 1) & 1); tc_fence_after(); }
        int ks0 = 0;
        if (a_trap) {
          int ti, tj, tjl;
          oz_tile<GEN>(a, tile, nt, ti, tj, tjl);
          ks0 = 4 * ti;
        }
        if constexpr (ATMEM) {
          // The A slices go shared memory -> TMEM once per k-step (tcgen05.cp) and every product reads them from
          // there: shared-memory traffic per k-step drops from S(S+1)/2 x 6 KB to S x 4 KB + S(S+1)/2 x 2 KB, which
          // takes the loop from shared-memory-bound (48 clk per product) to tensor-bound (32 clk).  The copy of
          // slice t for the NEXT k-step is issued right behind the last product that reads slice t; tcgen05
          // operations of one thread execute in issue order, so the single TMEM copy of A needs no extra barrier.
          const uint32_t ta = tbase + S * OZ_BN;
          {
            const int slot = it % OZ_ST;
            mbar_wait(&full[slot], (it / OZ_ST) & 1);
            tc_fence_after();
            if (dbg && tcount == 0) dbg[1] = clock64();
            const uint64_t da = dbase | (uint64_t)(smem_u32(sm + (size_t)slot * STAGE) >> 4);
#pragma unroll
            for (int t = 0; t < S; ++t) tc_cp_128x256b(ta + 8 * t, da + 16 * t);
          }
          for (int ks = ks0; ks < nk; ++ks, ++it) {
            const int slot = it % OZ_ST, nslot = (it + 1) % OZ_ST;
            const bool more = ks + 1 < nk;
            const uint64_t db = dbase | (uint64_t)((smem_u32(sm + (size_t)slot * STAGE) + A_BYTES) >> 4);
            const uint64_t dan = dbase | (uint64_t)(smem_u32(sm + (size_t)nslot * STAGE) >> 4);
#pragma unroll
            for (int t = 0; t < S; ++t) {
#pragma unroll
              for (int u = 0; u < S - t; ++u)
                tc_mma_i8_ts(tbase + (uint32_t)(t + u) * OZ_BN, ta + 8 * t, db + 16 * u, idesc, (ks > ks0 || t > 0) ? 1u : 0u);
              if (more) {
                if (t == 0) { mbar_wait(&full[nslot], ((it + 1) / OZ_ST) & 1); tc_fence_after(); }
                tc_cp_128x256b(ta + 8 * t, dan + 16 * t);
              }
            }
            tc_commit(&empty[slot]);
          }
        } else {
          for (int ks = ks0; ks < nk; ++ks, ++it) {
            const int slot = it % OZ_ST;
            mbar_wait(&full[slot], (it / OZ_ST) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(sm + (size_t)slot * STAGE);
            const uint64_t da = dbase | (uint64_t)(sa >> 4), db = dbase | (uint64_t)((sa + A_BYTES) >> 4);
#pragma unroll
            for (int t = 0; t < S; ++t) {
#pragma unroll
              for (int u = 0; u < S - t; ++u)
                tc_mma_i8(tbase + (uint32_t)(t + u) * OZ_BN, da + 16 * t, db + 16 * u, idesc, (ks > ks0 || t > 0) ? 1u : 0u);
            }
            tc_commit(&empty[slot]);
          }
        }
        tc_commit(&done);
        if (dbg) dbg[2] = clock64();
      }
    }
  } else {
    // epilogue warps: TMEM lanes 32*(warp%4).. = rows of the tile; warps 4..7 take the upper 32 columns.
    // The C tile is fetched BEFORE the accumulators are complete, so its latency hides under the MMA loop.
    const int q4 = warp & 3, chalf = (warp >> 2) * (OZ_BN / 2);
    const uint32_t tw = tbase + ((uint32_t)(q4 * 32) << 16) + chalf;
    int tcount = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++tcount) {
      int ti, tj, tjl;
      oz_tile<GEN>(a, tile, nt, ti, tj, tjl);
      const int gi = ti * 128 + q4 * 32 + lane, gj0 = tj * OZ_BN + chalf;      // global (mask, scales)
      const double si = a.sc[gi];
      double* crow = a.C + gi + (int64_t)(tjl * OZ_BN + chalf) * a.ldc;        // local packed column
      if constexpr (EPI == 0) {
        // pull this thread's 32 C entries towards L2 now; they are read after the accumulators are complete
#pragma unroll
        for (int q = 0; q < OZ_BN / 2; ++q)
          if (gi >= gj0 + q) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(crow + (int64_t)q * a.ldc));
        mbar_wait(&done, tcount & 1);
        tc_fence_after();
        if (dbg && tid == 0) dbg[3] = clock64();
#pragma unroll 1
        for (int c0 = 0; c0 < OZ_BN / 2; c0 += 8) {
          uint32_t v[S][8];
#pragma unroll
          for (int m = 0; m < S; ++m) tc_ld8(tw + (uint32_t)m * OZ_BN + c0, v[m]);
          double cv[8], sj[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            cv[q] = (gi >= gj0 + c0 + q) ? crow[(int64_t)(c0 + q) * a.ldc] : 0.0;
            sj[q] = __ldg(a.sc + gj0 + c0 + q);
          }
          tc_wait_ld();
          if (c0 + 8 == OZ_BN / 2) {           // all accumulator reads of this tile are done: hand TMEM back
            tc_fence_before();
            mbar_arrive(&tfree);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            // int32 -> fp64 without the (slow) conversion unit: 2^52 + 2^31 + v is exactly representable
            double acc = __hiloint2double(0x43300000, (int)(v[S - 1][q] ^ 0x80000000u)) - 4503601774854144.0;
#pragma unroll
            for (int m = S - 2; m >= 0; --m)
              acc = fma(acc, HORNER, __hiloint2double(0x43300000, (int)(v[m][q] ^ 0x80000000u)) - 4503601774854144.0);
            if (gi >= gj0 + c0 + q) crow[(int64_t)(c0 + q) * a.ldc] = cv[q] - (acc * si) * sj[q];
          }
        }
      } else {
        // EPI 1: software-pipelined drain.  The accumulators are read in rounds of 4 columns (7 x tcgen05.ld.x4); round
        // r+1's TMEM loads and round r+2's C loads are in flight while round r is converted and stored, so the TMEM
        // read (the longest part of the epilogue), the fp64 conversion and the global latency overlap.
        constexpr int NR = OZ_BN / 2 / 4;               // rounds per thread
        uint32_t v[2][S][4];
        double cvb[3][4];
        // the tile's 64 column scales go to shared memory once (ncu: the epilogue's largest stall was the global
        // load of the first column scale of every round)
        if (tid < OZ_BN) sjs[tid] = a.sc[tj * OZ_BN + tid];
        asm volatile("bar.sync 1, %0;\n" ::"n"(OZ_EPI_WARPS * 32) : "memory");
        const bool rd = (a_cmode != 1);                  // C is read (update) or only written (set)
        const double sg = (a_cmode == 0) ? 1.0 : -1.0;
        if (rd) {
#pragma unroll
          for (int q = 0; q < OZ_BN / 2; ++q)
            if (gi >= gj0 + q) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(crow + (int64_t)q * a.ldc));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) cvb[r][q] = (rd && gi >= gj0 + 4 * r + q) ? crow[(int64_t)(4 * r + q) * a.ldc] : 0.0;
        mbar_wait(&done, tcount & 1);
        tc_fence_after();
        if (dbg && tid == 0) dbg[3] = clock64();
#pragma unroll
        for (int m = 0; m < S; ++m) tc_ld4(tw + (uint32_t)m * OZ_BN, v[0][m]);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          tc_wait_ld();                                  // round r has landed
          if (r + 1 < NR) {
#pragma unroll
            for (int m = 0; m < S; ++m) tc_ld4(tw + (uint32_t)m * OZ_BN + 4 * (r + 1), v[(r + 1) & 1][m]);
          } else {                                       // all accumulator reads of this tile are done: hand TMEM back
            tc_fence_before();
            mbar_arrive(&tfree);
          }
          if (r + 2 < NR) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              cvb[(r + 2) % 3][q] = (rd && gi >= gj0 + 4 * (r + 2) + q) ? crow[(int64_t)(4 * (r + 2) + q) * a.ldc] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = 4 * r + q;
            double acc = __hiloint2double(0x43300000, (int)(v[r & 1][S - 1][q] ^ 0x80000000u)) - 4503601774854144.0;
#pragma unroll
            for (int m = S - 2; m >= 0; --m)
              acc = fma(acc, HORNER, __hiloint2double(0x43300000, (int)(v[r & 1][m][q] ^ 0x80000000u)) - 4503601774854144.0);
            if (gi >= gj0 + c) crow[(int64_t)c * a.ldc] = cvb[r % 3][q] - sg * ((acc * si) * sjs[chalf + c]);
          }
        }
      }
      if (dbg && tid == 0) dbg[4] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(TCOLS) : "memory");
}

struct OzCfg { int S, RB, tpc, epi; };
static OzCfg oz_cfg() {
  static OzCfg c{0, 0, 0, 0};
  if (c.S == 0) {
    c.RB = 8; c.S = 7; c.tpc = 1; c.epi = 1;   // epi 0: the first, unpipelined epilogue (kept for A/B runs)
    if (const char* e = getenv("GPK_OZAKI_EPI")) c.epi = atoi(e) ? 1 : 0;
    if (const char* e = getenv("GPK_OZAKI_RADIX")) { const int v = atoi(e); if (v == 7 || v == 8) c.RB = v; }
    if (c.RB == 7) c.S = 8;
    if (const char* e = getenv("GPK_OZAKI_SLICES")) { const int v = atoi(e); if (v >= 6 && v <= 8) c.S = v; }
    if (c.RB == 8 && c.S == 8) c.S = 7;
    if (c.RB == 7 && c.S == 6) c.S = 7;
    if (const char* e = getenv("GPK_OZAKI_TPC")) { const int v = atoi(e); if (v >= 1 && v <= 64) c.tpc = v; }
  }
  return c;
}

int oz_ensure(Handle* h, int which, int64_t n, int kw) {
  const size_t need = (size_t)n * kw * 8;
  if (h->ozCap[which] < need) {
    if (h->ozSl[which]) cudaFree(h->ozSl[which]);
    h->ozSl[which] = nullptr; h->ozCap[which] = 0;
    GPK_CK(h, cudaMalloc((void**)&h->ozSl[which], need));
    h->ozCap[which] = need;
  }
  if (h->ozScCap[which] < (size_t)n) {
    if (h->ozSc[which]) cudaFree(h->ozSc[which]);
    h->ozSc[which] = nullptr; h->ozScCap[which] = 0;
    GPK_CK(h, cudaMalloc((void**)&h->ozSc[which], 2 * (size_t)n * sizeof(double)));   // scales, then row-max scratch
    h->ozScCap[which] = (size_t)n;
  }
  return 0;
}

template <int S, int RB>
static int oz_slice_t(Handle* h, int which, cudaStream_t st, const double* P, int64_t lda, int n, int kw, int row0,
                      int ntot) {
  // Few rows and a long contraction (the FITC SYRK: 4096 x 16384; the Cholesky panels: ~15000 x 1152) would leave most
  // SMs idle with one CTA per 32 rows: split the k-range over gridDim.y CTAs, with the row maxima from a pre-pass.
  const int ctas = n / 32, nsteps = kw / 32;
  int ksplit = (ctas >= 1184) ? 1 : (1184 + ctas - 1) / ctas;
  if (ksplit > nsteps / 4) ksplit = nsteps / 4 > 0 ? nsteps / 4 : 1;
  if (env_int("GPK_OZAKI_KSPLIT", 1) == 0) ksplit = 1;
  if (ksplit <= 1) {
    oz_slice_kernel<S, RB><<<ctas, 256, 0, st>>>(P, lda, ntot, kw, h->ozSc[which], h->ozSl[which], row0, nullptr);
  } else {
    unsigned long long* mx = reinterpret_cast<unsigned long long*>(h->ozSc[which] + h->ozScCap[which]);   // second half
    GPK_CK(h, cudaMemsetAsync(mx, 0, (size_t)n * sizeof(unsigned long long), st));
    oz_rowmax_kernel<<<dim3(ctas, ksplit), 256, 0, st>>>(P, lda, kw, mx);
    oz_slice_kernel<S, RB><<<dim3(ctas, ksplit), 256, 0, st>>>(P, lda, ntot, kw, h->ozSc[which], h->ozSl[which], row0, mx);
  }
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// rows [row0, row0+n) of a slice buffer of ntot rows (0: n) <- P (n rows x kw columns, column-major, lda)
int launch_oz_slice(Handle* h, int which, cudaStream_t st, const double* P, int64_t lda, int n, int kw, int row0,
                    int ntot) {
  if (ntot <= 0) ntot = n;
  if (n % 128 != 0 || kw % 32 != 0 || row0 % 128 != 0 || row0 + n > ntot || (size_t)ntot * kw * 8 > h->ozCap[which] ||
      (size_t)ntot > h->ozScCap[which])
    return GPK_ERR_ARG;
  const OzCfg c = oz_cfg();
  if (c.RB == 7)
    return c.S == 8 ? oz_slice_t<8, 7>(h, which, st, P, lda, n, kw, row0, ntot) : oz_slice_t<7, 7>(h, which, st, P, lda, n, kw, row0, ntot);
  return c.S == 7 ? oz_slice_t<7, 8>(h, which, st, P, lda, n, kw, row0, ntot) : oz_slice_t<6, 8>(h, which, st, P, lda, n, kw, row0, ntot);
}

template <int S, int RB, int EPI, int GEN>
static int oz_syrk_g(Handle* h, cudaStream_t st, const OzArgs& a) {
  const size_t smem = (size_t)OZ_ST * (128 + OZ_BN) * S * 32;
  GPK_SMEM_ATTR(h, (oz_syrk_kernel<S, RB, EPI, GEN>), smem);
  oz_syrk_kernel<S, RB, EPI, GEN><<<(a.ntiles + a.tpc - 1) / a.tpc, OZ_THREADS, smem, st>>>(a);
  GPK_CK(h, cudaGetLastError());
  return 0;
}

template <int S, int RB, int EPI>
static int oz_syrk_e(Handle* h, cudaStream_t st, const OzArgs& a) {
  return (a.trap || a.cmode || a.rect_cols) ? oz_syrk_g<S, RB, EPI, 1>(h, st, a) : oz_syrk_g<S, RB, EPI, 0>(h, st, a);
}

template <int S, int RB>
static int oz_syrk_t(Handle* h, cudaStream_t st, const OzArgs& a) {
  return oz_cfg().epi ? oz_syrk_e<S, RB, 1>(h, st, a) : oz_syrk_e<S, RB, 0>(h, st, a);
}

// Block-cyclic columns (sharded factorisation, dist.cu): C holds `ncols` LOCAL 128-column blocks of this rank, local block c
// being global block cfirst + c*cs of the sliced rows; lower(C) -= P P' on the tiles ti >= global block (the others exit).
int launch_oz_cyclic(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int ncols, int cfirst,
                     int cs) {
  const OzCfg c = oz_cfg();
  const int nt = n / 128;
  if (ncols <= 0) return 0;
  if (c.tpc != 1 || !c.epi || cs < 1 || cfirst < 0 || cfirst + (ncols - 1) * cs >= nt) return GPK_ERR_ARG;
  if ((size_t)kw * 7 * 16384 >= 2147483648ull) return GPK_ERR_ARG;
  OzArgs a{h->ozSl[which], h->ozSc[which], C, ldc, n, kw, 0, ncols, 2 * nt * ncols, 1, 0, 0, 0, 0, ncols, cs, cfirst, h->ozDbg, 0};
  if (c.RB == 7) return c.S == 8 ? oz_syrk_t<8, 7>(h, st, a) : oz_syrk_t<7, 7>(h, st, a);
  return c.S == 7 ? oz_syrk_t<7, 8>(h, st, a) : oz_syrk_t<6, 8>(h, st, a);
}

// ---- fixed row scales: one slice buffer shared by every update of a factorisation level (potrf_device) ------------
int launch_oz_fixed_scales(Handle* h, cudaStream_t st, const double* A, int64_t lda, int n, double diag_const, double* sc) {
  const OzCfg c = oz_cfg();
  if (c.RB == 7) oz_fixed_scale_kernel<7><<<(n + 255) / 256, 256, 0, st>>>(A, lda, n, diag_const, sc);
  else oz_fixed_scale_kernel<8><<<(n + 255) / 256, 256, 0, st>>>(A, lda, n, diag_const, sc);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

template <int S, int RB>
static int oz_slice_fixed_t(cudaStream_t st, const double* P, int64_t lda, int nrows, int kw, int8_t* sl, double* sc,
                            int srows, int row0, int kstep0) {
  const int ctas = nrows / 32, nsteps = kw / 32;
  int ksplit = (ctas >= 1184) ? 1 : (1184 + ctas - 1) / ctas;
  if (ksplit > nsteps / 4) ksplit = nsteps / 4 > 0 ? nsteps / 4 : 1;
  int8_t* base = sl + (size_t)kstep0 * srows * S * 32;
  oz_slice_kernel<S, RB, true><<<dim3(ctas, ksplit), 256, 0, st>>>(P, lda, srows, kw, sc, base, row0, nullptr);
  return 0;
}

// rows [row0, row0 + nrows) x k-steps [kstep0, kstep0 + kw/32) of the slice buffer (sl, srows rows) <- P (nrows x kw), with
// the GIVEN scales sc[row0 ...] (buffer-row indexed)
int launch_oz_slice_fixed(Handle* h, cudaStream_t st, const double* P, int64_t lda, int nrows, int kw, int8_t* sl,
                          double* sc, int srows, int row0, int kstep0) {
  if (nrows % 128 != 0 || kw % 32 != 0 || row0 % 128 != 0 || row0 + nrows > srows) return GPK_ERR_ARG;
  const OzCfg c = oz_cfg();
  if (c.RB == 7) { if (c.S == 8) oz_slice_fixed_t<8, 7>(st, P, lda, nrows, kw, sl, sc, srows, row0, kstep0); else oz_slice_fixed_t<7, 7>(st, P, lda, nrows, kw, sl, sc, srows, row0, kstep0); }
  else { if (c.S == 7) oz_slice_fixed_t<7, 8>(st, P, lda, nrows, kw, sl, sc, srows, row0, kstep0); else oz_slice_fixed_t<6, 8>(st, P, lda, nrows, kw, sl, sc, srows, row0, kstep0); }
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int oz_slices() { return oz_cfg().S; }

// lower(C)(128-column blocks [jb0, jb1)) -= P P' from a shared slice buffer: sl / sc already point at the first row and
// the first k-step of this update; srows = rows of the whole buffer (the k-step stride)
int launch_oz_syrk_buf(Handle* h, cudaStream_t st, const int8_t* sl, const double* sc, int srows, double* C, int64_t ldc,
                       int n, int kw, int jb0, int jb1, int skip00) {
  const OzCfg c = oz_cfg();
  const int nt = n / 128;
  if (jb0 < 0 || jb1 > nt || jb0 >= jb1 || n > srows) return GPK_ERR_ARG;
  if (skip00 && (c.tpc != 1 || jb0 != 0)) return GPK_ERR_ARG;
  if ((size_t)kw * 7 * 16384 >= 2147483648ull) return GPK_ERR_ARG;
  long long nt64 = 0;
  for (int jb = jb0; jb < jb1; ++jb) nt64 += 2 * (nt - jb);
  OzArgs a{sl, sc, C, ldc, n, kw, jb0, jb1, (int)nt64, c.tpc, skip00, 0, 0, 0, 0, 0, 0, h->ozDbg, srows};
  if (c.RB == 7) return c.S == 8 ? oz_syrk_t<8, 7>(h, st, a) : oz_syrk_t<7, 7>(h, st, a);
  return c.S == 7 ? oz_syrk_t<7, 8>(h, st, a) : oz_syrk_t<6, 8>(h, st, a);
}

// C(lower triangle, 128-column blocks [jb0, jb1)) -= P·P' from the current slices
// General form: tiles {jb0 <= jb < jb1, ti >= max(jb, ti_min)} of C(stacked row, stacked column) (-)= P P' from the current
// slices (n rows).  trap: the contraction of row tile ti starts at k = 128 ti.  cmode 1 / 2: C = / += P P'.
int launch_oz_ex(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int jb0, int jb1,
                 int skip00, int ti_min, int trap, int cmode) {
  const OzCfg c = oz_cfg();
  const int nt = n / 128;
  if (jb0 < 0 || jb1 > nt || jb0 >= jb1 || ti_min < 0 || ti_min >= nt) return GPK_ERR_ARG;
  if (skip00 && (c.tpc != 1 || jb0 != 0 || ti_min != 0)) return GPK_ERR_ARG;
  if (trap && (kw != n || jb0 != 0 || ti_min != 0)) return GPK_ERR_ARG;
  if (cmode && !c.epi) return GPK_ERR_ARG;                               // only the pipelined epilogue knows set / add
  if ((size_t)kw * 7 * 16384 >= 2147483648ull) return GPK_ERR_ARG;      // int32 accumulators: 7 kw 2^14 < 2^31
  long long nt64 = 0;
  for (int jb = jb0; jb < jb1; ++jb) nt64 += 2 * (nt - (jb > ti_min ? jb : ti_min));
  const int ntiles = (int)nt64;
  OzArgs a{h->ozSl[which], h->ozSc[which], C, ldc, n, kw, jb0, jb1, ntiles, c.tpc, skip00, ti_min, trap, cmode, 0, 0, 0, h->ozDbg, 0};
  if (c.RB == 7) return c.S == 8 ? oz_syrk_t<8, 7>(h, st, a) : oz_syrk_t<7, 7>(h, st, a);
  return c.S == 7 ? oz_syrk_t<7, 8>(h, st, a) : oz_syrk_t<6, 8>(h, st, a);
}

// C(lower triangle, 128-column blocks [jb0, jb1)) -= P·P' from the current slices
int launch_oz_syrk(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int jb0, int jb1,
                   int skip00) {
  return launch_oz_ex(h, which, st, C, ldc, n, kw, jb0, jb1, skip00, 0, 0, 0);
}

// C(rows of A, rows of B) -= A·B' for two operands sliced into ONE buffer as [B (nb rows); A (na rows)]: in stacked
// indices every wanted tile lies below the diagonal, so the SYRK kernel computes exactly the nb x na rectangle.
// Cab points at the entry (first A row, first B row) of the target, column-major with pitch ldc.
int launch_oz_gemm_stacked(Handle* h, int which, cudaStream_t st, double* Cab, int64_t ldc, int nb, int na, int kw) {
  if (nb % 128 != 0 || na % 128 != 0 || nb <= 0 || na <= 0) return GPK_ERR_ARG;
  return launch_oz_ex(h, which, st, Cab - nb, ldc, nb + na, kw, 0, nb / 128, 0, nb / 128, 0, 0);
}

}  // namespace gpk

using namespace gpk;

// C (n x n, column-major, lower triangle) -= P (n x kw) · P'; host buffers; mode 0: int8 tensor-core path, 1: DMMA path.
// `ms` (optional) receives the device time of the update alone (slicing included for mode 0), averaged over `reps`.
extern "C" int gpk_dbg_oz_syrk(gpk_handle hh, int64_t n, int kw, const double* P, double* C, int mode, int reps,
                               double* ms) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!P || !C || n <= 0 || n % 128 != 0 || kw <= 0 || kw % 128 != 0 || reps < 1) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  double *dP = nullptr, *dC = nullptr;
  GPK_CK(h, cudaMalloc((void**)&dP, (size_t)n * kw * 8));
  GPK_CK(h, cudaMalloc((void**)&dC, (size_t)n * n * 8));
  int rc = 0;
  cudaMemcpyAsync(dP, P, (size_t)n * kw * 8, cudaMemcpyHostToDevice, st);
  if (mode == 0) rc = oz_ensure(h, 0, n, kw);
  long long* dDbg = nullptr;
  if (mode == 0 && getenv("GPK_OZ_TIMING")) {
    cudaMalloc((void**)&dDbg, 4096 * 5 * sizeof(long long));
    cudaMemset(dDbg, 0, 4096 * 5 * sizeof(long long));
    h->ozDbg = dDbg;
  }
  float total = 0.f;
  for (int r = 0; r < reps && rc == 0; ++r) {
    cudaMemcpyAsync(dC, C, (size_t)n * n * 8, cudaMemcpyHostToDevice, st);
    cudaEventRecord(h->t0, st);
    if (mode == 0) {
      rc = launch_oz_slice(h, 0, st, dP, n, (int)n, kw);
      if (rc == 0) rc = launch_oz_syrk(h, 0, st, dC, n, (int)n, kw, 0, (int)(n / 128));
    } else {
      GemmArgs u{};
      u.A = dP; u.B = dP; u.C = dC; u.lda = n; u.ldb = n; u.ldc = n; u.K = kw; u.tri = 1;
      rc = launch_gemm_nt(h, st, 1, u, (int)(n / NB), (int)(n / NB));
    }
    cudaEventRecord(h->t1, st);
    cudaStreamSynchronize(st);
    float t = 0.f;
    cudaEventElapsedTime(&t, h->t0, h->t1);
    total += t;
  }
  if (rc == 0) cudaMemcpyAsync(C, dC, (size_t)n * n * 8, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (dDbg) {
    std::vector<long long> v(4096 * 5);
    cudaMemcpy(v.data(), dDbg, v.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0; int cnt = 0;
    for (int b = 0; b < 4096; ++b) {
      const long long* d = &v[5 * b];
      if (!d[0] || !d[4]) continue;
      s1 += d[1] - d[0]; s2 += d[2] - d[0]; s3 += d[3] - d[0]; s4 += d[4] - d[0]; ++cnt;
    }
    if (cnt) fprintf(stderr, "oz timing (clk, mean over %d CTAs): first stage landed +%.0f, all MMAs issued +%.0f, accumulators complete +%.0f, epilogue end +%.0f\n", cnt, s1 / cnt, s2 / cnt, s3 / cnt, s4 / cnt);
    cudaFree(dDbg);
    h->ozDbg = nullptr;
  }
  cudaFree(dP); cudaFree(dC);
  if (rc) return rc;
  GPK_CK(h, e);
  GPK_CK(h, cudaGetLastError());
  if (ms) *ms = total / reps;
  return 0;
}
