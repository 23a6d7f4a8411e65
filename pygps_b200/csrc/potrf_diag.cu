// potrf_diag.cu - the latency-critical kernels of the blocked Cholesky: everything that sits ON the dependent panel chain.
//
//  potrf_diag_ovl_kernel : factor one 128x128 diagonal block (lower Cholesky) AND invert the factor, in shared memory,
//                          one CTA.  Warp 0 factors the 32x32 diagonal sub-blocks (square-root-free LDL', one matrix row
//                          per lane, columns travelling through shared memory); the sub-panel solves and trailing updates
//                          are small shared-memory GEMMs by all 8 warps (DMMA); the inverse is built by ROW blocks by
//                          warps 1-7 while warp 0 factors the next sub-block.  The default (GPK_DIAG_OVL=1).
//  potrf_diag_kernel     : the round-1 form of the same (inverse after the factorisation, GPK_DIAG_OVL=0), and with
//                          INV_ONLY the inverses + log-determinant shares of an UPLOADED factor, all blocks in one launch
//                          (gpk_set_factor).
//    The inverse turns the panel TRSM and every later triangular solve with this block into a tensor-pipe GEMM.
//    Replaces the unblocked part of LAPACK dpotrf called at /root/reference/pyGPs/Core/tools.py:61; emits
//    sum(log(diag L)) for Core/inf.py:370 and a LAPACK-style info for Core/tools.py:62-77.
//
//  small_nt_kernel       : the two products behind every diagonal block (tile (p+1,p) times Dinv_p', rank-128 update of
//                          tile (p+1,p+1)) and the rank-512 update of the next diagonal tile at a level-1 hand-over, as
//                          sixteen 32x32-block CTAs per 128x128 tile.
//
//  trsv_fwd_kernel / trsv_bwd_kernel / trsv_bwd_persistent_kernel : the forward / backward substitution with a single
//    right-hand side (solve_chol with B=(n,1), Core/tools.py:96, as used at Core/inf.py:363).
#include "gpk_internal.cuh"

namespace gpk {

constexpr unsigned FULL = 0xffffffffu;
constexpr int DB = NB;            // 128
constexpr int IB = DIAG_IB;       // 32
constexpr int LDS_ = DIAG_LDS;    // 132: == 4 mod 16 -> DMMA fragment loads from S are bank-conflict free
constexpr int LDT = DIAG_LDT;     // 36:  same property for the 32x32 inverses
constexpr int DW = DIAG_THREADS / 32;

// S: DBxDB column-major (pitch LDS_).  T: 4 blocks W_jj = inv(L_jj), 32x32 column-major, pitch LDT.
#define S_(r, c) S[(r) + (c) * LDS_]
#define T_(b, r, c) T[(b) * IB * LDT + (r) + (c) * LDT]

__device__ __forceinline__ void dmma8(double (&c)[4], const double (&a)[4], double b0, double b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}

// One warp: acc(32 x 8*NT) += A(32 x K) * Bop(K x 8*NT), operands in shared memory.
//   A(r,k)   = A[r + k*lda]                       (column-major rows of S)
//   Bop(k,n) = B[n*sbn + k*sbk]                   (sbn=1,sbk=ld: B^T of a column-major matrix; sbn=ld,sbk=1: B itself)
// *p = the smallest non-zero info seen (0 = none yet): LAPACK reports the FIRST failing pivot
__device__ __forceinline__ void info_min(int* p, int v) {
  int old = atomicCAS(p, 0, v);
  while (old != 0 && old > v) {
    const int prev = atomicCAS(p, old, v);
    if (prev == old) break;
    old = prev;
  }
}

template <int NT>
__device__ __forceinline__ void warp_mma32(double (&acc)[2][NT][4], const double* A, int lda, const double* B, int sbn,
                                           int sbk, int K, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int k0 = 0; k0 < K; k0 += 8) {
    double a[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      a[mi][0] = A[(mi * 16 + g) + (k0 + t) * lda];
      a[mi][1] = A[(mi * 16 + g + 8) + (k0 + t) * lda];
      a[mi][2] = A[(mi * 16 + g) + (k0 + t + 4) * lda];
      a[mi][3] = A[(mi * 16 + g + 8) + (k0 + t + 4) * lda];
    }
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      const double b0 = B[(ni * 8 + g) * sbn + (k0 + t) * sbk];
      const double b1 = B[(ni * 8 + g) * sbn + (k0 + t + 4) * sbk];
      dmma8(acc[0][ni], a[0], b0, b1);
      dmma8(acc[1][ni], a[1], b0, b1);
    }
  }
}

// C fragment <-> shared memory: element (mi*16+g(+8), ni*8+2t(+1)) of a 32 x 8*NT block at C[r + c*LDS_]
template <int NT>
__device__ __forceinline__ void frag_apply(double (&acc)[2][NT][4], double* C, int lane, double alpha, double beta) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      double* p = C + (mi * 16 + g) + (ni * 8 + 2 * t) * LDS_;
      p[0] = alpha * acc[mi][ni][0] + (beta != 0.0 ? beta * p[0] : 0.0);
      p[LDS_] = alpha * acc[mi][ni][1] + (beta != 0.0 ? beta * p[LDS_] : 0.0);
      p[8] = alpha * acc[mi][ni][2] + (beta != 0.0 ? beta * p[8] : 0.0);
      p[8 + LDS_] = alpha * acc[mi][ni][3] + (beta != 0.0 ? beta * p[8 + LDS_] : 0.0);
    }
}

template <int NT>
__device__ __forceinline__ void frag_zero(double (&acc)[2][NT][4]) {
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.0;
}

// INV_ONLY = true: the block already holds a lower Cholesky factor (gpk_set_factor: a factor uploaded by the caller);
// phase 1 is skipped, only 1/diag, the log-determinant share and the block inverse are produced, one CTA per diagonal
// block (blockIdx.x) so the whole factor is prepared in ONE launch.
template <bool INV_ONLY>
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ Ablk, int64_t lda, double* __restrict__ Dinv,
                  double* __restrict__ logdet_slot, int* __restrict__ info, int gidx0, long long* dbg_clk) {
  extern __shared__ __align__(16) double dsm[];
  if (INV_ONLY) {
    Ablk += (int64_t)blockIdx.x * DB * (1 + lda);
    Dinv += (int64_t)blockIdx.x * DB * DB;
    logdet_slot += blockIdx.x;
    gidx0 += blockIdx.x * DB;
  }
  int dbg_i = 0;
#define DBG_T() do { if (dbg_clk && threadIdx.x == 0) dbg_clk[dbg_i++] = clock64(); } while (0)
  DBG_T();
  double* S = dsm;                    // DB x LDS_
  double* T = dsm + DB * LDS_;        // 4 x (IB x LDT)
  double* rdiag = T + 4 * IB * LDT;   // DB reciprocals of the diagonal of L
  __shared__ double s_logdet;
  __shared__ int s_info;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // 16-byte loads, all 64 KB of the lower triangle requested before the first use (the block is L2-resident:
  // the update that produced it has just finished)
#pragma unroll 8
  for (int idx = tid; idx < DB * DB / 2; idx += DIAG_THREADS) {
    const int r = (idx % (DB / 2)) * 2, c = idx / (DB / 2);
    double2 v = make_double2(0.0, 0.0);
    if (r + 1 >= c) v = *reinterpret_cast<const double2*>(Ablk + r + (int64_t)c * lda);
    if (r < c) v.x = 0.0;
    *reinterpret_cast<double2*>(&S_(r, c)) = v;
  }
  if (tid == 0) { s_logdet = 0.0; s_info = 0; }
  __syncthreads();
  DBG_T();

  if (INV_ONLY) {
    if (tid < DB) {
      const double d = S_(tid, tid);
      rdiag[tid] = 1.0 / d;
      if (!(d > 0.0)) info_min(&s_info, gidx0 + tid + 1);
    }
    if (warp == 0) {
      double lg = 0.0;
      for (int i = lane; i < DB; i += 32) lg += log(S_(i, i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(FULL, lg, o);
      if (lane == 0) s_logdet = lg;
    }
    __syncthreads();
  }
  // ---------------- phase 1: blocked Cholesky, inner block 32 ----------------
  for (int jb = 0; jb < (INV_ONLY ? 0 : DB / IB); ++jb) {
    const int j0 = jb * IB;
    if (warp == 0) {
      // 32x32 diagonal block: ONE warp, one matrix row per lane in registers, square-root-free (LDL') elimination.
      // Measured on B200 (scripts/lat_probe.cu): DFMA latency 8 clk, but a warp issues only one SHFL per 6.5 clk
      // (37 clk latency) - the 992 shuffles per block of the first version cost 11k clk, and eight warps running it
      // redundantly shared one shuffle unit.  Here column j travels through shared memory instead: every lane stores
      // its (unscaled) entry straight into its final place S(:, j), and reads the column back with broadcast
      // 16-byte loads.  Dependent chain per column: STS -> LDS (pivot) -> MUFU seed + cubic Newton step -> multiplier
      // -> update of the next column, about 100 clk; the square roots are taken once at the end, one per lane.
      double a[IB];
#pragma unroll
      for (int c = 0; c < IB; ++c) a[c] = S_(j0 + lane, j0 + c);
      __syncwarp();
      double dmine = 1.0;
      int bad = 0;
#pragma unroll
      for (int j = 0; j < IB; ++j) {
        const double aj = (lane >= j) ? a[j] : 0.0;      // column j of L * sqrt(d_j); zero above the diagonal
        double* colj = &S_(j0, j0 + j);
        colj[lane] = aj;
        __syncwarp();
        const double d = colj[j];
        if (!(d > 0.0) && bad == 0) bad = j + 1;
        if (lane == j) dmine = d;
        double x0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
        const double e = fma(-d, x0, 1.0);
        const double inv = fma(x0, fma(e, e, e), x0);      // x0 (1 + e + e^2): relative error e^3 < 2^-58
        const double w = -aj * inv;
        if (((j + 1) & 1) && j + 1 < IB) a[j + 1] = fma(w, colj[j + 1], a[j + 1]);
#pragma unroll
        for (int c = (j + 2) & ~1; c < IB; c += 2) {
          const double2 lc = *reinterpret_cast<const double2*>(colj + c);
          a[c] = fma(w, lc.x, a[c]);
          a[c + 1] = fma(w, lc.y, a[c + 1]);
        }
      }
      const double rs = rsqrt(dmine);                      // 1 / L_jj of this lane's column
      double lg = 0.5 * log(dmine);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(FULL, lg, o);
      rdiag[j0 + lane] = rs;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < IB; c += 2) {
        const double2 rc = *reinterpret_cast<const double2*>(rdiag + j0 + c);
        // entries above the diagonal were stored as zeros already; the unscaled columns are re-read from S
        const double v0 = (lane == c) ? dmine * rs : S_(j0 + lane, j0 + c) * rc.x;
        const double v1 = (lane == c + 1) ? dmine * rs : S_(j0 + lane, j0 + c + 1) * rc.y;
        S_(j0 + lane, j0 + c) = v0;
        S_(j0 + lane, j0 + c + 1) = v1;
      }
      if (lane == 0) {
        s_logdet += lg;
        if (bad != 0 && s_info == 0) s_info = gidx0 + j0 + bad;
      }
    }
    __syncthreads();
    DBG_T();
    const int nrb = DB / IB - 1 - jb;  // 32-row blocks below the diagonal block
    if (nrb > 0) {
      // sub-panel: X * L_jj^T = Y, one row per thread kept in registers, right-looking so that the factor is read
      // down its columns (broadcast 16-byte loads) and the dependent chain per column is one DMUL + one DFMA
      if (tid < nrb * IB) {
        const int r = j0 + IB + tid;
        double x[IB];
#pragma unroll
        for (int c = 0; c < IB; ++c) x[c] = S_(r, j0 + c);
#pragma unroll
        for (int k = 0; k < IB; ++k) {
          const double xk = x[k] * rdiag[j0 + k];
          x[k] = xk;
          const double* colk = &S_(j0, j0 + k);
          if (((k + 1) & 1) && k + 1 < IB) x[k + 1] = fma(-xk, colk[k + 1], x[k + 1]);
#pragma unroll
          for (int c = (k + 2) & ~1; c < IB; c += 2) {
            const double2 lc = *reinterpret_cast<const double2*>(colk + c);
            x[c] = fma(-xk, lc.x, x[c]);
            x[c + 1] = fma(-xk, lc.y, x[c + 1]);
          }
        }
#pragma unroll
        for (int c = 0; c < IB; ++c) S_(r, j0 + c) = x[c];
      }
      __syncthreads();
      // trailing update of the lower block triangle on the tensor pipe: S[ib,cb] -= S[ib,jb] * S[cb,jb]^T
      const int npair = nrb * (nrb + 1) / 2;
      for (int pr = warp; pr < npair; pr += DW) {
        int ib = 0, cb = 0;  // decode pair index -> (ib >= cb), both in [0,nrb)
        {
          int q = pr;
          while (q > ib) { q -= (ib + 1); ++ib; }
          cb = q;
        }
        double acc[2][4][4];
        frag_zero<4>(acc);
        const int rr = j0 + IB + ib * IB, cc = j0 + IB + cb * IB;
        warp_mma32<4>(acc, &S_(rr, j0), LDS_, &S_(cc, j0), 1, LDS_, IB, lane);
        frag_apply<4>(acc, &S_(rr, cc), lane, -1.0, 1.0);
      }
      __syncthreads();
    }
    DBG_T();
  }

  // ---------------- phase 2: 32x32 inverses (warps 0-3) || L back to global (warps 4-7) ----------
  if (warp < 4) {
    // lane owns column `lane` of W = inv(L_ww) : forward substitution in registers
    const int j0 = warp * IB;
    double w[IB];
#pragma unroll
    for (int i = 0; i < IB; ++i) {
      double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) {
        const double lik = S_(j0 + i, j0 + k);
        if ((k & 3) == 0) s0 = fma(-lik, w[k], s0);
        else if ((k & 3) == 1) s1 = fma(-lik, w[k], s1);
        else if ((k & 3) == 2) s2 = fma(-lik, w[k], s2);
        else s3 = fma(-lik, w[k], s3);
      }
      w[i] = ((s0 + s1) + (s2 + s3)) * rdiag[j0 + i];
    }
#pragma unroll
    for (int i = 0; i < IB; ++i) T_(warp, i, lane) = w[i];
  } else {
    for (int idx = tid - 4 * 32; idx < DB * DB / 2; idx += DIAG_THREADS - 4 * 32) {
      const int r = (idx % (DB / 2)) * 2, c = idx / (DB / 2);
      double2 v = *reinterpret_cast<const double2*>(&S_(r, c));
      if (r < c) v.x = 0.0;
      if (r + 1 < c) v.y = 0.0;
      *reinterpret_cast<double2*>(Ablk + r + (int64_t)c * lda) = v;
    }
  }
  if (tid == 0) {
    *logdet_slot = s_logdet;
    if (s_info != 0) info_min(info, s_info);
  }
  __syncthreads();
  DBG_T();

  // ---------------- phase 3: in-place inverse of the block factor ----------------
  // diagonal blocks <- W_jj ; then W[>j, j] = -W22 * (L[>j, j] * W_jj), last block column first.
  for (int idx = tid; idx < 4 * IB * IB; idx += DIAG_THREADS) {
    const int b = idx / (IB * IB), e = idx % (IB * IB);
    const int r = e % IB, c = e / IB;
    S_(b * IB + r, b * IB + c) = T_(b, r, c);
  }
  __syncthreads();
  for (int jb = DB / IB - 2; jb >= 0; --jb) {
    const int j0 = jb * IB;
    const int nrb = DB / IB - 1 - jb;
    const int r0 = j0 + IB;
    // Work items are (row block rb, 16-column half): up to 6 warps busy.  Results are staged in registers
    // across a barrier because the products are formed in place.
    const int rb = warp >> 1, half = warp & 1;
    const bool active = rb < nrb;
    double acc[2][2][4];
    // (i) Y <- Y * W_jj
    if (active) {
      frag_zero<2>(acc);
      warp_mma32<2>(acc, &S_(r0 + rb * IB, j0), LDS_, &T_(jb, 0, half * 16), LDT, 1, IB, lane);
    }
    __syncthreads();
    if (active) frag_apply<2>(acc, &S_(r0 + rb * IB, j0 + half * 16), lane, 1.0, 0.0);
    __syncthreads();
    // (ii) X <- -W22 * Y : row block rb needs Y blocks 0..rb
    if (active) {
      frag_zero<2>(acc);
      warp_mma32<2>(acc, &S_(r0 + rb * IB, r0), LDS_, &S_(r0, j0 + half * 16), LDS_, 1, (rb + 1) * IB, lane);
    }
    __syncthreads();
    if (active) frag_apply<2>(acc, &S_(r0 + rb * IB, j0 + half * 16), lane, -1.0, 0.0);
    __syncthreads();
  }

  DBG_T();
  // ---------------- phase 4: inverse to global (dense 128x128, pitch 128) --------
  for (int idx = tid; idx < DB * DB / 2; idx += DIAG_THREADS) {
    const int r = (idx % (DB / 2)) * 2, c = idx / (DB / 2);
    *reinterpret_cast<double2*>(Dinv + r + c * DB) = *reinterpret_cast<const double2*>(&S_(r, c));
  }
  __syncthreads();
  DBG_T();
#undef DBG_T
}

// ---------------------------------------------------------------------------
// potrf_diag_ovl_kernel: the same factorisation with the block inverse taken OFF the end of the kernel.
// The phase clocks of potrf_diag_kernel (scripts/diag_clk.py, B200): load 3.3k | 4 x (LDL' of a 32x32 block by ONE warp
// 6.7-8.0k, sub-panel solve + trailing update 5.0-6.9k) | 32x32 inverses 4.0k | block inverse 18.3k | store 4.1k
// = 75.7k clk; while warp 0 factors a 32x32 block the other seven warps idle, and the whole inverse (22.3k) waits for
// the last block.  Here the inverse is built by ROW blocks, W[i,j] = -W_ii * Z[i,j], Z[i,j] = sum_{k=j}^{i-1} L[i,k] W[k,j],
// and row block i only needs rows <= i of L: warps 1-7 compute W_(i-1,i-1), row i-1 of the inverse and Z[i,.] in the
// shadow of LDL'(i) on warp 0 (named barrier 1 among themselves).  After the last block only W_33 and three 32x32
// products remain.  The strictly-upper 32x32 blocks of S are unused by the factorisation: block (j,i) of S holds W[i,j].
// ---------------------------------------------------------------------------
constexpr size_t DIAG_OVL_SMEM = DIAG_SMEM + size_t(3) * IB * LDT * sizeof(double);

template <int NT>
__device__ __forceinline__ void frag_store(double (&acc)[2][NT][4], double* C, int ldc, int lane, double alpha) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      double* p = C + (mi * 16 + g) + (ni * 8 + 2 * t) * ldc;
      p[0] = alpha * acc[mi][ni][0];
      p[ldc] = alpha * acc[mi][ni][1];
      p[8] = alpha * acc[mi][ni][2];
      p[8 + ldc] = alpha * acc[mi][ni][3];
    }
}

// one warp: T[b] = W = inv(L_bb).  Lane r solves x L_bb' = e_r by the right-looking substitution of the sub-panel solve
// (the factor is read down its columns with broadcast 16-byte loads; one DMUL + one DFMA on the chain per column):
// x = row r of L_bb^-T = column r of W.  The left-looking form of potrf_diag_kernel's phase 2 needs a 8-byte shared
// load per DFMA: 4.0k clk against 1.6k here.
__device__ __forceinline__ void inv32_warp(const double* S, double* T, const double* rdiag, int b, int lane) {
  const int j0 = b * IB;
  double x[IB];
#pragma unroll
  for (int c = 0; c < IB; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < IB; ++k) {
    const double xk = x[k] * rdiag[j0 + k];
    x[k] = xk;
    const double* colk = &S_(j0, j0 + k);
    if (((k + 1) & 1) && k + 1 < IB) x[k + 1] = fma(-xk, colk[k + 1], x[k + 1]);
#pragma unroll
    for (int c = (k + 2) & ~1; c < IB; c += 2) {
      const double2 lc = *reinterpret_cast<const double2*>(colk + c);
      x[c] = fma(-xk, lc.x, x[c]);
      x[c + 1] = fma(-xk, lc.y, x[c + 1]);
    }
  }
#pragma unroll
  for (int c = 0; c < IB; ++c) T_(b, c, lane) = x[c];
}

// rows [32*rb0, 32*rb1) of the block inverse to global (dense 128x128, pitch 128): diagonal blocks from T, block
// (bi > bj) from S block (bj, bi), zeros above; `nthr` threads with index `t`
__device__ __forceinline__ void inv_rows_to_global(const double* S, const double* T, double* __restrict__ Dinv, int rb0,
                                                   int rb1, int t, int nthr, int skipz) {
  const int hr = (rb1 - rb0) * IB / 2;               // 16-byte pieces per column
  for (int idx = t; idx < hr * DB; idx += nthr) {
    const int r = rb0 * IB + (idx % hr) * 2, c = idx / hr;
    const int bi = r / IB, bj = c / IB, ri = r % IB, ci = c % IB;
    if (skipz && bi < bj) continue;                   // a zero block the (zero-filled) buffer already holds
    double2 v = make_double2(0.0, 0.0);
    if (bi == bj) v = *reinterpret_cast<const double2*>(&T_(bi, ri, ci));
    else if (bi > bj) v = *reinterpret_cast<const double2*>(&S_(bj * IB + ri, bi * IB + ci));
    *reinterpret_cast<double2*>(Dinv + r + c * DB) = v;
  }
}

// shared -> global through the TMA engine (async proxy, bulk_group completion): the store neither occupies the
// load/store pipe of the issuing warp nor stalls later shared-memory instructions behind a queue of global stores
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ASYNC_ST = false: L and the inverse leave through ordinary 16-byte stores after the factorisation (the L tile by warps
// 4-7 beside W_33, the inverse by everybody at the end).  One SM drains about 32 bytes per clock: the 2 x 128 KB cost
// 8k clk, and shared-memory instructions issued behind them wait (measured: the 1.5k-clk last row block took 12.7k clk
// behind 224 KB of stores).  ASYNC_ST = true: every piece leaves by cp.async.bulk the moment it is final - block column
// jb of L after its sub-panel solve, row block i-1 of the inverse after step S2 of iteration i, the zero blocks above
// the diagonal right at the start - and only the last row block of the inverse is still to go when the arithmetic ends.
template <bool ASYNC_ST>
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_ovl_kernel(double* __restrict__ Ablk, int64_t lda, double* __restrict__ Dinv,
                      double* __restrict__ logdet_slot, int* __restrict__ info, int gidx0, long long* dbg_clk, int skipz) {
  // skipz: the six 32x32 blocks above the block diagonal are not stored - 48 KB of zeros of the inverse (its buffer is
  // zero-filled when allocated, ensure_zero) and 48 KB of the L tile (nobody reads a diagonal tile above its diagonal: the
  // diagonal blocks enter every later product through their inverses).  The kernel ends with the drain of its own stores
  // through one SM (profiles/r2_diag_clk.md); this takes 96 of the 256 KB away.
  extern __shared__ __align__(16) double dsm[];
  int dbg_i = 0;
#define DBG_T() do { if (dbg_clk && threadIdx.x == 0) dbg_clk[dbg_i++] = clock64(); } while (0)
#define SIDE_BAR() asm volatile("bar.sync 1, 224;\n" ::: "memory")
  DBG_T();
  double* S = dsm;                    // DB x LDS_
  double* T = dsm + DB * LDS_;        // 4 x (IB x LDT): W_jj
  double* rdiag = T + 4 * IB * LDT;   // DB reciprocals of the diagonal of L
  double* Zs = rdiag + DB;            // 3 x (IB x LDT): Z[i, 0..2] of the row block in progress
  __shared__ double s_logdet;
  __shared__ int s_info;
  __shared__ __align__(16) double zcol[IB];   // ASYNC_ST: source of the zero blocks

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

#pragma unroll 8
  for (int idx = tid; idx < DB * DB / 2; idx += DIAG_THREADS) {
    const int r = (idx % (DB / 2)) * 2, c = idx / (DB / 2);
    double2 v = make_double2(0.0, 0.0);
    if (r + 1 >= c) v = *reinterpret_cast<const double2*>(Ablk + r + (int64_t)c * lda);
    if (r < c) v.x = 0.0;
    *reinterpret_cast<double2*>(&S_(r, c)) = v;
  }
  if (tid == 0) { s_logdet = 0.0; s_info = 0; }
  if (ASYNC_ST) {
    if (tid < IB) zcol[tid] = 0.0;
    fence_async_smem();
  }
  __syncthreads();
  DBG_T();
  if (ASYNC_ST && !skipz && warp >= 1 && warp <= 6) {
    // the six 32x32 blocks above the block diagonal, of the L tile and of the inverse: zeros (lane = column)
    const int ub = warp - 1;                         // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    const int bj = (ub < 3) ? ub + 1 : (ub < 5 ? ub - 1 : 3), bi = (ub < 3) ? 0 : (ub < 5 ? 1 : 2);
    fence_async_smem();
    bulk_store(Ablk + bi * IB + (int64_t)(bj * IB + lane) * lda, zcol, IB * sizeof(double));
    bulk_store(Dinv + bi * IB + (bj * IB + lane) * DB, zcol, IB * sizeof(double));
    bulk_commit();
  }

  for (int jb = 0; jb < DB / IB; ++jb) {
    const int j0 = jb * IB;
    if (warp == 0) {
      // 32x32 diagonal block by ONE warp, square-root-free (see potrf_diag_kernel)
      double a[IB];
#pragma unroll
      for (int c = 0; c < IB; ++c) a[c] = S_(j0 + lane, j0 + c);
      __syncwarp();
      double dmine = 1.0;
      int bad = 0;
#pragma unroll
      for (int j = 0; j < IB; ++j) {
        const double aj = (lane >= j) ? a[j] : 0.0;
        double* colj = &S_(j0, j0 + j);
        colj[lane] = aj;
        __syncwarp();
        const double d = colj[j];
        if (!(d > 0.0) && bad == 0) bad = j + 1;
        if (lane == j) dmine = d;
        double x0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
        const double e = fma(-d, x0, 1.0);
        const double inv = fma(x0, fma(e, e, e), x0);
        const double w = -aj * inv;
        if (((j + 1) & 1) && j + 1 < IB) a[j + 1] = fma(w, colj[j + 1], a[j + 1]);
#pragma unroll
        for (int c = (j + 2) & ~1; c < IB; c += 2) {
          const double2 lc = *reinterpret_cast<const double2*>(colj + c);
          a[c] = fma(w, lc.x, a[c]);
          a[c + 1] = fma(w, lc.y, a[c + 1]);
        }
      }
      const double rs = rsqrt(dmine);
      double lg = 0.5 * log(dmine);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(FULL, lg, o);
      rdiag[j0 + lane] = rs;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < IB; c += 2) {
        const double2 rc = *reinterpret_cast<const double2*>(rdiag + j0 + c);
        const double v0 = (lane == c) ? dmine * rs : S_(j0 + lane, j0 + c) * rc.x;
        const double v1 = (lane == c + 1) ? dmine * rs : S_(j0 + lane, j0 + c + 1) * rc.y;
        S_(j0 + lane, j0 + c) = v0;
        S_(j0 + lane, j0 + c + 1) = v1;
      }
      if (lane == 0) {
        s_logdet += lg;
        if (bad != 0 && s_info == 0) s_info = gidx0 + j0 + bad;
      }
    } else if (jb >= 1) {
      // ---- in the shadow of LDL'(jb): the inverse up to row block jb-1, and Z of row block jb ----
      const int i = jb, sw = warp - 1;               // sw = 0..6
      const int j = sw >> 1, half = sw & 1;
      if (sw == 0) inv32_warp(S, T, rdiag, i - 1, lane);
      if (ASYNC_ST) fence_async_smem();
      SIDE_BAR();
      if (j < i - 1) {                               // W[i-1, j] = -W_(i-1,i-1) * Z[i-1, j]  ->  S block (j, i-1)
        double acc[2][2][4];
        frag_zero<2>(acc);
        warp_mma32<2>(acc, &T_(i - 1, 0, 0), LDT, Zs + j * IB * LDT + half * 16 * LDT, LDT, 1, IB, lane);
        frag_store<2>(acc, &S_(j * IB, (i - 1) * IB + half * 16), LDS_, lane, -1.0);
      }
      if (ASYNC_ST) fence_async_smem();
      SIDE_BAR();
      if (ASYNC_ST && sw == 6) {
        // row block i-1 of the inverse is final: W[i-1, j] from S block (j, i-1), W_(i-1,i-1) from T (lane = column)
        fence_async_smem();
        for (int jj = 0; jj < i - 1; ++jj)
          bulk_store(Dinv + (i - 1) * IB + (jj * IB + lane) * DB, &S_(jj * IB, (i - 1) * IB + lane), IB * sizeof(double));
        bulk_store(Dinv + (i - 1) * IB + ((i - 1) * IB + lane) * DB, &T_(i - 1, 0, lane), IB * sizeof(double));
        bulk_commit();
      }
      if (j < i) {                                   // Z[i, j] = sum_{k=j}^{i-1} L[i,k] W[k,j]  ->  Zs[j]
        double acc[2][2][4];
        frag_zero<2>(acc);
        warp_mma32<2>(acc, &S_(i * IB, j * IB), LDS_, &T_(j, 0, half * 16), LDT, 1, IB, lane);
        for (int k = j + 1; k < i; ++k)
          warp_mma32<2>(acc, &S_(i * IB, k * IB), LDS_, &S_(j * IB, k * IB + half * 16), LDS_, 1, IB, lane);
        frag_store<2>(acc, Zs + j * IB * LDT + half * 16 * LDT, LDT, lane, 1.0);
      }
    }
    if (ASYNC_ST) fence_async_smem();
    __syncthreads();
    DBG_T();
    const int nrb = DB / IB - 1 - jb;
    if (nrb > 0) {
      if (tid < nrb * IB) {
        const int r = j0 + IB + tid;
        double x[IB];
#pragma unroll
        for (int c = 0; c < IB; ++c) x[c] = S_(r, j0 + c);
#pragma unroll
        for (int k = 0; k < IB; ++k) {
          const double xk = x[k] * rdiag[j0 + k];
          x[k] = xk;
          const double* colk = &S_(j0, j0 + k);
          if (((k + 1) & 1) && k + 1 < IB) x[k + 1] = fma(-xk, colk[k + 1], x[k + 1]);
#pragma unroll
          for (int c = (k + 2) & ~1; c < IB; c += 2) {
            const double2 lc = *reinterpret_cast<const double2*>(colk + c);
            x[c] = fma(-xk, lc.x, x[c]);
            x[c + 1] = fma(-xk, lc.y, x[c + 1]);
          }
        }
#pragma unroll
        for (int c = 0; c < IB; ++c) S_(r, j0 + c) = x[c];
      }
      if (ASYNC_ST) fence_async_smem();
      __syncthreads();
    }
    if (ASYNC_ST && warp == DW - 1) {
      // block column jb of L is final (rows from its diagonal block down; zeros above the diagonal inside that block
      // were stored by the LDL' warp): one bulk store per column, issued by the warp the trailing update leaves idle
      fence_async_smem();
      bulk_store(Ablk + j0 + (int64_t)(j0 + lane) * lda, &S_(j0, j0 + lane), (DB - j0) * sizeof(double));
      bulk_commit();
    }
    if (nrb > 0) {
      const int npair = nrb * (nrb + 1) / 2;
      for (int pr = warp; pr < npair; pr += DW) {
        int ib = 0, cb = 0;
        {
          int q = pr;
          while (q > ib) { q -= (ib + 1); ++ib; }
          cb = q;
        }
        double acc[2][4][4];
        frag_zero<4>(acc);
        const int rr = j0 + IB + ib * IB, cc = j0 + IB + cb * IB;
        warp_mma32<4>(acc, &S_(rr, j0), LDS_, &S_(cc, j0), 1, LDS_, IB, lane);
        frag_apply<4>(acc, &S_(rr, cc), lane, -1.0, 1.0);
      }
      __syncthreads();
    }
    DBG_T();
  }

  // ---- W_33 (warp 0) || ordinary stores only: L back to global (warps 4-7) ----
  if (warp == 0) {
    inv32_warp(S, T, rdiag, DB / IB - 1, lane);
  } else if (!ASYNC_ST && warp >= 4) {
    for (int idx = tid - 4 * 32; idx < DB * DB / 2; idx += DIAG_THREADS - 4 * 32) {
      const int r = (idx % (DB / 2)) * 2, c = idx / (DB / 2);
      if (skipz && r / IB < c / IB) continue;
      double2 v = *reinterpret_cast<const double2*>(&S_(r, c));
      if (r < c) v.x = 0.0;
      if (r + 1 < c) v.y = 0.0;
      *reinterpret_cast<double2*>(Ablk + r + (int64_t)c * lda) = v;
    }
  }
  if (tid == 0) {
    *logdet_slot = s_logdet;
    if (s_info != 0) info_min(info, s_info);
  }
  __syncthreads();
  DBG_T();
  // ---- last row block of the inverse: W[3, j] = -W_33 * Z[3, j] -> S block (j, 3); twelve 32x8 tasks over 8 warps ----
  for (int task = warp; task < 3 * 4; task += DW) {
    const int j = task >> 2, q = task & 3;
    double acc[2][1][4];
    frag_zero<1>(acc);
    warp_mma32<1>(acc, &T_(DB / IB - 1, 0, 0), LDT, Zs + j * IB * LDT + q * 8 * LDT, LDT, 1, IB, lane);
    frag_store<1>(acc, &S_(j * IB, (DB / IB - 1) * IB + q * 8), LDS_, lane, -1.0);
  }
  if (ASYNC_ST) fence_async_smem();
  __syncthreads();
  DBG_T();
  if (ASYNC_ST) {
    // ---- last row block of the inverse: warp j < 3 sends W[3, j], warp 3 sends W_33; then every thread that issued
    //      bulk stores waits until the engine has read its shared-memory sources ----
    if (warp < DB / IB) {
      fence_async_smem();
      const double* src = (warp < DB / IB - 1) ? &S_(warp * IB, (DB / IB - 1) * IB + lane) : &T_(DB / IB - 1, 0, lane);
      bulk_store(Dinv + (DB / IB - 1) * IB + (warp * IB + lane) * DB, src, IB * sizeof(double));
      bulk_commit();
    }
    bulk_wait_read_all();
  } else {
    inv_rows_to_global(S, T, Dinv, 0, DB / IB, tid, DIAG_THREADS, skipz);
  }
  __syncthreads();
  DBG_T();
#undef DBG_T
#undef SIDE_BAR
}

// ---------------------------------------------------------------------------
// small_nt_kernel: the products ON the dependent chain, where the operands are one or two 128-row tiles and the time is
// launch + memory latency, not arithmetic:  C(32x32 block bi,bj) (op)= A(rows 32bi..) * B(rows 32bj..)'  over K.
// The chain timeline (GPK_CHAIN_DUMP) showed the two "head" products after every diagonal block - tile (p+1,p) times
// Dinv_p', then the rank-128 update of tile (p+1,p+1) - at 10-13 us EACH through the pipelined tile kernel of
// gemm_nt.cu (four 32-row-strip CTAs, a 4-stage ring of 16-column slabs: eight dependent slab round trips for K=128).
// Here a 128x128 tile is cut into sixteen 32x32 blocks, one CTA of four warps each; a CTA requests its WHOLE operand
// chunk (32 rows x 128 columns of A and of B) at once, waits once, and runs 16 DMMA k-steps per warp from shared
// memory.  K > 128 (the rank-512 update of the next diagonal tile at a level-1 hand-over) walks chunks of 128 through
// two stages.  Sixteen CTAs cannot work in place (a sibling would overwrite columns still being read), so the
// solved tile goes to a scratch tile and the CTAs of block column 0 of the FOLLOWING update copy it home
// (copy_dst) - the update reads the scratch tile as both operands anyway.
//   mode 0: C = A B'      mode 1: C -= A B'
//   tri   : only blocks bi >= bj, diagonal blocks store row >= col
// Replaces (on the chain-bound part of the factorisation) the same reference lines as launch_gemm_nt:
// LAPACK dpotrf called at /root/reference/pyGPs/Core/tools.py:61.
// ---------------------------------------------------------------------------
constexpr int SN_P = 36;                        // shared-memory pitch of a 32-row operand block: == 4 mod 16, see LDT
constexpr int SN_KC = 128;                      // contraction chunk
constexpr int SN_STAGE = 2 * SN_KC * SN_P;      // doubles per stage: A block, then B block
constexpr int SN_THREADS = 128;
constexpr size_t SN_SMEM1 = size_t(SN_STAGE) * sizeof(double);       // K == 128
constexpr size_t SN_SMEM2 = 2 * SN_SMEM1;                            // K  > 128

// 32 rows x 128 columns of a column-major matrix -> shared memory (pitch SN_P): 16 pieces of 16 bytes per column
__device__ __forceinline__ void sn_load(double* dst, const double* __restrict__ src, int64_t ld, int tid) {
#pragma unroll
  for (int i = 0; i < (32 * SN_KC / 2) / SN_THREADS; ++i) {
    const int piece = tid + i * SN_THREADS;
    const int c = piece >> 4, r = (piece & 15) * 2;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + r + c * SN_P);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + r + (int64_t)c * ld) : "memory");
  }
}

// BREG (K == 128 only): the B operand never touches shared memory.  A warp needs just 8 rows of B (its 8 output columns),
// 32 doubles per lane for the whole contraction, fetched straight into the DMMA fragment registers while the A block
// arrives by cp.async.  Shared memory per CTA drops from 74 KB to 37 KB, which is what lets these CTAs start BESIDE a
// resident trailing-update CTA (172 KB of the SM's 227 KB): with 74 KB the chain timeline showed the head products
// waiting up to 20 us for the first wave of a freshly launched update to retire.
template <bool BREG>
__global__ void __launch_bounds__(SN_THREADS) small_nt_kernel(SmallArgs a) {
  extern __shared__ __align__(16) double sn_sm[];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (a.tri && bj > bi) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* Ag = a.A + 32 * bi;
  const double* Bg = a.B + 32 * bj;
  const int nch = BREG ? 1 : a.K / SN_KC;
  const int nst = nch > 1 ? 2 : 1;
  double breg[BREG ? 2 * (SN_KC / 8) : 1];
  if (BREG) {
    sn_load(sn_sm, Ag, a.lda, tid);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const double* Bw = Bg + 8 * warp + g;           // row 8*warp + g of this block of B
#pragma unroll
    for (int ks = 0; ks < SN_KC / 8; ++ks) {
      breg[2 * ks] = Bw[(int64_t)(8 * ks + t) * a.ldb];
      breg[2 * ks + 1] = Bw[(int64_t)(8 * ks + t + 4) * a.ldb];
    }
  } else {
    for (int s = 0; s < nst; ++s) {
      double* stg = sn_sm + s * SN_STAGE;
      sn_load(stg, Ag + (int64_t)s * SN_KC * a.lda, a.lda, tid);
      sn_load(stg + SN_KC * SN_P, Bg + (int64_t)s * SN_KC * a.ldb, a.ldb, tid);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
  }
  // this thread's entries of the 32x32 block: rows mi*16+g (+8), columns 8*warp + 2t (+1)
  double* Cb = a.C + 32 * bi + (int64_t)(32 * bj + 8 * warp + 2 * t) * a.ldc;
  double cold[2][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    cold[mi][0] = cold[mi][1] = cold[mi][2] = cold[mi][3] = 0.0;
    if (a.mode == 1) {                            // requested while the operands are in flight
      const int r0 = mi * 16 + g;
      cold[mi][0] = Cb[r0]; cold[mi][1] = Cb[r0 + a.ldc];
      cold[mi][2] = Cb[r0 + 8]; cold[mi][3] = Cb[r0 + 8 + a.ldc];
    }
  }
  double acc[2][1][4];
  frag_zero<1>(acc);
  if (BREG) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < SN_KC / 8; ++ks) {
      const int k0 = 8 * ks;
      double af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        af[mi][0] = sn_sm[(mi * 16 + g) + (k0 + t) * SN_P];
        af[mi][1] = sn_sm[(mi * 16 + g + 8) + (k0 + t) * SN_P];
        af[mi][2] = sn_sm[(mi * 16 + g) + (k0 + t + 4) * SN_P];
        af[mi][3] = sn_sm[(mi * 16 + g + 8) + (k0 + t + 4) * SN_P];
      }
      dmma8(acc[0][0], af[0], breg[2 * ks], breg[2 * ks + 1]);
      dmma8(acc[1][0], af[1], breg[2 * ks], breg[2 * ks + 1]);
    }
  }
  for (int c = 0; !BREG && c < nch; ++c) {
    if (c + 1 < nch) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    const double* stg = sn_sm + (c & 1) * SN_STAGE;
    warp_mma32<1>(acc, stg, SN_P, stg + SN_KC * SN_P + 8 * warp, 1, SN_P, SN_KC, lane);
    if (c + 2 < nch) {
      __syncthreads();                            // every warp is done with this stage
      double* nxt = sn_sm + (c & 1) * SN_STAGE;
      sn_load(nxt, Ag + (int64_t)(c + 2) * SN_KC * a.lda, a.lda, tid);
      sn_load(nxt + SN_KC * SN_P, Bg + (int64_t)(c + 2) * SN_KC * a.ldb, a.ldb, tid);
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
  }
  const bool diag_blk = a.tri && bi == bj;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int r0 = mi * 16 + g, r1 = r0 + 8, c0 = 8 * warp + 2 * t, c1 = c0 + 1;
    const double v0 = (a.mode == 1) ? cold[mi][0] - acc[mi][0][0] : acc[mi][0][0];
    const double v1 = (a.mode == 1) ? cold[mi][1] - acc[mi][0][1] : acc[mi][0][1];
    const double v2 = (a.mode == 1) ? cold[mi][2] - acc[mi][0][2] : acc[mi][0][2];
    const double v3 = (a.mode == 1) ? cold[mi][3] - acc[mi][0][3] : acc[mi][0][3];
    if (!diag_blk || r0 >= c0) Cb[r0] = v0;
    if (!diag_blk || r0 >= c1) Cb[r0 + a.ldc] = v1;
    if (!diag_blk || r1 >= c0) Cb[r1] = v2;
    if (!diag_blk || r1 >= c1) Cb[r1 + a.ldc] = v3;
  }
  if (a.copy_dst && bj == 0) {
    // K == 128 (checked by the launcher): stage 0 still holds this CTA's 32 x 128 strip of A
    double* dst = a.copy_dst + 32 * bi;
#pragma unroll
    for (int i = 0; i < (32 * SN_KC / 2) / SN_THREADS; ++i) {
      const int piece = tid + i * SN_THREADS;
      const int c = piece >> 4, r = (piece & 15) * 2;
      *reinterpret_cast<double2*>(dst + r + (int64_t)c * a.ld_copy) = *reinterpret_cast<const double2*>(sn_sm + r + c * SN_P);
    }
  }
}

// one 128x128 tile of C: sixteen (tri: ten working) CTAs
int launch_small_nt(Handle* h, cudaStream_t st, const SmallArgs& a) {
  if (a.K <= 0 || a.K % SN_KC != 0 || (a.mode != 0 && a.mode != 1)) return GPK_ERR_ARG;
  if (a.copy_dst && a.K != SN_KC) return GPK_ERR_ARG;
  if (a.breg && a.K == SN_KC)
    small_nt_kernel<true><<<dim3(NB / 32, NB / 32), SN_THREADS, SN_SMEM1 / 2, st>>>(a);
  else
    small_nt_kernel<false><<<dim3(NB / 32, NB / 32), SN_THREADS, a.K > SN_KC ? SN_SMEM2 : SN_SMEM1, st>>>(a);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// single right-hand-side substitution steps.  512 threads; each 128x128 tile is pulled into shared
// memory with cp.async (all 128 KB in flight at once - these steps are pure latency), then reduced there.
// ---------------------------------------------------------------------------
constexpr int TRSV_THREADS = 512;
constexpr size_t TRSV_SMEM = size_t(NB) * NB * sizeof(double);

__device__ __forceinline__ void tile_to_smem(double* tile, const double* __restrict__ M, int64_t pitch) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < (NB * NB / 2) / TRSV_THREADS; ++i) {
    const int ch = tid + i * TRSV_THREADS;   // 16-byte chunk id: 64 per column
    const int c = ch >> 6, r = (ch & 63) * 2;
    unsigned sa = (unsigned)__cvta_generic_to_shared(tile + r + c * NB);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(M + r + (int64_t)c * pitch) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
}

// out[r] = sum_c tile[r + c*NB] * v[c]     (4 column groups of 32, then a shared-memory reduction)
__device__ __forceinline__ void smem_gemv_n(const double* tile, const double* v, double* out, double* red) {
  const int tid = threadIdx.x;
  const int r = tid & (NB - 1), cg = tid >> 7;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
  for (int c = cg * 32; c < cg * 32 + 32; c += 2) {
    a0 = fma(tile[r + c * NB], v[c], a0);
    a1 = fma(tile[r + (c + 1) * NB], v[c + 1], a1);
  }
  red[cg * NB + r] = a0 + a1;
  __syncthreads();
  if (tid < NB) out[tid] = (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]);
  __syncthreads();
}

// out[c] = sum_r tile[r + c*NB] * v[r]     (one warp per column, lanes along r, shuffle reduction)
__device__ __forceinline__ void smem_gemv_t(const double* tile, const double* v, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double v0 = v[lane], v1 = v[lane + 32], v2 = v[lane + 64], v3 = v[lane + 96];
#pragma unroll
  for (int c = warp; c < NB; c += TRSV_THREADS / 32) {
    const double* col = tile + c * NB;
    double s = fma(col[lane], v0, fma(col[lane + 32], v1, fma(col[lane + 64], v2, col[lane + 96] * v3)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) out[c] = s;
  }
  __syncthreads();
}

// forward step k (grid T-k CTAs): every CTA forms z_k = Dinv_k * b_k in shared memory; CTA 0 stores it;
// CTA i>0 updates b_{k+i} -= L[k+i, k] * z_k.
__global__ void __launch_bounds__(TRSV_THREADS, 1) trsv_fwd_kernel(const double* __restrict__ A, int64_t lda,
                                                                   const double* __restrict__ Dinv,
                                                                   double* __restrict__ b, double* __restrict__ z,
                                                                   int k) {
  extern __shared__ __align__(16) double tile[];
  __shared__ double sb[NB], sz[NB], so[NB], red[4 * NB];
  const int tid = threadIdx.x;
  if (tid < NB) sb[tid] = b[(int64_t)k * NB + tid];
  tile_to_smem(tile, Dinv + (int64_t)k * NB * NB, NB);
  smem_gemv_n(tile, sb, sz, red);
  const int i = blockIdx.x;
  if (i == 0) {
    if (tid < NB) z[(int64_t)k * NB + tid] = sz[tid];
  } else {
    tile_to_smem(tile, A + (int64_t)(k + i) * NB + (int64_t)k * NB * lda, lda);
    smem_gemv_n(tile, sz, so, red);
    if (tid < NB) b[(int64_t)(k + i) * NB + tid] -= so[tid];
  }
}

// backward step k (grid k+1 CTAs): every CTA forms x_k = Dinv_k^T * z_k; CTA k stores it;
// CTA j<k updates z_j -= L[k, j]^T * x_k.
__global__ void __launch_bounds__(TRSV_THREADS, 1) trsv_bwd_kernel(const double* __restrict__ A, int64_t lda,
                                                                   const double* __restrict__ Dinv,
                                                                   double* __restrict__ z, double* __restrict__ x,
                                                                   int k) {
  extern __shared__ __align__(16) double tile[];
  __shared__ double sz[NB], sx[NB], so[NB];
  const int tid = threadIdx.x;
  if (tid < NB) sz[tid] = z[(int64_t)k * NB + tid];
  tile_to_smem(tile, Dinv + (int64_t)k * NB * NB, NB);
  smem_gemv_t(tile, sz, sx);
  const int j = blockIdx.x;
  if (j == k) {
    if (tid < NB) x[(int64_t)k * NB + tid] = sx[tid];
  } else {
    tile_to_smem(tile, A + (int64_t)k * NB + (int64_t)j * NB * lda, lda);
    smem_gemv_t(tile, sx, so);
    if (tid < NB) z[(int64_t)j * NB + tid] -= so[tid];
  }
}

// ---------------------------------------------------------------------------
// Whole backward substitution x = L^-T z in ONE cooperative launch (T <= number of SMs).
// CTA j owns block j: it streams the tiles L[k, j], k = T-1 .. j+1, of its block column through a ring of
// 32-column chunks (cp.async, prefetched as far as the ring allows - the factor is complete, only the x_k are
// not), applies z_j -= L[k,j]' x_k as soon as CTA k has published x_k (epoch-stamped flag, release/acquire), and
// finishes with x_j = Dinv_j' z_j.  The dependent chain per block step is flag -> x_k -> two 128x128 transposed
// GEMVs from shared memory -> publish: a few microseconds, against ~10 us per step for one launch per step.
// ---------------------------------------------------------------------------
constexpr int BW_CH = 32;                                  // columns per chunk
constexpr int BW_RING = 4;                                 // L chunks in flight: one whole 128x128 tile
constexpr int BW_DINV = (128 + 96 + 64 + 32) * BW_CH;      // Dinv_j, lower triangle by 32-column chunks (rows >= 32q)
constexpr size_t BW_SMEM = (size_t(BW_RING) * BW_CH * NB + BW_DINV) * sizeof(double);   // 128 KB + 80 KB

__global__ void __launch_bounds__(TRSV_THREADS, 1) trsv_bwd_persistent_kernel(
    const double* __restrict__ A, int64_t lda, const double* __restrict__ Dinv, const double* __restrict__ z,
    double* __restrict__ x, int* __restrict__ flags, int epoch, int T) {
  extern __shared__ __align__(16) double bw_sm[];
  double* dinv_sm = bw_sm;                                 // resident from the start: nothing to fetch on the chain
  double* ring = bw_sm + BW_DINV;
  __shared__ double zacc[NB], zfin[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;
  const int ntiles = T - 1 - j;                            // tiles of L in this block column
  const int nchunks = 4 * ntiles;
  if (tid < NB) zacc[tid] = 0.0;

  // Dinv_j: chunk q = columns 32q..32q+31, rows 32q..127 (the block is lower triangular), pitch 128 - 32q
  {
    const double* src = Dinv + (int64_t)j * NB * NB;
    int off = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rows = NB - 32 * q, pieces = rows / 2;     // 16-byte pieces per column
      for (int ch = tid; ch < BW_CH * pieces; ch += TRSV_THREADS) {
        const int c = ch / pieces, r = (ch % pieces) * 2;
        unsigned sa = (unsigned)__cvta_generic_to_shared(dinv_sm + off + r + c * rows);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + (32 * q + r) + (int64_t)(32 * q + c) * NB) : "memory");
      }
      off += rows * BW_CH;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  auto issue = [&](int g) {
    if (g < nchunks) {
      const int t = g >> 2, q = g & 3;
      const int k = T - 1 - t;
      const double* src = A + (int64_t)k * NB + (int64_t)(j * NB + q * BW_CH) * lda;
      double* dst = ring + (size_t)(g % BW_RING) * BW_CH * NB;
#pragma unroll
      for (int i = 0; i < (BW_CH * NB / 2) / TRSV_THREADS; ++i) {
        const int ch = tid + i * TRSV_THREADS;             // 16-byte piece: 64 per column
        const int c = ch >> 6, r = (ch & 63) * 2;
        unsigned sa = (unsigned)__cvta_generic_to_shared(dst + r + c * NB);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + r + (int64_t)c * lda) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
#pragma unroll
  for (int g = 0; g < BW_RING - 1; ++g) issue(g);

  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;           // this lane's entries of x_k
  constexpr int CPW = BW_CH / (TRSV_THREADS / 32);         // columns per warp per chunk
  for (int g = 0; g < nchunks; ++g) {
    const int q = g & 3;
    asm volatile("cp.async.wait_group %0;\n" ::"n"(BW_RING - 2) : "memory");
    __syncthreads();                                       // chunk g has landed; chunk g-1's slot is free
    issue(g + BW_RING - 1);                                // requested BEFORE blocking on x_k: the whole tile is in
    if (q == 0) {                                          // flight (or here) when its x_k is published
      const int k = T - 1 - (g >> 2);
      if (tid == 0) {
        int f;
        do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(f) : "l"(flags + k) : "memory"); } while (f != epoch);
      }
      __syncthreads();
      const double* xk = x + (int64_t)k * NB;
      v0 = __ldcg(xk + lane); v1 = __ldcg(xk + lane + 32); v2 = __ldcg(xk + lane + 64); v3 = __ldcg(xk + lane + 96);
    }
    const double* ch = ring + (size_t)(g % BW_RING) * BW_CH * NB;
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
      const int c = warp * CPW + i;
      const double* col = ch + c * NB;
      double sacc = fma(col[lane], v0, fma(col[lane + 32], v1, fma(col[lane + 64], v2, col[lane + 96] * v3)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(FULL, sacc, o);
      if (lane == 0) zacc[q * BW_CH + c] += sacc;
    }
  }
  // x_j = Dinv_j' z_j from the resident copy
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  if (tid < NB) zfin[tid] = z[(int64_t)j * NB + tid] - zacc[tid];
  __syncthreads();
  {
    int off = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rows = NB - 32 * q;
#pragma unroll
      for (int i = 0; i < CPW; ++i) {
        const int c = warp * CPW + i;
        const double* col = dinv_sm + off + c * rows;      // col[r'] = Dinv[32q + r', 32q + c]
        double sacc = 0.0;
#pragma unroll
        for (int b = 0; b < 4 - q; ++b) sacc = fma(col[lane + 32 * b], zfin[32 * q + lane + 32 * b], sacc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(FULL, sacc, o);
        if (lane == 0) x[(int64_t)j * NB + q * BW_CH + c] = sacc;
      }
      off += rows * BW_CH;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(flags + j), "r"(epoch) : "memory");
  }
}

int launch_trsv_bwd_all(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* z,
                        double* x, int T) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const char* e = getenv("GPK_TRSV_PERSIST");            // 0: one launch per block step (A/B runs, tests)
  const int persist = e ? atoi(e) : 1;
  if (!persist || T > sms || T < 2) {
    for (int k = T - 1; k >= 0; --k) GPK_TRY(launch_trsv_bwd(h, st, A, lda, Dinv, z, x, k, T));
    return 0;
  }
  if (!h->dFlags) {
    GPK_CK(h, cudaMalloc((void**)&h->dFlags, 1024 * sizeof(int)));
    GPK_CK(h, cudaMemsetAsync(h->dFlags, 0, 1024 * sizeof(int), st));
    h->flag_epoch = 0;
  }
  int epoch = ++h->flag_epoch;
  int* flags = h->dFlags;
  void* args[] = {(void*)&A, (void*)&lda, (void*)&Dinv, (void*)&z, (void*)&x, (void*)&flags, (void*)&epoch, (void*)&T};
  // All T CTAs must be co-resident (they wait on each other's flags), which only a cooperative launch guarantees.
  // If the device cannot grant it (SMs held by another context, MPS limits), fall back to one launch per block step -
  // the same arithmetic on the same GPU, never a CPU path.
  const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)trsv_bwd_persistent_kernel, dim3(T),
                                                     dim3(TRSV_THREADS), args, BW_SMEM, st);
  if (ce != cudaSuccess) {
    (void)cudaGetLastError();
    for (int k = T - 1; k >= 0; --k) GPK_TRY(launch_trsv_bwd(h, st, A, lda, Dinv, z, x, k, T));
    return 0;
  }
  h->stats.launches++;
  return 0;
}

int diag_init(Handle* h) {
  GPK_CK(h, cudaFuncSetAttribute(potrf_diag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(potrf_diag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(potrf_diag_ovl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_OVL_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(potrf_diag_ovl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_OVL_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(trsv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSV_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(trsv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSV_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(trsv_bwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BW_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(small_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SN_SMEM2));
  return 0;
}

int launch_diag(Handle* h, cudaStream_t st, double* Ablk, int64_t lda, double* Dinv, double* logdet_slot, int* info,
                int gidx0, long long* dbg_clk) {
  const int ovl = env_int("GPK_DIAG_OVL", 1);             // 0: the inverse after the factorisation (A/B runs)
  const int skipz = env_int("GPK_DIAG_SKIPZ", 0);        // 1: the zero blocks above the block diagonal are not stored
  if (ovl >= 2)       // 2: every finished piece leaves at once by cp.async.bulk
    potrf_diag_ovl_kernel<true><<<1, DIAG_THREADS, DIAG_OVL_SMEM, st>>>(Ablk, lda, Dinv, logdet_slot, info, gidx0, dbg_clk, skipz);
  else if (ovl == 1)  // 1: the inverse by row blocks in the shadow of the 32x32 factorisations, ordinary stores
    potrf_diag_ovl_kernel<false><<<1, DIAG_THREADS, DIAG_OVL_SMEM, st>>>(Ablk, lda, Dinv, logdet_slot, info, gidx0, dbg_clk, skipz);
  else
    potrf_diag_kernel<false><<<1, DIAG_THREADS, DIAG_SMEM, st>>>(Ablk, lda, Dinv, logdet_slot, info, gidx0, dbg_clk);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// Dinv_k = inv(L_kk), logdet_parts[k] = sum(log diag L_kk) for all T diagonal blocks of a factor that is already there
int launch_diag_invert(Handle* h, cudaStream_t st, double* A, int64_t lda, double* Dinv, double* logdet_parts, int* info,
                       int T) {
  potrf_diag_kernel<true><<<T, DIAG_THREADS, DIAG_SMEM, st>>>(A, lda, Dinv, logdet_parts, info, 0, nullptr);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int launch_trsv_fwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* b, double* z,
                    int k, int T) {
  trsv_fwd_kernel<<<T - k, TRSV_THREADS, TRSV_SMEM, st>>>(A, lda, Dinv, b, z, k);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int launch_trsv_bwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* z, double* x,
                    int k, int T) {
  (void)T;
  trsv_bwd_kernel<<<k + 1, TRSV_THREADS, TRSV_SMEM, st>>>(A, lda, Dinv, z, x, k);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
