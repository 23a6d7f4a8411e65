// potrf_diag.cu - the latency-critical panel kernels of the blocked Cholesky.
//
//  potrf_diag_kernel : factor one 128x128 diagonal block (lower Cholesky) AND
//                      invert the factor, in shared memory, one CTA.
//     - 32x32 diagonal sub-blocks are factored by ONE WARP holding one matrix
//       row per lane in registers; pivots and multipliers travel by warp
//       shuffle (no shared-memory round trips, no block barriers).
//     - the sub-block inverse is a per-lane forward substitution (lane c owns
//       column c of the inverse), again in registers.
//     - sub-panel solves / trailing updates / the block inverse are small
//       shared-memory GEMMs by all 8 warps.
//    The inverse turns the panel TRSM and every later triangular solve with
//    this block into the tensor-pipe GEMM of gemm_nt.cu.
//    Replaces the unblocked part of LAPACK dpotrf called at
//    /root/reference/pyGPs/Core/tools.py:61; emits sum(log(diag L)) for
//    Core/inf.py:370 and a LAPACK-style info for Core/tools.py:62-77.
//
//  trsv_fwd_step / trsv_bwd_step : one block step of the forward / backward
//    substitution with a single right-hand side (solve_chol with B=(n,1),
//    Core/tools.py:96, as used at Core/inf.py:363).
#include "gpk_internal.cuh"

namespace gpk {

constexpr unsigned FULL = 0xffffffffu;
constexpr int DB = NB;            // 128
constexpr int IB = DIAG_IB;       // 32
constexpr int LDT = DIAG_LDT;     // 33
constexpr int DW = DIAG_THREADS / 32;
constexpr int NRD = (3 * (IB / 4) + DW - 1) / DW;  // register-staging rounds for a 96x32 sub-panel

// S: DBxDB column-major (pitch DB).  T: 4 blocks of IBxIB, column-major, pitch LDT.
#define S_(r, c) S[(r) + (c) * DB]
#define T_(b, r, c) T[(b) * IB * LDT + (r) + (c) * LDT]

__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(double* __restrict__ Ablk, int64_t lda, double* __restrict__ Dinv,
                  double* __restrict__ logdet_slot, int* __restrict__ info, int gidx0) {
  extern __shared__ __align__(16) double dsm[];
  double* S = dsm;
  double* T = dsm + DB * DB;
  __shared__ double s_logdet;
  __shared__ int s_info;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int idx = tid; idx < DB * DB; idx += DIAG_THREADS) {
    const int r = idx % DB, c = idx / DB;
    S[idx] = (r >= c) ? Ablk[r + (int64_t)c * lda] : 0.0;
  }
  if (tid == 0) { s_logdet = 0.0; s_info = 0; }
  __syncthreads();

  // ---------------- phase 1: blocked Cholesky, inner block 32 ----------------
  for (int jb = 0; jb < DB / IB; ++jb) {
    const int j0 = jb * IB;
    if (warp == 0) {
      double a[IB];
#pragma unroll
      for (int c = 0; c < IB; ++c) a[c] = S_(j0 + lane, j0 + c);
      double lsum = 0.0;
      int bad = 0;
#pragma unroll
      for (int j = 0; j < IB; ++j) {
        const double d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0.0) && bad == 0) bad = j + 1;
        const double l = sqrt(d);
        const double rinv = 1.0 / l;
        lsum += log(l);
        const double v = (lane == j) ? l : a[j] * rinv;
        a[j] = (lane >= j) ? v : 0.0;
#pragma unroll
        for (int c = j + 1; c < IB; ++c) {
          const double lc = __shfl_sync(FULL, a[j], c);
          a[c] = fma(-a[j], lc, a[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < IB; ++c) S_(j0 + lane, j0 + c) = (lane >= c) ? a[c] : 0.0;
      __syncwarp();
      // inverse of the 32x32 factor: lane owns column `lane` of W = L^-1
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
        w[i] = ((s0 + s1) + (s2 + s3)) / S_(j0 + i, j0 + i);
      }
#pragma unroll
      for (int i = 0; i < IB; ++i) T_(jb, i, lane) = w[i];
      if (lane == 0) {
        s_logdet += lsum;
        if (bad != 0 && s_info == 0) s_info = gidx0 + j0 + bad;
      }
    }
    __syncthreads();
    const int nrb = DB / IB - 1 - jb;  // 32-row blocks below the diagonal block
    if (nrb > 0) {
      // sub-panel: S[r, j0+c] <- sum_{k<=c} S[r, j0+k] * W(c,k)   (row-wise in place -> stage in registers)
      double outv[NRD][4];
      const int ntile = nrb * (IB / 4);
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int r = j0 + IB + (wt / (IB / 4)) * IB + lane;
          const int c0 = (wt % (IB / 4)) * 4;
          double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
          for (int k = 0; k < c0 + 4; ++k) {
            const double x = S_(r, j0 + k);
            acc0 = fma(x, T_(jb, c0 + 0, k), acc0);
            acc1 = fma(x, T_(jb, c0 + 1, k), acc1);
            acc2 = fma(x, T_(jb, c0 + 2, k), acc2);
            acc3 = fma(x, T_(jb, c0 + 3, k), acc3);
          }
          outv[rd][0] = acc0; outv[rd][1] = acc1; outv[rd][2] = acc2; outv[rd][3] = acc3;
        }
      }
      __syncthreads();
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int r = j0 + IB + (wt / (IB / 4)) * IB + lane;
          const int c0 = (wt % (IB / 4)) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) S_(r, j0 + c0 + q) = outv[rd][q];
        }
      }
      __syncthreads();
      // trailing update of the lower block triangle: S[ib,cb] -= S[ib,jb] * S[cb,jb]^T
      const int npair = nrb * (nrb + 1) / 2;
      for (int wt = warp; wt < npair * (IB / 4); wt += DW) {
        const int pr = wt / (IB / 4);
        const int cg = wt % (IB / 4);
        int ib = 0, cb = 0;  // decode pair index -> (ib >= cb), both in [0,nrb)
        {
          int q = pr;
          while (q > ib) { q -= (ib + 1); ++ib; }
          cb = q;
        }
        const int r = j0 + IB + ib * IB + lane;
        const int c0 = j0 + IB + cb * IB + cg * 4;
        double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll 8
        for (int k = 0; k < IB; ++k) {
          const double x = S_(r, j0 + k);
          acc0 = fma(x, S_(c0 + 0, j0 + k), acc0);
          acc1 = fma(x, S_(c0 + 1, j0 + k), acc1);
          acc2 = fma(x, S_(c0 + 2, j0 + k), acc2);
          acc3 = fma(x, S_(c0 + 3, j0 + k), acc3);
        }
        S_(r, c0 + 0) -= acc0;
        S_(r, c0 + 1) -= acc1;
        S_(r, c0 + 2) -= acc2;
        S_(r, c0 + 3) -= acc3;
      }
      __syncthreads();
    }
  }

  // ---------------- phase 2: L back to global (strict upper of the block = 0) ----
  for (int idx = tid; idx < DB * DB; idx += DIAG_THREADS) {
    const int r = idx % DB, c = idx / DB;
    Ablk[r + (int64_t)c * lda] = (r >= c) ? S[idx] : 0.0;
  }
  if (tid == 0) {
    *logdet_slot = s_logdet;
    if (s_info != 0) atomicCAS(info, 0, s_info);
  }
  __syncthreads();

  // ---------------- phase 3: in-place inverse of the block factor ----------------
  // W[>j, j] = -W22 * L[>j, j] * W_jj, from the last block column to the first.
  for (int jb = DB / IB - 1; jb >= 0; --jb) {
    const int j0 = jb * IB;
    const int nrb = DB / IB - 1 - jb;
    const int r0 = j0 + IB;
    if (nrb > 0) {
      const int ntile = nrb * (IB / 4);
      double outv[NRD][4];
      // (i) Y <- Y * W_jj :  Y(r,c) = sum_{k>=c} Y(r,k) * W_jj(k,c)
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int r = r0 + (wt / (IB / 4)) * IB + lane;
          const int c0 = (wt % (IB / 4)) * 4;
          double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
          for (int k = c0; k < IB; ++k) {
            const double x = S_(r, j0 + k);
            acc0 = fma(x, T_(jb, k, c0 + 0), acc0);
            acc1 = fma(x, T_(jb, k, c0 + 1), acc1);
            acc2 = fma(x, T_(jb, k, c0 + 2), acc2);
            acc3 = fma(x, T_(jb, k, c0 + 3), acc3);
          }
          outv[rd][0] = acc0; outv[rd][1] = acc1; outv[rd][2] = acc2; outv[rd][3] = acc3;
        }
      }
      __syncthreads();
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int r = r0 + (wt / (IB / 4)) * IB + lane;
          const int c0 = (wt % (IB / 4)) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) S_(r, j0 + c0 + q) = outv[rd][q];
        }
      }
      __syncthreads();
      // (ii) X <- -W22 * Y : X(r,c) = -sum_{r0<=k<=r} S(r,k) * Y(k,c)   (S upper entries are exact zeros)
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int rb = wt / (IB / 4);
          const int r = r0 + rb * IB + lane;
          const int c0 = j0 + (wt % (IB / 4)) * 4;
          double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
          const int kend = r0 + (rb + 1) * IB;
#pragma unroll 8
          for (int k = r0; k < kend; ++k) {
            const double x = S_(r, k);
            acc0 = fma(x, S_(k, c0 + 0), acc0);
            acc1 = fma(x, S_(k, c0 + 1), acc1);
            acc2 = fma(x, S_(k, c0 + 2), acc2);
            acc3 = fma(x, S_(k, c0 + 3), acc3);
          }
          outv[rd][0] = -acc0; outv[rd][1] = -acc1; outv[rd][2] = -acc2; outv[rd][3] = -acc3;
        }
      }
      __syncthreads();
#pragma unroll
      for (int rd = 0; rd < NRD; ++rd) {
        const int wt = warp + rd * DW;
        if (wt < ntile) {
          const int r = r0 + (wt / (IB / 4)) * IB + lane;
          const int c0 = j0 + (wt % (IB / 4)) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) S_(r, c0 + q) = outv[rd][q];
        }
      }
    }
    // (iii) diagonal block <- W_jj
    for (int idx = tid; idx < IB * IB; idx += DIAG_THREADS) {
      const int r = idx % IB, c = idx / IB;
      S_(j0 + r, j0 + c) = T_(jb, r, c);
    }
    __syncthreads();
  }

  // ---------------- phase 4: inverse to global (dense 128x128, pitch 128) --------
  for (int idx = tid; idx < DB * DB; idx += DIAG_THREADS) Dinv[idx] = S[idx];
}

int diag_init(Handle* h) {
  GPK_CK(h, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  return 0;
}

int launch_diag(Handle* h, cudaStream_t st, double* Ablk, int64_t lda, double* Dinv, double* logdet_slot, int* info,
                int gidx0) {
  potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM, st>>>(Ablk, lda, Dinv, logdet_slot, info, gidx0);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// single right-hand-side substitution steps
// ---------------------------------------------------------------------------
// forward step k (grid T-k CTAs of 128 threads):
//   every CTA forms z_k = Dinv_k * b_k in shared memory; CTA 0 stores it;
//   CTA i>0 updates b_{k+i} -= L[k+i, k] * z_k.
__global__ void __launch_bounds__(NB) trsv_fwd_kernel(const double* __restrict__ A, int64_t lda,
                                                      const double* __restrict__ Dinv, double* __restrict__ b,
                                                      double* __restrict__ z, int k) {
  __shared__ double sb[NB], sz[NB];
  const int tid = threadIdx.x;
  const double* Dk = Dinv + (int64_t)k * NB * NB;
  sb[tid] = b[(int64_t)k * NB + tid];
  __syncthreads();
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < NB; ++c) acc = fma(Dk[tid + c * NB], sb[c], acc);  // lower-triangular: zeros above the diagonal
  sz[tid] = acc;
  __syncthreads();
  const int i = blockIdx.x;
  if (i == 0) {
    z[(int64_t)k * NB + tid] = acc;
  } else {
    const double* Lik = A + (int64_t)(k + i) * NB + (int64_t)k * NB * lda;
    double s = 0.0;
#pragma unroll 8
    for (int c = 0; c < NB; ++c) s = fma(Lik[tid + (int64_t)c * lda], sz[c], s);
    b[(int64_t)(k + i) * NB + tid] -= s;
  }
}

// backward step k (grid k+1 CTAs of 128 threads):
//   every CTA forms x_k = Dinv_k^T * z_k; CTA k stores it; CTA j<k updates z_j -= L[k, j]^T * x_k.
__global__ void __launch_bounds__(NB) trsv_bwd_kernel(const double* __restrict__ A, int64_t lda,
                                                      const double* __restrict__ Dinv, double* __restrict__ z,
                                                      double* __restrict__ x, int k) {
  __shared__ double sz[NB], sx[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Dk = Dinv + (int64_t)k * NB * NB;
  sz[tid] = z[(int64_t)k * NB + tid];
  __syncthreads();
  // x[c] = sum_r Dk[r,c] * z[r]; warp per column group so that loads run along r (contiguous)
  for (int c = warp; c < NB; c += NB / 32) {
    double s = 0.0;
#pragma unroll
    for (int r = lane; r < NB; r += 32) s = fma(Dk[r + c * NB], sz[r], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) sx[c] = s;
  }
  __syncthreads();
  const int j = blockIdx.x;
  if (j == k) {
    x[(int64_t)k * NB + tid] = sx[tid];
  } else {
    const double* Lkj = A + (int64_t)k * NB + (int64_t)j * NB * lda;
    for (int c = warp; c < NB; c += NB / 32) {
      double s = 0.0;
#pragma unroll
      for (int r = lane; r < NB; r += 32) s = fma(Lkj[r + (int64_t)c * lda], sx[r], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      if (lane == 0) z[(int64_t)j * NB + c] -= s;
    }
  }
}

int launch_trsv_fwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* b, double* z,
                    int k, int T) {
  trsv_fwd_kernel<<<T - k, NB, 0, st>>>(A, lda, Dinv, b, z, k);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int launch_trsv_bwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* z, double* x,
                    int k, int T) {
  (void)T;
  trsv_bwd_kernel<<<k + 1, NB, 0, st>>>(A, lda, Dinv, z, x, k);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
