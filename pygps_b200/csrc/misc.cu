// misc.cu - small bandwidth-bound kernels around the factorisation: vector
// scaling, the nlZ reduction (Core/inf.py:370), predictive column reductions
// (Core/gp.py:412,416), padding/compaction copies and the fused dnlZ reduction
// (Core/inf.py:373-377) that never materialises a derivative matrix.
#include "gpk_internal.cuh"

namespace gpk {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < nw) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLM, t, o);
  }
  return t;  // valid on warp 0
}

// alpha = x * inv_sn2 ; res[0] = sum_i r_i*alpha_i ; res[1] = sum_k logdet_parts[k]   (one CTA, deterministic)
__global__ void __launch_bounds__(1024) finish_alpha_kernel(const double* __restrict__ x, const double* __restrict__ r,
                                                            double inv_sn2, int64_t np, double* __restrict__ alpha,
                                                            const double* __restrict__ parts, int T,
                                                            double* __restrict__ res) {
  __shared__ double sh[32];
  double dot = 0.0, ld = 0.0;
  for (int64_t i = threadIdx.x; i < np; i += blockDim.x) {
    const double a = x[i] * inv_sn2;
    alpha[i] = a;
    dot = fma(r[i], a, dot);
  }
  for (int k = threadIdx.x; k < T; k += blockDim.x) ld += parts[k];
  const double d = block_sum(dot, sh);
  const double l = block_sum(ld, sh);
  if (threadIdx.x == 0) { res[0] = d; res[1] = l; }
}

__global__ void sum_parts_kernel(const double* __restrict__ parts, int T, double* __restrict__ res) {
  __shared__ double sh[32];
  double ld = 0.0;
  for (int k = threadIdx.x; k < T; k += blockDim.x) ld += parts[k];
  const double l = block_sum(ld, sh);
  if (threadIdx.x == 0) res[0] = l;
}

// dst (pn x pn, column-major) <- src (n x n, any symmetric order) with identity padding
__global__ void pad_sym_kernel(const double* __restrict__ src, int64_t n, double* __restrict__ dst, int64_t pn) {
  const int64_t total = pn * pn;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % pn, c = i / pn;
    double v;
    if (r < n && c < n) v = src[r + c * n];
    else v = (r == c) ? 1.0 : 0.0;
    dst[i] = v;
  }
}

// dst (n x n compact, pitch n) <- lower triangle of src (pitch ld), zeros above.
// Read as a C-order numpy array this is the UPPER factor R = L^T (Core/inf.py:362).
__global__ void compact_lower_kernel(const double* __restrict__ src, int64_t ld, int64_t n, double* __restrict__ dst) {
  const int64_t total = n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % n, c = i / n;
    dst[i] = (r >= c) ? src[r + c * ld] : 0.0;
  }
}

// dst (n x n) full symmetric from the lower triangle of src
__global__ void compact_sym_kernel(const double* __restrict__ src, int64_t ld, int64_t n, double* __restrict__ dst) {
  const int64_t total = n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % n, c = i / n;
    dst[i] = (r >= c) ? src[r + c * ld] : src[c + r * ld];
  }
}

// M (rows x cols, pitch ld) <- identity pattern (1 on r==c, else 0)
__global__ void set_identity_kernel(double* __restrict__ M, int64_t ld, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % rows, c = i / rows;
    M[r + c * ld] = (r == c) ? 1.0 : 0.0;
  }
}

// Column reductions over a (rows x cols) column-major matrix P with pitch ld, thread per ROW index c:
//   mode 0: part[split][c] = sum_r P[c + r*ld] * v[r]      (Ks' alpha)
//   mode 1: part[split][c] = sum_r P[c + r*ld]^2           (colsum(V*V))
// grid = (rows/128, nsplit)
__global__ void __launch_bounds__(128) rowdot_kernel(const double* __restrict__ P, int64_t ld, int64_t cols,
                                                     const double* __restrict__ v, int mode, int nsplit,
                                                     double* __restrict__ part, int64_t rows) {
  const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int sp = blockIdx.y;
  const int64_t per = (cols + nsplit - 1) / nsplit;
  const int64_t r0 = sp * per, r1 = min(cols, r0 + per);
  if (c < rows) {
    // eight independent accumulators: eight loads in flight per thread instead of one dependent FMA chain
    // (the FITC V r product, 4096 x 262144, took 13.5 ms as a single chain per thread)
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.0;
    int64_t r = r0;
    for (; r + 8 <= r1; r += 8) {
      double x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = P[c + (r + q) * ld];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(x[q], (mode == 0) ? v[r + q] : x[q], a[q]);
    }
    for (; r < r1; ++r) { const double x = P[c + r * ld]; a[0] = fma(x, (mode == 0) ? v[r] : x, a[0]); }
    part[(int64_t)sp * rows + c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
}

// out[c] = post( sum_split part[split][c] ):  mode 0: scale*sum ; mode 1: max(kss - sum, 0)
__global__ void rowdot_finish_kernel(const double* __restrict__ part, int nsplit, int64_t rows, int mode, double scale,
                                     double kss, double* __restrict__ out, int64_t nvalid,
                                     const double* __restrict__ kss_vec) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nvalid) return;
  double s = 0.0;
  for (int sp = 0; sp < nsplit; ++sp) s += part[(int64_t)sp * rows + c];
  out[c] = (mode == 0) ? s * scale : fmax((kss_vec ? kss_vec[c] : kss) - s, 0.0);
}

// ---------------------------------------------------------------------------
// fused dnlZ reduction over the lower triangle (Core/inf.py:373-377):
//   Q_ij = Ainv_ij * inv_sn2 - alpha_i alpha_j
//   part[cta][h]  = sum_{i>=j} w_ij Q_ij dK^h_ij   (w = 1 on the diagonal, 2 below)
//   part[cta][nh] = sum_i Q_ii
// dK is recomputed from the scaled inputs; no derivative matrix is stored.
// One 64x64 tile per CTA, 256 threads; a thread walks 16 pairs one at a time.
// ---------------------------------------------------------------------------
constexpr int DT_ = 64;
constexpr int DMAXD = 32;  // dimensions handled per pass for the ARD kernel

struct DnlzArgs {
  const double* Xs; int64_t n; int D;
  const double* Ainv; int64_t ld;
  const double* alpha;
  const double* sw;  // EP: Q_ij = Ainv_ij*sw_i*sw_j - alpha_i alpha_j (Core/inf.py:788); null: Ainv_ij*inv_sn2
  double inv_sn2, sf2;
  int kind, matern_d;
  int d_begin;      // ARD: first length-scale index handled by this pass
  int nacc;         // accumulators produced by this pass (<= DMAXD+2)
  double* part;     // [ctas][nacc]
  int rect;         // 1: Ainv holds FULL ROWS [i_off, i_off + rows) of the inverse (rows x n, pitch ld): every pair of the
  int64_t i_off;    //    rectangle counts once (sharded derivatives: the ranks' rectangles tile the whole matrix)
  int64_t rows;
};

__global__ void __launch_bounds__(256) dnlz_kernel(const DnlzArgs a) {
  extern __shared__ double dsh[];
  const int bi = blockIdx.x, bj = blockIdx.y;
  const int cta = bi + bj * gridDim.x;
  const int tid = threadIdx.x;
  double acc[DMAXD + 2];
#pragma unroll
  for (int q = 0; q < DMAXD + 2; ++q) acc[q] = 0.0;

  if (a.rect || bi >= bj) {
    double* Xi = dsh;                 // [64][D]
    double* Xj = dsh + DT_ * a.D;     // [64][D]
    // i: GLOBAL row index; li: row inside Ainv (rect mode: Ainv row 0 is global row i_off)
    const int64_t i0 = (int64_t)bi * DT_ + (a.rect ? a.i_off : 0), j0 = (int64_t)bj * DT_;
    const int64_t ilim = a.rect ? min(a.n, a.i_off + a.rows) : a.n;
    for (int idx = tid; idx < DT_ * a.D; idx += 256) {
      const int p = idx / a.D, d = idx % a.D;
      Xi[idx] = (i0 + p < ilim) ? a.Xs[(i0 + p) * a.D + d] : 0.0;
      Xj[idx] = (j0 + p < a.n) ? a.Xs[(j0 + p) * a.D + d] : 0.0;
    }
    __syncthreads();
    const int ti = tid & 63, tj0 = tid >> 6;  // i fastest across lanes -> coalesced Ainv reads
    const int64_t i = i0 + ti;
    const int64_t li = i - (a.rect ? a.i_off : 0);
    const double ai = (i < ilim) ? a.alpha[i] : 0.0;
    const double swi = (a.sw && i < ilim) ? a.sw[i] : 0.0;
    for (int jj = tj0; jj < DT_; jj += 4) {
      const int64_t j = j0 + jj;
      if (i >= ilim || j >= a.n || (!a.rect && i < j)) continue;
      const double q = a.Ainv[li + j * a.ld] * (a.sw ? swi * a.sw[j] : a.inv_sn2) - ai * a.alpha[j];
      const double w = (a.rect || i == j) ? 1.0 : 2.0;
      double d2 = 0.0;
      for (int d = 0; d < a.D; ++d) { const double df = Xi[ti * a.D + d] - Xj[jj * a.D + d]; d2 = fma(df, df, d2); }
      if (a.kind == GPK_COV_MATERN) {
        const double t = sqrt(d2), e = exp(-t);
        double f, df;
        switch (a.matern_d) {
          case 1: f = 1.0; df = 1.0; break;
          case 3: f = 1.0 + t; df = t; break;
          case 5: f = 1.0 + t + t * t / 3.0; df = (t + t * t) / 3.0; break;
          default: f = 1.0 + t + 2.0 * t * t / 5.0 + t * t * t / 15.0; df = (3.0 * t + 3.0 * t * t + t * t * t) / 15.0; break;
        }
        acc[0] = fma(w * q, a.sf2 * df * t * e, acc[0]);
        acc[1] = fma(w * q, 2.0 * a.sf2 * f * e, acc[1]);
        if (i == j) acc[2] += q;
      } else {
        const double k = a.sf2 * exp(-0.5 * d2);
        const double wqk = w * q * k;
        if (a.kind == GPK_COV_RBF) {
          acc[0] = fma(wqk, d2, acc[0]);
          acc[1] = fma(wqk, 2.0, acc[1]);
          if (i == j) acc[2] += q;
        } else {
          // ARD pass: accumulators [0, nd) are length scales d_begin.., then (first pass only) sf and trace
          const int nd = min(DMAXD, a.D - a.d_begin);
#pragma unroll
          for (int d = 0; d < DMAXD; ++d) {
            if (d < nd) {
              const double df = Xi[ti * a.D + a.d_begin + d] - Xj[jj * a.D + a.d_begin + d];
              acc[d] = fma(wqk, df * df, acc[d]);
            }
          }
          if (a.d_begin == 0) {
            acc[DMAXD] = fma(wqk, 2.0, acc[DMAXD]);
            if (i == j) acc[DMAXD + 1] += q;
          }
        }
      }
    }
  }
  // CTA reduction of every accumulator, deterministic
  __shared__ double red[32];
#pragma unroll
  for (int q = 0; q < DMAXD + 2; ++q) {
    if (q < a.nacc) {
      const double t = block_sum(acc[q], red);
      if (tid == 0) a.part[(int64_t)cta * a.nacc + q] = t;
    }
  }
}

// res[q] = sum_cta part[cta][q]
__global__ void __launch_bounds__(1024) dnlz_finish_kernel(const double* __restrict__ part, int64_t nctas, int nacc,
                                                           double* __restrict__ res) {
  __shared__ double sh[32];
  for (int q = 0; q < nacc; ++q) {
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < nctas; c += blockDim.x) s += part[c * nacc + q];
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0) res[q] = t;
  }
}

// dst[c + r*ldd] = src[r + c*lds] for `batch` matrices (rows x cols) spaced sstride / dstride apart: 32x32 tiles through
// shared memory so that both sides are coalesced
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ src, int64_t lds, int64_t sstride,
                                                        double* __restrict__ dst, int64_t ldd, int64_t dstride,
                                                        int64_t rows, int64_t cols) {
  __shared__ double t[32][33];
  src += (int64_t)blockIdx.z * sstride;
  dst += (int64_t)blockIdx.z * dstride;
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + tx, c = c0 + ty + 8 * k;
    t[ty + 8 * k][tx] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t c = c0 + tx, r = r0 + ty + 8 * k;
    if (r < rows && c < cols) dst[c + r * ldd] = t[tx][ty + 8 * k];
  }
}

int launch_transpose(Handle* h, cudaStream_t st, const double* src, int64_t lds, int64_t sstride, double* dst,
                     int64_t ldd, int64_t dstride, int64_t rows, int64_t cols, int batch) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return 0;
  if ((cols + 31) / 32 > 65535 || batch > 65535) return GPK_ERR_ARG;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)batch);
  transpose_kernel<<<grid, 256, 0, st>>>(src, lds, sstride, dst, ldd, dstride, rows, cols);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

static inline int grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_finish_alpha(Handle* h, cudaStream_t st, const double* x, const double* r, double inv_sn2, int64_t np,
                        double* alpha, const double* parts, int T, double* res) {
  finish_alpha_kernel<<<1, 1024, 0, st>>>(x, r, inv_sn2, np, alpha, parts, T, res);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_sum_parts(Handle* h, cudaStream_t st, const double* parts, int T, double* res) {
  sum_parts_kernel<<<1, 256, 0, st>>>(parts, T, res);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_pad_sym(Handle* h, cudaStream_t st, const double* src, int64_t n, double* dst, int64_t pn) {
  pad_sym_kernel<<<grid_for(pn * pn), 256, 0, st>>>(src, n, dst, pn);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_compact_lower(Handle* h, cudaStream_t st, const double* src, int64_t ld, int64_t n, double* dst) {
  compact_lower_kernel<<<grid_for(n * n), 256, 0, st>>>(src, ld, n, dst);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_compact_sym(Handle* h, cudaStream_t st, const double* src, int64_t ld, int64_t n, double* dst) {
  compact_sym_kernel<<<grid_for(n * n), 256, 0, st>>>(src, ld, n, dst);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_set_identity(Handle* h, cudaStream_t st, double* M, int64_t ld, int64_t rows, int64_t cols) {
  set_identity_kernel<<<grid_for(rows * cols), 256, 0, st>>>(M, ld, rows, cols);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}
int launch_rowdot(Handle* h, cudaStream_t st, const double* P, int64_t ld, int64_t rows, int64_t cols, const double* v,
                  int mode, double scale, double kss, double* part, int nsplit, double* out, int64_t nvalid,
                  const double* kss_vec) {
  dim3 grid((unsigned)((rows + 127) / 128), (unsigned)nsplit);
  rowdot_kernel<<<grid, 128, 0, st>>>(P, ld, cols, v, mode, nsplit, part, rows);
  rowdot_finish_kernel<<<(unsigned)((nvalid + 255) / 256), 256, 0, st>>>(part, nsplit, rows, mode, scale, kss, out,
                                                                         nvalid, kss_vec);
  h->stats.launches += 2;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// Runs the fused reduction; res gets [dcov_0 .. dcov_{nhyp-1}, trace(Q)] (before the 1/2 and sn2 factors).
static int launch_dnlz_impl(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv,
                            int64_t ld, const double* alpha, const double* sw, double inv_sn2, double sf2, int kind,
                            int matern_d, double* part, int64_t part_cap, double* res, int rect = 0, int64_t rect_off = 0,
                            int64_t rect_rows = 0);

int launch_dnlz(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv, int64_t ld,
                const double* alpha, double inv_sn2, double sf2, int kind, int matern_d, double* part,
                int64_t part_cap, double* res) {
  return launch_dnlz_impl(h, st, Xs, n, D, Ainv, ld, alpha, nullptr, inv_sn2, sf2, kind, matern_d, part, part_cap, res);
}
// Ainv = rows [i_off, i_off + rows) of the inverse, full width (rows x n, pitch ld); every pair counts once
int launch_dnlz_rect(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv_rows, int64_t ld,
                     int64_t i_off, int64_t rows, const double* alpha, double inv_sn2, double sf2, int kind, int matern_d,
                     double* part, int64_t part_cap, double* res) {
  return launch_dnlz_impl(h, st, Xs, n, D, Ainv_rows, ld, alpha, nullptr, inv_sn2, sf2, kind, matern_d, part, part_cap,
                          res, 1, i_off, rows);
}
int launch_dnlz_sw(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv, int64_t ld,
                   const double* alpha, const double* sw, double sf2, int kind, int matern_d, double* part,
                   int64_t part_cap, double* res) {
  return launch_dnlz_impl(h, st, Xs, n, D, Ainv, ld, alpha, sw, 1.0, sf2, kind, matern_d, part, part_cap, res);
}

static int launch_dnlz_impl(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv,
                            int64_t ld, const double* alpha, const double* sw, double inv_sn2, double sf2, int kind,
                            int matern_d, double* part, int64_t part_cap, double* res, int rect, int64_t rect_off,
                            int64_t rect_rows) {
  const int64_t g = (n + DT_ - 1) / DT_;
  const int64_t gi = rect ? (rect_rows + DT_ - 1) / DT_ : g;     // row tiles (rect mode: the rectangle's rows)
  if (g > 65535) return GPK_ERR_ARG;
  const int64_t nctas = gi * g;
  const size_t smem = size_t(2) * DT_ * D * sizeof(double);
  if (smem > 200 * 1024) return GPK_ERR_ARG;
  GPK_CK(h, cudaFuncSetAttribute(dnlz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DnlzArgs a;
  a.Xs = Xs; a.n = n; a.D = D; a.Ainv = Ainv; a.ld = ld; a.alpha = alpha; a.sw = sw; a.inv_sn2 = inv_sn2; a.sf2 = sf2;
  a.kind = kind; a.matern_d = matern_d; a.part = part;
  a.rect = rect; a.i_off = rect_off; a.rows = rect_rows;
  dim3 grid((unsigned)gi, (unsigned)g);
  if (kind != GPK_COV_RBFARD) {
    a.d_begin = 0; a.nacc = 3;
    if (nctas * a.nacc > part_cap) return GPK_ERR_ARG;
    dnlz_kernel<<<grid, 256, smem, st>>>(a);
    dnlz_finish_kernel<<<1, 1024, 0, st>>>(part, nctas, 3, res);
    h->stats.launches += 2;
  } else {
    // res layout for ARD: [ell_0..ell_{D-1}, sf, trace]
    for (int d0 = 0; d0 < D; d0 += DMAXD) {
      a.d_begin = d0;
      a.nacc = DMAXD + 2;
      if (nctas * a.nacc > part_cap) return GPK_ERR_ARG;
      dnlz_kernel<<<grid, 256, smem, st>>>(a);
      // scratch result: [DMAXD length scales, sf, trace]
      dnlz_finish_kernel<<<1, 1024, 0, st>>>(part, nctas, a.nacc, res + D + 2);
      const int nd = (D - d0 < DMAXD) ? D - d0 : DMAXD;
      GPK_CK(h, cudaMemcpyAsync(res + d0, res + D + 2, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
      if (d0 == 0)
        GPK_CK(h, cudaMemcpyAsync(res + D, res + D + 2 + DMAXD, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
      h->stats.launches += 2;
    }
  }
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
__global__ void copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
int launch_copy(Handle* h, cudaStream_t st, const double* src, double* dst, int64_t n) {
  copy_kernel<<<148 * 16, 256, 0, st>>>((const double4*)src, (double4*)dst, n / 4);
  GPK_CK(h, cudaGetLastError());
  return 0;
}

__global__ void fill_random_kernel(double* __restrict__ p, int64_t n, unsigned seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i * 6364136223846793005ULL + seed * 1442695040888963407ULL + 1ULL;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
    p[i] = ((double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 0.01;
  }
}
int launch_fill_random(Handle* h, cudaStream_t st, double* p, int64_t n, unsigned seed) {
  fill_random_kernel<<<148 * 8, 256, 0, st>>>(p, n, seed);
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
