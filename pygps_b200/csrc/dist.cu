// dist.cu - one exact-GP evaluation sharded over the GPUs of one box (BASELINE config 3).
//
// One process per GPU.  The (np x np) matrix K/sn2+I is distributed by BLOCK COLUMNS, block-cyclic with block
// 128: rank r owns global column blocks r, r+G, r+2G, ... stored packed, full height.  Each panel (one block
// column) is therefore local to its owner: diag + TRSM need no communication.  The ONE exchange per step is the
// broadcast of the solved panel over NVLink (ncclBroadcast); every rank then updates the columns it owns with the
// same DMMA kernel (GemmArgs::cstride = G maps its packed columns to global ones).  K tiles are built where they
// live (CovArgs::s_bstride), X is replicated (<= 16 MiB).
//
// The forward substitution needs no kernel and no message of its own: y-m rides along as one extra ROW of the matrix
// (the Cholesky factor of [A b; b' c] carries z' = b'L^-T in its last row), so the panel TRSM and the trailing updates
// produce z.  The backward substitution is the dot-product form over the owner's local column with the solution
// block broadcast (1 KiB) per step.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 PyTorch ships) so libgpk.so has no link-time dependency.
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include "gpk_internal.cuh"

namespace gpk {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

static struct {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_bcast bcast = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

constexpr int NCCL_F64 = 8, NCCL_I32 = 2, NCCL_SUM = 0, NCCL_MAX = 2;

static int nccl_load(const char* path) {
  if (g_nccl.lib) return 0;
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return GPK_ERR_STATE;
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.bcast = (fn_bcast)dlsym(g_nccl.lib, "ncclBroadcast");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.bcast || !g_nccl.allreduce) return GPK_ERR_STATE;
  return 0;
}

#define NCCL_CK(h, call)                                                                       \
  do {                                                                                         \
    int r__ = (call);                                                                          \
    if (r__ != 0) {                                                                            \
      (h)->last_msg = std::string(#call) + ": " + (g_nccl.errstr ? g_nccl.errstr(r__) : "nccl error"); \
      return GPK_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// row np of the local columns <- (y-m) of the matching global columns; the other 127 rows of the extra tile row <- 0
__global__ void fill_aug_kernel(double* __restrict__ A, int64_t ld, int64_t np, int64_t ncols_loc, int G, int r,
                                const double* __restrict__ ymm, int64_t n) {
  const int64_t total = ncols_loc * NB;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = k / NB, i = k % NB;           // local column, row within the extra tile row
    const int64_t gc = (c / NB) * G * NB + (int64_t)r * NB + c % NB;
    A[np + i + c * ld] = (i == 0 && gc < n) ? ymm[gc] : 0.0;
  }
}

// panels per block of the blocked (int8) sharded factorisation: measured on 8 GPUs at N=65536: 4 -> 352 ms, 8 -> 325 ms,
// 16 -> 319 ms per evaluation (profiles/r2_dist8_c3.md); small problems keep 8 so that the blocked path still applies
static inline int dist_wd(int T) { return env_int("GPK_DIST_WD", T >= 256 ? 16 : 8); }

constexpr int DT_THREADS = 512;
constexpr size_t DT_SMEM = size_t(NB) * NB * sizeof(double);

// partial[i][c] = sum_r L[(k+1+i)*128 + r, c] * x[(k+1+i)*128 + r]   for tile i of the owner's local column k
__global__ void __launch_bounds__(DT_THREADS, 1) bwd_partial_kernel(const double* __restrict__ Acol, int64_t ld,
                                                                    const double* __restrict__ x, int k,
                                                                    double* __restrict__ partial) {
  extern __shared__ __align__(16) double tile[];
  __shared__ double sx[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)(k + 1 + blockIdx.x) * NB;
  if (tid < NB) sx[tid] = x[row0 + tid];
#pragma unroll
  for (int i = 0; i < (NB * NB / 2) / DT_THREADS; ++i) {
    const int ch = tid + i * DT_THREADS;
    const int c = ch >> 6, r = (ch & 63) * 2;
    unsigned sa = (unsigned)__cvta_generic_to_shared(tile + r + c * NB);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(Acol + row0 + r + (int64_t)c * ld) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const double v0 = sx[lane], v1 = sx[lane + 32], v2 = sx[lane + 64], v3 = sx[lane + 96];
  for (int c = warp; c < NB; c += DT_THREADS / 32) {
    const double* col = tile + c * NB;
    double s = fma(col[lane], v0, fma(col[lane + 32], v1, fma(col[lane + 64], v2, col[lane + 96] * v3)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) partial[(int64_t)blockIdx.x * NB + c] = s;
  }
}

// x_k = Dinv_k^T (z_k - sum_i partial[i]) ; z_k is row `zrow` of the local column block (pitch ld)
__global__ void __launch_bounds__(NB) bwd_finish_kernel(const double* __restrict__ Acol, int64_t ld, int64_t zrow,
                                                        const double* __restrict__ Dk, const double* __restrict__ partial,
                                                        int ntiles, double* __restrict__ xk) {
  __shared__ double s[NB];
  const int c = threadIdx.x;
  double acc = Acol[zrow + (int64_t)c * ld];
  for (int i = 0; i < ntiles; ++i) acc -= partial[(int64_t)i * NB + c];
  s[c] = acc;
  __syncthreads();
  double o = 0.0;
  for (int r = c; r < NB; ++r) o = fma(Dk[r + c * NB], s[r], o);   // (Dinv^T s)_c = sum_{r>=c} Dinv[r,c] s_r
  xk[c] = o;
}

// acc[c][j] += sum_r L[k*128 + r, (local column block c) j] * x_k[r]  for the owned column blocks c = 0 .. gridDim.x-1 (all
// those whose global index is below k): the contribution of the freshly broadcast x_k to every later step of this rank, so
// that the step of column k-1 has nothing left to do but the 128x128 finish once x_k has arrived.
__global__ void __launch_bounds__(DT_THREADS, 1) bwd_update_kernel(const double* __restrict__ gA, int64_t ld, int k,
                                                                   const double* __restrict__ xk,
                                                                   double* __restrict__ acc) {
  extern __shared__ __align__(16) double tile[];
  __shared__ double sx[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  const double* src = gA + (int64_t)k * NB + (int64_t)c * NB * ld;
  if (tid < NB) sx[tid] = xk[tid];
#pragma unroll
  for (int i = 0; i < (NB * NB / 2) / DT_THREADS; ++i) {
    const int ch = tid + i * DT_THREADS;
    const int cc = ch >> 6, r = (ch & 63) * 2;
    unsigned sa = (unsigned)__cvta_generic_to_shared(tile + r + cc * NB);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + r + (int64_t)cc * ld) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  const double v0 = sx[lane], v1 = sx[lane + 32], v2 = sx[lane + 64], v3 = sx[lane + 96];
  for (int j = warp; j < NB; j += DT_THREADS / 32) {
    const double* col = tile + j * NB;
    double sacc = fma(col[lane], v0, fma(col[lane + 32], v1, fma(col[lane + 64], v2, col[lane + 96] * v3)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (lane == 0) acc[(int64_t)c * NB + j] += sacc;
  }
}

// alpha = x/sn2 ; res[0] = r'alpha
__global__ void __launch_bounds__(1024) dist_finish_kernel(const double* __restrict__ x, const double* __restrict__ r,
                                                           double inv_sn2, int64_t n, double* __restrict__ alpha,
                                                           double* __restrict__ res) {
  __shared__ double sh[32];
  double dot = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = x[i] * inv_sn2;
    alpha[i] = a;
    dot = fma(r[i], a, dot);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane == 0) sh[warp] = dot;
  __syncthreads();
  if (warp == 0) {
    double t = sh[lane];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) res[0] = t;
  }
}

// sum-all-reduce of a device buffer over the handle's communicator (no-op on one rank); used by fitc.cu
int dist_allreduce_sum(Handle* h, double* buf, size_t count, cudaStream_t st) {
  if (h->world <= 1 || !h->nccl_comm) return 0;
  NCCL_CK(h, g_nccl.allreduce(buf, buf, count, NCCL_F64, NCCL_SUM, h->nccl_comm, st));
  return 0;
}

}  // namespace gpk

using namespace gpk;

extern "C" {

int gpk_dist_unique_id(const char* nccl_path, char* out128) {
  if (!out128) return GPK_ERR_ARG;
  if (nccl_load(nccl_path) != 0) return GPK_ERR_STATE;
  nccl_uid id;
  if (g_nccl.get_uid(&id) != 0) return GPK_ERR_CUDA;
  std::memcpy(out128, id.internal, 128);
  return 0;
}

int gpk_dist_init(gpk_handle hh, const char* nccl_path, int rank, int world, const char* id128) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (world < 1 || rank < 0 || rank >= world) return GPK_ERR_ARG;
  h->rank = rank; h->world = world;
  if (world == 1) return 0;
  if (!id128) return GPK_ERR_ARG;
  if (nccl_load(nccl_path) != 0) { h->last_msg = "cannot load libnccl.so.2"; return GPK_ERR_STATE; }
  nccl_uid id;
  std::memcpy(id.internal, id128, 128);
  // connect every channel inside ncclCommInitRank, not lazily inside the first collective: with the ranks as threads of
  // one process a lazy connect (which allocates) could land behind a peer's already-enqueued kernel
  setenv("NCCL_RUNTIME_CONNECT", "0", 0);
  NCCL_CK(h, g_nccl.init_rank(&h->nccl_comm, world, id, rank));
  GPK_CK(h, cudaFuncSetAttribute(bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));
  return 0;
}

int gpk_dist_finalize(gpk_handle hh) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (h->nccl_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
  h->nccl_comm = nullptr;
  h->world = 1; h->rank = 0;
  return 0;
}

// Every device allocation the sharded evaluation (level 0), the factor gather (level 1) and the sharded derivatives
// (level 2) need for the current problem size.  When the ranks are THREADS of one process (pygps_b200.ShardedEngine),
// cudaMalloc / cudaFree on one device synchronise with its NCCL peers (peer mappings), so an allocation issued after a
// peer has already enqueued a collective that waits for this rank dead-locks: the caller reserves on every rank, joins,
// and only then starts the collective call, which then allocates nothing.
static int dist_reserve(Handle* h, int level) {
  const int G = h->world, r = h->rank;
  const int64_t n = h->n, np = h->np;
  if (n <= 0 || np <= 0) return GPK_ERR_STATE;
  const int T = (int)(np / NB);
  const int64_t ld = np + NB;
  const int nloc = (T > r) ? (T - 1 - r) / G + 1 : 0;
  const int64_t ncl = (int64_t)(nloc > 0 ? nloc : 1) * NB;
  GPK_TRY(ensure(h, &h->gA, &h->cgA, ld * ncl));
  GPK_TRY(ensure_zero(h, &h->gDinv, &h->cgDinv, ncl * NB));
  GPK_TRY(ensure(h, &h->gVec, &h->cgVec, np + T + 16 + (int64_t)T * NB + NB));
  GPK_TRY(ensure(h, &h->gPack, &h->cgPack, 2 * ld * NB));
  while (h->ev.size() < 5 * (size_t)T + 8) {
    cudaEvent_t e;
    GPK_CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ev.push_back(e);
  }
  const int WD = dist_wd(T);
  const bool doz = env_int("GPK_DIST_OZAKI", 1) != 0 && env_int("GPK_OZAKI", 1) != 0 && T >= 4 * WD && WD >= 1 && WD <= 16;
  if (doz) {
    GPK_TRY(ensure(h, &h->gBlk, &h->cgBlk, 2 * ld * (int64_t)WD * NB));
    GPK_TRY(oz_ensure(h, 0, ld, WD * NB));
    GPK_TRY(oz_ensure(h, 1, ld, WD * NB));
  }
  if (level >= 1) GPK_TRY(ensure(h, &h->dA, &h->capA, np * np));
  if (level >= 2) {
    const int t0 = (int)(((int64_t)T * r) / G), t1 = (int)(((int64_t)T * (r + 1)) / G);
    const int64_t rows = (int64_t)(t1 - t0) * NB;
    const int64_t g = (n + 63) / 64, gi = (rows + 63) / 64 > 0 ? (rows + 63) / 64 : 1;
    GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, gi * g * 34 + 64));
    if (rows > 0) {
      GPK_TRY(ensure(h, &h->dU, &h->capU, np * np));
      GPK_TRY(ensure(h, &h->dDinvT, &h->capDinvT, np * NB));
      GPK_TRY(ensure(h, &h->dW, &h->capW, rows * np));
    }
  }
  if (!h->dFlags) {                                  // (not used by the sharded path; allocated here for completeness)
    GPK_CK(h, cudaMalloc((void**)&h->dFlags, 1024 * sizeof(int)));
    GPK_CK(h, cudaMemsetAsync(h->dFlags, 0, 1024 * sizeof(int), h->s_main));
    h->flag_epoch = 0;
  }
  GPK_CK(h, cudaStreamSynchronize(h->s_main));
  return 0;
}

// Sharded counterpart of gpk_exact_eval (no derivatives): every rank calls it with the same arguments after
// gpk_set_data with the same X; every rank gets the same nlZ and the full alpha.
static int exact_eval_dist_impl(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                                const double* ymm, double* nlZ, double* alpha) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->dX || h->n <= 0) return GPK_ERR_STATE;
  if (!hyp || !ymm || !nlZ || !alpha) return GPK_ERR_ARG;
  const int G = h->world, r = h->rank;
  if (G > 1 && !h->nccl_comm) return GPK_ERR_STATE;
  const int64_t n = h->n, np = h->np;
  const int D = h->D, T = (int)(np / NB);
  const int64_t ld = np + NB;                       // one extra tile row carries y-m
  const int nloc = (T > r) ? (T - 1 - r) / G + 1 : 0;   // owned block columns
  std::vector<double> scale;
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  GPK_TRY(kind_scale(kind, matern_d, hyp, nhyp, D, scale, &divide, &premul, &sf2));
  if (D > 1900) return GPK_ERR_ARG;
  const double sn2 = std::exp(2.0 * log_sn);
  stats_begin(h);
  h->has_post = false; h->has_fitc = false; h->pn = 0; h->dist_post = false;
  cudaStream_t st = h->s_main;
  GPK_CK(h, cudaFuncSetAttribute(bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));

  const int64_t ncl = (int64_t)(nloc > 0 ? nloc : 1) * NB;
  GPK_TRY(ensure(h, &h->gA, &h->cgA, ld * ncl));
  GPK_TRY(ensure_zero(h, &h->gDinv, &h->cgDinv, ncl * NB));
  // vectors: x (np) | parts (T) | res (16) | partial (T*NB) | xk staging (NB)
  GPK_TRY(ensure(h, &h->gVec, &h->cgVec, np + T + 16 + (int64_t)T * NB + NB));
  double* x = h->gVec;
  double* parts = x + np;
  double* res = parts + T;
  double* partial = res + 16;

  GPK_CK(h, cudaEventRecord(h->t0, st));
  std::memcpy(h->hPinned, scale.data(), D * sizeof(double));
  GPK_CK(h, cudaMemcpyAsync(h->dScale, h->hPinned, D * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  GPK_CK(h, cudaMemsetAsync(parts, 0, (size_t)(T + 16) * sizeof(double), st));
  GPK_CK(h, cudaMemsetAsync(x, 0, (size_t)np * sizeof(double), st));
  GPK_CK(h, cudaMemsetAsync(h->dR, 0, (size_t)np * sizeof(double), st));
  GPK_CK(h, cudaMemcpyAsync(h->dR, ymm, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_TRY(launch_prescale(h, st, h->dX, n, np, D, h->dScale, divide, premul, h->dXs));
  if (nloc > 0) {
    CovArgs c{};
    c.F = h->dXs; c.S = h->dXs; c.out = h->gA; c.ld = ld;
    c.nF = n; c.nS = n; c.pF = np; c.pS = (int64_t)nloc * NB; c.D = D;
    c.kind = kind; c.matern_d = matern_d; c.epi = EPI_COV;
    c.sf2 = sf2; c.scale = 1.0 / sn2; c.diag_add = 1.0;
    c.same_set = 1; c.lower_only = 1; c.pad_identity = 1; c.padded128 = 1;
    c.s_bstride = G; c.s_boff = r;
    GPK_TRY(launch_cov(h, st, c));
    fill_aug_kernel<<<148, 256, 0, st>>>(h->gA, ld, np, (int64_t)nloc * NB, G, r, h->dR, n);
    h->stats.launches++;
  }
  GPK_CK(h, cudaEventRecord(h->t1, st));

  // ---- right-looking factorisation, one panel broadcast per step, look-ahead 1 --------------------------------
  //   s_panel : owner(k): diag(k) -> trsm(k) -> pack into Pack[k%2]
  //   s_aux   : every rank: ncclBroadcast(k) in step order (the communicator's stream)
  //   s_main  : every rank: update(k) - the column that becomes panel k+1 first (its owner can then factor it and
  //             get the next broadcast on the wire while everybody is still updating), then the other owned columns
  // Pack is double-buffered: broadcast k+1 lands while update k still reads Pack[k%2].
  GPK_TRY(ensure(h, &h->gPack, &h->cgPack, 2 * ld * NB));
  {
    // event pool layout: [0,T) packed  [T,2T) bcast done  [2T,3T) update done  [3T,4T) column ready  [4T] fork
    // [4T+4, 5T+4) far update of block b done (blocked variant)  [5T+4] block sliced
    while (h->ev.size() < 5 * (size_t)T + 8) {
      cudaEvent_t e;
      GPK_CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev.push_back(e);
    }
  }
  cudaEvent_t* ev_packed = h->ev.data();
  cudaEvent_t* ev_bcast = h->ev.data() + T;
  cudaEvent_t* ev_upd = h->ev.data() + 2 * T;
  cudaEvent_t* ev_col = h->ev.data() + 3 * T;
  cudaEvent_t ev_fork = h->ev[4 * T];
  cudaEvent_t* ev_far = h->ev.data() + 4 * T + 4;
  cudaEvent_t ev_sliced = h->ev[5 * T + 4];
  cudaStream_t sp = h->s_panel, sc = h->s_aux, sf = h->s_tail;
  GPK_CK(h, cudaEventRecord(ev_fork, st));
  GPK_CK(h, cudaStreamWaitEvent(sp, ev_fork, 0));
  GPK_CK(h, cudaStreamWaitEvent(sc, ev_fork, 0));
  // Blocked variant (the default; GPK_DIST_OZAKI=0 selects the rank-128 DMMA updates): inside a block of WD
  // panels only the block's own columns get the immediate rank-128 DMMA updates; the broadcast panels are collected
  // (full height, pitch ld) in a double-buffered block buffer, and after the block every rank slices it once and
  // applies ONE rank-(WD*128) update on the int8 tensor cores to the columns it owns beyond the block
  // (launch_oz_cyclic) - the column that becomes the next panel first.
  const int WD = dist_wd(T);
  const bool doz = env_int("GPK_DIST_OZAKI", 1) != 0 && env_int("GPK_OZAKI", 1) != 0 && T >= 4 * WD && WD >= 1 && WD <= 16;
  double* PB = nullptr;
  if (doz) {
    GPK_TRY(ensure(h, &h->gBlk, &h->cgBlk, 2 * ld * (int64_t)WD * NB));
    PB = h->gBlk;
    GPK_TRY(oz_ensure(h, 0, ld, WD * NB));
    GPK_TRY(oz_ensure(h, 1, ld, WD * NB));
    GPK_CK(h, cudaStreamWaitEvent(sf, ev_fork, 0));
  }
  // Blocked variant, far/near split (GPK_DIST_SPLIT, default on).  The rank-(WD*128) int8 update after block b is applied
  // in two parts: NEAR = the owned columns of the NEXT block (on the main stream: the next block's panels and immediate
  // updates need them, the column that becomes the next panel first), FAR = every owned column beyond the next block, on
  // its own stream (s_tail) where the far updates of consecutive blocks queue behind one another.  The far update of
  // block b therefore runs UNDER the dependent chain (diag -> TRSM -> broadcast -> column update) of block b+1 instead of
  // in front of it - measured before the split: per block 4.1 ms of update + 4.8 ms of chain, strictly one after the other.
  const bool dsplit = env_int("GPK_DIST_SPLIT", 1) != 0;
  // Stream priorities matter here: the far update is a grid of tens of thousands of CTAs, the chain-side kernels
  // (copy into the block buffer, immediate rank-128 updates, slicing, near update) must get SMs ahead of it.  s_main has the
  // LOWEST priority of the handle's streams, s_tail the second highest: in split mode the chain side runs on s_tail and the
  // far update on s_main (a first version had them the other way round and gained nothing).
  cudaStream_t su = (doz && dsplit) ? sf : st;     // chain-side updates
  cudaStream_t sfar = st;                          // far updates (split mode)
  for (int k = 0; k < T; ++k) {
    const int o = k % G, lk = k / G;
    const int rem = T - k;                          // tile rows below the diagonal block, incl. the extra row
    const int64_t prow = (int64_t)rem * NB;         // packed panel height
    double* pack = h->gPack + (int64_t)(k & 1) * ld * NB;
    if (r == o) {
      double* Akk = h->gA + (int64_t)k * NB + (int64_t)lk * NB * ld;
      double* Dk = h->gDinv + (int64_t)lk * NB * NB;
      if (k > 0) GPK_CK(h, cudaStreamWaitEvent(sp, ev_col[k], 0));          // column k has all its updates
      if (k > 1) GPK_CK(h, cudaStreamWaitEvent(sp, ev_upd[k - 2], 0));      // Pack[k%2] no longer read
      GPK_TRY(launch_diag(h, sp, Akk, ld, Dk, parts + k, h->dInfo, k * NB));
      GemmArgs t{};
      t.A = Akk + NB; t.B = Dk; t.C = Akk + NB; t.lda = ld; t.ldb = NB; t.ldc = ld; t.K = NB; t.tri = 0;
      GPK_TRY(launch_gemm_nt(h, sp, 0, t, rem, 1));
      GPK_CK(h, cudaMemcpy2DAsync(pack, (size_t)prow * sizeof(double), Akk + NB, (size_t)ld * sizeof(double),
                                  (size_t)prow * sizeof(double), NB, cudaMemcpyDeviceToDevice, sp));
      GPK_CK(h, cudaEventRecord(ev_packed[k], sp));
    }
    if (G > 1) {
      if (r == o) GPK_CK(h, cudaStreamWaitEvent(sc, ev_packed[k], 0));
      else if (k > 1) GPK_CK(h, cudaStreamWaitEvent(sc, ev_upd[k - 2], 0));  // receive buffer free
      NCCL_CK(h, g_nccl.bcast(pack, pack, (size_t)prow * NB, NCCL_F64, o, h->nccl_comm, sc));
      GPK_CK(h, cudaEventRecord(ev_bcast[k], sc));
      GPK_CK(h, cudaStreamWaitEvent(su, ev_bcast[k], 0));
    } else {
      GPK_CK(h, cudaStreamWaitEvent(su, ev_packed[k], 0));
    }
    if (doz) {
      const int kb = (k / WD) * WD, ke = (kb + WD < T) ? kb + WD : T;     // the block of panel k: [kb, ke)
      double* blk = PB + (int64_t)((k / WD) & 1) * ld * WD * NB;
      // panel k, full height, into its slot of the block buffer (rows (k+1)*NB .. ld)
      GPK_CK(h, cudaMemcpy2DAsync(blk + (int64_t)(k - kb) * NB * ld + (int64_t)(k + 1) * NB, (size_t)ld * sizeof(double),
                                  pack, (size_t)prow * sizeof(double), (size_t)prow * sizeof(double), NB,
                                  cudaMemcpyDeviceToDevice, su));
      // immediate updates: owned columns j in (k, ke) only
      const int j0 = k + 1 + (((r - (k + 1)) % G) + G) % G;
      int jstart = j0;
      if (j0 == k + 1 && j0 < ke) {
        GemmArgs u{};
        u.A = pack; u.B = pack; u.C = h->gA + (int64_t)(k + 1) * NB + (int64_t)(j0 / G) * NB * ld;
        u.lda = prow; u.ldb = prow; u.ldc = ld; u.K = NB; u.tri = 1; u.ti_off = k + 1; u.tj_off = j0; u.cstride = G;
        GPK_TRY(launch_gemm_nt(h, su, 1, u, rem, 1));
        GPK_CK(h, cudaEventRecord(ev_col[k + 1], su));
        jstart = j0 + G;
      }
      if (jstart < ke) {
        const int ncols = (ke - 1 - jstart) / G + 1;
        GemmArgs u{};
        u.A = pack; u.B = pack + (int64_t)(jstart - (k + 1)) * NB;
        u.C = h->gA + (int64_t)(k + 1) * NB + (int64_t)(jstart / G) * NB * ld;
        u.lda = prow; u.ldb = prow; u.ldc = ld; u.K = NB; u.tri = 1; u.ti_off = k + 1; u.tj_off = jstart; u.cstride = G;
        GPK_TRY(launch_gemm_nt(h, su, 1, u, rem, ncols));
      }
      if (k == ke - 1 && ke < T) {
        // the block is complete: one sliced rank-(ke-kb)*NB update of the owned columns >= ke, rows >= ke*NB
        const int b = k / WD, which = dsplit ? (b & 1) : 0;
        const int nrows = (T + 1 - ke) * NB, kw = (ke - kb) * NB;
        // slice buffer `which` was last read by the far update of block b-2
        if (dsplit && b >= 2) GPK_CK(h, cudaStreamWaitEvent(su, ev_far[b - 2], 0));
        GPK_TRY(launch_oz_slice(h, which, su, blk + (int64_t)ke * NB, ld, nrows, kw));
        const int jf = ke + (((r - ke) % G) + G) % G;                     // first owned column >= ke
        int nfar = 0, cfar = 0;
        double* Cfar = nullptr;
        if (jf < T) {
          int ncols = (T - 1 - jf) / G + 1, cfirst = jf - ke;
          double* C = h->gA + (int64_t)ke * NB + (int64_t)(jf / G) * NB * ld;
          int nnear = ncols;
          if (dsplit) {
            // near = owned columns inside the next block [ke, ke + WD)
            nnear = (jf < ke + WD) ? (std::min(ke + WD, T) - 1 - jf) / G + 1 : 0;
            nfar = ncols - nnear;
            Cfar = C + (int64_t)nnear * NB * ld;
            cfar = cfirst + nnear * G;
            // the near columns received the far update of block b-1: it must be complete
            if (b >= 1 && nnear > 0) GPK_CK(h, cudaStreamWaitEvent(su, ev_far[b - 1], 0));
          }
          if (nnear > 0) {
            if (jf == ke) {                                                // this rank owns the next panel: that column first
              GPK_TRY(launch_oz_cyclic(h, which, su, C, ld, nrows, kw, 1, cfirst, G));
              GPK_CK(h, cudaEventRecord(ev_col[ke], su));
              C += (int64_t)NB * ld; cfirst += G; --nnear;
            }
            GPK_TRY(launch_oz_cyclic(h, which, su, C, ld, nrows, kw, nnear, cfirst, G));
          }
        }
        if (dsplit) {
          GPK_CK(h, cudaEventRecord(ev_sliced, su));
          GPK_CK(h, cudaStreamWaitEvent(sfar, ev_sliced, 0));
          if (nfar > 0) GPK_TRY(launch_oz_cyclic(h, which, sfar, Cfar, ld, nrows, kw, nfar, cfar, G));
          GPK_CK(h, cudaEventRecord(ev_far[b], sfar));
        }
      }
      GPK_CK(h, cudaEventRecord(ev_upd[k], su));
      continue;
    }
    // update the owned columns j > k:  C[:, j] -= P[rows >= j] * P[j]^T
    const int j0 = k + 1 + (((r - (k + 1)) % G) + G) % G;
    int jstart = j0;
    if (j0 == k + 1 && j0 < T) {
      // this rank owns the next panel: bring that one column up to date first
      GemmArgs u{};
      u.A = pack; u.B = pack; u.C = h->gA + (int64_t)(k + 1) * NB + (int64_t)(j0 / G) * NB * ld;
      u.lda = prow; u.ldb = prow; u.ldc = ld; u.K = NB; u.tri = 1; u.ti_off = k + 1; u.tj_off = j0; u.cstride = G;
      GPK_TRY(launch_gemm_nt(h, st, 1, u, rem, 1));
      GPK_CK(h, cudaEventRecord(ev_col[k + 1], st));
      jstart = j0 + G;
    }
    if (jstart < T) {
      const int ncols = (T - 1 - jstart) / G + 1;
      GemmArgs u{};
      u.A = pack; u.B = pack + (int64_t)(jstart - (k + 1)) * NB;
      u.C = h->gA + (int64_t)(k + 1) * NB + (int64_t)(jstart / G) * NB * ld;
      u.lda = prow; u.ldb = prow; u.ldc = ld; u.K = NB; u.tri = 1; u.ti_off = k + 1; u.tj_off = jstart; u.cstride = G;
      GPK_TRY(launch_gemm_nt(h, st, 1, u, rem, ncols));
    }
    GPK_CK(h, cudaEventRecord(ev_upd[k], st));
  }
  // join the helper streams
  GPK_CK(h, cudaEventRecord(ev_fork, sf));
  GPK_CK(h, cudaStreamWaitEvent(st, ev_fork, 0));
  GPK_CK(h, cudaEventRecord(ev_fork, sp));
  GPK_CK(h, cudaStreamWaitEvent(st, ev_fork, 0));
  GPK_CK(h, cudaEventRecord(ev_fork, sc));
  GPK_CK(h, cudaStreamWaitEvent(st, ev_fork, 0));
  GPK_CK(h, cudaEventRecord(h->t2, st));

  // ---- backward substitution: x_k = Dinv_k^T (z_k - sum_{i>k} L[i,k]^T x_i), x_k broadcast ------------------------
  // Right-looking: as soon as x_k has arrived, every rank adds its contribution L[k, c]' x_k to the running sums of ALL
  // the columns c < k it owns (one launch, one CTA per owned column block), so the dependent chain per block step is
  // finish (one CTA) -> 1 KiB broadcast -> that one update launch, instead of a launch over the whole column height.
  // GPK_DIST_BWD=0 selects the round-1 left-looking form (one launch over the column's tiles per step).
  if (env_int("GPK_DIST_BWD", 1) != 0) {
    GPK_CK(h, cudaFuncSetAttribute(bwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM));
    double* acc = partial;                                   // nloc x 128 running sums (the buffer holds T x 128)
    GPK_CK(h, cudaMemsetAsync(acc, 0, (size_t)(nloc > 0 ? nloc : 1) * NB * sizeof(double), st));
    for (int k = T - 1; k >= 0; --k) {
      const int o = k % G, lk = k / G;
      double* xk = x + (int64_t)k * NB;
      if (r == o) {
        const double* Acol = h->gA + (int64_t)lk * NB * ld;
        bwd_finish_kernel<<<1, NB, 0, st>>>(Acol, ld, np, h->gDinv + (int64_t)lk * NB * NB, acc + (int64_t)lk * NB, 1, xk);
        h->stats.launches++;
      }
      if (G > 1) NCCL_CK(h, g_nccl.bcast(xk, xk, NB, NCCL_F64, o, h->nccl_comm, st));
      const int ncols = (k > r) ? (k - r + G - 1) / G : 0;   // owned column blocks with global index < k
      if (ncols > 0) {
        bwd_update_kernel<<<ncols, DT_THREADS, DT_SMEM, st>>>(h->gA, ld, k, xk, acc);
        h->stats.launches++;
      }
    }
  } else {
  for (int k = T - 1; k >= 0; --k) {
    const int o = k % G, lk = k / G;
    double* xk = x + (int64_t)k * NB;
    if (r == o) {
      const double* Acol = h->gA + (int64_t)lk * NB * ld;
      const int nt = T - 1 - k;
      if (nt > 0) bwd_partial_kernel<<<nt, DT_THREADS, DT_SMEM, st>>>(Acol, ld, x, k, partial);
      bwd_finish_kernel<<<1, NB, 0, st>>>(Acol, ld, np, h->gDinv + (int64_t)lk * NB * NB, partial, nt, xk);
      h->stats.launches += 2;
    }
    if (G > 1) NCCL_CK(h, g_nccl.bcast(xk, xk, NB, NCCL_F64, o, h->nccl_comm, st));
  }
  }
  // log-det parts and info: every entry has exactly one non-zero contributor
  if (G > 1) {
    NCCL_CK(h, g_nccl.allreduce(parts, parts, (size_t)T, NCCL_F64, NCCL_SUM, h->nccl_comm, st));
    NCCL_CK(h, g_nccl.allreduce(h->dInfo, h->dInfo, 1, NCCL_I32, NCCL_MAX, h->nccl_comm, st));
  }
  dist_finish_kernel<<<1, 1024, 0, st>>>(x, h->dR, 1.0 / sn2, np, h->dAlpha, res);
  GPK_TRY(launch_sum_parts(h, st, parts, T, res + 1));
  GPK_CK(h, cudaGetLastError());
  GPK_CK(h, cudaEventRecord(h->t3, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned, res, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(alpha, h->dAlpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t3); h->stats.total_ms = ms;
  cudaEventElapsedTime(&ms, h->t0, h->t1); h->stats.kbuild_ms = ms;
  cudaEventElapsedTime(&ms, h->t1, h->t2); h->stats.potrf_ms = ms;
  cudaEventElapsedTime(&ms, h->t2, h->t3); h->stats.solve_ms = ms;
  h->stats.h2d_bytes = (n + D) * (int64_t)sizeof(double);
  h->stats.d2h_bytes = (n + 2) * (int64_t)sizeof(double) + 4;
  const int info = *reinterpret_cast<int*>(h->hPinned + 2048);
  *nlZ = h->hPinned[0] / 2.0 + h->hPinned[1] + (double)n * std::log(2.0 * M_PI * sn2) / 2.0;
  h->kind = kind; h->matern_d = matern_d; h->nhyp = nhyp; h->sn2 = sn2; h->sf2 = sf2;
  h->hyp.assign(hyp, hyp + nhyp);
  h->dist_post = (info == 0);            // the distributed factor + alpha describe a posterior (gpk_dist_gather_factor)
  if (info != 0) return info;
  return 0;
}

int gpk_exact_eval_dist(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                        const double* ymm, double* nlZ, double* alpha) {
  return exact_eval_dist_impl(hh, kind, matern_d, hyp, nhyp, log_sn, ymm, nlZ, alpha);
}

int gpk_dist_reserve(gpk_handle hh, int level) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->dX) return GPK_ERR_STATE;
  return dist_reserve(h, level);
}

// Replicate the distributed factor on every rank: L into dA (np x np, lower), the block inverses into dDinv.  One packed
// trapezoid (rows >= k*128 of block column k) is broadcast per block column by its owner - 4 N^2 bytes in total over
// NVLink - and the 128x128 block inverses travel as one sum-all-reduce of a buffer in which every block has exactly one
// non-zero contributor.  Afterwards the handle is in the same state as after gpk_exact_eval: gpk_predict and
// gpk_get_factor work on every rank (each for its own share of the test points: no further communication).
int gpk_dist_gather_factor(gpk_handle hh) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (h->has_post && h->dist_post) return 0;          // already gathered for this posterior
  if (!h->dist_post) return GPK_ERR_STATE;
  const int G = h->world, r = h->rank;
  const int64_t np = h->np, ld = np + NB;
  const int T = (int)(np / NB);
  cudaStream_t st = h->s_main;
  GPK_TRY(ensure(h, &h->dA, &h->capA, np * np));
  GPK_TRY(ensure(h, &h->gPack, &h->cgPack, 2 * ld * NB));
  GPK_CK(h, cudaMemsetAsync(h->dDinv, 0, (size_t)np * NB * sizeof(double), st));
  for (int k = 0; k < T; ++k) {
    const int o = k % G, lk = k / G;
    const int64_t rows = np - (int64_t)k * NB;        // the lower trapezoid of block column k
    double* pack = h->gPack + (int64_t)(k & 1) * ld * NB;
    double* dst = h->dA + (int64_t)k * NB * np + (int64_t)k * NB;
    if (r == o) {
      const double* src = h->gA + (int64_t)lk * NB * ld + (int64_t)k * NB;
      GPK_CK(h, cudaMemcpy2DAsync(dst, (size_t)np * sizeof(double), src, (size_t)ld * sizeof(double),
                                  (size_t)rows * sizeof(double), NB, cudaMemcpyDeviceToDevice, st));
      GPK_CK(h, cudaMemcpyAsync(h->dDinv + (int64_t)k * NB * NB, h->gDinv + (int64_t)lk * NB * NB,
                                (size_t)NB * NB * sizeof(double), cudaMemcpyDeviceToDevice, st));
      if (G > 1)
        GPK_CK(h, cudaMemcpy2DAsync(pack, (size_t)rows * sizeof(double), src, (size_t)ld * sizeof(double),
                                    (size_t)rows * sizeof(double), NB, cudaMemcpyDeviceToDevice, st));
    }
    if (G > 1) {
      NCCL_CK(h, g_nccl.bcast(pack, pack, (size_t)rows * NB, NCCL_F64, o, h->nccl_comm, st));
      if (r != o)
        GPK_CK(h, cudaMemcpy2DAsync(dst, (size_t)np * sizeof(double), pack, (size_t)rows * sizeof(double),
                                    (size_t)rows * sizeof(double), NB, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (G > 1) NCCL_CK(h, g_nccl.allreduce(h->dDinv, h->dDinv, (size_t)np * NB, NCCL_F64, NCCL_SUM, h->nccl_comm, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  h->has_post = true; h->post_ep = false; h->has_fitc = false; h->pn = 0;
  return 0;
}

// Sharded counterpart of gpk_exact_eval(want_der = 1).  After the sharded factorisation and the gather, rank r owns the
// ROWS [i0, i1) of the inverse (a contiguous, tile-aligned share): they are the solutions of A X = E[:, i0:i1], computed
// as right-hand sides held transposed by the two triangular sweeps (forward from block column i0/128: everything before
// it is zero), and the fused Q o dK reduction runs over that rectangle (every pair of the full matrix is in exactly one
// rank's rectangle).  ONE all-reduce of nhyp+1 partial sums.
int gpk_exact_eval_dist_der(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                            const double* ymm, double* nlZ, double* alpha, double* dcov, double* dlik) {
  if (!dcov || !dlik) return GPK_ERR_ARG;
  GPK_TRY(exact_eval_dist_impl(hh, kind, matern_d, hyp, nhyp, log_sn, ymm, nlZ, alpha));
  GPK_TRY(gpk_dist_gather_factor(hh));
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  const int G = h->world, r = h->rank;
  const int64_t n = h->n, np = h->np;
  const int D = h->D, T = (int)(np / NB);
  cudaStream_t st = h->s_main;
  const int t0 = (int)(((int64_t)T * r) / G), t1 = (int)(((int64_t)T * (r + 1)) / G);
  const int64_t i0 = (int64_t)t0 * NB, rows = (int64_t)(t1 - t0) * NB;
  const int64_t g = (n + 63) / 64;
  const int64_t gi = (rows + 63) / 64 > 0 ? (rows + 63) / 64 : 1;
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, gi * g * 34 + 64));
  GPK_CK(h, cudaEventRecord(h->t3, st));
  double* res = h->gVec + np + T;                       // the result slots of the sharded evaluation (16 doubles)
  double* dres = h->dScal + T + 8;                      // [dcov..., trace]
  GPK_CK(h, cudaMemsetAsync(dres, 0, (size_t)(nhyp + 1) * sizeof(double), st));
  (void)res;
  if (rows > 0) {
    GPK_TRY(ensure(h, &h->dU, &h->capU, np * np));                 // L' (upper)
    GPK_TRY(ensure(h, &h->dDinvT, &h->capDinvT, np * NB));
    GPK_TRY(ensure(h, &h->dW, &h->capW, rows * np));               // this rank's rows of the inverse
    GPK_TRY(launch_transpose(h, st, h->dA, np, 0, h->dU, np, 0, np, np, 1));
    GPK_TRY(launch_transpose(h, st, h->dDinv, NB, (int64_t)NB * NB, h->dDinvT, NB, (int64_t)NB * NB, NB, NB, T));
    GPK_CK(h, cudaMemsetAsync(h->dW, 0, (size_t)rows * np * sizeof(double), st));
    GPK_TRY(launch_set_identity(h, st, h->dW + i0 * rows, rows, rows, rows));   // P[li, i0 + li] = 1
    GPK_TRY(sweep_forward(h, st, h->dW, rows, (int)(rows / NB), h->dA, np, h->dDinv, T, t0));
    GPK_TRY(sweep_backward(h, st, h->dW, rows, (int)(rows / NB), h->dU, np, h->dDinvT, T));
    GPK_TRY(launch_dnlz_rect(h, st, h->dXs, n, D, h->dW, rows, i0, rows, h->dAlpha, 1.0 / h->sn2, h->sf2, kind,
                             matern_d, h->dTmp, h->capTmp, dres));
  }
  if (G > 1) NCCL_CK(h, g_nccl.allreduce(dres, dres, (size_t)(nhyp + 1), NCCL_F64, NCCL_SUM, h->nccl_comm, st));
  GPK_CK(h, cudaEventRecord(h->t4, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 8, dres, (size_t)(nhyp + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t3, h->t4); h->stats.deriv_ms = ms;
  h->stats.total_ms += ms;
  for (int i = 0; i < nhyp; ++i) dcov[i] = h->hPinned[8 + i] / 2.0;
  dlik[0] = h->sn2 * h->hPinned[8 + nhyp];
  return 0;
}

}  // extern "C"
