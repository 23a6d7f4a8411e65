// Internal declarations shared by the libgpk translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>
#include "../../include/gpk.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgpk is written for sm_100a (B200) only"
#endif

namespace gpk {

constexpr int NB = 128;          // tile edge == panel width of the blocked Cholesky
constexpr int GEMM_LDS = NB + 4; // smem column pitch (doubles); == 4 mod 16 -> conflict-free DMMA fragment loads

constexpr int DIAG_THREADS = 256;
constexpr int DIAG_IB = 32;
constexpr int DIAG_LDS = NB + 4;
constexpr int DIAG_LDT = DIAG_IB + 4;
constexpr size_t DIAG_SMEM = (size_t(NB) * DIAG_LDS + 4 * DIAG_IB * DIAG_LDT + NB) * sizeof(double);

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
inline int env_int(const char* name, int dflt) {   // GPK_* switches for A/B measurements; read at every call
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---- argument blocks ------------------------------------------------------
// C(tile ti,tj) (+)= A(rows ti) * B(rows tj)^T ; all column-major, all tile-aligned.
struct GemmArgs {
  const double* A; const double* B; double* C;
  int64_t lda, ldb, ldc;
  int K;            // contraction length (multiple of 32)
  int ti_off;       // global tile index of grid row 0   (triangular tests)
  int tj_off;       // global tile index of grid column 0
  int tri;          // 0: all tiles; 1: only gi>=gj, diagonal tiles store row>=col;
                    // 2: as 1, and the contraction starts at k = gi*NB (U*U^T of an upper-triangular U)
  int strips;       // 1: compute in 32-row strips (four CTAs per 128-row tile) even when the product is not in place
  int cstride;      // block-cyclic columns (multi-GPU): grid column tile tc is GLOBAL column tile tj_off + tc*cstride;
                    // B rows advance by cstride*NB per tc while C columns stay packed.  0/1 = contiguous.
};

// one 128x128 tile of C (op)= A(128 x K) * B(128 x K)' by sixteen 32x32-block CTAs (potrf_diag.cu: small_nt_kernel)
struct SmallArgs {
  const double* A; const double* B; double* C;
  int64_t lda, ldb, ldc;
  int K;               // multiple of 128
  int mode;            // 0: C = A B' ; 1: C -= A B'
  int tri;             // only blocks on / below the diagonal; diagonal blocks store row >= col
  double* copy_dst;    // K == 128 only: block column 0 also stores its 32 x 128 strip of A here (pitch ld_copy)
  int64_t ld_copy;
  int breg;            // K == 128 only: B goes global -> registers, 37 KB of shared memory per CTA instead of 74 KB
};

enum CovEpi { EPI_COV = 0, EPI_DER_ELL = 1, EPI_DER_SF = 2, EPI_DER_ARD = 3 };

// ---- covariance programs: composite kernels evaluated on the device (covprog.cu) ---------------------------------
// leaf ops 0..2 coincide with GPK_COV_RBF / RBFARD / MATERN
enum CovOp {
  OP_RBF = 0, OP_RBFARD = 1, OP_MATERN = 2, OP_RBFUNIT = 3, OP_RQ = 4, OP_RQARD = 5, OP_PERIODIC = 6, OP_PIECEPOLY = 7,
  OP_GABOR = 8, OP_NOISE = 9, OP_CONST = 10, OP_LINEAR = 11, OP_POLY = 12, OP_PRE = 13,
  OP_SUM = 32, OP_PROD = 33, OP_SCALE = 34
};
constexpr int PROG_MAX_NODES = 32, PROG_MAX_HYP = 96, PROG_MAX_ARD = 2, PROG_MAX_D = 64;
constexpr int GPK_COV_PROG = 3;      // Handle::kind of a posterior built from a program
struct ProgNode {
  int op, a, b, h0;                  // children (node indices, post-order: < own index); first hyper-parameter index
  int ard, ipar;                     // ARD weight set; integer parameter (Matern d, PiecePoly v, Poly order)
  double p0, p1, p2, p3;             // derived parameters (see prog_compile)
};
struct CovProg {
  int n_nodes, nhyp, n_ard, D;
  ProgNode node[PROG_MAX_NODES];
  double ardw[PROG_MAX_ARD][PROG_MAX_D];   // 1/ell_d^2 per ARD leaf
};

struct CovArgs {
  const double* F;  // scaled inputs indexed by the FAST output index (nF, D) row-major
  const double* S;  // scaled inputs indexed by the SLOW output index (nS, D)
  double* out;      // out[f + s*ld]
  int64_t ld;
  int64_t nF, nS;   // valid extents
  int64_t pF, pS;   // padded extents actually written (>= nF/nS)
  int D;
  int kind;         // GPK_COV_*
  int matern_d;
  int epi;          // CovEpi
  int ard_dim;      // for EPI_DER_ARD
  double sf2;
  double scale;     // out = value*scale (+ diag_add on f==s)
  double diag_add;
  int same_set;     // F and S are the same point set (train mode): f==s is the diagonal
  int lower_only;   // skip tiles entirely above the diagonal (f-tile < s-tile); zero strict upper inside diagonal tiles
  int pad_identity; // padded diagonal entries (f==s>=nF) get 1.0 instead of 0.0
  const CovProg* prog;   // non-null: evaluate this program (DEVICE pointer) on the RAW inputs F, S instead of `kind`
  int prog_der1;         //   0: the covariance; h+1: the derivative w.r.t. hyper-parameter h (getDerMatrix)
  const double* pre;     //   OP_PRE leaf: uploaded training matrix (cov.Pre), entry (f, s) at pre[f + s*pre_ld]
  int64_t pre_ld;
  int padded128;    // F and S are allocated (and zero beyond nF / nS) up to a multiple of 128 points: enables
                    // cov_tile_kernel (bulk-copied 128-point blocks)
  int s_bstride;    // block-cyclic slow index (multi-GPU): local s maps to the GLOBAL point
  int s_boff;       //   (s/128)*s_bstride*128 + s_boff*128 + s%128 ; 0 = identity.  nS bounds the global index.
};

// ---- handle ---------------------------------------------------------------
struct Handle {
  int device = 0;
  cudaStream_t s_main = nullptr, s_panel = nullptr, s_aux = nullptr, s_tail = nullptr;
  std::vector<cudaEvent_t> ev;          // dependency events (no timing)
  cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr, t3 = nullptr, t4 = nullptr;
  cudaError_t last_cuda = cudaSuccess;
  std::string last_msg;
  gpk_stats stats{};
  std::set<const void*> smem_attr;      // kernels whose dynamic shared-memory limit was raised ON THIS DEVICE
  int profile = 0;
  std::vector<cudaEvent_t> prof_ev;
  int prof_pairs = 0;

  // training data
  int64_t n = 0, np = 0; int D = 0;
  double* dX = nullptr;      // (n,D) as given
  double* dXs = nullptr;     // (np,D) scaled by the kernel's length scales, zero padded
  double* dScale = nullptr;  // (D) per-dimension multipliers
  // factor storage
  double* dA = nullptr; int64_t capA = 0;    // (np,np) column-major, lower = L
  double* dDinv = nullptr;   // (np,NB): inverses of the diagonal blocks, block k at rows k*NB
  double* dB = nullptr;      // (np) forward-solve work vector
  double* dZ = nullptr;      // (np) z = L^-1 r
  double* dAlpha = nullptr;  // (np) alpha
  double* dR = nullptr;      // (np) y - m
  double* dScal = nullptr;   // scalars: [0..T) logdet parts, then results
  int* dInfo = nullptr;
  void* epGraphExec = nullptr; const void* epGraphSig[8] = {};   // the EP block's site launches as a CUDA graph (ep.cu)
  int* dFlags = nullptr; int flag_epoch = 0;   // epoch-stamped ready flags of the persistent backward substitution
  double* hPinned = nullptr; // small pinned staging
  double* dHead = nullptr;   // (128,128) scratch tile of the panel chain: the solved head tile between its two products
  // posterior state
  bool has_post = false; int kind = 0, matern_d = 3, nhyp = 0; double sn2 = 1.0, sf2 = 1.0;
  std::vector<double> hyp;
  // derivative / predict / fitc work buffers (lazy)
  double* dU = nullptr; int64_t capU = 0;
  double* dW = nullptr; int64_t capW = 0;
  double* dP = nullptr; int64_t capP = 0;
  double* dTmp = nullptr; int64_t capTmp = 0;
  double* dXtmp = nullptr; int64_t capXtmp = 0;
  // covariance program of the current evaluation / posterior (covprog.cu)
  CovProg hprog{}; CovProg* dProg = nullptr; CovProg* hProgPinned = nullptr;
  double* dPre = nullptr; int64_t capPre = 0; int64_t preN = 0;     // cov.Pre: uploaded training matrix (n x n)
  // standalone potrf state
  int64_t pn = 0;
  bool lt_valid = false;                       // dU holds L' and dDinvT the transposed block inverses of the CURRENT factor
  double* dDinvT = nullptr; int64_t capDinvT = 0;
  // EP (classification) state; see ep.cu
  bool post_ep = false;
  double *eK = nullptr, *eSig = nullptr, *eVec = nullptr, *eSW = nullptr;
  int64_t ceK = 0, ceSig = 0, ceVec = 0, ceSW = 0;
  // multi-GPU (block-cyclic columns over NCCL); see dist.cu
  void* nccl_comm = nullptr; int rank = 0, world = 1;
  bool dist_post = false;                       // the distributed factor (gA) + dAlpha describe the current posterior
  double* gA = nullptr; int64_t cgA = 0;        // local columns of the (np+128) x np augmented matrix
  double* gDinv = nullptr; int64_t cgDinv = 0;  // inverses of the owned diagonal blocks
  double* gPack = nullptr; int64_t cgPack = 0;  // packed panel being broadcast
  double* gBlk = nullptr; int64_t cgBlk = 0;    // blocked variant: the panels of the current block, full height (x2)
  double* gVec = nullptr; int64_t cgVec = 0;
  // int8 tensor-core trailing update (ozaki.cu): slices + row exponents of the current panel block
  // (two sets: [0] level-1 updates on s_main, [1] level-2 updates on s_panel, which run concurrently)
  int8_t* ozSl[2] = {nullptr, nullptr}; size_t ozCap[2] = {0, 0}; double* ozSc[2] = {nullptr, nullptr}; size_t ozScCap[2] = {0, 0};
  long long* ozDbg = nullptr;
  double* ozFix = nullptr; size_t ozFixCap = 0;      // fixed row scales of the running factorisation (one per matrix row)
  // fitc state
  bool has_fitc = false; int64_t M = 0, Mp = 0;
  double* dUin = nullptr; double* dUs = nullptr; double* dLpost = nullptr; double* dAlphaU = nullptr;
  int64_t capUin = 0, capUs = 0, capLpost = 0, capAlphaU = 0;
  double *fKuu = nullptr, *fDinvU = nullptr, *fA2 = nullptr, *fDinv2 = nullptr, *fVt = nullptr, *fVs = nullptr,
         *fVec = nullptr, *fWt = nullptr;
  int64_t cKuu = 0, cDinvU = 0, cA2 = 0, cDinv2 = 0, cVt = 0, cVs = 0, cVec = 0, cWt = 0;
};

#define GPK_CK(h, call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      (h)->last_cuda = e__;                                                    \
      (h)->last_msg = std::string(#call) + ": " + cudaGetErrorString(e__);     \
      return (e__ == cudaErrorMemoryAllocation) ? GPK_ERR_NOMEM : GPK_ERR_CUDA; \
    }                                                                          \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember it per handle, not per process
#define GPK_SMEM_ATTR(h, func, bytes)                                                                          \
  do {                                                                                                         \
    if ((h)->smem_attr.find((const void*)(func)) == (h)->smem_attr.end()) {                                    \
      GPK_CK(h, cudaFuncSetAttribute((func), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      (h)->smem_attr.insert((const void*)(func));                                                              \
    }                                                                                                          \
  } while (0)

#define GPK_TRY(expr)                 \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != 0) return rc__;       \
  } while (0)

// ---- launchers (defined in the .cu files) ----------------------------------
int launch_gemm_nt(Handle* h, cudaStream_t st, int mode /*0 set, 1 sub*/, const GemmArgs& a, int tiles_m, int tiles_n);
int launch_cov(Handle* h, cudaStream_t st, const CovArgs& a);
int prog_compile(const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp, int D, CovProg* out);
bool prog_has_op(const CovProg& p, int op);
int prog_upload(Handle* h, cudaStream_t st, const CovProg& p);
int launch_cov_prog(Handle* h, cudaStream_t st, const CovArgs& a);
int launch_cov_prog_diag(Handle* h, cudaStream_t st, const CovProg* dprog, const double* Z, int64_t m, int D, int der,
                         double* out);
int launch_dnlz_prog(Handle* h, cudaStream_t st, const CovProg* dprog, int nhyp, const double* X, int64_t n, int D,
                     const double* Ainv, int64_t ld, const double* alpha, double inv_sn2, const double* pre,
                     int64_t pre_ld, double* part, int64_t part_cap, double* res);
int launch_prescale(Handle* h, cudaStream_t st, const double* X, int64_t n, int64_t np, int D,
                    const double* scale, int divide, double premul, double* out);
int launch_diag(Handle* h, cudaStream_t st, double* Ablk, int64_t lda, double* Dinv, double* logdet_slot,
                int* info, int gidx0, long long* dbg_clk = nullptr);
int launch_diag_invert(Handle* h, cudaStream_t st, double* A, int64_t lda, double* Dinv, double* logdet_parts, int* info,
                       int T);
int launch_trsv_fwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* b,
                    double* z, int k, int T);
int launch_trsv_bwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* z,
                    double* x, int k, int T);
int launch_trsv_bwd_all(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* z,
                        double* x, int T);   // all T steps; one cooperative launch when T <= #SMs
int launch_finish_alpha(Handle* h, cudaStream_t st, const double* x, const double* r, double inv_sn2, int64_t np,
                        double* alpha, const double* parts, int T, double* res);
int launch_sum_parts(Handle* h, cudaStream_t st, const double* parts, int T, double* res);
int launch_pad_sym(Handle* h, cudaStream_t st, const double* src, int64_t n, double* dst, int64_t pn);
int launch_compact_lower(Handle* h, cudaStream_t st, const double* src, int64_t ld, int64_t n, double* dst);
int launch_compact_sym(Handle* h, cudaStream_t st, const double* src, int64_t ld, int64_t n, double* dst);
int launch_set_identity(Handle* h, cudaStream_t st, double* M, int64_t ld, int64_t rows, int64_t cols);
int launch_rowdot(Handle* h, cudaStream_t st, const double* P, int64_t ld, int64_t rows, int64_t cols, const double* v,
                  int mode, double scale, double kss, double* part, int nsplit, double* out, int64_t nvalid,
                  const double* kss_vec = nullptr /*mode 1: per-point prior variances instead of the scalar kss*/);
int launch_dnlz(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv, int64_t ld,
                const double* alpha, double inv_sn2, double sf2, int kind, int matern_d, double* part,
                int64_t part_cap, double* res);
int launch_dnlz_sw(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv, int64_t ld,
                   const double* alpha, const double* sw, double sf2, int kind, int matern_d, double* part,
                   int64_t part_cap, double* res);
int launch_colscale_inplace(Handle* h, cudaStream_t st, double* P, int64_t ld, int64_t rows, int64_t cols,
                            const double* s);
int launch_copy(Handle* h, cudaStream_t st, const double* src, double* dst, int64_t n);
int launch_transpose(Handle* h, cudaStream_t st, const double* src, int64_t lds, int64_t sstride, double* dst,
                     int64_t ldd, int64_t dstride, int64_t rows, int64_t cols, int batch);
int launch_fill_random(Handle* h, cudaStream_t st, double* p, int64_t n, unsigned seed);
int bench_dmma(Handle* h, int shape, int warps, int iters, double* tflops, double* ms_out);
int gemm_init(Handle* h);
int diag_init(Handle* h);
int launch_small_nt(Handle* h, cudaStream_t st, const SmallArgs& a);

int ensure(Handle* h, double** p, int64_t* cap, int64_t need_elems);
int ensure_zero(Handle* h, double** p, int64_t* cap, int64_t need_elems);   // zero-filled when (re)allocated
int oz_ensure(Handle* h, int which, int64_t n, int kw);
int launch_oz_slice(Handle* h, int which, cudaStream_t st, const double* P, int64_t lda, int n, int kw, int row0 = 0,
                    int ntot = 0);
int launch_oz_ex(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int jb0, int jb1,
                 int skip00, int ti_min, int trap, int cmode /*0: C -= PP', 1: C = PP', 2: C += PP'*/);
int launch_oz_cyclic(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int ncols, int cfirst,
                     int cs);
int launch_oz_gemm_stacked(Handle* h, int which, cudaStream_t st, double* Cab, int64_t ldc, int nb, int na, int kw);
int launch_oz_syrk(Handle* h, int which, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int jb0, int jb1,
                   int skip00 = 0);
int launch_oz_fixed_scales(Handle* h, cudaStream_t st, const double* A, int64_t lda, int n, double diag_const, double* sc);
int launch_oz_slice_fixed(Handle* h, cudaStream_t st, const double* P, int64_t lda, int nrows, int kw, int8_t* sl,
                          double* sc, int srows, int row0, int kstep0);
int launch_oz_syrk_buf(Handle* h, cudaStream_t st, const int8_t* sl, const double* sc, int srows, double* C, int64_t ldc,
                       int n, int kw, int jb0, int jb1, int skip00);
int oz_slices();
int dist_allreduce_sum(Handle* h, double* buf, size_t count, cudaStream_t st);
int kind_scale(int kind, int matern_d, const double* hyp, int nhyp, int D, std::vector<double>& scale, int* divide,
               double* premul, double* sf2);
void stats_begin(Handle* h);
int check_handle(gpk_handle hh, Handle** out);
int sweep_forward(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* A, int64_t lda,
                  const double* Dinv, int T, int kstart = 0);
int launch_dnlz_rect(Handle* h, cudaStream_t st, const double* Xs, int64_t n, int D, const double* Ainv_rows, int64_t ld,
                     int64_t i_off, int64_t rows, const double* alpha, double inv_sn2, double sf2, int kind, int matern_d,
                     double* part, int64_t part_cap, double* res);
int sweep_backward(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* Lt, int64_t ldt,
                   const double* DinvT, int T);
int inverse_factor_T(Handle* h, cudaStream_t st, double* U, const double* A, int64_t np, const double* Dinv);
int inverse_factor_T_oz(Handle* h, cudaStream_t st, double* U, const double* A, int64_t np, const double* Dinv);
int oz_gemm_nt(Handle* h, cudaStream_t st, double* C, int64_t ldc, const double* A, int64_t lda, int na, const double* B,
               int64_t ldb, int nb, int K, int cmode);
int sweep_forward_oz(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* A, int64_t lda,
                     const double* Dinv, int T);
int potrf_device(Handle* h, double* A, int64_t np, double* Dinv, double* logdet_parts, int* info,
                 double* b_fwd /*nullable: fused forward solve in/out*/, double* z_out,
                 const CovArgs* lazy_cov = nullptr /*generate the matrix inside, overlapped with the first panels*/);

}  // namespace gpk
