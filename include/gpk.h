/*
 * gpk.h - C ABI of libgpk.so: the B200 (sm_100a) exact-GP hot path behind the
 * pyGPs plugin API.
 *
 * The reference (marionmari/pyGPs @ 792f3c6) is pure Python and has NO FFI
 * boundary on this path; native code is entered only inside numpy/scipy.  Each
 * entry point below therefore names the reference *Python* interface whose work
 * it takes over (file:line under /root/reference/pyGPs/).  The ctypes binding a
 * maintainer would add to the reference is shown in INTEGRATION.md; this
 * repository's own binding is pygps_b200/_lib.py.
 *
 * Conventions
 *   - all functions return int: 0 = ok; > 0 = LAPACK-style `info` (1-based index
 *     of the first non-positive pivot: the matrix is not positive definite; the
 *     Python wrapper raises np.linalg.LinAlgError exactly where
 *     Core/tools.py:67,77 does); < 0 = GPK_ERR_* (bad argument / CUDA failure),
 *     text from gpk_strerror().  Nothing aborts; NaN/Inf propagate into outputs.
 *   - every pointer is a HOST pointer to caller-owned, C-contiguous float64
 *     memory unless the name says otherwise.  The handle owns all device memory,
 *     streams and events.  A handle is not re-entrant; use one per thread/GPU.
 *   - hyper-parameters are the reference's LOG values, in the reference's order
 *     (cov.X.hyp lists, Core/cov.py:793,882,1089; lik.Gauss.hyp[0], Core/lik.py:132).
 *   - there is no CPU fallback: without a usable CUDA device every call fails
 *     with GPK_ERR_CUDA.
 */
#ifndef GPK_H_
#define GPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpk_handle_s* gpk_handle;

/* covariance kinds: cov.RBF (Core/cov.py:786), cov.RBFard (:872), cov.Matern (:1078) */
enum { GPK_COV_RBF = 0, GPK_COV_RBFARD = 1, GPK_COV_MATERN = 2 };
/* getCovMatrix modes (Core/cov.py:82-93) */
enum { GPK_MODE_TRAIN = 0, GPK_MODE_CROSS = 1, GPK_MODE_SELF_TEST = 2 };

enum {
  GPK_OK = 0,
  GPK_ERR_ARG = -1,      /* bad argument (NULL, size, kind, der index ...)     */
  GPK_ERR_CUDA = -2,     /* CUDA runtime error; see gpk_strerror / last_error  */
  GPK_ERR_STATE = -3,    /* call order: no data / no posterior on the handle   */
  GPK_ERR_NOMEM = -4     /* device allocation failed                           */
};

/* per-call device timings (CUDA events on the library's own streams), ms */
typedef struct gpk_stats {
  double total_ms;       /* whole call on device, first kernel to last          */
  double kbuild_ms;      /* covariance-matrix build (cov.*.getCovMatrix)        */
  double potrf_ms;       /* blocked Cholesky incl. fused forward solve          */
  double solve_ms;       /* backward solve + nlZ reduction                      */
  double deriv_ms;       /* inverse + dnlZ reduction (want_der only)            */
  double syrk_ms;        /* sum of trailing-update launches (when profiled)     */
  double syrk_flops;     /* algorithmic flops of those launches                 */
  int64_t launches;      /* kernels launched by the call                        */
  int64_t h2d_bytes;     /* host->device bytes moved by the call                */
  int64_t d2h_bytes;     /* device->host bytes moved by the call                */
} gpk_stats;

/* ---- library / device --------------------------------------------------- */
int         gpk_version(void);
const char* gpk_strerror(int code);
int         gpk_device_count(int* count);
/* free / total bytes of device memory: the host side routes an evaluation whose N x N factor does not fit one GPU to
 * the sharded path (gpk_exact_eval_dist). */
int         gpk_device_memory(int device, int64_t* free_bytes, int64_t* total_bytes);

/* One handle = one GPU.  `device` is the CUDA ordinal. */
int gpk_create(int device, gpk_handle* out);
int gpk_destroy(gpk_handle h);
const char* gpk_last_error(gpk_handle h);          /* text of the last CUDA error */
int gpk_last_stats(gpk_handle h, gpk_stats* out);
/* profile != 0: two timing events are recorded on the trailing-update stream
 * around each level-1 update (gpk_stats.syrk_ms / syrk_flops).  Nothing is
 * synchronised and no stream dependency is added, so the look-ahead schedule -
 * and the evaluation time - are the same with and without it. */
int gpk_set_profile(gpk_handle h, int profile);

/* ---- cov.*.getCovMatrix / getDerMatrix ---------------------------------- *
 * Replaces Kernel.getCovMatrix(x,z,mode) (Core/cov.py:796-808, 887-904,
 * 1124-1148) and getDerMatrix (:811-828, :906-938, :1150-1182) including the
 * scipy cdist('sqeuclidean') + np.exp they call.
 *   train:     X (n,D)            -> out (n,n)
 *   cross:     X (n,D), Z (m,D)   -> out (n,m)
 *   self_test: Z (m,D)            -> out (m,1)
 * der < 0: covariance; der >= 0: derivative w.r.t. hyp[der].
 * matern_d in {1,3,5,7} (ignored for the RBF kinds).                        */
int gpk_cov_matrix(gpk_handle h, int kind, int matern_d,
                   const double* hyp, int nhyp,
                   const double* X, int64_t n, const double* Z, int64_t m, int D,
                   int mode, int der, double* out);

/* ---- composite kernels on the device (Core/cov.py:230-328 and the other stationary kernels) --------------------- *
 * A covariance function is a PROGRAM: the expression tree of SumOfKernel (:265) / ProductOfKernel (:230) /
 * ScaleOfKernel (:299) over leaf kernels, as an array of nodes in post-order (children before parents, root last).
 * `hyp` is the composite's flat list of LOG hyper-parameters exactly as the reference concatenates it (cov1.hyp +
 * cov2.hyp; ScaleOfKernel: [scalar] + cov.hyp); node.hyp0 is the index of the node's first own entry in it.
 * node.para: Matern d (:1078), PiecePoly v (:683), Poly order (:623); unused otherwise.
 * Every thread evaluates the program for its own matrix entries from one pass over the pair's coordinates; the same
 * program run backwards yields all hyper-parameter derivatives for the fused dnlZ reduction (Core/inf.py:376-377).  */
enum {
  GPK_OP_RBF = 0, GPK_OP_RBFARD = 1, GPK_OP_MATERN = 2, GPK_OP_RBFUNIT = 3, GPK_OP_RQ = 4, GPK_OP_RQARD = 5,
  GPK_OP_PERIODIC = 6, GPK_OP_PIECEPOLY = 7, GPK_OP_GABOR = 8, GPK_OP_NOISE = 9, GPK_OP_CONST = 10, GPK_OP_LINEAR = 11,
  GPK_OP_POLY = 12, GPK_OP_PRE = 13,
  GPK_OP_SUM = 32, GPK_OP_PROD = 33, GPK_OP_SCALE = 34
};
typedef struct gpk_cov_node {
  int32_t op;      /* GPK_OP_*                                                          */
  int32_t a, b;    /* child node indices (SUM, PROD: a and b; SCALE: a); -1 for leaves */
  int32_t hyp0;    /* index of this node's first hyper-parameter in `hyp`              */
  double para;     /* Matern d / PiecePoly v / Poly order                               */
} gpk_cov_node;

/* getCovMatrix / getDerMatrix of a composite (same modes and layouts as gpk_cov_matrix).                          */
int gpk_cov_matrix_prog(gpk_handle h, const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp,
                        const double* X, int64_t n, const double* Z, int64_t m, int D, int mode, int der, double* out);
/* cov.Pre (Core/cov.py:1429-1455): upload the precomputed TRAINING matrix M2 (n,n) for the GPK_OP_PRE leaf of the
 * programs evaluated afterwards on this handle (the cross/self-test parts of M1 stay with the caller).           */
int gpk_set_pre(gpk_handle h, const double* Ktrain, int64_t n);
/* inf.Exact.evaluate (Core/inf.py:353-384) for a composite: as gpk_exact_eval, dcov has nhyp entries in the
 * reference's order.  gpk_predict / gpk_get_factor work afterwards (programs without a PRE leaf).                */
int gpk_exact_eval_prog(gpk_handle h, const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp,
                        double log_sn, const double* ymm, int want_der,
                        double* nlZ, double* alpha, double* dcov, double* dlik);

/* ---- tools.jitchol / tools.solve_chol (Core/tools.py:31-97) ------------- *
 * gpk_potrf: A (n,n) symmetric (only the lower triangle is read) -> R (n,n)
 * C-order UPPER factor with an exactly zero strict lower triangle, A = R'R.
 * This is jitchol(A).T as the reference uses it (Core/inf.py:362).  The factor
 * stays resident on the handle for gpk_potrs.  logdet_half = sum(log(diag R)).
 * gpk_set_factor: make a factor computed elsewhere resident (solve_chol(R,B)
 * with an R that is not the last gpk_potrf result; post.L of a posterior kept
 * on the host, Core/gp.py:404-416): R (n,n) C-order upper; info > 0 = first
 * non-positive diagonal entry.
 * gpk_potrs: X = (R'R)^-1 B for B (n,nrhs) - solve_chol(R,B) - by two
 * triangular sweeps over the resident factor.  n must equal the resident
 * factor's order (GPK_ERR_ARG otherwise): B and X_out are n x nrhs.          */
int gpk_potrf(gpk_handle h, const double* A, int64_t n, double* R_out, double* logdet_half);
int gpk_set_factor(gpk_handle h, const double* R, int64_t n, double* logdet_half);
int gpk_potrs(gpk_handle h, const double* B, int64_t n, int64_t nrhs, double* X_out);

/* ---- inf.Exact.evaluate (Core/inf.py:353-384) --------------------------- *
 * gpk_set_data uploads the training inputs once (GP.setData, Core/gp.py:131). */
int gpk_set_data(gpk_handle h, const double* X, int64_t n, int D);
/* One evaluation: K build -> chol(K/sn2+I) -> alpha -> nlZ [-> dnlZ].
 *   ymm      (n)    y - m(x), the mean-subtracted targets (Core/inf.py:358,363)
 *   alpha    (n)    out: post.alpha                                (:364)
 *   nlZ      (1)    out                                            (:370)
 *   dcov     (nhyp) out if want_der: dnlZ.cov                      (:376-377)
 *   dlik     (1)    out if want_der: dnlZ.lik                      (:374)
 * post.sW = 1/sn is formed by the caller (:366); dnlZ.mean = -dm'alpha is an
 * O(n) host product (:378-381).  The factor stays on the device.            */
int gpk_exact_eval(gpk_handle h, int kind, int matern_d,
                   const double* hyp, int nhyp, double log_sn,
                   const double* ymm, int want_der,
                   double* nlZ, double* alpha, double* dcov, double* dlik);
/* post.L: the (n,n) C-order upper factor of K/sn2+I (Core/inf.py:367), copied
 * out only when the caller touches it. */
int gpk_get_factor(gpk_handle h, double* R_out);

/* ---- GP.predict solves (Core/gp.py:404-419, Cholesky branch) ------------ *
 * Uses the posterior left on the handle by gpk_exact_eval.
 *   Xs (ns,D) -> ks_alpha (ns) = Ks'alpha ;  fs2 (ns) = max(kss - colsum(V*V), 0)
 * with V = R'^-1 (Ks/sn).  The caller adds the prior mean and lik.Gauss's sn2. */
int gpk_predict(gpk_handle h, const double* Xs, int64_t ns, double* ks_alpha, double* fs2);

/* ---- inf.FITC_Exact.evaluate (Core/inf.py:398-455) ---------------------- *
 *   U (M,D) inducing inputs (cov.FITCOfKernel.inducingInput, Core/cov.py:339)
 *   alpha (M) out: post.alpha ; Lpost (M,M) out: post.L (dense, C-order)
 *   dcov/dlik as for gpk_exact_eval; al (n) out if want_der: (Kt+sn2 I)^-1 (y-m)
 *   (needed by the caller for dnlZ.mean, :449-451).                          */
int gpk_fitc_eval(gpk_handle h, int kind, int matern_d,
                  const double* hyp, int nhyp, double log_sn,
                  const double* U, int64_t M, const double* ymm, int want_der,
                  double* nlZ, double* alpha, double* Lpost,
                  double* dcov, double* dlik, double* al);
/* FITC branch of GP.predict (Core/gp.py:418): fs2 = kss + colsum(Ks*(L Ks)). */
int gpk_fitc_predict(gpk_handle h, const double* Xs, int64_t ns, double* ks_alpha, double* fs2);

/* ---- inf.EP.evaluate with lik.Erf (Core/inf.py:731-806; BASELINE config 5) -- *
 * Binary GP classification by Expectation Propagation, sites visited in the reference's fixed order.
 *   mvec (n) prior mean m(x) ; y (n) labels (sign is used, 0 -> +1)
 *   ttau_io, tnu_io (n): site parameters; read as the warm start when use_last != 0 (EP.last_ttau / last_tnu,
 *                        Core/inf.py:738-753), always written back (:776)
 *   alpha (n), sW (n) out: post.alpha, post.sW ; post.L through gpk_get_factor ; gpk_predict works afterwards
 *   dcov (nhyp) and dlz_out (n) if want_der: dnlZ.cov (:789-791) and d lZ / d mu of the cavity (:796) from which
 *   the caller forms dnlZ.mean = -dlZ'dm (:797-800).  lik.Erf has no hyper-parameters.                          */
int gpk_ep_eval(gpk_handle h, int kind, int matern_d, const double* hyp, int nhyp, const double* mvec,
                const double* y, double* ttau_io, double* tnu_io, int use_last, int want_der, double* nlZ,
                double* alpha, double* sW, double* dcov, double* dlz_out, int* sweeps_out);

/* ---- one evaluation sharded over the GPUs of a box (BASELINE config 3) ---- *
 * One process per GPU.  The matrix K/sn2+I is distributed by block columns (block-cyclic, block 128); each step
 * broadcasts the solved panel over NVLink with NCCL (loaded at run time from `nccl_path`, e.g. the libnccl.so.2
 * PyTorch ships); trailing updates, the K build and the substitutions run where the columns live.
 *   gpk_dist_unique_id: rank 0 obtains the 128-byte NCCL id; the caller distributes it (torch.distributed).
 *   gpk_dist_init:      collective; binds the handle to (rank, world).  world == 1 needs no NCCL.
 *   gpk_exact_eval_dist: collective counterpart of gpk_exact_eval without derivatives; same X (gpk_set_data),
 *                       hyper-parameters and y-m on every rank; every rank receives nlZ and the full alpha.    */
int gpk_dist_unique_id(const char* nccl_path, char* out128);
int gpk_dist_init(gpk_handle h, const char* nccl_path, int rank, int world, const char* id128);
int gpk_dist_finalize(gpk_handle h);
/* Allocate everything the collective calls below need for the current data (level 0: gpk_exact_eval_dist; 1: + the
 * factor gather; 2: + the sharded derivatives).  Not collective.  When the ranks are threads of ONE process, call it on
 * every rank and join before the collective call: a device allocation inside a collective phase can dead-lock against
 * a peer's enqueued NCCL kernel (peer-mapped allocations synchronise the devices).  One process per GPU: optional. */
int gpk_dist_reserve(gpk_handle h, int level);
int gpk_exact_eval_dist(gpk_handle h, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                        const double* ymm, double* nlZ, double* alpha);
/* Sharded GP.predict (Core/gp.py:404-419) and post.L: collective; replicates the distributed factor on every rank
 * (one packed NCCL broadcast per block column, 4 N^2 bytes in total).  Afterwards gpk_predict / gpk_get_factor work
 * on every rank WITHOUT further communication - the caller gives every rank its own share of the test points.   */
int gpk_dist_gather_factor(gpk_handle h);
/* Sharded inf.Exact.evaluate with derivatives (Core/inf.py:371-382; what optimize() calls): collective; as
 * gpk_exact_eval_dist, then every rank forms its rows of (K/sn2+I)^-1 by two triangular sweeps against its replica
 * of the factor and reduces Q o dK over them; one all-reduce of nhyp+1 scalars.  Outputs as gpk_exact_eval.       */
int gpk_exact_eval_dist_der(gpk_handle h, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                            const double* ymm, double* nlZ, double* alpha, double* dcov, double* dlik);

/* ---- measurement helpers (bench.py / tests only) ------------------------ */
/* fp64 tensor-pipe micro-benchmark: shape 0=m8n8k4 1=m16n8k4 2=m16n8k8
 * 3=m16n8k16 4=DFMA (no tensor pipe); returns TFLOP/s over `iters` inner loops. */
int gpk_bench_dmma(gpk_handle h, int shape, int warps_per_cta, int iters, double* tflops, double* ms);
/* device-resident DGEMM-NT (C -= A*B' lower / C = A*B') on random data; returns ms
 * per launch and TFLOP/s of the hot kernel in isolation.                    */
int gpk_bench_syrk(gpk_handle h, int64_t n, int k, int reps, double* ms, double* tflops);
/* stream copy bandwidth GB/s (read+write bytes), for a same-box HBM denominator */
int gpk_bench_copy(gpk_handle h, int64_t bytes, int reps, double* gbs);
/* debug: C(M,N) = beta*C + alpha*A(M,K)*B(N,K)' with column-major host arrays,
 * through the production tile kernel.  mode 0: C=A*B' ; 1: C-=A*B' full;
 * 2: C-=A*B' lower tiles only (M==N); 3: C=A*B' lower tiles, contraction from
 * k = 128*tile_row (upper-triangular operands); 4: A <- A*B' in place (N==K), the
 * panel-TRSM form.  M,N multiples of 128, K of 32.
 * Modes 5-7 run the chain products (sixteen 32x32-block CTAs per 128x128 tile; M = N = 128, K a multiple of 128):
 * 5: C = A*B' ; 6: C -= A*B' on the blocks on / below the diagonal ; 7 (K = 128): the head pair of the panel chain,
 * X = A*B' , C -= X*X' (lower), and A is OVERWRITTEN with X.                                                   */
int gpk_dbg_gemm_nt(gpk_handle h, int mode, int64_t M, int64_t N, int64_t K,
                    const double* A, const double* B, double* C);
/* debug: factor + invert one 128x128 block (column-major): L and inv(L). */
int gpk_dbg_diag(gpk_handle h, const double* A128, double* L128, double* Linv128,
                 double* logdet_half, int* info);

/* debug: C(128,N) int32 = A(128,K) int8 * B(N,K)' int8 (row-major host arrays) through one tcgen05.mma.kind::i8 tile
 * (TMEM accumulator, shared-memory descriptors) - the building block of the planned int8 emulation of the fp64
 * trailing update.  N in {64,128,256}, K a multiple of 32, K <= 256.  fmt bit 0 / bit 1: the bytes of A / B are
 * unsigned (u8) instead of signed (s8); bit 2 / bit 3: A goes shared memory -> TMEM (tcgen05.cp) and the MMA reads
 * it from there (bit 3: every k-step reuses the same TMEM columns).                                        */
int gpk_dbg_i8_tile(gpk_handle h, int N, int K, const int8_t* A, const int8_t* B, int32_t* C, int fmt);

/* debug/bench: C (n,n column-major, lower triangle) -= P (n,kw column-major) * P' - the trailing update of the
 * blocked Cholesky (reference: the BLAS-3 inside np.linalg.cholesky, tools.py:86).  mode 0: int8 tensor-core
 * (Ozaki split, tcgen05 + TMEM) path, mode 1: fp64 DMMA path.  n, kw multiples of 128.  *ms = mean device time
 * of the update over `reps` repetitions (each starts from the given C).                                          */
/* bench: issue rate of tcgen05.mma.kind::i8 (128 x N x 32) for a given shared-memory descriptor geometry
 * (leading/stride byte offsets, per-MMA operand step), max clocks per MMA over `ctas` concurrent CTAs; *tops =
 * int8 tensor throughput (2*128*N*32 ops per MMA) of the whole launch by CUDA events, in TOP/s.             */
int gpk_bench_i8_rate(gpk_handle h, int N, int iters, int lbo, int sbo, int astep, int same_acc, int ctas,
                      double* clk_per_mma, double* tops);

int gpk_dbg_oz_syrk(gpk_handle h, int64_t n, int kw, const double* P, double* C, int mode, int reps, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* GPK_H_ */
