// api.cu - the extern "C" surface of libgpk.so (include/gpk.h) and the host-side
// orchestration of the blocked right-looking Cholesky with one-panel look-ahead.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>
#include "gpk_internal.cuh"

namespace gpk {

static const double LOG2PI = 1.8378770664093454835606594728112;

int ensure(Handle* h, double** p, int64_t* cap, int64_t need) {
  if (*cap >= need && *p) return 0;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  GPK_CK(h, cudaMalloc((void**)p, (size_t)need * sizeof(double)));
  *cap = need;
  return 0;
}

// ensure() for the buffers that hold inverses of diagonal blocks: a fresh allocation is zero-filled, because the
// factorisation's diagonal kernel does not store the zero blocks above the block diagonal of an inverse (GPK_DIAG_SKIPZ)
// while every consumer multiplies by the whole 128x128 block.
int ensure_zero(Handle* h, double** p, int64_t* cap, int64_t need) {
  if (*cap >= need && *p) return 0;
  GPK_TRY(ensure(h, p, cap, need));
  GPK_CK(h, cudaMemset(*p, 0, (size_t)need * sizeof(double)));
  GPK_CK(h, cudaStreamSynchronize(0));
  return 0;
}

static int ensure_events(Handle* h, size_t count) {
  while (h->ev.size() < count) {
    cudaEvent_t e;
    // GPK_CHAIN_DUMP (diagnostic): the dependency events carry time stamps so that potrf_device can print the
    // per-panel timeline of the dependent chain
    GPK_CK(h, cudaEventCreateWithFlags(&e, getenv("GPK_CHAIN_DUMP") ? cudaEventDefault : cudaEventDisableTiming));
    h->ev.push_back(e);
  }
  return 0;
}
static int ensure_prof_events(Handle* h, size_t count) {
  while (h->prof_ev.size() < count) {
    cudaEvent_t e;
    GPK_CK(h, cudaEventCreate(&e));
    h->prof_ev.push_back(e);
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Blocked right-looking Cholesky, in place on the lower triangle of A (np x np, column-major, np a multiple
// of NB), with up to three levels of blocking:
//   level 1: a block of W1 panels; when it is done the whole trailing matrix gets ONE update with contraction
//            length W1*128.  On the int8 tensor-core path (ozaki.cu) the cost of an update tile's epilogue (TMEM ->
//            fp64 -> C) is fixed, so long contractions are what makes it efficient: W1 = 9 while the trailing
//            matrix is large; once the factorization is bound by the panel chain the blocks are ONE sub-block of
//            W2B = 4 panels.
//   level 2: sub-blocks of W2 panels inside the block; after each, the REST OF THE BLOCK'S columns are updated
//            (contraction W2*128), on the panel stream.
//   level 3: single panels of 128 columns: diag -> trsm -> rank-128 update of the remaining columns of the
//            sub-block (fp64 DMMA).
// Streams:
//   s_panel (high priority): the dependent chain: diag(p), then the HEAD of panel p (tile (p+1,p) solved, tile
//                            (p+1,p+1) updated: all diag(p+1) needs), the level-2 updates, and - for the small blocks -
//                            the level-1 update of the next block's first diagonal tile
//   s_tail                 : the TAIL of panel p, one panel behind: TRSM of the rows below, rest of the rank-128 update
//   s_main                 : level-1 update: first the next block's columns (so its panels can start), then
//                            the rest of the trailing matrix; with lazy_cov also the generation of the matrix itself
//                            (first block's columns, then the rest under the first panels)
//   s_aux                  : single-right-hand-side forward substitution step p, off the critical path
// Level-1 block j+1 therefore overlaps the bulk of trailing update j (look-ahead 1).
// ---------------------------------------------------------------------------
static int g_potrf_w = 0;   // 0: choose from the problem size (measured on B200: 3 panels at T>64, else 2)


int potrf_device(Handle* h, double* A, int64_t np, double* Dinv, double* logdet_parts, int* info, double* b_fwd,
                 double* z_out, const CovArgs* lazy_cov) {
  const int T = (int)(np / NB);
  const int64_t lda = np;
  if (const char* e = getenv("GPK_POTRF_W")) { g_potrf_w = atoi(e); }
  const int W2 = (g_potrf_w < 1) ? ((T > 64) ? 3 : (T > 8 ? 2 : 1)) : (g_potrf_w > 8 ? 8 : g_potrf_w);
  // trailing updates with at least oz_min tile rows go to the int8 tensor cores (GPK_OZAKI=0 keeps everything on DMMA)
  const int oz = env_int("GPK_OZAKI", 1), oz_min = env_int("GPK_OZAKI_MIN", 16);
  int W1 = env_int("GPK_POTRF_W1", 9);
  const int w1_minrem = env_int("GPK_POTRF_W1_MINREM", 48);
  if (W1 > 16) W1 = 16;
  W1 = (W1 / W2) * W2;                                  // a whole number of sub-blocks
  // plan the level-1 blocks
  std::vector<int> bstart, bsub, bbig;                       // block starts; sub-block width of each block
  // small blocks (the panel-chain-bound tail of the factorisation) are ONE sub-block of W2B panels: fewer
  // hand-overs between the panel chain and the trailing update per panel
  int W2B = env_int("GPK_POTRF_W2B", T > 64 ? 4 : W2);   // measured at N=16384: 4 (3: +0.15 ms, 6: +0.3 ms)
  if (W2B < 1) W2B = 1;
  if (W2B > 8) W2B = 8;
  int64_t need0 = 0, need1 = 0; int rows0 = 0, rows1 = 0, kw0 = 0;
  // optional short first block (GPK_POTRF_FIRST panels): nothing overlaps the first block's panel chain
  const int wfirst = env_int("GPK_POTRF_FIRST", 0);   // 0 = off (measured: a short first block costs more in the low-K update than it saves)
  for (int pos = 0; pos < T;) {
    const bool big = oz && W1 > W2 && T - (pos + W1) >= w1_minrem;
    int w = big ? W1 : W2B;
    if (pos == 0 && big && wfirst >= 1 && wfirst < w) w = ((wfirst + W2 - 1) / W2) * W2;
    bstart.push_back(pos);
    bsub.push_back(big ? W2 : w);
    bbig.push_back(big ? 1 : 0);
    const int pe = (pos + w < T) ? pos + w : T;
    if (oz && T - pe >= oz_min) {
      const int64_t nd = (int64_t)(T - pe) * NB * (pe - pos) * NB;
      if (nd > need0) need0 = nd;
      if ((T - pe) * NB > rows0) rows0 = (T - pe) * NB;   // row scales: the tallest sliced panel block
    }
    if (big && T - (pos + W2) >= oz_min) { need1 = 1; if ((T - pos - W2) * NB > rows1) rows1 = (T - pos - W2) * NB; }
    pos = pe;
  }
  bstart.push_back(T);
  const int nblk = (int)bstart.size() - 1;
  // Fixed row scales (GPK_OZ_FIXED, default on): |L_ij| <= sqrt(A_ii), so ONE scale per matrix row is valid for every
  // panel.  Each sub-block is then sliced once, right when it is complete, on the panel stream, into a buffer indexed by
  // GLOBAL row and by the k-position inside the level-1 block; the level-2 update of the block's other columns and the
  // level-1 update of the trailing matrix read the same digits.  This removes the row-maximum passes, the second read of
  // every panel and - the point - the slicing of the whole block from the critical stream of the level-1 hand-over.
  // Two buffers alternate between consecutive blocks (block j+1's sub-blocks are sliced while update j still runs).
  const bool fixed = oz && (need0 || need1) && env_int("GPK_OZ_FIXED", 1) != 0;
  if (fixed) {
    int wmax = 0;
    for (size_t b = 0; b + 1 < bstart.size(); ++b) wmax = std::max(wmax, bstart[b + 1] - bstart[b]);
    GPK_TRY(oz_ensure(h, 0, np, wmax * NB));
    GPK_TRY(oz_ensure(h, 1, np, wmax * NB));
    if (h->ozFixCap < (size_t)np) {
      if (h->ozFix) cudaFree(h->ozFix);
      h->ozFix = nullptr; h->ozFixCap = 0;
      GPK_CK(h, cudaMalloc((void**)&h->ozFix, (size_t)np * sizeof(double)));
      h->ozFixCap = (size_t)np;
    }
  } else {
    if (need0) {
      kw0 = (int)((need0 + rows0 - 1) / rows0);             // capacity rows0 x kw0 covers every block's slices
      GPK_TRY(oz_ensure(h, 0, rows0, kw0));
    }
    if (need1) GPK_TRY(oz_ensure(h, 1, rows1, W2 * NB));
  }
  const size_t rowbytes = (size_t)oz_slices() * 32;        // bytes per buffer row per k-step
  auto fx_sl = [&](int j, int row_tile, int kstep0) {        // slice-buffer address of (row tile, k-step) of block j
    return h->ozSl[j & 1] + ((size_t)kstep0 * np + (size_t)row_tile * NB) * rowbytes;
  };
  const bool chain_dump = getenv("GPK_CHAIN_DUMP") != nullptr;
  GPK_TRY(ensure_events(h, (chain_dump ? 6 : 5) * (size_t)T + 5));
  cudaEvent_t* ev_dstart = chain_dump ? h->ev.data() + 5 * T + 5 : nullptr;   // [T] the panel stream reaches diag(p)
  std::vector<char> was_split(T, 0);
  if (h->profile) GPK_TRY(ensure_prof_events(h, 2 * (size_t)T + 2));
  cudaEvent_t* ev_panel = h->ev.data();      // [T]   panel p factored and solved
  cudaEvent_t* ev_col = h->ev.data() + T;    // [T]   columns of level-1 block j up to date
  cudaEvent_t ev_fork = h->ev[2 * T], ev_join = h->ev[2 * T + 1], ev_aux = h->ev[2 * T + 2], ev_fix = h->ev[2 * T + 3];
  bool fix_pending = false;                   // the panel stream still has to wait for the fixed scales (ev_fix)
  cudaEvent_t ev_built = h->ev[5 * T + 4];    // lazy_cov: the WHOLE matrix exists (the columns behind the first block too)
  bool built_pending = false;
  cudaEvent_t* ev_diag = h->ev.data() + 2 * T + 4;   // [T] diagonal block p factored and inverted   (split chain)
  cudaEvent_t* ev_head = ev_diag + T;                // [T] head of panel p: tile (p+1,p) solved, tile (p+1,p+1) updated
  cudaEvent_t* ev_tail = ev_head + T;                // [T] tail of panel p: rows p+2.. solved, rest of the sub-block updated
  // Split panel chain (GPK_POTRF_SPLIT, default on).  The next diagonal block only needs tile (p+1,p) solved and tile
  // (p+1,p+1) updated, so those two small products (the "head", 4 CTAs each) run on the panel stream right behind the
  // diagonal kernel, while the TRSM of the rows below and the rest of the rank-128 update (the "tail") run on s_tail,
  // one panel behind.  The dependent chain per panel shrinks from diag + full TRSM + full inner update to
  // diag + two 32-row-strip products.
  const int split = env_int("GPK_POTRF_SPLIT", 1);

  h->stats.syrk_flops = 0.0;
  h->prof_pairs = 0;

  if (lazy_cov) {
    // The matrix is still to be generated (cov_kernel writes K/sn2 + I straight into A): build the columns of the
    // first level-1 block, let the panel stream start on them, and build the rest underneath the first panels.
    const int64_t c1 = (int64_t)bstart[1] * NB;
    CovArgs c = *lazy_cov;
    c.pS = c1;
    // stationary kernels: A_ii = sf2*scale + diag_add for every i - the fixed scales need no look at the matrix
    const bool const_diag = fixed && !lazy_cov->prog;
    if (const_diag)
      GPK_TRY(launch_oz_fixed_scales(h, h->s_main, nullptr, 0, (int)np, lazy_cov->sf2 * lazy_cov->scale + lazy_cov->diag_add,
                                     h->ozFix));
    GPK_TRY(launch_cov(h, h->s_main, c));
    GPK_CK(h, cudaEventRecord(ev_fork, h->s_main));
    if (c1 < np) {
      CovArgs r = *lazy_cov;
      r.out = lazy_cov->out + c1 * lazy_cov->ld;
      r.s_bstride = 1; r.s_boff = (int)(c1 / NB);
      r.pS = np - c1;
      GPK_TRY(launch_cov(h, h->s_main, r));
    }
    GPK_CK(h, cudaEventRecord(h->t1, h->s_main));
    if (c1 < np) {
      // The panel stream works inside the first block's columns until the block is done; the first thing it touches
      // outside them is the head update of the next diagonal tile at the hand-over - it must not run before the builder
      // has written that tile (a slow builder - a long composite program - used to lose this race).
      GPK_CK(h, cudaEventRecord(ev_built, h->s_main));
      built_pending = true;
    }
    if (fixed && !const_diag) {                 // composite kernels: the diagonal is read once the whole matrix exists
      GPK_TRY(launch_oz_fixed_scales(h, h->s_main, A, lda, (int)np, 0.0, h->ozFix));
      GPK_CK(h, cudaEventRecord(ev_fix, h->s_main));
      fix_pending = true;
    }
  } else {
    if (fixed) GPK_TRY(launch_oz_fixed_scales(h, h->s_main, A, lda, (int)np, 0.0, h->ozFix));
    GPK_CK(h, cudaEventRecord(ev_fork, h->s_main));
  }
  GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_fork, 0));
  if (split) GPK_CK(h, cudaStreamWaitEvent(h->s_tail, ev_fork, 0));
  if (b_fwd) GPK_CK(h, cudaStreamWaitEvent(h->s_aux, ev_fork, 0));
  const int head_l1_on = env_int("GPK_POTRF_HEADL1", 1);
  // the two products behind every diagonal block through small_nt_kernel.  0: the pipelined strip kernel of gemm_nt.cu;
  // 1 / 2: small_nt_kernel in the chain-bound blocks / everywhere, both operands in shared memory; 3 / 4: the same with
  // the B operand in registers (37 KB of shared memory: the CTAs fit beside a resident trailing-update CTA).
  // Measured at N=16384 on one B200 (same box, best of 8): 0: 20.45-20.56 ms, 1: 19.75-20.12, 3: 19.78, 4: 19.78.
  const int headk = env_int("GPK_POTRF_HEADK", 3);
  bool head_l1 = false;                                 // the last hand-over updated the next diagonal tile itself
  for (int j = 0; j < nblk; ++j) {
    const int pb = bstart[j], pe = bstart[j + 1];      // the level-1 block: panels [pb, pe)
    // block j's columns are up to date after ev_col[j]; when the hand-over already updated the first diagonal tile on
    // the panel stream (head_l1), only the kernels after the first diagonal block have to wait for it
    bool col_pending = false;
    if (j > 0) {
      if (head_l1) col_pending = true;
      else GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_col[j], 0));
    }
    const int w2 = bsub[j];
    // the products on the dependent chain through small_nt_kernel: 1 = in the chain-bound (small) blocks, 2 = everywhere
    // (3 / 4: as 1 / 2 with the B operand in registers, see small_nt_kernel)
    const bool small_heads = headk == 2 || headk == 4 || ((headk == 1 || headk == 3) && !bbig[j]);
    for (int sb = pb; sb < pe; sb += w2) {
      const int se = (sb + w2 < pe) ? sb + w2 : pe;    // the level-2 sub-block: panels [sb, se)
      for (int p = sb; p < se; ++p) {
        double* App = A + (int64_t)p * NB * (1 + lda);
        double* Dp = Dinv + (int64_t)p * NB * NB;
        const int rem = T - p - 1;                     // tile rows below panel p
        const int inner = se - p - 1;                  // remaining columns of this sub-block
        if (split && p >= sb + 2) GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_tail[p - 2], 0));
        if (chain_dump) GPK_CK(h, cudaEventRecord(ev_dstart[p], h->s_panel));
        GPK_TRY(launch_diag(h, h->s_panel, App, lda, Dp, logdet_parts + p, info, p * NB));
        if (col_pending) {
          GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_col[j], 0));
          GPK_CK(h, cudaStreamWaitEvent(h->s_tail, ev_col[j], 0));
          col_pending = false;
        }
        if (!split || inner == 0 || rem < 2) {
          // whole panel on the panel stream: TRSM of all rows below, then the rank-128 update of the sub-block
          if (split && p > sb) GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_tail[p - 1], 0));
          if (rem > 0) {
            GemmArgs t{};
            t.A = App + NB; t.B = Dp; t.C = App + NB;
            t.lda = lda; t.ldb = NB; t.ldc = lda; t.K = NB; t.tri = 0;
            GPK_TRY(launch_gemm_nt(h, h->s_panel, 0, t, rem, 1));
          }
          GPK_CK(h, cudaEventRecord(ev_panel[p], h->s_panel));
          if (inner > 0) {
            GemmArgs u{};
            u.A = App + NB; u.B = App + NB; u.C = A + (int64_t)(p + 1) * NB * (1 + lda);
            u.lda = lda; u.ldb = lda; u.ldc = lda; u.K = NB; u.tri = 1;
            GPK_TRY(launch_gemm_nt(h, h->s_panel, 1, u, rem, inner));
          }
          if (split) GPK_CK(h, cudaEventRecord(ev_tail[p], h->s_panel));   // nothing of panel p is left for s_tail
        } else {
          GPK_CK(h, cudaEventRecord(ev_diag[p], h->s_panel));
          was_split[p] = 1;
          // head (panel stream): tile (p+1,p) <- tile * Dinv_p', then tile (p+1,p+1) -= L(p+1,p) L(p+1,p)'
          if (p > sb) GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_tail[p - 1], 0));
          GemmArgs t{};
          t.A = App + NB; t.B = Dp; t.C = App + NB;
          t.lda = lda; t.ldb = NB; t.ldc = lda; t.K = NB; t.tri = 0;
          if (small_heads) {
            // sixteen 32x32-block CTAs per product, whole operands requested at once (small_nt_kernel): the solved
            // tile goes to the scratch tile, the update reads it from there and copies it home
            SmallArgs s1{};
            s1.A = App + NB; s1.lda = lda; s1.B = Dp; s1.ldb = NB; s1.C = h->dHead; s1.ldc = NB; s1.K = NB; s1.mode = 0;
            s1.breg = headk >= 3;
            GPK_TRY(launch_small_nt(h, h->s_panel, s1));
            SmallArgs s2{};
            s2.A = h->dHead; s2.lda = NB; s2.B = h->dHead; s2.ldb = NB; s2.C = A + (int64_t)(p + 1) * NB * (1 + lda);
            s2.ldc = lda; s2.K = NB; s2.mode = 1; s2.tri = 1; s2.copy_dst = App + NB; s2.ld_copy = lda;
            s2.breg = headk >= 3;
            GPK_TRY(launch_small_nt(h, h->s_panel, s2));
          } else {
            GPK_TRY(launch_gemm_nt(h, h->s_panel, 0, t, 1, 1));
            GemmArgs hs{};
            hs.A = App + NB; hs.B = App + NB; hs.C = A + (int64_t)(p + 1) * NB * (1 + lda);
            hs.lda = lda; hs.ldb = lda; hs.ldc = lda; hs.K = NB; hs.tri = 1; hs.strips = 1;
            GPK_TRY(launch_gemm_nt(h, h->s_panel, 1, hs, 1, 1));
          }
          GPK_CK(h, cudaEventRecord(ev_head[p], h->s_panel));
          // tail (s_tail): TRSM of rows p+2.., then column p+1 below its diagonal tile, then columns p+2..se-1
          GPK_CK(h, cudaStreamWaitEvent(h->s_tail, ev_diag[p], 0));
          GemmArgs tt = t;
          tt.A = App + 2 * NB; tt.C = App + 2 * NB;
          GPK_TRY(launch_gemm_nt(h, h->s_tail, 0, tt, rem - 1, 1));
          GPK_CK(h, cudaStreamWaitEvent(h->s_tail, ev_head[p], 0));
          GPK_CK(h, cudaEventRecord(ev_panel[p], h->s_tail));
          GemmArgs u1{};
          u1.A = App + 2 * NB; u1.B = App + NB; u1.C = A + (int64_t)(p + 2) * NB + (int64_t)(p + 1) * NB * lda;
          u1.lda = lda; u1.ldb = lda; u1.ldc = lda; u1.K = NB; u1.tri = 0;
          GPK_TRY(launch_gemm_nt(h, h->s_tail, 1, u1, rem - 1, 1));
          if (inner > 1) {
            GemmArgs u2{};
            u2.A = App + 2 * NB; u2.B = App + 2 * NB; u2.C = A + (int64_t)(p + 2) * NB * (1 + lda);
            u2.lda = lda; u2.ldb = lda; u2.ldc = lda; u2.K = NB; u2.tri = 1;
            GPK_TRY(launch_gemm_nt(h, h->s_tail, 1, u2, rem - 1, inner - 1));
          }
          GPK_CK(h, cudaEventRecord(ev_tail[p], h->s_tail));
        }
        if (b_fwd) {
          GPK_CK(h, cudaStreamWaitEvent(h->s_aux, ev_panel[p], 0));
          GPK_TRY(launch_trsv_fwd(h, h->s_aux, A, lda, Dinv, b_fwd, z_out, p, T));
        }
      }
      // the sub-block is complete when its last panel's tail is (the panel stream runs the level-2 / level-1 hand-over)
      if (split && se - 1 >= sb) GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_tail[se - 1], 0));
      // fixed scales: this sub-block's digits, once, for the level-2 update below AND the level-1 update of the block
      const bool fx_blk = fixed && bbig[j] && T - pe >= oz_min;
      if (fx_blk && T - se > 0) {
        if (fix_pending) { GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_fix, 0)); fix_pending = false; }
        GPK_TRY(launch_oz_slice_fixed(h, h->s_panel, A + (int64_t)se * NB + (int64_t)sb * NB * lda, lda, (T - se) * NB,
                                      (se - sb) * NB, h->ozSl[j & 1], h->ozFix, (int)np, se * NB, (sb - pb) * 4));
      }
      if (se < pe) {
        // level-2 update: the block's remaining columns [se, pe), all rows below the sub-block
        const int rows = T - se, cols = pe - se, kw2 = (se - sb) * NB;
        double* P2 = A + (int64_t)se * NB + (int64_t)sb * NB * lda;
        double* C2 = A + (int64_t)se * NB * (1 + lda);
        if (fx_blk) {
          GPK_TRY(launch_oz_syrk_buf(h, h->s_panel, fx_sl(j, se, (sb - pb) * 4), h->ozFix + (int64_t)se * NB, (int)np, C2,
                                     lda, rows * NB, kw2, 0, cols, 0));
        } else if (oz && rows >= oz_min) {
          GPK_TRY(launch_oz_slice(h, 1, h->s_panel, P2, lda, rows * NB, kw2));
          GPK_TRY(launch_oz_syrk(h, 1, h->s_panel, C2, lda, rows * NB, kw2, 0, cols));
        } else {
          GemmArgs u{};
          u.A = P2; u.B = P2; u.C = C2;
          u.lda = lda; u.ldb = lda; u.ldc = lda; u.K = kw2; u.tri = 1;
          GPK_TRY(launch_gemm_nt(h, h->s_panel, 1, u, rows, cols));
        }
      }
    }
    const int rem = T - pe;                            // trailing tiles after the block
    if (rem > 0) {
      GPK_CK(h, cudaEventRecord(ev_join, h->s_panel));
      GPK_CK(h, cudaStreamWaitEvent(h->s_main, ev_join, 0));
      if (h->profile) GPK_CK(h, cudaEventRecord(h->prof_ev[2 * h->prof_pairs], h->s_main));
      const int kw = (pe - pb) * NB;
      const int nextw = bstart[j + 2 <= nblk ? j + 2 : nblk] - pe;
      const int first = (nextw < rem) ? nextw : rem;   // the next block's columns go first
      double* Pblk = A + (int64_t)pe * NB + (int64_t)pb * NB * lda;
      double* Ctr = A + (int64_t)pe * NB * (1 + lda);
      // Panel-chain-bound blocks: the next block's first diagonal tile gets its update on the PANEL stream (four
      // 32-row strips), so the next diagonal kernel starts at once and overlaps the update of the other tiles.
      head_l1 = split && head_l1_on && !bbig[j];
      if (head_l1 && built_pending) { GPK_CK(h, cudaStreamWaitEvent(h->s_panel, ev_built, 0)); built_pending = false; }
      if (head_l1 && small_heads) {
        SmallArgs hl{};
        hl.A = Pblk; hl.lda = lda; hl.B = Pblk; hl.ldb = lda; hl.C = Ctr; hl.ldc = lda; hl.K = kw; hl.mode = 1; hl.tri = 1;
        GPK_TRY(launch_small_nt(h, h->s_panel, hl));
      } else if (head_l1) {
        GemmArgs hl{};
        hl.A = Pblk; hl.B = Pblk; hl.C = Ctr;
        hl.lda = lda; hl.ldb = lda; hl.ldc = lda; hl.K = kw; hl.tri = 1; hl.strips = 1;
        GPK_TRY(launch_gemm_nt(h, h->s_panel, 1, hl, 1, 1));
      }
      if (fixed && rem >= oz_min) {
        // int8 tensor cores, fixed row scales.  Big blocks: every sub-block was sliced on the panel stream when it was
        // complete - nothing to do here but the update.  Small (single sub-block) blocks: slice here, without a
        // row-maximum pass.
        if (!bbig[j]) {
          if (fix_pending) { GPK_CK(h, cudaStreamWaitEvent(h->s_main, ev_fix, 0)); }
          GPK_TRY(launch_oz_slice_fixed(h, h->s_main, Pblk, lda, rem * NB, kw, h->ozSl[j & 1], h->ozFix, (int)np, pe * NB, 0));
        }
        const int8_t* sl1 = fx_sl(j, pe, 0);
        const double* sc1 = h->ozFix + (int64_t)pe * NB;
        GPK_TRY(launch_oz_syrk_buf(h, h->s_main, sl1, sc1, (int)np, Ctr, lda, rem * NB, kw, 0, first, head_l1 ? 1 : 0));
        GPK_CK(h, cudaEventRecord(ev_col[j + 1], h->s_main));
        if (rem > first) GPK_TRY(launch_oz_syrk_buf(h, h->s_main, sl1, sc1, (int)np, Ctr, lda, rem * NB, kw, first, rem, 0));
      } else if (oz && rem >= oz_min) {
        // int8 tensor-core path (ozaki.cu): slice the panel block once, then the same two launches
        GPK_TRY(launch_oz_slice(h, 0, h->s_main, Pblk, lda, rem * NB, kw));
        GPK_TRY(launch_oz_syrk(h, 0, h->s_main, Ctr, lda, rem * NB, kw, 0, first, head_l1 ? 1 : 0));
        GPK_CK(h, cudaEventRecord(ev_col[j + 1], h->s_main));
        if (rem > first) GPK_TRY(launch_oz_syrk(h, 0, h->s_main, Ctr, lda, rem * NB, kw, first, rem));
      } else {
        GemmArgs u{};
        u.A = Pblk; u.B = u.A;
        u.C = Ctr;
        u.lda = lda; u.ldb = lda; u.ldc = lda; u.K = kw; u.tri = 1; u.ti_off = 0; u.tj_off = 0;
        if (!head_l1) {
          GPK_TRY(launch_gemm_nt(h, h->s_main, 1, u, rem, first));
        } else if (rem > 1) {
          // everything but tile (0,0): column 0 below its diagonal tile, then columns 1..first-1 (lower triangle)
          GemmArgs c0 = u;
          c0.A = Pblk + NB; c0.C = Ctr + NB; c0.tri = 0;
          GPK_TRY(launch_gemm_nt(h, h->s_main, 1, c0, rem - 1, 1));
          if (first > 1) {
            GemmArgs c1 = u;
            c1.A = Pblk + NB; c1.B = Pblk + NB; c1.C = Ctr + (int64_t)NB * (1 + lda);
            GPK_TRY(launch_gemm_nt(h, h->s_main, 1, c1, rem - 1, first - 1));
          }
        }
        GPK_CK(h, cudaEventRecord(ev_col[j + 1], h->s_main));
        if (rem > first) {
          GemmArgs v = u;
          v.B = u.B + (int64_t)first * NB; v.C = u.C + (int64_t)first * NB * lda; v.tj_off = first;
          GPK_TRY(launch_gemm_nt(h, h->s_main, 1, v, rem, rem - first));
        }
      }
      if (h->profile) { GPK_CK(h, cudaEventRecord(h->prof_ev[2 * h->prof_pairs + 1], h->s_main)); h->prof_pairs++; }
      const double nt = (double)rem * NB;
      h->stats.syrk_flops += (double)kw * nt * nt;
    }
  }
  GPK_CK(h, cudaEventRecord(ev_join, h->s_panel));
  GPK_CK(h, cudaStreamWaitEvent(h->s_main, ev_join, 0));
  if (b_fwd) {
    GPK_CK(h, cudaEventRecord(ev_aux, h->s_aux));
    GPK_CK(h, cudaStreamWaitEvent(h->s_main, ev_aux, 0));
  }
  if (chain_dump && h->profile) {
    // per-panel timeline in microseconds since the fork: when the panel stream reached diag(p), diag done, head done,
    // panel solved (ev_panel), tail done; per level-1 block: its columns up to date (ev_col), update start / end
    GPK_CK(h, cudaStreamSynchronize(h->s_main));
    auto us = [&](cudaEvent_t e) {
      float ms = 0.f;
      return cudaEventElapsedTime(&ms, ev_fork, e) == cudaSuccess ? (double)ms * 1e3 : -1.0;
    };
    (void)cudaGetLastError();
    int pairs = 0;
    for (int j = 0; j < nblk; ++j) {
      const int pe = bstart[j + 1];
      fprintf(stderr, "block %d panels [%d,%d) big=%d col_ready %.1f", j, bstart[j], pe, bbig[j], j > 0 ? us(ev_col[j]) : 0.0);
      if (T - pe > 0 && h->profile && 2 * pairs + 1 < (int)h->prof_ev.size()) {
        fprintf(stderr, " update [%.1f, %.1f]", us(h->prof_ev[2 * pairs]), us(h->prof_ev[2 * pairs + 1]));
        ++pairs;
      }
      fprintf(stderr, "\n");
      for (int p = bstart[j]; p < pe; ++p)
        fprintf(stderr, "  p %3d dstart %9.1f diag %9.1f head %9.1f panel %9.1f tail %9.1f\n", p, us(ev_dstart[p]),
                was_split[p] ? us(ev_diag[p]) : -1.0, was_split[p] ? us(ev_head[p]) : -1.0, us(ev_panel[p]),
                split ? us(ev_tail[p]) : -1.0);
    }
    (void)cudaGetLastError();
  }
  return 0;
}

static int collect_profile(Handle* h, int T) {
  h->stats.syrk_ms = 0.0;
  (void)T;
  if (!h->profile) return 0;
  const bool dump = getenv("GPK_PROFILE_DUMP") != nullptr;
  for (int k = 0; k < h->prof_pairs; ++k) {
    float ms = 0.f;
    GPK_CK(h, cudaEventElapsedTime(&ms, h->prof_ev[2 * k], h->prof_ev[2 * k + 1]));
    h->stats.syrk_ms += ms;
    if (dump) {
      float gap = 0.f;
      if (k + 1 < h->prof_pairs) cudaEventElapsedTime(&gap, h->prof_ev[2 * k + 1], h->prof_ev[2 * k + 2]);
      fprintf(stderr, "step %d update %.3f ms, wait-for-next-panel %.3f ms\n", k, ms, gap);
    }
  }
  return 0;
}

// per-kind input scaling; returns sf2
int kind_scale(int kind, int matern_d, const double* hyp, int nhyp, int D, std::vector<double>& scale,
                      int* divide, double* premul, double* sf2) {
  scale.assign(D, 1.0);
  *premul = 1.0;
  if (kind == GPK_COV_RBF) {
    if (nhyp != 2) return GPK_ERR_ARG;
    for (int d = 0; d < D; ++d) scale[d] = std::exp(hyp[0]);
    *divide = 1;
    *sf2 = std::exp(2.0 * hyp[1]);
  } else if (kind == GPK_COV_RBFARD) {
    if (nhyp != D + 1) return GPK_ERR_ARG;
    for (int d = 0; d < D; ++d) scale[d] = 1.0 / std::exp(hyp[d]);
    *divide = 0;
    *sf2 = std::exp(2.0 * hyp[D]);
  } else if (kind == GPK_COV_MATERN) {
    if (nhyp != 2) return GPK_ERR_ARG;
    if (!(matern_d == 1 || matern_d == 3 || matern_d == 5 || matern_d == 7)) return GPK_ERR_ARG;
    for (int d = 0; d < D; ++d) scale[d] = std::exp(hyp[0]);
    *divide = 1;
    *premul = std::sqrt((double)matern_d);
    *sf2 = std::exp(2.0 * hyp[1]);
  } else {
    return GPK_ERR_ARG;
  }
  return 0;
}

static int free_all(Handle* h) {
  double** ptrs[] = {&h->dX, &h->dXs, &h->dScale, &h->dA, &h->dDinv, &h->dB, &h->dZ, &h->dAlpha, &h->dR, &h->dScal,
                     &h->dU, &h->dW, &h->dP, &h->dTmp, &h->dUin, &h->dUs, &h->dLpost, &h->dAlphaU,
                     &h->dDinvT, &h->fKuu, &h->fDinvU, &h->fA2, &h->fDinv2, &h->fVt, &h->fVs, &h->fVec, &h->fWt, &h->dXtmp, &h->gA, &h->gDinv, &h->gPack, &h->gBlk, &h->gVec, &h->eK, &h->eSig, &h->eVec, &h->eSW};
  for (auto p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (h->dInfo) cudaFree(h->dInfo);
  h->dInfo = nullptr;
  if (h->dFlags) cudaFree(h->dFlags);
  h->dFlags = nullptr;
  if (h->epGraphExec) cudaGraphExecDestroy((cudaGraphExec_t)h->epGraphExec);
  h->epGraphExec = nullptr;
  if (h->ozFix) cudaFree(h->ozFix);
  h->ozFix = nullptr; h->ozFixCap = 0;
  if (h->dProg) cudaFree(h->dProg);
  h->dProg = nullptr;
  if (h->hProgPinned) cudaFreeHost(h->hProgPinned);
  h->hProgPinned = nullptr;
  if (h->dPre) cudaFree(h->dPre);
  h->dPre = nullptr; h->capPre = 0; h->preN = 0;
  for (int w = 0; w < 2; ++w) {
    if (h->ozSl[w]) cudaFree(h->ozSl[w]);
    if (h->ozSc[w]) cudaFree(h->ozSc[w]);
    h->ozSl[w] = nullptr; h->ozSc[w] = nullptr; h->ozCap[w] = h->ozScCap[w] = 0;
  }
  h->capA = h->capU = h->capW = h->capP = h->capTmp = h->capUin = h->capDinvT = 0;
  h->lt_valid = false;
  h->cKuu = h->cDinvU = h->cA2 = h->cDinv2 = h->cVt = h->cVs = h->cVec = h->cWt = h->capXtmp = h->cgA = h->cgDinv = h->cgPack = h->cgBlk = h->cgVec = h->ceK = h->ceSig = h->ceVec = h->ceSW = h->capUs = h->capLpost = h->capAlphaU = 0;
  return 0;
}

// (re)allocate everything that depends on the padded problem size
static int alloc_problem(Handle* h, int64_t n, int D) {
  const int64_t np = round_up(n, NB);
  const int T = (int)(np / NB);
  if (np != h->np || D != h->D || !h->dXs) {
    double** ptrs[] = {&h->dXs, &h->dScale, &h->dDinv, &h->dB, &h->dZ, &h->dAlpha, &h->dR, &h->dScal};
    for (auto p : ptrs) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    // the (np x np) factor storage is allocated by the entry points that need it (exact_eval, potrf): FITC and the
    // sharded evaluation never hold an N x N matrix on one GPU
    GPK_CK(h, cudaMalloc((void**)&h->dXs, (size_t)np * D * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dScale, (size_t)(D + 8) * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dDinv, (size_t)np * NB * sizeof(double)));
    GPK_CK(h, cudaMemset(h->dDinv, 0, (size_t)np * NB * sizeof(double)));     // see ensure_zero
    GPK_CK(h, cudaStreamSynchronize(0));
    GPK_CK(h, cudaMalloc((void**)&h->dB, (size_t)np * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dZ, (size_t)np * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dAlpha, (size_t)np * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dR, (size_t)np * sizeof(double)));
    GPK_CK(h, cudaMalloc((void**)&h->dScal, (size_t)(T + 4 * D + 256) * sizeof(double)));
    if (!h->dInfo) GPK_CK(h, cudaMalloc((void**)&h->dInfo, 4 * sizeof(int)));
  }
  h->n = n; h->np = np; h->D = D;
  return 0;
}

void stats_begin(Handle* h) {
  std::memset(&h->stats, 0, sizeof(h->stats));
  h->lt_valid = false;     // every entry point may overwrite the factor or the work buffers gpk_potrs caches things in
}

int check_handle(gpk_handle hh, Handle** out) {
  if (!hh) return GPK_ERR_ARG;
  Handle* h = reinterpret_cast<Handle*>(hh);
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) { h->last_cuda = e; h->last_msg = cudaGetErrorString(e); return GPK_ERR_CUDA; }
  *out = h;
  return 0;
}

// Solve with many right-hand sides held TRANSPOSED: P is (rows x np) column-major with pitch ldp
// (one right-hand side per ROW).  Forward: P <- P * L^-T, block column by block column.
int sweep_forward(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* A,
                         int64_t lda, const double* Dinv, int T, int kstart) {
  // kstart > 0: the first kstart block columns of P are known to be zero (unit right-hand sides e_i, i >= kstart*NB)
  for (int k = kstart; k < T; ++k) {
    GemmArgs t{};
    t.A = P + (int64_t)k * NB * ldp; t.B = Dinv + (int64_t)k * NB * NB; t.C = P + (int64_t)k * NB * ldp;
    t.lda = ldp; t.ldb = NB; t.ldc = ldp; t.K = NB; t.tri = 0;
    GPK_TRY(launch_gemm_nt(h, st, 0, t, row_tiles, 1));
    if (k + 1 < T) {
      GemmArgs u{};
      u.A = P + (int64_t)k * NB * ldp; u.B = A + (int64_t)(k + 1) * NB + (int64_t)k * NB * lda;
      u.C = P + (int64_t)(k + 1) * NB * ldp;
      u.lda = ldp; u.ldb = lda; u.ldc = ldp; u.K = NB; u.tri = 0;
      GPK_TRY(launch_gemm_nt(h, st, 1, u, row_tiles, T - k - 1));
    }
  }
  return 0;
}

// Backward: P <- P * L^-1 for right-hand sides held transposed (one per ROW of P), block column by block column from the
// last.  The products need L[k, 0:k] as the B operand of an NT product, i.e. rows of L' : Lt is the explicit transpose
// of the factor (upper triangular, pitch ldt) and DinvT the transposed block inverses, both made once per factor.
int sweep_backward(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* Lt, int64_t ldt,
                   const double* DinvT, int T) {
  for (int k = T - 1; k >= 0; --k) {
    GemmArgs t{};                                     // X_k = P_k * L_kk^-1 = P_k * (DinvT_k)'
    t.A = P + (int64_t)k * NB * ldp; t.B = DinvT + (int64_t)k * NB * NB; t.C = P + (int64_t)k * NB * ldp;
    t.lda = ldp; t.ldb = NB; t.ldc = ldp; t.K = NB; t.tri = 0;
    GPK_TRY(launch_gemm_nt(h, st, 0, t, row_tiles, 1));
    if (k > 0) {
      GemmArgs u{};                                   // P[:, 0:k] -= X_k * L[k, 0:k] ;  B(j, kk) = L[k*NB+kk, j] = Lt[j, k*NB+kk]
      u.A = P + (int64_t)k * NB * ldp; u.B = Lt + (int64_t)k * NB * ldt; u.C = P;
      u.lda = ldp; u.ldb = ldt; u.ldc = ldp; u.K = NB; u.tri = 0;
      GPK_TRY(launch_gemm_nt(h, st, 1, u, row_tiles, k));
    }
  }
  return 0;
}

// U <- L^-T (upper triangular, np x np column-major pitch np) by the same sweep applied to the identity,
// touching only the tiles on or above the diagonal.
int inverse_factor_T(Handle* h, cudaStream_t st, double* U, const double* A, int64_t np, const double* Dinv) {
  const int T = (int)(np / NB);
  GPK_TRY(launch_set_identity(h, st, U, np, np, np));
  for (int k = 0; k < T; ++k) {
    GemmArgs t{};
    t.A = U + (int64_t)k * NB * np; t.B = Dinv + (int64_t)k * NB * NB; t.C = U + (int64_t)k * NB * np;
    t.lda = np; t.ldb = NB; t.ldc = np; t.K = NB; t.tri = 0;
    GPK_TRY(launch_gemm_nt(h, st, 0, t, k + 1, 1));
    if (k + 1 < T) {
      GemmArgs u{};
      u.A = U + (int64_t)k * NB * np; u.B = A + (int64_t)(k + 1) * NB + (int64_t)k * NB * np;
      u.C = U + (int64_t)(k + 1) * NB * np;
      u.lda = np; u.ldb = np; u.ldc = np; u.K = NB; u.tri = 0;
      GPK_TRY(launch_gemm_nt(h, st, 1, u, k + 1, T - k - 1));
    }
  }
  return 0;
}

// C (na x nb, column-major ldc) = / += / -= A (na x K) * B (nb x K)' on the int8 tensor cores: both operands are sliced
// into one buffer as [B rows; A rows] per contraction chunk of <= 16384 (int32 accumulators), see launch_oz_gemm_stacked.
// cmode as in launch_oz_ex (0: C -= , 1: C = , 2: C +=).
int oz_gemm_nt(Handle* h, cudaStream_t st, double* C, int64_t ldc, const double* A, int64_t lda, int na, const double* B,
               int64_t ldb, int nb, int K, int cmode) {
  const int KC = 16384;
  if (na % NB || nb % NB || K % NB || na <= 0 || nb <= 0 || K <= 0) return GPK_ERR_ARG;
  GPK_TRY(oz_ensure(h, 0, (int64_t)na + nb, K < KC ? K : KC));
  for (int c0 = 0; c0 < K; c0 += KC) {
    const int kw = (K - c0 < KC) ? K - c0 : KC;
    GPK_TRY(launch_oz_slice(h, 0, st, B + (int64_t)c0 * ldb, ldb, nb, kw, 0, nb + na));
    GPK_TRY(launch_oz_slice(h, 0, st, A + (int64_t)c0 * lda, lda, na, kw, nb, nb + na));
    const int cm = (c0 == 0) ? cmode : (cmode == 0 ? 0 : 2);
    GPK_TRY(launch_oz_ex(h, 0, st, C - nb, ldc, nb + na, kw, 0, nb / NB, 0, nb / NB, 0, cm));
  }
  return 0;
}

// sweep_forward with the O(M^2 n) part on the int8 tensor cores: column blocks of WB panels are finished with the DMMA
// sweep restricted to the block, then ALL later columns get one update of contraction length WB*128,
//   P[:, ke:] -= P[:, kb:ke] * L[ke:, kb:ke]',   operands stacked as [L rows; P rows] (launch_oz_gemm_stacked).
int sweep_forward_oz(Handle* h, cudaStream_t st, double* P, int64_t ldp, int row_tiles, const double* A, int64_t lda,
                     const double* Dinv, int T) {
  // block width: the in-block DMMA share of the work is ~WB/T, the int8 update wants a long contraction
  int WB = env_int("GPK_SWEEP_WB", T <= 48 ? 4 : 8);
  if (WB < 1) WB = 1;
  if (WB > 16) WB = 16;
  const int na = row_tiles * NB;
  if (T <= WB) return sweep_forward(h, st, P, ldp, row_tiles, A, lda, Dinv, T);
  GPK_TRY(oz_ensure(h, 0, (int64_t)(T - WB) * NB + na, WB * NB));
  for (int kb = 0; kb < T; kb += WB) {
    const int ke = (kb + WB < T) ? kb + WB : T;
    for (int k = kb; k < ke; ++k) {
      GemmArgs t{};
      t.A = P + (int64_t)k * NB * ldp; t.B = Dinv + (int64_t)k * NB * NB; t.C = P + (int64_t)k * NB * ldp;
      t.lda = ldp; t.ldb = NB; t.ldc = ldp; t.K = NB; t.tri = 0;
      GPK_TRY(launch_gemm_nt(h, st, 0, t, row_tiles, 1));
      if (k + 1 < ke) {
        GemmArgs u{};
        u.A = P + (int64_t)k * NB * ldp; u.B = A + (int64_t)(k + 1) * NB + (int64_t)k * NB * lda;
        u.C = P + (int64_t)(k + 1) * NB * ldp;
        u.lda = ldp; u.ldb = lda; u.ldc = ldp; u.K = NB; u.tri = 0;
        GPK_TRY(launch_gemm_nt(h, st, 1, u, row_tiles, ke - k - 1));
      }
    }
    if (ke < T) {
      const int nb = (T - ke) * NB, kw = (ke - kb) * NB;
      GPK_TRY(launch_oz_slice(h, 0, st, A + (int64_t)ke * NB + (int64_t)kb * NB * lda, lda, nb, kw, 0, nb + na));
      GPK_TRY(launch_oz_slice(h, 0, st, P + (int64_t)kb * NB * ldp, ldp, na, kw, nb, nb + na));
      GPK_TRY(launch_oz_gemm_stacked(h, 0, st, P + (int64_t)ke * NB * ldp, ldp, nb, na, kw));
    }
  }
  return 0;
}

// The same U = L^-T with the O(N^3) part on the int8 tensor cores: column blocks of WB panels are finished with the DMMA
// sweep above (restricted to the block), then ALL later columns get one update of contraction length WB*128,
//   U[0:ke, ke:] -= U[0:ke, kb:ke] * L[ke:, kb:ke]',
// computed by the sliced SYRK kernel on the two operands stacked as [L rows; U rows] (launch_oz_gemm_stacked).
int inverse_factor_T_oz(Handle* h, cudaStream_t st, double* U, const double* A, int64_t np, const double* Dinv) {
  const int T = (int)(np / NB);
  const int WB = 8;
  GPK_TRY(oz_ensure(h, 0, np, WB * NB));
  GPK_TRY(launch_set_identity(h, st, U, np, np, np));
  for (int kb = 0; kb < T; kb += WB) {
    const int ke = (kb + WB < T) ? kb + WB : T;
    for (int k = kb; k < ke; ++k) {
      GemmArgs t{};
      t.A = U + (int64_t)k * NB * np; t.B = Dinv + (int64_t)k * NB * NB; t.C = U + (int64_t)k * NB * np;
      t.lda = np; t.ldb = NB; t.ldc = np; t.K = NB; t.tri = 0;
      GPK_TRY(launch_gemm_nt(h, st, 0, t, k + 1, 1));
      if (k + 1 < ke) {
        GemmArgs u{};
        u.A = U + (int64_t)k * NB * np; u.B = A + (int64_t)(k + 1) * NB + (int64_t)k * NB * np;
        u.C = U + (int64_t)(k + 1) * NB * np;
        u.lda = np; u.ldb = np; u.ldc = np; u.K = NB; u.tri = 0;
        GPK_TRY(launch_gemm_nt(h, st, 1, u, k + 1, ke - k - 1));
      }
    }
    if (ke < T) {
      const int nb = (T - ke) * NB, na = ke * NB, kw = (ke - kb) * NB;
      GPK_TRY(launch_oz_slice(h, 0, st, A + (int64_t)ke * NB + (int64_t)kb * NB * np, np, nb, kw, 0, nb + na));
      GPK_TRY(launch_oz_slice(h, 0, st, U + (int64_t)kb * NB * np, np, na, kw, nb, nb + na));
      GPK_TRY(launch_oz_gemm_stacked(h, 0, st, U + (int64_t)ke * NB * np, np, nb, na, kw));
    }
  }
  return 0;
}

}  // namespace gpk

using namespace gpk;

extern "C" {

int gpk_version(void) { return 100; }

const char* gpk_strerror(int code) {
  switch (code) {
    case GPK_OK: return "ok";
    case GPK_ERR_ARG: return "invalid argument";
    case GPK_ERR_CUDA: return "CUDA error (see gpk_last_error)";
    case GPK_ERR_STATE: return "invalid call order: no data / posterior on the handle";
    case GPK_ERR_NOMEM: return "device out of memory";
    default: return code > 0 ? "matrix not positive definite (info = failing pivot)" : "unknown error";
  }
}

int gpk_device_count(int* count) {
  if (!count) return GPK_ERR_ARG;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *count = 0; return GPK_ERR_CUDA; }
  *count = c;
  return 0;
}

int gpk_device_memory(int device, int64_t* free_bytes, int64_t* total_bytes) {
  if (!free_bytes || !total_bytes) return GPK_ERR_ARG;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return GPK_ERR_ARG;
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return GPK_ERR_CUDA;
  size_t f = 0, t = 0;
  const cudaError_t e = cudaMemGetInfo(&f, &t);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return GPK_ERR_CUDA;
  *free_bytes = (int64_t)f; *total_bytes = (int64_t)t;
  return 0;
}

int gpk_create(int device, gpk_handle* out) {
  if (!out) return GPK_ERR_ARG;
  *out = nullptr;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) return GPK_ERR_CUDA;
  if (device < 0 || device >= cnt) return GPK_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return GPK_ERR_CUDA;
  Handle* h = new (std::nothrow) Handle();
  if (!h) return GPK_ERR_NOMEM;
  h->device = device;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  auto fail = [&](int rc) { delete h; return rc; };
  if (cudaStreamCreateWithPriority(&h->s_main, cudaStreamNonBlocking, lo) != cudaSuccess) return fail(GPK_ERR_CUDA);
  if (cudaStreamCreateWithPriority(&h->s_panel, cudaStreamNonBlocking, hi) != cudaSuccess) return fail(GPK_ERR_CUDA);
  if (cudaStreamCreateWithPriority(&h->s_aux, cudaStreamNonBlocking, (hi < lo) ? hi + 1 : lo) != cudaSuccess)
    return fail(GPK_ERR_CUDA);
  if (cudaStreamCreateWithPriority(&h->s_tail, cudaStreamNonBlocking, (hi < lo) ? hi + 1 : lo) != cudaSuccess)
    return fail(GPK_ERR_CUDA);
  cudaEvent_t* te[] = {&h->t0, &h->t1, &h->t2, &h->t3, &h->t4};
  for (auto e : te)
    if (cudaEventCreate(e) != cudaSuccess) return fail(GPK_ERR_CUDA);
  if (cudaMallocHost((void**)&h->hPinned, 4096 * sizeof(double)) != cudaSuccess) return fail(GPK_ERR_CUDA);
  if (cudaMalloc((void**)&h->dHead, (size_t)NB * NB * sizeof(double)) != cudaSuccess) return fail(GPK_ERR_NOMEM);
  int rc = gemm_init(h);
  if (rc == 0) rc = diag_init(h);
  if (rc != 0) return fail(rc);
  *out = reinterpret_cast<gpk_handle>(h);
  return 0;
}

int gpk_destroy(gpk_handle hh) {
  if (!hh) return GPK_ERR_ARG;
  Handle* h = reinterpret_cast<Handle*>(hh);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  gpk_dist_finalize(hh);
  free_all(h);
  for (auto e : h->ev) cudaEventDestroy(e);
  for (auto e : h->prof_ev) cudaEventDestroy(e);
  cudaEvent_t te[] = {h->t0, h->t1, h->t2, h->t3, h->t4};
  for (auto e : te)
    if (e) cudaEventDestroy(e);
  if (h->hPinned) cudaFreeHost(h->hPinned);
  if (h->dHead) cudaFree(h->dHead);
  if (h->s_main) cudaStreamDestroy(h->s_main);
  if (h->s_panel) cudaStreamDestroy(h->s_panel);
  if (h->s_aux) cudaStreamDestroy(h->s_aux);
  if (h->s_tail) cudaStreamDestroy(h->s_tail);
  delete h;
  return 0;
}

const char* gpk_last_error(gpk_handle hh) {
  if (!hh) return "null handle";
  return reinterpret_cast<Handle*>(hh)->last_msg.c_str();
}

int gpk_last_stats(gpk_handle hh, gpk_stats* out) {
  if (!hh || !out) return GPK_ERR_ARG;
  *out = reinterpret_cast<Handle*>(hh)->stats;
  return 0;
}

int gpk_set_profile(gpk_handle hh, int profile) {
  if (!hh) return GPK_ERR_ARG;
  reinterpret_cast<Handle*>(hh)->profile = profile;
  return 0;
}

// ---------------------------------------------------------------------------
int gpk_set_data(gpk_handle hh, const double* X, int64_t n, int D) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!X || n <= 0 || D <= 0) return GPK_ERR_ARG;
  if (n != h->n || D != h->D || !h->dX) {
    if (h->dX) cudaFree(h->dX);
    h->dX = nullptr;
    GPK_CK(h, cudaMalloc((void**)&h->dX, (size_t)n * D * sizeof(double)));
  }
  GPK_TRY(alloc_problem(h, n, D));
  GPK_CK(h, cudaMemcpyAsync(h->dX, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
  GPK_CK(h, cudaStreamSynchronize(h->s_main));
  h->has_post = false;
  h->has_fitc = false;
  h->stats.h2d_bytes = n * D * (int64_t)sizeof(double);
  return 0;
}

// kind == GPK_COV_PROG: the program has been compiled into h->hprog and uploaded to h->dProg by the caller
static int exact_eval_core(Handle* h, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                           const double* ymm, int want_der, double* nlZ, double* alpha, double* dcov, double* dlik) {
  if (!h->dX || h->n <= 0) return GPK_ERR_STATE;
  if ((nhyp > 0 && !hyp) || !ymm || !nlZ || !alpha) return GPK_ERR_ARG;
  if (want_der && (!dcov || !dlik)) return GPK_ERR_ARG;
  const int64_t n = h->n, np = h->np;
  const int D = h->D, T = (int)(np / NB);
  const bool prog = (kind == GPK_COV_PROG);
  std::vector<double> scale(D, 1.0);
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  if (!prog) GPK_TRY(kind_scale(kind, matern_d, hyp, nhyp, D, scale, &divide, &premul, &sf2));
  if (D > 1900) return GPK_ERR_ARG;  // pinned staging layout
  const double sn2 = std::exp(2.0 * log_sn);
  stats_begin(h);
  h->has_post = false; h->post_ep = false;
  cudaStream_t st = h->s_main;
  GPK_TRY(ensure(h, &h->dA, &h->capA, np * np));

  GPK_CK(h, cudaEventRecord(h->t0, st));
  std::memcpy(h->hPinned, scale.data(), D * sizeof(double));
  GPK_CK(h, cudaMemcpyAsync(h->dScale, h->hPinned, D * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  GPK_TRY(launch_prescale(h, st, h->dX, n, np, D, h->dScale, divide, premul, h->dXs));
  CovArgs lazy{};
  {
    CovArgs c{};
    c.F = h->dXs; c.S = h->dXs; c.out = h->dA; c.ld = np;
    c.nF = n; c.nS = n; c.pF = np; c.pS = np; c.D = D;
    c.kind = kind; c.matern_d = matern_d; c.epi = EPI_COV; c.ard_dim = 0;
    c.sf2 = sf2; c.scale = 1.0 / sn2; c.diag_add = 1.0;
    c.same_set = 1; c.lower_only = 1; c.pad_identity = 1; c.padded128 = 1;
    if (prog) {                      // composite kernel: the program works on the RAW inputs
      c.F = h->dX; c.S = h->dX; c.prog = h->dProg; c.prog_der1 = 0; c.padded128 = 0;
      c.pre = (h->preN == n) ? h->dPre : nullptr; c.pre_ld = n;
    }
    lazy = c;
  }
  // right-hand side y - m, zero padded; dB is the forward-solve work copy
  GPK_CK(h, cudaMemsetAsync(h->dR, 0, (size_t)np * sizeof(double), st));
  GPK_CK(h, cudaMemcpyAsync(h->dR, ymm, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemcpyAsync(h->dB, h->dR, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, st));
  h->stats.h2d_bytes += (n + D) * (int64_t)sizeof(double);

  // the matrix build is issued by potrf_device (first block's columns, then the rest under the first panels);
  // stats.kbuild_ms is the time until the whole matrix exists, potrf_ms the rest of the factorisation
  if (env_int("GPK_LAZY_COV", 1)) {
    GPK_TRY(potrf_device(h, h->dA, np, h->dDinv, h->dScal, h->dInfo, h->dB, h->dZ, &lazy));
  } else {                 // A/B switch: whole matrix first, then the factorisation
    GPK_TRY(launch_cov(h, st, lazy));
    GPK_CK(h, cudaEventRecord(h->t1, st));
    GPK_TRY(potrf_device(h, h->dA, np, h->dDinv, h->dScal, h->dInfo, h->dB, h->dZ));
  }
  GPK_CK(h, cudaEventRecord(h->t2, st));
  GPK_TRY(launch_trsv_bwd_all(h, st, h->dA, np, h->dDinv, h->dZ, h->dB, T));
  double* res = h->dScal + T;  // [0]=r'alpha [1]=logdet ; [8..] derivative results
  GPK_TRY(launch_finish_alpha(h, st, h->dB, h->dR, 1.0 / sn2, np, h->dAlpha, h->dScal, T, res));
  GPK_CK(h, cudaEventRecord(h->t3, st));

  if (want_der) {
    GPK_TRY(ensure(h, &h->dU, &h->capU, np * np));
    GPK_TRY(ensure(h, &h->dW, &h->capW, np * np));
    const int64_t g = (n + 63) / 64;
    const int64_t part_need = g * g * 34;
    GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, part_need));
    // (K/sn2+I)^-1 = U U' with U = L^-T.  Both O(N^3) products run on the int8 tensor cores when the matrix is large
    // enough for it to pay and the int32 accumulators cannot overflow (7 N 2^14 < 2^31); else on DMMA.
    const bool der_oz = env_int("GPK_OZAKI", 1) && env_int("GPK_OZAKI_DER", 1) && T >= 16 && np < 18432;
    if (der_oz) {
      GPK_TRY(inverse_factor_T_oz(h, st, h->dU, h->dA, np, h->dDinv));
      GPK_TRY(oz_ensure(h, 1, np, (int)np));
      GPK_TRY(launch_oz_slice(h, 1, st, h->dU, np, (int)np, (int)np));
      GPK_TRY(launch_oz_ex(h, 1, st, h->dW, np, (int)np, (int)np, 0, T, 0, 0, /*trap*/ 1, /*set*/ 1));
    } else {
      GPK_TRY(inverse_factor_T(h, st, h->dU, h->dA, np, h->dDinv));
      GemmArgs v{};
      v.A = h->dU; v.B = h->dU; v.C = h->dW; v.lda = np; v.ldb = np; v.ldc = np;
      v.K = (int)np; v.tri = 2;
      GPK_TRY(launch_gemm_nt(h, st, 0, v, T, T));
    }
    if (prog) {
      GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, g * g * (nhyp + 1) + 64));
      GPK_TRY(launch_dnlz_prog(h, st, h->dProg, nhyp, h->dX, n, D, h->dW, np, h->dAlpha, 1.0 / sn2,
                               (h->preN == n) ? h->dPre : nullptr, n, h->dTmp, h->capTmp, res + 8));
    } else {
      GPK_TRY(launch_dnlz(h, st, h->dXs, n, D, h->dW, np, h->dAlpha, 1.0 / sn2, sf2, kind, matern_d, h->dTmp,
                          h->capTmp, res + 8));
    }
  }
  GPK_CK(h, cudaEventRecord(h->t4, st));

  const int nres = 8 + (want_der ? nhyp + 1 : 0);
  GPK_CK(h, cudaMemcpyAsync(h->hPinned, res, nres * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(alpha, h->dAlpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  h->stats.d2h_bytes = (n + nres) * (int64_t)sizeof(double) + 4;

  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t4); h->stats.total_ms = ms;
  cudaEventElapsedTime(&ms, h->t0, h->t1); h->stats.kbuild_ms = ms;
  cudaEventElapsedTime(&ms, h->t1, h->t2); h->stats.potrf_ms = ms;
  cudaEventElapsedTime(&ms, h->t2, h->t3); h->stats.solve_ms = ms;
  cudaEventElapsedTime(&ms, h->t3, h->t4); h->stats.deriv_ms = ms;
  GPK_TRY(collect_profile(h, T));

  const int info = *reinterpret_cast<int*>(h->hPinned + 2048);
  const double dot = h->hPinned[0], logdet = h->hPinned[1];
  *nlZ = dot / 2.0 + logdet + (double)n * std::log(2.0 * M_PI * sn2) / 2.0;
  if (want_der) {
    for (int i = 0; i < nhyp; ++i) dcov[i] = h->hPinned[8 + i] / 2.0;
    dlik[0] = sn2 * h->hPinned[8 + nhyp];
  }
  h->kind = kind; h->matern_d = matern_d; h->nhyp = nhyp; h->sn2 = sn2; h->sf2 = sf2;
  h->hyp.assign(hyp, hyp + nhyp);
  h->pn = 0;
  if (info != 0) return info;  // not positive definite: outputs are NaN-contaminated, as LAPACK would leave them
  h->has_post = true;
  return 0;
}

int gpk_exact_eval(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, double log_sn,
                   const double* ymm, int want_der, double* nlZ, double* alpha, double* dcov, double* dlik) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!hyp || kind < 0 || kind > GPK_COV_MATERN) return GPK_ERR_ARG;
  return exact_eval_core(h, kind, matern_d, hyp, nhyp, log_sn, ymm, want_der, nlZ, alpha, dcov, dlik);
}

int gpk_exact_eval_prog(gpk_handle hh, const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp,
                        double log_sn, const double* ymm, int want_der, double* nlZ, double* alpha, double* dcov,
                        double* dlik) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->dX || h->n <= 0) return GPK_ERR_STATE;
  GPK_TRY(prog_compile(nodes, nnodes, hyp, nhyp, h->D, &h->hprog));
  if (prog_has_op(h->hprog, OP_PRE) && h->preN != h->n) return GPK_ERR_STATE;     // gpk_set_pre first
  if (nhyp + 12 > 4 * h->D + 240) return GPK_ERR_ARG;                             // result slots behind the log-det parts
  GPK_TRY(prog_upload(h, h->s_main, h->hprog));
  static const double none = 0.0;
  return exact_eval_core(h, GPK_COV_PROG, 0, nhyp > 0 ? hyp : &none, nhyp, log_sn, ymm, want_der, nlZ, alpha, dcov, dlik);
}

int gpk_set_pre(gpk_handle hh, const double* Ktrain, int64_t n) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!Ktrain || n <= 0) return GPK_ERR_ARG;
  GPK_TRY(ensure(h, &h->dPre, &h->capPre, n * n));
  GPK_CK(h, cudaMemcpyAsync(h->dPre, Ktrain, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
  GPK_CK(h, cudaStreamSynchronize(h->s_main));
  h->preN = n;
  h->stats.h2d_bytes = n * n * (int64_t)sizeof(double);
  return 0;
}

int gpk_get_factor(gpk_handle hh, double* R_out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!R_out) return GPK_ERR_ARG;
  if (!h->has_post && h->pn == 0) return GPK_ERR_STATE;
  const int64_t n = h->has_post ? h->n : h->pn;
  const int64_t np = round_up(n, NB);
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, n * n));
  GPK_TRY(launch_compact_lower(h, h->s_main, h->dA, np, n, h->dTmp));
  GPK_CK(h, cudaMemcpyAsync(R_out, h->dTmp, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
  GPK_CK(h, cudaStreamSynchronize(h->s_main));
  h->stats.d2h_bytes = n * n * (int64_t)sizeof(double);
  return 0;
}

int gpk_predict(gpk_handle hh, const double* Xs, int64_t ns, double* ks_alpha, double* fs2) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->has_post) return GPK_ERR_STATE;
  if (!Xs || ns <= 0 || !ks_alpha || !fs2) return GPK_ERR_ARG;
  const int64_t n = h->n, np = h->np;
  const int D = h->D, T = (int)(np / NB);
  cudaStream_t st = h->s_main;
  stats_begin(h);
  GPK_CK(h, cudaEventRecord(h->t0, st));
  const int64_t chunk = 8192;
  const int nsplit = 16;
  const int64_t cp_max = round_up(ns < chunk ? ns : chunk, NB);
  // dP: transposed cross-covariance chunk (cp_max x np); dTmp: raw + scaled test inputs, partial sums, outputs
  GPK_TRY(ensure(h, &h->dP, &h->capP, cp_max * np));
  const int64_t tmp_need = 2 * cp_max * D + (int64_t)nsplit * cp_max + 3 * cp_max;
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, tmp_need));
  double* dXraw = h->dTmp;
  double* dXsc = dXraw + cp_max * D;
  double* dPart = dXsc + cp_max * D;
  double* dOut = dPart + (int64_t)nsplit * cp_max;
  // same input scaling as the posterior's kernel
  const bool prog = (h->kind == GPK_COV_PROG);
  if (prog && prog_has_op(h->hprog, OP_PRE)) return GPK_ERR_STATE;   // cov.Pre: cross-covariances are the caller's (M1)
  std::vector<double> scale(D, 1.0);
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  if (!prog) GPK_TRY(kind_scale(h->kind, h->matern_d, h->hyp.data(), h->nhyp, D, scale, &divide, &premul, &sf2));
  const double sn = std::sqrt(h->sn2);
  for (int64_t lo = 0; lo < ns; lo += chunk) {
    const int64_t m = (ns - lo < chunk) ? ns - lo : chunk;
    const int64_t mp = round_up(m, NB);
    GPK_CK(h, cudaMemcpyAsync(dXraw, Xs + lo * D, (size_t)m * D * sizeof(double), cudaMemcpyHostToDevice, st));
    GPK_TRY(launch_prescale(h, st, dXraw, m, mp, D, h->dScale, divide, premul, dXsc));
    CovArgs c{};
    c.F = dXsc; c.S = h->dXs; c.out = h->dP; c.ld = mp;
    c.nF = m; c.nS = n; c.pF = mp; c.pS = np; c.D = D;
    c.kind = h->kind; c.matern_d = h->matern_d; c.epi = EPI_COV;
    const bool ep = h->post_ep;                   // EP posterior: sW is a vector (sqrt of the site precisions)
    c.sf2 = sf2; c.scale = ep ? 1.0 : 1.0 / sn; c.diag_add = 0.0; c.same_set = 0; c.lower_only = 0; c.pad_identity = 0;
    c.padded128 = 1;
    const double* kss_vec = nullptr;
    if (prog) {
      // composite kernel: cross-covariances from the program on the raw inputs; prior variances k(z,z) per test point
      GPK_CK(h, cudaMemsetAsync(h->dP, 0, (size_t)mp * np * sizeof(double), st));   // padding rows / columns
      c.F = dXraw; c.S = h->dX; c.prog = h->dProg; c.prog_der1 = 0; c.padded128 = 0;
      c.pF = m; c.pS = n;
      GPK_TRY(launch_cov_prog_diag(h, st, h->dProg, dXraw, m, D, -1, dOut + 2 * cp_max));
      kss_vec = dOut + 2 * cp_max;
    }
    GPK_TRY(launch_cov(h, st, c));
    // Ks' alpha  (undo the 1/sn scaling)
    GPK_TRY(launch_rowdot(h, st, h->dP, mp, mp, np, h->dAlpha, 0, ep ? 1.0 : sn, 0.0, dPart, nsplit, dOut, m));
    if (ep) GPK_TRY(launch_colscale_inplace(h, st, h->dP, mp, mp, np, h->eSW));   // sW .* Ks  (Core/gp.py:415)
    // the multi-right-hand-side forward solve (mp x np x np flops): blocked, with the updates on the int8 tensor
    // cores, when there are enough test points and panels for it to pay (GPK_OZAKI_PREDICT=0: fp64 DMMA sweep)
    if (env_int("GPK_OZAKI", 1) && env_int("GPK_OZAKI_PREDICT", 1) && T >= 16 && mp >= 1024)
      GPK_TRY(sweep_forward_oz(h, st, h->dP, mp, (int)(mp / NB), h->dA, np, h->dDinv, T));
    else
      GPK_TRY(sweep_forward(h, st, h->dP, mp, (int)(mp / NB), h->dA, np, h->dDinv, T));
    GPK_TRY(launch_rowdot(h, st, h->dP, mp, mp, np, nullptr, 1, 1.0, sf2, dPart, nsplit, dOut + mp, m, kss_vec));
    GPK_CK(h, cudaMemcpyAsync(ks_alpha + lo, dOut, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPK_CK(h, cudaMemcpyAsync(fs2 + lo, dOut + mp, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPK_CK(h, cudaStreamSynchronize(st));
  }
  GPK_CK(h, cudaEventRecord(h->t1, st));
  GPK_CK(h, cudaEventSynchronize(h->t1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t1);
  h->stats.total_ms = ms;
  h->stats.h2d_bytes = ns * D * (int64_t)sizeof(double);
  h->stats.d2h_bytes = 2 * ns * (int64_t)sizeof(double);
  return 0;
}

// ---------------------------------------------------------------------------
int gpk_cov_matrix(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, const double* X, int64_t n,
                   const double* Z, int64_t m, int D, int mode, int der, double* out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!hyp || !out || D <= 0) return GPK_ERR_ARG;
  std::vector<double> scale;
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  GPK_TRY(kind_scale(kind, matern_d, hyp, nhyp, D, scale, &divide, &premul, &sf2));
  if (der >= nhyp) return GPK_ERR_ARG;
  int epi = EPI_COV, ard_dim = 0;
  if (der >= 0) {
    if (kind == GPK_COV_RBFARD) {
      if (der < D) { epi = EPI_DER_ARD; ard_dim = der; } else epi = EPI_DER_SF;
    } else {
      epi = (der == 0) ? EPI_DER_ELL : EPI_DER_SF;
    }
  }
  stats_begin(h);
  if (mode == GPK_MODE_SELF_TEST) {
    if (!Z || m <= 0) return GPK_ERR_ARG;
    // k(z,z) at zero distance: sf2 for all three kinds; derivatives: 2*sf2 w.r.t. log sf, 0 otherwise
    const double v = (epi == EPI_COV) ? sf2 : (epi == EPI_DER_SF ? 2.0 * sf2 : 0.0);
    for (int64_t i = 0; i < m; ++i) out[i] = v;
    return 0;
  }
  const bool train = (mode == GPK_MODE_TRAIN);
  if (!train && mode != GPK_MODE_CROSS) return GPK_ERR_ARG;
  if (!X || n <= 0) return GPK_ERR_ARG;
  if (!train && (!Z || m <= 0)) return GPK_ERR_ARG;
  const int64_t mm = train ? n : m;
  cudaStream_t st = h->s_main;
  double *dXr = nullptr, *dXq = nullptr, *dZq = nullptr, *dOut = nullptr, *dSc = nullptr;
  auto cleanup = [&]() {
    if (dXr) cudaFree(dXr);
    if (dXq) cudaFree(dXq);
    if (dZq) cudaFree(dZq);
    if (dOut) cudaFree(dOut);
    if (dSc) cudaFree(dSc);
  };
#define CKC(call)                                                          \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      h->last_cuda = e__;                                                  \
      h->last_msg = std::string(#call) + ": " + cudaGetErrorString(e__);   \
      cleanup();                                                           \
      return (e__ == cudaErrorMemoryAllocation) ? GPK_ERR_NOMEM : GPK_ERR_CUDA; \
    }                                                                      \
  } while (0)
  const int64_t big = (n > mm ? n : mm);
  // scaled inputs padded (zero rows) to whole 128-point blocks: the tile kernel bulk-copies 128 x D blocks
  const int64_t n128 = round_up(n, NB), m128 = round_up(train ? n : m, NB);
  CKC(cudaMalloc((void**)&dXr, (size_t)big * D * sizeof(double)));
  CKC(cudaMalloc((void**)&dXq, (size_t)n128 * D * sizeof(double)));
  if (!train) CKC(cudaMalloc((void**)&dZq, (size_t)m128 * D * sizeof(double)));
  CKC(cudaMalloc((void**)&dOut, (size_t)n * mm * sizeof(double)));
  CKC(cudaMalloc((void**)&dSc, (size_t)D * sizeof(double)));
  CKC(cudaMemcpyAsync(dSc, scale.data(), D * sizeof(double), cudaMemcpyHostToDevice, st));
  CKC(cudaMemcpyAsync(dXr, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, st));
  int rc = launch_prescale(h, st, dXr, n, n128, D, dSc, divide, premul, dXq);
  if (rc == 0 && !train) {
    CKC(cudaMemcpyAsync(dXr, Z, (size_t)m * D * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = launch_prescale(h, st, dXr, m, m128, D, dSc, divide, premul, dZq);
  }
  if (rc == 0) {
    CovArgs c{};
    // C-order (n, mm) output: fast index = column j (second point set), slow = row i
    c.F = train ? dXq : dZq; c.S = dXq; c.out = dOut; c.ld = mm;
    c.nF = mm; c.nS = n; c.pF = mm; c.pS = n; c.D = D;
    c.kind = kind; c.matern_d = matern_d; c.epi = epi; c.ard_dim = ard_dim;
    c.sf2 = sf2; c.scale = 1.0; c.diag_add = 0.0; c.same_set = train ? 1 : 0; c.lower_only = 0; c.pad_identity = 0;
    c.padded128 = 1;
    rc = launch_cov(h, st, c);
  }
  if (rc != 0) { cleanup(); return rc; }
  CKC(cudaMemcpyAsync(out, dOut, (size_t)n * mm * sizeof(double), cudaMemcpyDeviceToHost, st));
  CKC(cudaStreamSynchronize(st));
#undef CKC
  cleanup();
  h->stats.h2d_bytes = (n + (train ? 0 : m)) * D * (int64_t)sizeof(double);
  h->stats.d2h_bytes = n * mm * (int64_t)sizeof(double);
  return 0;
}

// getCovMatrix / getDerMatrix of a composite kernel (program), all three modes
int gpk_cov_matrix_prog(gpk_handle hh, const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp,
                        const double* X, int64_t n, const double* Z, int64_t m, int D, int mode, int der, double* out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!out || D <= 0 || der >= nhyp) return GPK_ERR_ARG;
  CovProg prog;
  GPK_TRY(prog_compile(nodes, nnodes, hyp, nhyp, D, &prog));
  const bool has_pre = prog_has_op(prog, OP_PRE);
  stats_begin(h);
  cudaStream_t st = h->s_main;
  const bool train = (mode == GPK_MODE_TRAIN), selft = (mode == GPK_MODE_SELF_TEST);
  if (!train && !selft && mode != GPK_MODE_CROSS) return GPK_ERR_ARG;
  if (selft ? (!Z || m <= 0) : (!X || n <= 0)) return GPK_ERR_ARG;
  if (mode == GPK_MODE_CROSS && (!Z || m <= 0)) return GPK_ERR_ARG;
  if (has_pre && (!train || h->preN != n)) return GPK_ERR_STATE;     // cov.Pre: only the training matrix lives here
  GPK_TRY(prog_upload(h, st, prog));
  const int64_t mm = train ? n : m;
  const int64_t nout = selft ? m : n * mm;
  double *dXr = nullptr, *dZr = nullptr, *dOut = nullptr;
  auto cleanup = [&]() {
    if (dXr) cudaFree(dXr);
    if (dZr) cudaFree(dZr);
    if (dOut) cudaFree(dOut);
  };
  int rc = 0;
  cudaError_t e = cudaSuccess;
  do {
    if (!selft) {
      if ((e = cudaMalloc((void**)&dXr, (size_t)n * D * sizeof(double))) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(dXr, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    }
    if (!train) {
      if ((e = cudaMalloc((void**)&dZr, (size_t)m * D * sizeof(double))) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(dZr, Z, (size_t)m * D * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    }
    if ((e = cudaMalloc((void**)&dOut, (size_t)nout * sizeof(double))) != cudaSuccess) break;
    if (selft) {
      rc = launch_cov_prog_diag(h, st, h->dProg, dZr, m, D, der, dOut);
    } else {
      CovArgs c{};
      // C-order (n, mm) output: fast index = column j (second point set), slow = row i
      c.F = train ? dXr : dZr; c.S = dXr; c.out = dOut; c.ld = mm;
      c.nF = mm; c.nS = n; c.pF = mm; c.pS = n; c.D = D;
      c.scale = 1.0; c.diag_add = 0.0; c.same_set = train ? 1 : 0;
      c.prog = h->dProg; c.prog_der1 = der + 1;
      c.pre = has_pre ? h->dPre : nullptr; c.pre_ld = n;
      rc = launch_cov(h, st, c);
    }
    if (rc != 0) break;
    if ((e = cudaMemcpyAsync(out, dOut, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  cleanup();
  if (rc != 0) return rc;
  GPK_CK(h, e);
  h->stats.h2d_bytes = ((selft ? 0 : n) + (train ? 0 : m)) * D * (int64_t)sizeof(double);
  h->stats.d2h_bytes = nout * (int64_t)sizeof(double);
  return 0;
}

// ---------------------------------------------------------------------------
int gpk_potrf(gpk_handle hh, const double* A, int64_t n, double* R_out, double* logdet_half) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!A || n <= 0) return GPK_ERR_ARG;
  const int64_t pn = round_up(n, NB);
  const int T = (int)(pn / NB);
  cudaStream_t st = h->s_main;
  stats_begin(h);
  h->has_post = false; h->has_fitc = false;
  // the standalone factor reuses the posterior's storage; sizes follow this call
  if (pn != h->np || !h->dXs) {
    h->n = 0;
    GPK_TRY(alloc_problem(h, n, h->D > 0 ? h->D : 1));
    h->n = 0;
  }
  GPK_TRY(ensure(h, &h->dA, &h->capA, pn * pn));
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, n * n));
  GPK_CK(h, cudaMemcpyAsync(h->dTmp, A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  GPK_CK(h, cudaEventRecord(h->t0, st));
  GPK_TRY(launch_pad_sym(h, st, h->dTmp, n, h->dA, pn));
  GPK_TRY(potrf_device(h, h->dA, pn, h->dDinv, h->dScal, h->dInfo, nullptr, nullptr));
  GPK_TRY(launch_sum_parts(h, st, h->dScal, T, h->dScal + T));
  GPK_CK(h, cudaEventRecord(h->t1, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned, h->dScal + T, sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t1);
  h->stats.total_ms = ms; h->stats.potrf_ms = ms;
  GPK_TRY(collect_profile(h, T));
  const int info = *reinterpret_cast<int*>(h->hPinned + 2048);
  if (logdet_half) *logdet_half = h->hPinned[0];
  h->pn = n;
  if (R_out) GPK_TRY(gpk_get_factor(hh, R_out));
  if (info != 0) { h->pn = 0; return info; }
  return 0;
}

// Upload a factor computed elsewhere (post.L of a posterior kept on the host, Core/gp.py:404-416) as the resident factor:
// R (n,n) C-order upper, A = R'R.  Only the block inverses and the log-determinant are computed (one launch).
int gpk_set_factor(gpk_handle hh, const double* R, int64_t n, double* logdet_half) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!R || n <= 0) return GPK_ERR_ARG;
  const int64_t pn = round_up(n, NB);
  const int T = (int)(pn / NB);
  cudaStream_t st = h->s_main;
  stats_begin(h);
  h->has_post = false; h->has_fitc = false; h->pn = 0;
  if (pn != h->np || !h->dXs) {
    h->n = 0;
    GPK_TRY(alloc_problem(h, n, h->D > 0 ? h->D : 1));
    h->n = 0;
  }
  GPK_TRY(ensure(h, &h->dA, &h->capA, pn * pn));
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, n * n));
  GPK_CK(h, cudaMemcpyAsync(h->dTmp, R, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  // the C-order upper factor read as column-major IS the lower factor L = R'; padding rows/columns are the identity
  GPK_TRY(launch_pad_sym(h, st, h->dTmp, n, h->dA, pn));
  GPK_TRY(launch_diag_invert(h, st, h->dA, pn, h->dDinv, h->dScal, h->dInfo, T));
  GPK_TRY(launch_sum_parts(h, st, h->dScal, T, h->dScal + T));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned, h->dScal + T, sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  h->stats.h2d_bytes = n * n * (int64_t)sizeof(double);
  const int info = *reinterpret_cast<int*>(h->hPinned + 2048);
  if (logdet_half) *logdet_half = h->hPinned[0];
  if (info != 0) return info;          // non-positive diagonal entry: not a Cholesky factor
  h->pn = n;
  return 0;
}

// X = (R'R)^-1 B by two triangular sweeps over the resident factor (never the explicit inverse): with the right-hand
// sides transposed (one per row), Xt = Bt L^-T L^-1.  The backward sweep reads L' (one bandwidth-bound transpose),
// which is cached on the handle until the factor changes, so repeated solves cost 2 n^2 nrhs flops each.
int gpk_potrs(gpk_handle hh, const double* B, int64_t n_rows, int64_t nrhs, double* X_out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (h->pn <= 0) return GPK_ERR_STATE;
  if (!B || !X_out || nrhs <= 0) return GPK_ERR_ARG;
  if (n_rows != h->pn) return GPK_ERR_ARG;            // B must have exactly as many rows as the resident factor
  const int64_t n = h->pn, pn = round_up(n, NB);
  const int T = (int)(pn / NB);
  cudaStream_t st = h->s_main;
  const bool cached = h->lt_valid;
  stats_begin(h);
  GPK_TRY(ensure(h, &h->dU, &h->capU, pn * pn));
  GPK_TRY(ensure(h, &h->dDinvT, &h->capDinvT, pn * NB));
  if (!cached) {
    GPK_TRY(launch_transpose(h, st, h->dA, pn, 0, h->dU, pn, 0, pn, pn, 1));
    GPK_TRY(launch_transpose(h, st, h->dDinv, NB, (int64_t)NB * NB, h->dDinvT, NB, (int64_t)NB * NB, NB, NB, T));
  }
  // B (n,nrhs) C-order == (nrhs x n) column-major with pitch nrhs: exactly the transposed layout the sweeps use
  const int64_t chunk = 8192;                          // right-hand sides per pass (bounds the work buffer)
  const int64_t rp_max = round_up(nrhs < chunk ? nrhs : chunk, NB);
  GPK_TRY(ensure(h, &h->dP, &h->capP, rp_max * pn));
  for (int64_t lo = 0; lo < nrhs; lo += chunk) {
    const int64_t m = (nrhs - lo < chunk) ? nrhs - lo : chunk, rp = round_up(m, NB);
    double* P0 = h->dP;
    GPK_CK(h, cudaMemsetAsync(P0, 0, (size_t)rp * pn * sizeof(double), st));
    GPK_CK(h, cudaMemcpy2DAsync(P0, (size_t)rp * sizeof(double), B + lo, (size_t)nrhs * sizeof(double),
                                (size_t)m * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, st));
    GPK_TRY(sweep_forward(h, st, P0, rp, (int)(rp / NB), h->dA, pn, h->dDinv, T));
    GPK_TRY(sweep_backward(h, st, P0, rp, (int)(rp / NB), h->dU, pn, h->dDinvT, T));
    GPK_CK(h, cudaMemcpy2DAsync(X_out + lo, (size_t)nrhs * sizeof(double), P0, (size_t)rp * sizeof(double),
                                (size_t)m * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  GPK_CK(h, cudaStreamSynchronize(st));
  h->lt_valid = true;
  h->stats.h2d_bytes = n * nrhs * (int64_t)sizeof(double);
  h->stats.d2h_bytes = n * nrhs * (int64_t)sizeof(double);
  return 0;
}

// ---------------------------------------------------------------------------
int gpk_bench_dmma(gpk_handle hh, int shape, int warps_per_cta, int iters, double* tflops, double* ms) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!tflops || !ms) return GPK_ERR_ARG;
  return bench_dmma(h, shape, warps_per_cta, iters, tflops, ms);
}

int gpk_bench_syrk(gpk_handle hh, int64_t n, int k, int reps, double* ms_out, double* tflops) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (n <= 0 || n % NB != 0 || k <= 0 || k % 32 != 0 || reps <= 0 || !ms_out || !tflops) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  double *C = nullptr, *P = nullptr;
  GPK_CK(h, cudaMalloc((void**)&C, (size_t)n * n * sizeof(double)));
  if (cudaMalloc((void**)&P, (size_t)n * k * sizeof(double)) != cudaSuccess) { cudaFree(C); return GPK_ERR_NOMEM; }
  launch_fill_random(h, st, C, n * n, 1u);
  launch_fill_random(h, st, P, n * k, 2u);
  GemmArgs u{};
  u.A = P; u.B = P; u.C = C; u.lda = n; u.ldb = n; u.ldc = n; u.K = k; u.tri = 1;
  const int T = (int)(n / NB);
  int rc = launch_gemm_nt(h, st, 1, u, T, T);  // warm-up
  cudaStreamSynchronize(st);
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps && rc == 0; ++r) {
    cudaEventRecord(h->t0, st);
    rc = launch_gemm_nt(h, st, 1, u, T, T);
    cudaEventRecord(h->t1, st);
    cudaEventSynchronize(h->t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->t0, h->t1);
    sum += ms;
    if (ms < best) best = ms;
  }
  cudaFree(C);
  cudaFree(P);
  if (rc != 0) return rc;
  GPK_CK(h, cudaGetLastError());
  const double avg = sum / reps;
  *ms_out = avg;
  *tflops = (double)k * (double)n * (double)n / (avg * 1e-3) / 1e12;  // algorithmic: k*n^2 for the lower triangle
  (void)best;
  return 0;
}

int gpk_bench_copy(gpk_handle hh, int64_t bytes, int reps, double* gbs) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (bytes < 1024 || reps <= 0 || !gbs) return GPK_ERR_ARG;
  const int64_t n = bytes / 8 / 4 * 4;
  double *a = nullptr, *b = nullptr;
  GPK_CK(h, cudaMalloc((void**)&a, (size_t)n * 8));
  if (cudaMalloc((void**)&b, (size_t)n * 8) != cudaSuccess) { cudaFree(a); return GPK_ERR_NOMEM; }
  cudaStream_t st = h->s_main;
  launch_fill_random(h, st, a, n, 3u);
  launch_copy(h, st, a, b, n);
  cudaStreamSynchronize(st);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(h->t0, st);
    launch_copy(h, st, a, b, n);
    cudaEventRecord(h->t1, st);
    cudaEventSynchronize(h->t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->t0, h->t1);
    if (ms < best) best = ms;
  }
  cudaFree(a);
  cudaFree(b);
  GPK_CK(h, cudaGetLastError());
  *gbs = 2.0 * n * 8 / (best * 1e-3) / 1e9;
  return 0;
}

int gpk_dbg_gemm_nt(gpk_handle hh, int mode, int64_t M, int64_t N, int64_t K, const double* A, const double* B,
                    double* C) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!A || !B || !C || M % NB || N % NB || K % 32 || mode < 0 || mode > 7) return GPK_ERR_ARG;
  if (mode >= 5 && (M != NB || N != NB || K % NB)) return GPK_ERR_ARG;
  if (mode == 7 && K != NB) return GPK_ERR_ARG;
  if ((mode == 2 || mode == 3) && M != N) return GPK_ERR_ARG;
  if (mode == 4 && N != K) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  GPK_CK(h, cudaMalloc((void**)&dA, (size_t)M * K * 8));
  GPK_CK(h, cudaMalloc((void**)&dB, (size_t)N * K * 8));
  GPK_CK(h, cudaMalloc((void**)&dC, (size_t)M * N * 8));
  cudaMemcpyAsync(dA, A, (size_t)M * K * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dB, B, (size_t)N * K * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dC, C, (size_t)M * N * 8, cudaMemcpyHostToDevice, st);
  GemmArgs u{};
  u.A = dA; u.B = dB; u.C = dC; u.lda = M; u.ldb = N; u.ldc = M; u.K = (int)K;
  u.tri = (mode == 2) ? 1 : (mode == 3 ? 2 : 0);
  if (mode == 4) u.C = dA;   // in place over A, as the panel TRSM runs
  int rc = 0;
  if (mode >= 5) {
    // the chain products (small_nt_kernel).  5: C = A B' ; 6: C -= A B', blocks on / below the diagonal ;
    // 7: the head pair of the panel chain: X = A B' into the scratch tile, C -= X X' (lower), A <- X
    SmallArgs s1{};
    s1.A = dA; s1.lda = M; s1.B = dB; s1.ldb = N; s1.C = (mode == 7) ? h->dHead : dC; s1.ldc = NB; s1.K = (int)K;
    s1.mode = (mode == 6) ? 1 : 0; s1.tri = (mode == 6) ? 1 : 0;
    s1.breg = env_int("GPK_SMALL_BREG", 0);
    rc = launch_small_nt(h, st, s1);
    if (rc == 0 && mode == 7) {
      SmallArgs s2{};
      s2.A = h->dHead; s2.lda = NB; s2.B = h->dHead; s2.ldb = NB; s2.C = dC; s2.ldc = NB; s2.K = NB; s2.mode = 1; s2.tri = 1;
      s2.copy_dst = dA; s2.ld_copy = M; s2.breg = s1.breg;
      rc = launch_small_nt(h, st, s2);
      cudaMemcpyAsync(const_cast<double*>(A), dA, (size_t)M * K * 8, cudaMemcpyDeviceToHost, st);
    }
  } else {
    rc = launch_gemm_nt(h, st, (mode == 0 || mode >= 3) ? 0 : 1, u, (int)(M / NB), (int)(N / NB));
  }
  cudaMemcpyAsync(C, (mode == 4) ? dA : dC, (size_t)M * N * 8, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  if (rc != 0) return rc;
  GPK_CK(h, e);
  return 0;
}

int gpk_dbg_diag(gpk_handle hh, const double* A128, double* L128, double* Linv128, double* logdet_half, int* info) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!A128 || !L128 || !Linv128 || !logdet_half || !info) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  double *dA = nullptr, *dI = nullptr, *dS = nullptr;
  int* dInfo = nullptr;
  GPK_CK(h, cudaMalloc((void**)&dA, NB * NB * 8));
  GPK_CK(h, cudaMalloc((void**)&dI, NB * NB * 8));
  GPK_CK(h, cudaMalloc((void**)&dS, 64 * 8));
  GPK_CK(h, cudaMalloc((void**)&dInfo, 16));
  // only the lower triangle goes up (the kernel reads nothing else), and the inverse starts from zeros (see ensure_zero):
  // with GPK_DIAG_SKIPZ the kernel leaves the blocks above the block diagonal of both tiles as it finds them
  std::vector<double> lower((size_t)NB * NB, 0.0);
  for (int c = 0; c < NB; ++c)
    for (int r = c; r < NB; ++r) lower[r + (size_t)c * NB] = A128[r + (size_t)c * NB];
  cudaMemsetAsync(dI, 0, NB * NB * 8, st);
  cudaMemcpyAsync(dA, lower.data(), NB * NB * 8, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);                     // `lower` is pageable: the copy has left the host buffer after this
  cudaMemsetAsync(dInfo, 0, 16, st);
  cudaMemsetAsync(dS, 0, 64 * 8, st);
  int rc = launch_diag(h, st, dA, NB, dI, dS, dInfo, 0, reinterpret_cast<long long*>(dS + 8));
  if (getenv("GPK_DBG_DIAG_CLK")) {
    long long clk[32] = {0};
    cudaStreamSynchronize(st);
    cudaMemcpy(clk, dS + 8, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
    fprintf(stderr, "diag clk deltas:");
    for (int i = 1; i < 32 && clk[i]; ++i) fprintf(stderr, " %lld", clk[i] - clk[i - 1]);
    fprintf(stderr, "  total %lld\n", clk[0] ? 0LL : 0LL);
  }
  cudaMemcpyAsync(L128, dA, NB * NB * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(Linv128, dI, NB * NB * 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(logdet_half, dS, 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(info, dInfo, 4, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dA); cudaFree(dI); cudaFree(dS); cudaFree(dInfo);
  if (rc != 0) return rc;
  GPK_CK(h, e);
  return 0;
}

}  // extern "C"
