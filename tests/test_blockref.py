"""CPU checks of the device orchestration logic through its numpy model (tests/blockref.py)."""
import numpy as np
import scipy.linalg as sla

import blockref as br


def _spd(n, seed=0, cond_shift=1.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 4))
    d = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d / 4.0) / 0.01 + cond_shift * np.eye(n)


def test_diag_block_matches_lapack():
    A = _spd(128, 1)
    L, Li, ld, info = br.diag_block(np.asfortranarray(A))
    Lref = np.linalg.cholesky(A)
    assert info == 0
    np.testing.assert_allclose(L, Lref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(Li @ Lref, np.eye(128), atol=1e-9)
    assert np.all(np.triu(Li, 1) == 0) and np.all(np.triu(L, 1) == 0)
    np.testing.assert_allclose(ld, np.log(np.diag(Lref)).sum(), rtol=1e-13)


def test_diag_block_reports_first_bad_pivot():
    A = _spd(128, 2)
    A[40, 40] = -5.0
    _, _, _, info = br.diag_block(np.asfortranarray(A))
    assert info == 41


import pytest


@pytest.mark.parametrize("n,W", [(300, 1), (300, 2), (700, 2), (700, 3), (520, 4)])
def test_blocked_potrf_with_lookahead_split_and_fused_forward_solve(n, W):
    A = _spd(n, 3)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(n)
    P = br.pad_spd(A)
    b = np.zeros(P.shape[0]); b[:n] = y
    Dinv, parts, info, z = br.potrf_device(P, b, W=W)
    Lref = np.linalg.cholesky(A)
    assert info == 0
    np.testing.assert_allclose(np.tril(P)[:n, :n], Lref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(parts.sum(), np.log(np.diag(Lref)).sum(), rtol=1e-12)
    np.testing.assert_allclose(z[:n], sla.solve_triangular(Lref, y, lower=True), rtol=1e-9, atol=1e-9)
    # backward steps
    T = P.shape[0] // br.NB
    x = np.zeros_like(z)
    for k in range(T - 1, -1, -1):
        br.trsv_bwd(P, Dinv, z, x, k, T)
    np.testing.assert_allclose(x[:n], np.linalg.solve(A, y), rtol=1e-8, atol=1e-8)
    assert np.all(x[n:] == 0)


def test_transposed_forward_sweep_is_predict_solve():
    n, ns = 260, 130
    A = _spd(n, 4)
    P = br.pad_spd(A)
    Dinv, _, _, _ = br.potrf_device(P)
    rng = np.random.default_rng(1)
    Ks = rng.standard_normal((n, ns))
    nsp = 256
    Pt = np.asfortranarray(np.zeros((nsp, P.shape[0])))
    Pt[:ns, :n] = Ks.T
    br.sweep_forward(Pt, P, Dinv, P.shape[0] // br.NB)
    V = sla.solve_triangular(np.linalg.cholesky(A), Ks, lower=True)
    np.testing.assert_allclose(Pt[:ns, :n], V.T, rtol=1e-9, atol=1e-9)


def test_inverse_through_upper_factor():
    n = 384
    A = _spd(n, 5)
    P = br.pad_spd(A)
    Dinv, _, _, _ = br.potrf_device(P)
    U = br.inverse_factor_T(P, Dinv)
    Lref = np.linalg.cholesky(A)
    np.testing.assert_allclose(np.triu(U)[:n, :n], np.linalg.inv(Lref).T, rtol=1e-8, atol=1e-9)
    assert np.all(np.tril(U, -1) == 0)
    W = br.inverse_lower(U)
    Ainv = np.linalg.inv(A)
    np.testing.assert_allclose(np.tril(W)[:n, :n], np.tril(Ainv), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n,W,W1", [(1100, 2, 4), (1100, 1, 3), (900, 3, 6)])
def test_three_level_potrf_with_int8_sliced_updates(n, W, W1):
    """api.cu potrf_device with level-1 blocks of W1 panels and the trailing updates through the Ozaki split."""
    A = _spd(n, 5)
    rng = np.random.default_rng(1)
    y = rng.standard_normal(n)
    for oz in (False, True):
        P = br.pad_spd(A)
        b = np.zeros(P.shape[0]); b[:n] = y
        Dinv, parts, info, z = br.potrf_device(P, b, W=W, W1=W1, w1_minrem=0, oz=oz)
        Lref = np.linalg.cholesky(A)
        assert info == 0
        np.testing.assert_allclose(np.tril(P)[:n, :n], Lref, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(parts.sum(), np.log(np.diag(Lref)).sum(), rtol=1e-12)
        np.testing.assert_allclose(z[:n], sla.solve_triangular(Lref, y, lower=True), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("S,RB", [(7, 8), (8, 7), (7, 7), (6, 8)])
def test_ozaki_split_is_exact_to_the_last_slice(S, RB):
    """ozaki.cu: balanced radix-2^RB digits reconstruct every entry to 2^(e_i - S*RB - 1), digits fit int8, and the
    leading digit never overflows (rows whose maximum is just below a power of two included)."""
    rng = np.random.default_rng(S * 10 + RB)
    P = rng.standard_normal((64, 96)) * np.exp(rng.uniform(-30, 30, size=(64, 1)))
    P[0, 0] = np.nextafter(2.0 ** 7, 0.0)          # largest mantissa
    P[1, :] = 0.0                                   # an all-zero row
    P[2, 5] = -np.nextafter(2.0 ** -3, 0.0)
    d, sc = br.oz_split(P, S, RB)
    assert d.dtype == np.int8 and d.min() >= -(1 << (RB - 1)) and d.max() <= (1 << (RB - 1)) - 1
    rec = sum(d[t].astype(np.float64) * 2.0 ** (-RB * t) for t in range(S)) * sc[:, None]
    tol = sc * 2.0 ** RB * 2.0 ** (-S * RB - 1)     # 2^e_i * 2^(-S*RB-1)
    assert np.all(np.abs(rec - P) <= tol[:, None] * (1 + 1e-12))


def test_ozaki_syrk_matches_fp64_product():
    """The int8-sliced update is at least as accurate as an fp64 accumulation of the same contraction."""
    rng = np.random.default_rng(3)
    n, k = 96, 384
    P = rng.standard_normal((n, k)) * np.exp(rng.uniform(-6, 6, size=(n, 1)))
    C0 = rng.standard_normal((n, n)); C0 = C0 + C0.T
    C = C0.copy()
    br.oz_syrk(C, P)
    # exact reference in long double is not enough (x87 64-bit mantissa at best): use exact rationals on a sample
    from fractions import Fraction
    idx = [(0, 0), (5, 3), (40, 17), (95, 95), (70, 2), (33, 32)]
    den = np.abs(P) @ np.abs(P).T
    for (i, j) in idx:
        exact = sum(Fraction(P[i, q]) * Fraction(P[j, q]) for q in range(k))
        want = float(Fraction(C0[i, j]) - exact)
        assert abs(C[i, j] - want) <= 4e-16 * den[i, j] + 2e-16 * abs(C0[i, j])
    assert np.array_equal(np.triu(C, 1), np.triu(C0, 1))      # strictly upper part untouched


@pytest.mark.parametrize("nt,jb0,jb1", [(1, 0, 1), (5, 0, 5), (40, 0, 40), (40, 0, 3), (40, 3, 40), (33, 9, 33),
                                         (17, 0, 17), (16, 0, 16), (119, 0, 9), (119, 9, 119)])
def test_update_tile_rasterisation_covers_each_tile_once(nt, jb0, jb1):
    n = br.oz_ntiles(nt, jb0, jb1)
    got = [br.oz_decode(i, nt, jb0, jb1) for i in range(n)]
    want = {(ti, tj) for jb in range(jb0, jb1) for ti in range(jb, nt) for tj in (2 * jb, 2 * jb + 1)}
    assert len(set(got)) == n and set(got) == want


def test_ldl_block_with_newton_reciprocal_matches_cholesky():
    """The 32x32 sub-block factorisation of potrf_diag_kernel (square-root-free elimination, 20-bit reciprocal seed + one
    cubic Newton step, square roots at the end) against LAPACK, also on a badly scaled block."""
    rng = np.random.default_rng(5)
    for scale in (1.0, 1e-8, 1e8):
        B = rng.standard_normal((32, 40))
        A = (B @ B.T + 0.5 * np.eye(32)) * scale
        L, lg, bad = br.ldl_block(A)
        Lref = np.linalg.cholesky(A)
        assert bad == 0
        assert np.max(np.abs(L - Lref)) <= 1e-13 * np.max(np.abs(Lref))
        assert abs(lg - np.log(np.diag(Lref)).sum()) <= 1e-12 * max(1.0, abs(lg))
    x = rng.uniform(1e-3, 1e3, size=1000)
    assert np.max(np.abs(br.rcp_newton(x) * x - 1.0)) < 2.0 ** -50


def test_ldl_block_flags_first_bad_pivot():
    A = np.eye(32)
    A[7, 7] = -1.0
    assert br.ldl_block(A)[2] == 8


@pytest.mark.parametrize("n,W,W1,W2B", [(1536, 3, 0, None), (2048, 3, 9, 4), (1280, 2, 4, 3)])
def test_split_panel_chain_and_head_update_give_the_same_factor(n, W, W1, W2B):
    """The split panel chain (head on the panel stream, tail one panel behind) and the head of the level-1 hand-over
    touch every tile exactly once: same factor as the unsplit order, and as LAPACK."""
    A0 = _spd(n, seed=3)
    Ls = []
    for split in (False, True):
        A = br.pad_spd(A0)
        Dinv, parts, info, _ = br.potrf_device(A, W=W, W1=W1, w1_minrem=2, split=split, W2B=W2B)
        assert info == 0
        Ls.append(np.tril(A[:n, :n]))
    Lref = np.linalg.cholesky(A0)
    assert np.max(np.abs(Ls[1] - Lref)) <= 1e-11 * np.max(np.abs(Lref))
    assert np.max(np.abs(Ls[1] - Ls[0])) <= 1e-12 * np.max(np.abs(Lref))


@pytest.mark.parametrize("nt,jb0,jb1,ti_min", [(40, 0, 5, 5), (37, 0, 12, 12), (70, 0, 3, 3), (20, 0, 6, 2), (33, 2, 9, 4)])
def test_rectangular_tile_enumeration_of_the_stacked_product(nt, jb0, jb1, ti_min):
    """With ti_min the same band rasterisation enumerates {jb0 <= jb < jb1, ti >= max(jb, ti_min)} exactly once -
    for ti_min >= jb1 that is the nb x na rectangle of launch_oz_gemm_stacked."""
    n = br.oz_ntiles(nt, jb0, jb1, ti_min)
    got = [br.oz_decode(i, nt, jb0, jb1, ti_min=ti_min) for i in range(n)]
    want = {(ti, tj) for jb in range(jb0, jb1) for ti in range(max(jb, ti_min), nt) for tj in (2 * jb, 2 * jb + 1)}
    assert len(set(got)) == n and set(got) == want


def test_stacked_operand_product_matches_fp64():
    """A B' from ONE split of [B; A] (different row scales per operand row) against the fp64 product."""
    rng = np.random.default_rng(9)
    A = rng.standard_normal((256, 128)) * np.exp(rng.uniform(-5, 5, size=(256, 1)))
    B = rng.standard_normal((128, 128)) * np.exp(rng.uniform(-5, 5, size=(128, 1)))
    C0 = rng.standard_normal((256, 128))
    C = C0.copy()
    br.oz_gemm_stacked(C, A, B)
    ref = C0 - A @ B.T
    den = np.abs(A) @ np.abs(B).T
    assert np.max(np.abs(C - ref) / den) < 1e-14


@pytest.mark.parametrize("nt,ncols,cfirst,cs", [(20, 5, 0, 3), (37, 9, 2, 4), (9, 9, 0, 1), (18, 2, 1, 8)])
def test_block_cyclic_tile_enumeration(nt, ncols, cfirst, cs):
    """launch_oz_cyclic: every tile {local block c, row tile ti >= cfirst + c*cs} exactly once, with the global column
    tile for slices / masks and the packed local one for the C address."""
    got = br.oz_cyclic_tiles(nt, ncols, cfirst, cs)
    want = {(ti, 2 * (cfirst + c * cs) + hh, 2 * c + hh) for c in range(ncols) for hh in (0, 1)
            for ti in range(cfirst + c * cs, nt)}
    assert len(got) == len(set(got)) and set(got) == want


@pytest.mark.parametrize("G,WD,n", [(1, 3, 1280), (2, 3, 1536), (3, 4, 1408), (4, 2, 1100)])
def test_blocked_sharded_factorisation_matches_lapack(G, WD, n):
    """The blocked variant of the sharded factorisation (immediate rank-128 updates inside a block of WD panels, one
    delayed rank-(WD*128) update of the block-cyclic columns per block), G ranks simulated in one process."""
    A0 = _spd(n, seed=4)
    L = br.potrf_dist_blocked(br.pad_spd(A0), G, WD)
    Lref = np.linalg.cholesky(A0)
    assert np.max(np.abs(L[:n, :n] - Lref)) <= 1e-11 * np.max(np.abs(Lref))


@pytest.mark.parametrize("n,W,W1", [(1536, 3, 6), (1408, 2, 4)])
def test_fixed_row_scales_with_shared_slices_specification(n, W, W1):
    """Executable specification of the next slicing scheme: fixed row scales from the
    diagonal, every sub-block sliced once, the same digits used by the level-2 and the level-1 updates - as accurate
    as LAPACK on an RBF matrix with condition number ~1e5."""
    A0 = _spd(n, seed=6)
    A = br.pad_spd(A0)
    Dinv, parts, info, _ = br.potrf_device(A, W=W, W1=W1, w1_minrem=2, oz=True, fixed_scale=True, split=False)
    assert info == 0
    Lref = np.linalg.cholesky(A0)
    assert np.max(np.abs(np.tril(A[:n, :n]) - Lref)) <= 1e-11 * np.max(np.abs(Lref))
    assert abs(parts.sum() - np.log(np.diag(Lref)).sum()) <= 1e-10 * abs(parts.sum())


def test_overlapped_diag_block_builds_the_same_inverse_by_row_blocks():
    """Specification of potrf_diag_ovl_kernel: W[i,j] = -W_ii sum_k L[i,k] W[k,j], row block by row block, with the
    inverse parked in the strictly-upper 32x32 blocks of the tile the factorisation works in."""
    A = _spd(128, seed=4)
    L, Li, ld, info = br.diag_block_ovl(np.asfortranarray(A))
    L0, Li0, ld0, info0 = br.diag_block(np.asfortranarray(A))
    Lref = np.linalg.cholesky(A)
    assert info == 0 and info0 == 0
    assert np.allclose(L, Lref, rtol=1e-12, atol=1e-12)
    assert np.all(np.triu(L, 1) == 0) and np.all(np.triu(Li, 1) == 0)
    assert np.allclose(Li @ Lref, np.eye(128), atol=1e-10)
    assert np.allclose(Li, Li0, rtol=1e-9, atol=1e-12)
    assert abs(ld - ld0) < 1e-12 * abs(ld0)
    B = A.copy()
    B[70, 70] = -3.0
    assert br.diag_block_ovl(np.asfortranarray(B))[3] == 71


@pytest.mark.parametrize("K", [128, 512])
def test_chain_products_as_sixteen_independent_blocks(K):
    rng = np.random.default_rng(K)
    A = np.asfortranarray(rng.standard_normal((128, K))); B = np.asfortranarray(rng.standard_normal((128, K)))
    C = np.asfortranarray(rng.standard_normal((128, 128)))
    got = C.copy(order="F")
    br.small_nt(got, A, B, K, mode=1, tri=True)
    lo = np.tril(np.ones((128, 128), bool))
    assert np.allclose(got[lo], (C - A @ B.T)[lo], rtol=1e-12, atol=1e-12)
    assert np.array_equal(got[~lo], C[~lo])


def test_chain_head_pair_through_the_scratch_tile():
    """X = A W' -> scratch; C -= X X' (lower) from the scratch tile; X copied home by block column 0 of the update."""
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.standard_normal((128, 128))); W = np.asfortranarray(np.tril(rng.standard_normal((128, 128))))
    C = np.asfortranarray(rng.standard_normal((128, 128)))
    a, c = A.copy(order="F"), C.copy(order="F")
    br.chain_head_pair(a, W, c)
    X = A @ W.T
    lo = np.tril(np.ones((128, 128), bool))
    assert np.allclose(a, X, rtol=1e-12, atol=1e-12)
    assert np.allclose(c[lo], (C - X @ X.T)[lo], rtol=1e-12, atol=1e-11)
    assert np.array_equal(c[~lo], C[~lo])


@pytest.mark.parametrize("n,W,W1,W2B,headk", [(1100, 2, 0, None, 2), (1500, 3, 6, 4, 1), (1500, 3, 6, 4, 2), (900, 2, 4, 3, 1)])
def test_chain_through_small_products_and_overlapped_diag_gives_the_same_factor(n, W, W1, W2B, headk):
    """api.cu with GPK_POTRF_HEADK / GPK_DIAG_OVL: head products through the scratch tile, the hand-over tile in 32x32
    blocks, the diagonal kernel with the row-block inverse - same factor, same forward solve."""
    A0 = _spd(n, seed=n + headk)
    rng = np.random.default_rng(n)
    b0 = rng.standard_normal(n)
    A, b = br.pad_spd(A0), np.concatenate([b0, np.zeros(-n % 128)])
    Dinv, parts, info, z = br.potrf_device(A, b.copy(), W=W, W1=W1, w1_minrem=2, W2B=W2B, headk=headk, ovl=True)
    Lref = np.linalg.cholesky(A0)
    assert info == 0
    assert np.allclose(np.tril(A)[:n, :n], Lref, rtol=1e-10, atol=1e-10)
    assert np.allclose(z[:n], sla.solve_triangular(Lref, b0, lower=True), rtol=1e-9, atol=1e-9)
    assert abs(parts.sum() - np.log(np.diag(Lref)).sum()) < 1e-10 * abs(parts.sum())
