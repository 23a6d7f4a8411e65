"""Kernel-level parity on the B200: each CUDA kernel through the C ABI against numpy/scipy."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from pygps_b200 import _lib, build
    build.build()
    return _lib.Engine(0)


def _report(name, got, ref):
    err = np.abs(got - ref)
    i = np.unravel_index(np.nanargmax(err), err.shape)
    return "%s: max abs err %.3e at %s (got %r ref %r), nan=%d" % (
        name, err[i], i, got[i], ref[i], int(np.isnan(got).sum()))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 384, 128), (512, 128, 256)])
def test_gemm_nt_full(eng, mode, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K + mode)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
    got = eng.dbg_gemm_nt(mode, A, B, C)
    ref = A @ B.T if mode == 0 else C - A @ B.T
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-11), _report("gemm mode %d" % mode, got, ref)


def test_gemm_nt_lower_only_leaves_upper_untouched(eng):
    n, K = 384, 128
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, K)); C = rng.standard_normal((n, n))
    got = eng.dbg_gemm_nt(2, A, A, C)
    ref = C - A @ A.T
    lo = np.tril(np.ones((n, n), bool))
    assert np.allclose(got[lo], ref[lo], rtol=1e-12, atol=1e-11), _report("syrk lower", np.where(lo, got, 0), np.where(lo, ref, 0))
    assert np.array_equal(got[~lo], C[~lo]), "strict upper triangle was written"


def test_gemm_nt_trapezoid(eng):
    n = 384
    rng = np.random.default_rng(6)
    U = np.triu(rng.standard_normal((n, n)))
    C = np.full((n, n), 7.0)
    got = eng.dbg_gemm_nt(3, U, U, C)
    ref = U @ U.T
    lo = np.tril(np.ones((n, n), bool))
    assert np.allclose(got[lo], ref[lo], rtol=1e-12, atol=1e-11), _report("U U^T", np.where(lo, got, 0), np.where(lo, ref, 0))


def test_gemm_nt_in_place_is_race_free(eng):
    """Regression: the panel TRSM overwrites its own A operand.  With column-split tiles two CTAs shared a
    row tile and one could overwrite columns the other was still reading (seen as a wrong nlZ at N=16384)."""
    rng = np.random.default_rng(11)
    M = 128 * 96
    A = rng.standard_normal((M, 128)); B = np.tril(rng.standard_normal((128, 128)))
    for _ in range(3):
        got = eng.dbg_gemm_nt(4, A, B, np.zeros((M, 128)))
        assert np.allclose(got, A @ B.T, rtol=1e-12, atol=1e-11), _report("in-place A*B^T", got, A @ B.T)


@pytest.mark.parametrize("breg", ["0", "1"])
@pytest.mark.parametrize("K", [128, 256, 512, 1024])
def test_chain_products_match_numpy(eng, monkeypatch, K, breg):
    """small_nt_kernel (sixteen 32x32-block CTAs per tile): C = A B' and the lower-block C -= A B'; for K = 128 with
    both operands in shared memory and with the B operand in registers."""
    monkeypatch.setenv("GPK_SMALL_BREG", breg)
    rng = np.random.default_rng(K)
    A = rng.standard_normal((128, K)); B = rng.standard_normal((128, K)); C = rng.standard_normal((128, 128))
    got = eng.dbg_gemm_nt(5, A, B, C)
    assert np.allclose(got, A @ B.T, rtol=1e-12, atol=1e-11), _report("chain product, set", got, A @ B.T)
    got = eng.dbg_gemm_nt(6, A, B, C)
    ref = C - A @ B.T
    lo = np.tril(np.ones((128, 128), bool))
    assert np.allclose(got[lo], ref[lo], rtol=1e-12, atol=1e-11), _report("chain product, update", np.where(lo, got, 0), np.where(lo, ref, 0))
    assert np.array_equal(got[~lo], C[~lo]), "strict upper triangle was written"


@pytest.mark.parametrize("breg", ["0", "1"])
def test_chain_head_pair_solves_updates_and_copies_home(eng, monkeypatch, breg):
    """The two products behind every diagonal block: X = A W' (W lower triangular), C -= X X' (lower), A <- X."""
    monkeypatch.setenv("GPK_SMALL_BREG", breg)
    rng = np.random.default_rng(77)
    A = rng.standard_normal((128, 128)); W = np.tril(rng.standard_normal((128, 128))); C = rng.standard_normal((128, 128))
    for _ in range(3):
        gotC, gotA = eng.dbg_gemm_nt(7, A, W, C)
        X = A @ W.T
        ref = C - X @ X.T
        lo = np.tril(np.ones((128, 128), bool))
        assert np.allclose(gotA, X, rtol=1e-12, atol=1e-11), _report("head tile", gotA, X)
        assert np.allclose(gotC[lo], ref[lo], rtol=1e-12, atol=1e-10), _report("head update", np.where(lo, gotC, 0), np.where(lo, ref, 0))
        assert np.array_equal(gotC[~lo], C[~lo]), "strict upper triangle was written"


def _spd128(seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((128, 3))
    d = ((X[:, None] - X[None]) ** 2).sum(-1)
    return np.exp(-0.5 * d / 4.0) / 0.01 + np.eye(128)


@pytest.mark.parametrize("ovl", ["0", "1", "2"])
def test_diag_block_factor_and_inverse(eng, monkeypatch, ovl):
    """ovl 0: inverse after the factorisation; 1: inverse by row blocks under the 32x32 factorisations; 2: as 1, every
    finished piece stored at once by cp.async.bulk."""
    monkeypatch.setenv("GPK_DIAG_OVL", ovl)
    A = _spd128(1)
    L, Li, ld, info = eng.dbg_diag(A)
    Lref = np.linalg.cholesky(A)
    assert info == 0
    assert np.allclose(L, Lref, rtol=1e-10, atol=1e-10), _report("diag L", L, Lref)
    assert np.all(np.triu(L, 1) == 0) and np.all(np.triu(Li, 1) == 0)
    assert np.allclose(Li @ Lref, np.eye(128), atol=1e-8), _report("Linv*L", Li @ Lref, np.eye(128))
    assert abs(ld - np.log(np.diag(Lref)).sum()) < 1e-10 * abs(ld)


@pytest.mark.parametrize("ovl", ["0", "1", "2"])
def test_diag_block_flags_first_bad_pivot(eng, monkeypatch, ovl):
    monkeypatch.setenv("GPK_DIAG_OVL", ovl)
    A = _spd128(2)
    A[70, 70] = -3.0
    _, _, _, info = eng.dbg_diag(A)
    assert info == 71


def test_cov_matrices_against_reference_vectors(eng, golden):
    from pygps_b200 import _lib
    g = golden("cov_vectors")
    x, z = g["x"], g["z"]
    kinds = {"rbf": (_lib.COV_RBF, 3), "ard": (_lib.COV_RBFARD, 3), "mat1": (_lib.COV_MATERN, 1),
             "mat3": (_lib.COV_MATERN, 3), "mat5": (_lib.COV_MATERN, 5), "mat7": (_lib.COV_MATERN, 7)}
    for name, (kind, d) in kinds.items():
        hyp = g[name + "_hyp"]
        for mode, key, args in (("train", "_train", (x, None)), ("cross", "_cross", (x, z)), ("self_test", "_self", (None, z))):
            got = eng.cov_matrix(kind, d, hyp, args[0], args[1], mode)
            ref = g[name + key]
            assert got.shape == ref.shape
            assert np.allclose(got, ref, rtol=1e-12, atol=1e-14), _report(name + key, got, ref)
        K = eng.cov_matrix(kind, d, hyp, x, None, "train")
        assert np.array_equal(K, K.T), "train matrix must be bit-symmetric like cdist's"
        assert np.all(np.diag(K) == np.exp(2 * hyp[-1] if kind != _lib.COV_RBFARD else 2 * hyp[-1]))
    for name in ("rbf", "ard"):
        kind, d = kinds[name]
        hyp = g[name + "_hyp"]
        for i in range(len(hyp)):
            got = eng.cov_matrix(kind, d, hyp, x, None, "train", i)
            assert np.allclose(got, g["%s_dtrain%d" % (name, i)], rtol=1e-12, atol=1e-14), _report("d%s%d" % (name, i), got, g["%s_dtrain%d" % (name, i)])
            got = eng.cov_matrix(kind, d, hyp, x, z, "cross", i)
            assert np.allclose(got, g["%s_dcross%d" % (name, i)], rtol=1e-12, atol=1e-14)


def test_cov_matrix_ragged_sizes(eng):
    from pygps_b200 import _lib
    from oracle import gp_oracle as go
    rng = np.random.default_rng(3)
    for n, m, D in ((1, 1, 1), (65, 130, 5), (200, 63, 17), (129, 1, 33)):
        x = rng.standard_normal((n, D)); z = rng.standard_normal((m, D))
        hyp = [0.3, -0.2]
        got = eng.cov_matrix(_lib.COV_RBF, 3, hyp, x, z, "cross")
        ref = go.cov_matrix(("rbf", hyp), x=x, z=z, mode="cross")
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-15), _report("ragged %s" % ((n, m, D),), got, ref)


@pytest.mark.parametrize("n", [20, 128, 300, 1000])
def test_potrf_potrs_match_lapack(eng, n):
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 4))
    d = ((X[:, None] - X[None]) ** 2).sum(-1)
    A = np.exp(-0.5 * d / 4.0) / 0.01 + np.eye(n)
    R, ld = eng.potrf(A)
    Rref = np.linalg.cholesky(A).T
    assert np.all(np.tril(R, -1) == 0), "strict lower triangle of the upper factor must be exactly zero"
    assert np.allclose(R, Rref, rtol=1e-9, atol=1e-9), _report("potrf n=%d" % n, R, Rref)
    assert abs(ld - np.log(np.diag(Rref)).sum()) < 1e-10 * max(1.0, abs(ld))
    B = rng.standard_normal((n, 3))
    Xs = eng.potrs(B)
    ref = sla.cho_solve((Rref, False), B)
    assert np.allclose(Xs, ref, rtol=1e-7, atol=1e-9), _report("potrs n=%d" % n, Xs, ref)


@pytest.mark.parametrize("headk,ovl", [("0", "0"), ("1", "1"), ("2", "1"), ("3", "2"), ("4", "2")])
def test_chain_variants_give_the_same_factor(eng, monkeypatch, headk, ovl):
    """The dependent chain of the blocked factorisation through every combination of its kernels (GPK_POTRF_HEADK:
    strip kernel / small_nt_kernel with B in shared memory / in registers, in the small blocks / everywhere;
    GPK_DIAG_OVL: the three diagonal-block kernels): same factor as LAPACK, N large enough for split panels."""
    n = 1500
    rng = np.random.default_rng(15)
    X = rng.standard_normal((n, 3))
    d = ((X[:, None] - X[None]) ** 2).sum(-1)
    A = np.exp(-0.5 * d / 2.0) / 0.04 + np.eye(n)
    monkeypatch.setenv("GPK_POTRF_HEADK", headk)
    monkeypatch.setenv("GPK_DIAG_OVL", ovl)
    R, ld = eng.potrf(A)
    Rref = np.linalg.cholesky(A).T
    assert np.all(np.tril(R, -1) == 0)
    assert np.allclose(R, Rref, rtol=1e-9, atol=1e-9), _report("potrf headk=%s ovl=%s" % (headk, ovl), R, Rref)
    assert abs(ld - np.log(np.diag(Rref)).sum()) < 1e-10 * abs(ld)


def test_potrf_not_positive_definite_raises(eng):
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        eng.potrf(A)
    A = np.ones((140, 140)) + 1e-3 * np.eye(140)      # PD diagonal, indefinite after elimination? no: rank-1 + eps, PD
    A[100, 130] = A[130, 100] = 5.0                   # breaks positive definiteness with positive diagonal
    with pytest.raises(np.linalg.LinAlgError) as ei:
        eng.potrf(A)
    assert "even with jitter" in str(ei.value)


def test_tools_jitchol_solve_chol_drop_in(eng, golden):
    import pygps_b200 as pg
    g = golden("cov_vectors")
    L = pg.tools.jitchol(g["chol_A"])
    assert np.allclose(L, g["chol_L"], rtol=1e-11, atol=1e-12)
    X = pg.tools.solve_chol(L.T, g["chol_B"])
    assert np.allclose(X, g["chol_X"], rtol=1e-8, atol=1e-10)
    X2 = pg.tools.solve_chol(np.array(L.T), g["chol_B"])          # factor not resident: still correct
    assert np.allclose(X2, g["chol_X"], rtol=1e-8, atol=1e-10)
    with pytest.raises(Exception):
        pg.tools.solve_chol(L.T, np.zeros((3, 1)))


@pytest.mark.parametrize("N,K", [(64, 32), (64, 128), (128, 64), (256, 256)])
def test_tcgen05_int8_tile_is_exact(eng, N, K):
    """tcgen05.mma.kind::i8 through hand-built shared-memory/instruction descriptors and a TMEM accumulator:
    exact int32 result (the primitive of the int8 emulation of the fp64 trailing update)."""
    rng = np.random.default_rng(N + K)
    A = rng.integers(-64, 65, size=(128, K), dtype=np.int8)
    B = rng.integers(-64, 65, size=(N, K), dtype=np.int8)
    C = eng.dbg_i8_tile(A, B)
    assert np.array_equal(C, A.astype(np.int32) @ B.astype(np.int32).T)


@pytest.mark.parametrize("a_tmem", [1, 2])
@pytest.mark.parametrize("N,K", [(64, 32), (64, 256), (128, 64)])
def test_tcgen05_int8_a_operand_through_tmem(eng, N, K, a_tmem):
    """A staged shared memory -> TMEM by tcgen05.cp.128x256b and read from there by the MMA (mode 2: one TMEM
    region reused by every k-step, the production pattern of oz_syrk_kernel)."""
    rng = np.random.default_rng(N * K + a_tmem)
    A = rng.integers(-128, 128, size=(128, K)).astype(np.int8)
    B = rng.integers(-128, 128, size=(N, K)).astype(np.int8)
    C = eng.dbg_i8_tile(A, B, a_tmem=a_tmem)
    assert np.array_equal(C, A.astype(np.int32) @ B.astype(np.int32).T)


def test_tcgen05_int8_mixed_signedness(eng):
    rng = np.random.default_rng(5)
    A = rng.integers(0, 256, size=(128, 64)).astype(np.uint8)
    B = rng.integers(-128, 128, size=(64, 64)).astype(np.int8)
    assert np.array_equal(eng.dbg_i8_tile(A, B), A.astype(np.int32) @ B.astype(np.int32).T)
    assert np.array_equal(eng.dbg_i8_tile(B[:, :32].repeat(2, 0), A[:64, :32]),
                          B[:, :32].repeat(2, 0).astype(np.int32) @ A[:64, :32].astype(np.int32).T)


@pytest.mark.parametrize("n,kw", [(128, 128), (384, 256), (1024, 384), (2048, 1152)])
def test_int8_sliced_trailing_update_matches_fp64(eng, n, kw):
    """ozaki.cu: lower(C) -= P P' on the int8 tensor cores (7 balanced radix-256 slices, exact int32 accumulation in
    TMEM) against numpy fp64 and against the DMMA kernel; rows of very different scale; upper triangle untouched."""
    rng = np.random.default_rng(n + kw)
    P = rng.standard_normal((n, kw)) * np.exp(rng.uniform(-8, 8, size=(n, 1)))
    P[3, :] = 0.0
    C = rng.standard_normal((n, n)); C = C + C.T
    ref = C - P @ P.T
    den = np.abs(P) @ np.abs(P).T + np.abs(C)
    out, _ = eng.dbg_oz_syrk(P, C, mode=0)
    dm, _ = eng.dbg_oz_syrk(P, C, mode=1)
    L = np.tril_indices(n)
    # numpy's own fp64 dot carries ~sqrt(kw)*u relative to |P||P|'; the sliced product itself is exact to 2^-56
    assert np.max(np.abs(out[L] - ref[L]) / den[L]) < 2e-14
    assert np.max(np.abs(out[L] - dm[L]) / den[L]) < 2e-14
    assert np.array_equal(np.triu(out, 1), np.triu(C, 1))


def test_int8_update_propagates_nan_and_inf(eng):
    """include/gpk.h: "NaN/Inf propagate into outputs".  The int8 slicing cannot represent a non-finite entry, so a
    row that holds one gets a NaN scale: every entry of C in that row and column comes out NaN, exactly the footprint
    a NaN has in the fp64 product; every other entry still equals the fp64 result."""
    n, kw = 512, 256
    rng = np.random.default_rng(77)
    P = rng.standard_normal((n, kw))
    P[37, 5] = np.nan
    P[300, 100] = np.inf
    P[411, 255] = -np.inf
    C = rng.standard_normal((n, n)); C = C + C.T
    out, _ = eng.dbg_oz_syrk(P, C, mode=0)
    bad = np.zeros(n, bool); bad[[37, 300, 411]] = True
    hit = bad[:, None] | bad[None, :]
    lo = np.tril(np.ones((n, n), bool))
    assert not np.isfinite(out[lo & hit]).any(), "a non-finite panel entry was silently dropped by the slicing"
    Pz = np.where(np.isfinite(P), P, 0.0)
    ref = C - Pz @ Pz.T
    ok = lo & ~hit
    assert np.isfinite(out[ok]).all()
    assert np.allclose(out[ok], ref[ok], rtol=1e-12, atol=1e-12)


def test_nan_reaches_the_factor_through_the_int8_path_at_n4096(eng):
    """N=4096: the level-1 trailing updates run on the int8 tensor cores (oz_slice / oz_syrk).  A NaN in an OFF-diagonal
    entry of A travels L[3000,17] -> (sliced panel row 3000) -> C[3000,3000] -> pivot 3001: info > 0, LinAlgError - what
    LAPACK's dpotrf reports.  Before the slicing carried NaN the row was sliced as zeros and the factorisation
    "succeeded" with garbage."""
    n = 4096
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 8))
    from pygps_b200 import _lib
    K = eng.cov_matrix(_lib.COV_RBF, 3, [np.log(2.0), 0.0], X, None, "train")
    A = K / 0.01 + np.eye(n)
    R, _ = eng.potrf(A)                                   # sane matrix: factorises
    assert np.isfinite(R).all()
    assert eng.stats()["launches"] > 0
    A[3000, 17] = A[17, 3000] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        eng.potrf(A)
    A[3000, 17] = A[17, 3000] = np.inf
    with pytest.raises(np.linalg.LinAlgError):
        eng.potrf(A)
    # NaN in a target: the matrix is fine, nlZ and alpha become NaN (no abort, no exception)
    y = np.sin(X.sum(1)); y[2500] = np.nan
    eng.set_data(X)
    nlZ, alpha, _, _ = eng.exact_eval(_lib.COV_RBF, 3, [np.log(2.0), 0.0], np.log(0.1), y, False)
    assert np.isnan(nlZ) and np.isnan(alpha).any()


def test_solve_chol_only_trusts_the_exact_resident_factor(eng):
    """a slice / copy / edited copy of jitchol's result must not be solved with the resident factor."""
    import pygps_b200 as pg
    rng = np.random.default_rng(4)
    n, k = 300, 170
    X = rng.standard_normal((n, 3))
    A = np.exp(-0.5 * ((X[:, None] - X[None]) ** 2).sum(-1)) + 0.1 * np.eye(n)
    B = rng.standard_normal((n, 4))
    L = pg.tools.jitchol(A)
    assert not L.flags.writeable
    ref = sla.cho_solve((np.asarray(L), True), B)
    assert np.allclose(pg.tools.solve_chol(L.T, B), ref, rtol=1e-8, atol=1e-10)           # resident
    # leading principal sub-factor: the factor of A[:k,:k]; must be solved as a k x k system
    Xk = pg.tools.solve_chol(L[:k, :k].T, B[:k])
    assert Xk.shape == (k, 4)
    assert np.allclose(Xk, np.linalg.solve(A[:k, :k], B[:k]), rtol=1e-7, atol=1e-9)
    # an edited copy is a different matrix
    L2 = np.array(L) * 2.0
    assert np.allclose(pg.tools.solve_chol(L2.T, B), ref / 4.0, rtol=1e-8, atol=1e-10)
    # and the engine itself refuses right-hand sides of the wrong height
    eng.potrf(A)
    with pytest.raises(Exception):
        eng.potrs(B[:k])


@pytest.mark.parametrize("n,nrhs", [(130, 1), (700, 300), (1500, 1500)])
def test_set_factor_and_sweeps_match_lapack(eng, n, nrhs):
    """gpk_set_factor (upload R, block inverses in one launch) + gpk_potrs as two triangular sweeps."""
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 4))
    A = np.exp(-0.5 * ((X[:, None] - X[None]) ** 2).sum(-1) / 4.0) / 0.05 + np.eye(n)
    R = np.linalg.cholesky(A).T
    ld = eng.set_factor(np.ascontiguousarray(R))
    assert abs(ld - np.log(np.diag(R)).sum()) < 1e-10 * max(1.0, abs(ld))
    B = rng.standard_normal((n, nrhs))
    ref = sla.cho_solve((R, False), B)
    for _ in range(2):                                   # second call reuses the cached transpose
        got = eng.potrs(B)
        assert np.allclose(got, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max()), _report("potrs", got, ref)
    Rbad = R.copy(); Rbad[5, 5] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        eng.set_factor(Rbad)


@pytest.mark.parametrize("n,m,D", [(1, 1, 1), (130, 77, 5), (300, 300, 8), (515, 129, 32), (257, 1000, 16)])
def test_cov_tile_kernel_matches_generic_kernel_and_numpy(eng, monkeypatch, n, m, D):
    """cov_tile_kernel (128x128 tiles, inputs by cp.async.bulk, inline table-based exp) against the generic kernel
    (libdevice exp) and numpy, all kinds, train and cross modes, ragged sizes; exact diagonal, bit-symmetric."""
    from pygps_b200 import _lib
    from oracle import gp_oracle as go
    rng = np.random.default_rng(n + m + D)
    x = rng.standard_normal((n, D)) * 1.5
    z = rng.standard_normal((m, D)) * 1.5
    specs = [(_lib.COV_RBF, 3, [0.3, -0.2], ("rbf", [0.3, -0.2])),
             (_lib.COV_RBFARD, 3, list(rng.uniform(-0.3, 0.6, D)) + [0.1], None)]
    specs[1] = (specs[1][0], 3, specs[1][2], ("rbfard", specs[1][2]))
    for d in (1, 3, 5, 7):
        specs.append((_lib.COV_MATERN, d, [0.4, 0.2], ("matern", [0.4, 0.2], d)))
    for kind, md, hyp, ospec in specs:
        for mode in ("train", "cross"):
            monkeypatch.setenv("GPK_COV_TILE", "1")
            fast = eng.cov_matrix(kind, md, hyp, x, z if mode == "cross" else None, mode)
            monkeypatch.setenv("GPK_COV_TILE", "0")
            slow = eng.cov_matrix(kind, md, hyp, x, z if mode == "cross" else None, mode)
            ref = go.cov_matrix(ospec, x=x, z=z if mode == "cross" else None, mode=mode)
            assert fast.shape == ref.shape
            assert np.allclose(fast, slow, rtol=2e-15, atol=1e-300), _report("tile vs generic %s" % (ospec,), fast, slow)
            assert np.allclose(fast, ref, rtol=1e-12, atol=1e-15), _report("tile vs numpy %s" % (ospec,), fast, ref)
            if mode == "train":
                assert np.array_equal(fast, fast.T), "K must be bit-symmetric"
                assert np.all(np.diag(fast) == np.exp(2 * hyp[-1])), "exact diagonal sf2"
    monkeypatch.delenv("GPK_COV_TILE")
