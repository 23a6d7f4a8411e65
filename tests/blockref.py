"""numpy model of the DEVICE algorithms (block level), for CPU tests of the host logic.

Every function mirrors one launcher/kernel of pygps_b200/csrc with the same
argument meaning, so the orchestration in api.cu (look-ahead Cholesky, the
transposed multi-right-hand-side sweeps, the inverse through U = L^-T, the
in-block inversion of potrf_diag_kernel) can be checked against scipy without a
GPU.  Matrices are Fortran-ordered numpy arrays == the device's column-major.
"""
import numpy as np

NB = 128
IB = 32


def gemm_nt(mode, C, A, B, K, tiles_m, tiles_n, tri=0, ti_off=0, tj_off=0):
    """dgemm_nt_kernel: C(ti,tj) = A(ti)·B(tj)' (mode 0) or C -= A·B' (mode 1).
    A, B, C are views whose [0,0] is the tile-grid origin (like the device pointers)."""
    for tj in range(tiles_n):
        for ti in range(tiles_m):
            gi, gj = ti + ti_off, tj + tj_off
            if tri and gi < gj:
                continue
            kbeg = gi * NB if tri == 2 else 0
            a = A[ti * NB:(ti + 1) * NB, kbeg:K]
            b = B[tj * NB:(tj + 1) * NB, kbeg:K]
            p = a @ b.T
            c = C[ti * NB:(ti + 1) * NB, tj * NB:(tj + 1) * NB]
            new = p if mode == 0 else c - p
            if tri and gi == gj:
                mask = np.tril(np.ones((NB, NB), dtype=bool))
                c[mask] = new[mask]
            else:
                c[...] = new


def rcp_newton(d):
    """The diagonal kernel's reciprocal: a 20-bit seed (MUFU.RCP64H reads only the upper word of d; modelled by
    rounding 1/d to 20 mantissa bits) and ONE cubic Newton step x0 (1 + e + e^2), e = 1 - d x0: relative error e^3."""
    with np.errstate(all="ignore"):
        m, ex = np.frexp(1.0 / d)
        x0 = np.ldexp(np.round(m * 2.0 ** 21) / 2.0 ** 21, ex)
        e = 1.0 - d * x0
        return x0 + x0 * (e * e + e)


def ldl_block(a):
    """One 32x32 diagonal sub-block as potrf_diag_kernel's warp 0 factors it: square-root-free elimination (row r's
    multiplier is a[r,j] / d_j), the unscaled column kept until the end, then L[:,c] = column / sqrt(d_c).
    Returns (L lower, sum log diag L, 1-based index of the first non-positive pivot or 0)."""
    n = a.shape[0]
    a = a.copy()
    d = np.ones(n)
    bad = 0
    for j in range(n):
        d[j] = a[j, j]
        if not d[j] > 0 and bad == 0:
            bad = j + 1
        a[:j, j] = 0.0
        w = a[:, j] * rcp_newton(d[j])
        for c in range(j + 1, n):
            a[:, c] -= w * a[c, j]
    with np.errstate(all="ignore"):
        rs = 1.0 / np.sqrt(d)
        L = np.tril(a * rs[None, :])
        L[np.arange(n), np.arange(n)] = d * rs
        lg = 0.5 * np.sum(np.log(d))
    return L, lg, bad


def diag_block(Ablk):
    """potrf_diag_kernel: returns (L, inv(L), sum log diag, info) for one 128x128 block,
    following the kernel's 32-wide inner blocking and its in-place inversion order."""
    S = np.tril(Ablk).copy(order="F")
    T = []
    logdet = 0.0
    info = 0
    nblk = NB // IB
    for jb in range(nblk):
        j0 = jb * IB
        a = S[j0:j0 + IB, j0:j0 + IB].copy()
        a, lg, bad = ldl_block(a)
        logdet += lg
        if bad and info == 0:
            info = j0 + bad
        a = np.tril(a)
        S[j0:j0 + IB, j0:j0 + IB] = a
        W = np.zeros((IB, IB))
        for col in range(IB):          # per-lane forward substitution
            for i in range(IB):
                s = (1.0 if i == col else 0.0) - a[i, :i] @ W[:i, col]
                W[i, col] = s / a[i, i]
        T.append(W)
        if jb < nblk - 1:
            r0 = j0 + IB
            S[r0:, j0:j0 + IB] = S[r0:, j0:j0 + IB] @ W.T
            P = S[r0:, j0:j0 + IB]
            upd = P @ P.T
            nrb = nblk - 1 - jb
            for ib in range(nrb):
                for cb in range(ib + 1):
                    S[r0 + ib * IB:r0 + (ib + 1) * IB, r0 + cb * IB:r0 + (cb + 1) * IB] -= \
                        upd[ib * IB:(ib + 1) * IB, cb * IB:(cb + 1) * IB]
    L = np.tril(S).copy(order="F")
    # phase 3: in-place inverse, last block column first
    S = np.tril(S)                      # the kernel never reads the garbage above sub-block diagonals
    for jb in range(nblk - 1, -1, -1):
        j0 = jb * IB
        r0 = j0 + IB
        if jb < nblk - 1:
            Y = S[r0:, j0:j0 + IB] @ T[jb]
            S[r0:, j0:j0 + IB] = -(np.tril(S[r0:, r0:]) @ Y)
        S[j0:j0 + IB, j0:j0 + IB] = T[jb]
    return L, S.copy(order="F"), logdet, info


def diag_block_ovl(Ablk):
    """potrf_diag_ovl_kernel: the same factorisation with the block inverse built by ROW blocks in the shadow of the
    32x32 factorisations.  `S` is the kernel's shared-memory tile: its strictly-upper 32x32 blocks, unused by the
    factorisation, hold the inverse - block (j, i) of S is W[i, j] - and `Zs` holds Z[i, j] = sum_{k=j}^{i-1} L[i,k] W[k,j]
    of the row block in progress; W[i, j] = -W_ii Z[i, j].  Steps S1-S3 are what warps 1-7 do while warp 0 factors
    diagonal block i.  Returns (L, inv(L), sum log diag, info)."""
    S = np.tril(Ablk).copy(order="F")
    nblk = NB // IB
    T = [None] * nblk
    Zs = [None] * (nblk - 1)
    logdet = 0.0
    info = 0

    def blk(bi, bj):
        return S[bi * IB:(bi + 1) * IB, bj * IB:(bj + 1) * IB]

    def inv32(b):                       # inv32_warp: lane r solves x L' = e_r, right-looking; x = column r of W
        Lb = np.tril(blk(b, b))
        W = np.zeros((IB, IB))
        for r in range(IB):
            x = np.zeros(IB)
            x[r] = 1.0
            for k in range(IB):
                x[k] = x[k] * (1.0 / Lb[k, k])
                x[k + 1:] -= x[k] * Lb[k + 1:, k]
            W[:, r] = x
        return W

    for jb in range(nblk):
        j0 = jb * IB
        # warp 0: LDL' of diagonal block jb
        a, lg, bad = ldl_block(blk(jb, jb).copy())
        # warps 1-7, meanwhile (jb >= 1): nothing below touches block (jb, jb)
        if jb >= 1:
            i = jb
            T[i - 1] = inv32(i - 1)                                     # S1
            for j in range(i - 1):                                      # S2: row block i-1 of the inverse
                blk(j, i - 1)[...] = -(T[i - 1] @ Zs[j])
            for j in range(i):                                          # S3: Z of row block i
                z = blk(i, j) @ T[j]
                for k in range(j + 1, i):
                    z = z + blk(i, k) @ blk(j, k)
                Zs[j] = z
        logdet += lg
        if bad and info == 0:
            info = j0 + bad
        blk(jb, jb)[...] = np.tril(a)
        if jb < nblk - 1:
            r0 = j0 + IB
            Ljj = np.tril(blk(jb, jb))
            S[r0:, j0:j0 + IB] = np.linalg.solve(Ljj, S[r0:, j0:j0 + IB].T).T      # sub-panel solve X Ljj' = Y
            P = S[r0:, j0:j0 + IB]
            upd = P @ P.T
            nrb = nblk - 1 - jb
            for ib in range(nrb):
                for cb in range(ib + 1):
                    S[r0 + ib * IB:r0 + (ib + 1) * IB, r0 + cb * IB:r0 + (cb + 1) * IB] -= \
                        upd[ib * IB:(ib + 1) * IB, cb * IB:(cb + 1) * IB]
    L = np.tril(S).copy(order="F")
    last = nblk - 1
    T[last] = inv32(last)
    for j in range(last):
        blk(j, last)[...] = -(T[last] @ Zs[j])
    Dinv = np.zeros((NB, NB), order="F")
    for bi in range(nblk):
        for bj in range(bi + 1):
            Dinv[bi * IB:(bi + 1) * IB, bj * IB:(bj + 1) * IB] = T[bi] if bi == bj else blk(bj, bi)
    return L, Dinv, logdet, info


def small_nt(C, A, B, K, mode, tri=False, copy_dst=None):
    """small_nt_kernel: one 128x128 tile of C (op)= A(128 x K) B(128 x K)' as sixteen independent 32x32 blocks, each
    walking the contraction in chunks of 128; with copy_dst the CTAs of block column 0 also store their 32 x 128 strip
    of A there (K == 128).  Every block reads ONLY A, B and its own block of C - asserted by working on snapshots."""
    A0, B0, C0 = A.copy(), B.copy(), C.copy()
    for bj in range(4):
        for bi in range(4):
            if tri and bj > bi:
                continue
            acc = np.zeros((32, 32))
            for c0 in range(0, K, 128):
                acc += A0[32 * bi:32 * bi + 32, c0:c0 + 128] @ B0[32 * bj:32 * bj + 32, c0:c0 + 128].T
            new = acc if mode == 0 else C0[32 * bi:32 * bi + 32, 32 * bj:32 * bj + 32] - acc
            cb = C[32 * bi:32 * bi + 32, 32 * bj:32 * bj + 32]
            if tri and bi == bj:
                m = np.tril(np.ones((32, 32), dtype=bool))
                cb[m] = new[m]
            else:
                cb[...] = new
            if copy_dst is not None and bj == 0:
                assert K == 128
                copy_dst[32 * bi:32 * bi + 32, :] = A0[32 * bi:32 * bi + 32, :128]


def chain_head_pair(tile_p1p, Dinv_p, tile_p1p1):
    """The two products behind diag(p) on the panel stream (api.cu, small_heads): the solved head tile goes to a
    scratch tile (sixteen CTAs cannot work in place), the update of the next diagonal tile reads the scratch tile as
    both operands and copies it home."""
    scratch = np.zeros((NB, NB), order="F")
    small_nt(scratch, tile_p1p, Dinv_p, NB, mode=0)
    small_nt(tile_p1p1, scratch, scratch, NB, mode=1, tri=True, copy_dst=tile_p1p)


def trsv_fwd(A, Dinv, b, z, k, T):
    zk = Dinv[k] @ b[k * NB:(k + 1) * NB]
    z[k * NB:(k + 1) * NB] = zk
    for i in range(1, T - k):
        r = (k + i) * NB
        b[r:r + NB] -= A[r:r + NB, k * NB:(k + 1) * NB] @ zk


def trsv_bwd(A, Dinv, z, x, k, T):
    xk = Dinv[k].T @ z[k * NB:(k + 1) * NB]
    x[k * NB:(k + 1) * NB] = xk
    for j in range(k):
        z[j * NB:(j + 1) * NB] -= A[k * NB:(k + 1) * NB, j * NB:(j + 1) * NB].T @ xk


def oz_decode(idx, nt, jb0, jb1, band=16, ti_min=0):
    """ozaki.cu oz_decode: CTA index -> (128-row tile, 64-column tile) of the tile set
    {jb0 <= jb < jb1, ti >= max(jb, ti_min)}, rasterised in bands of `band` row tiles (L2 reuse of the operand slices).
    ti_min = 0: lower triangle (SYRK); ti_min >= jb1: the rectangle of a stacked-operand product."""
    r_lo = max(jb0, ti_min)
    while True:
        r_hi = min(r_lo + band, nt)
        rows = r_hi - r_lo
        nfull = max(min(jb1, r_lo + 1) - jb0, 0)
        cnt_full = 2 * rows * nfull
        t0, t1 = max(jb0, r_lo + 1), min(jb1, r_hi)
        n_ = max(t1 - t0, 0)
        cnt_tri = 2 * n_ * r_hi - (t0 + t1 - 1) * n_
        if idx < cnt_full + cnt_tri or r_hi >= nt:
            if idx < cnt_full:
                return r_lo + idx % rows, 2 * jb0 + idx // rows
            idx -= cnt_full
            jb = t0
            while jb < t1 - 1 and idx >= 2 * (r_hi - jb):
                idx -= 2 * (r_hi - jb)
                jb += 1
            c = r_hi - jb
            return jb + idx % c, 2 * jb + idx // c
        idx -= cnt_full + cnt_tri
        r_lo = r_hi


def oz_ntiles(nt, jb0, jb1, ti_min=0):
    """launch_oz_ex: number of 128x64 tiles (= CTAs) of a launch."""
    return sum(2 * (nt - max(jb, ti_min)) for jb in range(jb0, jb1))


def oz_gemm_stacked(C, A, B, S=7, RB=8):
    """launch_oz_gemm_stacked: C (rows of A x rows of B) -= A B' through the SYRK machinery on the operands stacked
    as [B; A]: one split (row scales per stacked row), tiles {ti >= nb/128, jb < nb/128}."""
    nb = B.shape[0]
    P = np.vstack([B, A])
    full = np.zeros((P.shape[0], P.shape[0]))
    oz_syrk(full, P, S, RB)                       # full -= P P'
    C += full[nb:, :nb]


def oz_split(P, S=7, RB=8):
    """ozaki.cu oz_slice_kernel: per-row exponent, then balanced radix-2^RB digits (all int8).
    Returns (digits [S, n, k] int8, row scale 2^(e_i-RB))."""
    m = np.abs(P).max(axis=1)
    _, ex = np.frexp(m)                      # m = f * 2^ex, f in [0.5, 1)  ->  ilogb(m) = ex - 1
    e = np.where(m > 0, ex + 1, 0).astype(np.int64)
    e = e + ((m > 0) & (np.ldexp(m, -e) > 0.48))
    e = np.maximum(e, -500)
    x = np.ldexp(P, -e[:, None])             # exact
    q = np.rint(x * 2.0 ** (S * RB)).astype(np.int64)
    d = np.empty((S,) + P.shape, dtype=np.int8)
    for t in range(S - 1, -1, -1):
        lo = q & ((1 << RB) - 1)
        dd = np.where(lo >= (1 << (RB - 1)), lo - (1 << RB), lo)
        d[t] = dd
        q = (q - dd) >> RB
    assert (q == 0).all(), "carry out of the leading digit"
    return d, np.ldexp(1.0, e - RB)


def oz_syrk(C, P, S=7, RB=8):
    """ozaki.cu oz_syrk_kernel: lower(C) -= P P' through exact integer products of the digit slices; the S
    accumulators G_m = sum_{t+u=m} d_t d_u' are int32 on the device (checked here), Horner in fp64."""
    d, sc = oz_split(P, S, RB)
    di = d.astype(np.int64)
    acc = np.zeros(C.shape)
    for m in range(S - 1, -1, -1):
        G = sum(di[t] @ di[m - t].T for t in range(m + 1))
        assert np.abs(G).max() < 2 ** 31
        acc = acc * 2.0 ** -RB + G
    upd = (acc * sc[:, None]) * sc[None, :]
    L = np.tril_indices(C.shape[0])
    C[L] -= upd[L]


def oz_split_fixed(P, e, S=7, RB=8):
    """Digits of P with GIVEN row exponents e (|P[i,:]| <= 2^(e_i-1)): the fixed-scale slicing of DESIGN section 8 item 2."""
    x = np.ldexp(P, -e[:, None])
    assert np.abs(x).max() <= 0.5
    q = np.rint(x * 2.0 ** (S * RB)).astype(np.int64)
    d = np.empty((S,) + P.shape, dtype=np.int8)
    for t in range(S - 1, -1, -1):
        lo = q & ((1 << RB) - 1)
        dd = np.where(lo >= (1 << (RB - 1)), lo - (1 << RB), lo)
        d[t] = dd
        q = (q - dd) >> RB
    assert (q == 0).all()
    return d


def oz_update_from_digits(C, dr, dc, sr, sc_, S=7, RB=8, lower=True):
    """C (rows x cols) -= (sum of digit products) scaled: dr [S, rows, k], dc [S, cols, k] digits, sr / sc_ row scales."""
    acc = np.zeros(C.shape)
    for m in range(S - 1, -1, -1):
        G = sum(dr[t].astype(np.int64) @ dc[m - t].astype(np.int64).T for t in range(m + 1))
        assert np.abs(G).max() < 2 ** 31
        acc = acc * 2.0 ** -RB + G
    upd = (acc * sr[:, None]) * sc_[None, :]
    if lower:
        C -= np.tril(upd)
    else:
        C -= upd


def potrf_device(A, b=None, W=2, W1=0, w1_minrem=0, oz=False, split=True, W2B=None, fixed_scale=False, headk=0, ovl=False):
    """api.cu potrf_device (stream order flattened): three-level blocking - level-1 blocks of W1 panels (W1=0: same as
    the sub-blocks), sub-blocks of W panels, single panels.  oz: trailing updates through oz_syrk.
    fixed_scale (with oz; NOT on the device yet - the executable specification of DESIGN section 8 item 2): row i is
    sliced with the scale 2^ceil(log2 sqrt(A_ii)) known before the factorisation (|L_ij| <= sqrt(A_ii)), every sub-block's
    panels are sliced ONCE when they are finished, and the same digits serve the level-2 update of the block's other
    columns and the level-1 update of the trailing matrix (no per-update slicing, no row-maximum pass).
    A: padded (np,np) F-order, lower triangle valid; the factor overwrites it.
    Returns Dinv list, logdet parts, info, z."""
    np_ = A.shape[0]
    T = np_ // NB
    fix_e = None
    if fixed_scale:
        _, ex = np.frexp(np.sqrt(np.diag(A).copy()))
        fix_e = (ex + 1).astype(np.int64)
    digits = {}                          # (k0, k1) -> digits [S, rows >= k1*NB, (k1-k0)*NB] of a finished sub-block
    Dinv = [None] * T
    parts = np.zeros(T)
    info = 0
    z = np.zeros(np_) if b is not None else None
    W1 = (W1 // W) * W
    W2B = W if W2B is None else W2B      # small (chain-bound) blocks are ONE sub-block of W2B panels
    bstart, bsub, bbig = [], [], []
    pos = 0
    while pos < T:
        big = W1 > W and T - (pos + W1) >= w1_minrem
        bstart.append(pos)
        bsub.append(W if big else W2B)
        bbig.append(big)
        pos = min(pos + (W1 if big else W2B), T)
    bstart.append(T)

    def update(rows0, cols_end, k0, k1):
        """lower tiles of A[rows0:, rows0:cols_end] -= A[rows0:, k0:k1] A[rows0:cols_end, k0:k1]' (tile units)"""
        pan = A[rows0 * NB:, k0 * NB:k1 * NB]
        Ct = A[rows0 * NB:, rows0 * NB:]
        ncols = cols_end - rows0
        if oz and fixed_scale:
            # digits of every finished sub-block inside [k0, k1), rows >= rows0*NB (a suffix of what was sliced)
            parts_ = []
            for (a0, a1), d in sorted(digits.items()):
                if a0 >= k0 and a1 <= k1:
                    parts_.append(d[:, (rows0 - a1) * NB:, :])
            d = np.concatenate(parts_, axis=2)
            assert d.shape[2] == (k1 - k0) * NB
            scl = np.ldexp(1.0, fix_e[rows0 * NB:] - 8)
            for tj in range(ncols):
                cs_ = slice(tj * NB, (tj + 1) * NB)
                rs_ = slice(tj * NB, None)
                blk = Ct[rs_, cs_]
                oz_update_from_digits(blk, d[:, rs_, :], d[:, cs_, :], scl[rs_], scl[cs_], lower=False)
                # the diagonal tile only keeps its lower triangle (the strict upper part of A is never read)
        elif oz:
            full = Ct.copy()
            oz_syrk(full, pan)
            for tj in range(ncols):
                for ti in range(tj, T - rows0):
                    blk = (slice(ti * NB, (ti + 1) * NB), slice(tj * NB, (tj + 1) * NB))
                    if ti == tj:
                        mask = np.tril(np.ones((NB, NB), dtype=bool))
                        Ct[blk][mask] = full[blk][mask]
                    else:
                        Ct[blk] = full[blk]
        else:
            gemm_nt(1, Ct, pan, pan, (k1 - k0) * NB, T - rows0, ncols, tri=1)

    head_l1 = False
    for j in range(len(bstart) - 1):
        pb, pe = bstart[j], bstart[j + 1]
        w2 = bsub[j]
        for sb in range(pb, pe, w2):
            se = min(sb + w2, pe)
            for p in range(sb, se):
                s = slice(p * NB, (p + 1) * NB)
                L, Li, ld, inf_p = (diag_block_ovl if ovl else diag_block)(A[s, s])
                A[s, s] = L
                Dinv[p], parts[p] = Li, ld
                if inf_p and not info:
                    info = p * NB + inf_p
                rem = T - p - 1
                inner = se - p - 1
                if not split or inner == 0 or rem < 2:
                    if rem > 0:
                        pan = A[(p + 1) * NB:, s]
                        gemm_nt(0, pan, pan.copy(), Li, NB, rem, 1)
                    if inner > 0:
                        pan = A[(p + 1) * NB:, s]
                        gemm_nt(1, A[(p + 1) * NB:, (p + 1) * NB:], pan, pan, NB, rem, inner, tri=1)
                else:
                    # head (panel stream): tile (p+1,p) solved, tile (p+1,p+1) updated
                    h = slice((p + 1) * NB, (p + 2) * NB)
                    if headk == 2 or (headk == 1 and not bbig[j]):
                        chain_head_pair(A[h, s], Li, A[h, h])          # small_nt_kernel x 2 through the scratch tile
                    else:
                        gemm_nt(0, A[h, s], A[h, s].copy(), Li, NB, 1, 1)
                        gemm_nt(1, A[h, h], A[h, s], A[h, s], NB, 1, 1, tri=1)
                    # tail (s_tail): rows p+2.. solved; column p+1 below its diagonal tile; columns p+2..se-1
                    t0 = (p + 2) * NB
                    pan = A[t0:, s]
                    gemm_nt(0, pan, pan.copy(), Li, NB, rem - 1, 1)
                    gemm_nt(1, A[t0:, h], pan, A[h, s], NB, rem - 1, 1)
                    if inner > 1:
                        gemm_nt(1, A[t0:, t0:], pan, pan, NB, rem - 1, inner - 1, tri=1)
                if b is not None:
                    trsv_fwd(A, Dinv, b, z, p, T)
            if oz and fixed_scale and se < T:
                digits[(sb, se)] = oz_split_fixed(A[se * NB:, sb * NB:se * NB], fix_e[se * NB:])
            if se < pe:
                update(se, pe, sb, se)          # level-2: the rest of the block's columns
        rem = T - pe
        if rem > 0:
            head_l1 = split and not bbig[j]
            if head_l1:
                # the next diagonal tile is updated on the panel stream; the level-1 update leaves it alone
                h = slice(pe * NB, (pe + 1) * NB)
                pan0 = A[h, pb * NB:pe * NB]
                saved = A[h, h].copy()
                if headk:
                    small_nt(A[h, h], pan0, pan0, (pe - pb) * NB, mode=1, tri=True)
                else:
                    gemm_nt(1, A[h, h], pan0, pan0, (pe - pb) * NB, 1, 1, tri=1)
                mine = A[h, h].copy()
                A[h, h] = saved
                update(pe, T, pb, pe)
                A[h, h] = mine                  # (device: oz_syrk skip00 / the split DMMA launches never touch the tile)
            else:
                update(pe, T, pb, pe)           # level-1: the whole trailing matrix (device: next block's columns first)
            digits.clear()
    return Dinv, parts, info, z


def sweep_forward(P, A, Dinv, T):
    """api.cu sweep_forward: P (rows x np) <- P · L^-T."""
    rt = P.shape[0] // NB
    for k in range(T):
        blk = P[:, k * NB:(k + 1) * NB]
        gemm_nt(0, blk, blk.copy(), Dinv[k], NB, rt, 1)
        if k + 1 < T:
            gemm_nt(1, P[:, (k + 1) * NB:], blk, A[(k + 1) * NB:, k * NB:(k + 1) * NB], NB, rt, T - k - 1)


def inverse_factor_T(A, Dinv):
    """api.cu inverse_factor_T: U = L^-T, touching only tiles on/above the diagonal."""
    np_ = A.shape[0]
    T = np_ // NB
    U = np.asfortranarray(np.eye(np_))
    for k in range(T):
        blk = U[:(k + 1) * NB, k * NB:(k + 1) * NB]
        gemm_nt(0, blk, blk.copy(), Dinv[k], NB, k + 1, 1)
        if k + 1 < T:
            gemm_nt(1, U[:(k + 1) * NB, (k + 1) * NB:], blk, A[(k + 1) * NB:, k * NB:(k + 1) * NB],
                    NB, k + 1, T - k - 1)
    return U


def inverse_lower(U):
    """Ainv (lower tiles) = U·U' with the trapezoidal contraction start (tri=2)."""
    np_ = U.shape[0]
    T = np_ // NB
    W = np.asfortranarray(np.full((np_, np_), np.nan))
    gemm_nt(0, W, U, U, np_, T, T, tri=2)
    return W


def pad_spd(Amat):
    n = Amat.shape[0]
    np_ = (n + NB - 1) // NB * NB
    P = np.asfortranarray(np.eye(np_))
    P[:n, :n] = Amat
    return P


def oz_decode_rect(idx, nt, ncols, band=16):
    """ozaki.cu oz_decode_rect: CTA index -> (row tile, virtual 64-column tile) over the full nt x ncols rectangle."""
    r_lo = 0
    while True:
        rows = min(band, nt - r_lo)
        cnt = 2 * rows * ncols
        if idx < cnt or r_lo + rows >= nt:
            return r_lo + idx % rows, idx // rows
        idx -= cnt
        r_lo += rows


def oz_cyclic_tiles(nt, ncols, cfirst, cs):
    """launch_oz_cyclic / oz_tile: the tiles a block-cyclic launch computes: (row tile, GLOBAL 64-column tile, LOCAL
    64-column tile) for every CTA whose row tile is not above its global column block."""
    out = []
    for idx in range(2 * nt * ncols):
        ti, tjv = oz_decode_rect(idx, nt, ncols)
        jb = cfirst + (tjv >> 1) * cs
        if ti >= jb:
            out.append((ti, 2 * jb + (tjv & 1), tjv))
    return out


def potrf_dist_blocked(A, G, WD):
    """dist.cu, blocked variant (GPK_DIST_OZAKI=1), all G ranks simulated in one process: block columns are owned
    cyclically (column j by rank j % G, packed locally), panels are factored by their owner and "broadcast"; inside a
    block of WD panels only the block's own columns get the immediate rank-128 updates, after the block every rank applies
    ONE rank-(WD*128) update to the columns it owns beyond the block, tile by tile as launch_oz_cyclic enumerates them.
    A: padded (np,np) F-order, lower triangle valid.  Returns the factor (np,np) assembled from the ranks' columns."""
    np_ = A.shape[0]
    T = np_ // NB
    loc = [np.asfortranarray(A[:, [c for j in range(r, T, G) for c in range(j * NB, (j + 1) * NB)]]) for r in range(G)]

    def col(r, j):                                  # rank r's packed storage of global block column j
        lj = j // G
        return loc[r][:, lj * NB:(lj + 1) * NB]

    for kb in range(0, T, WD):
        ke = min(kb + WD, T)
        blk = np.zeros((np_, (ke - kb) * NB))       # the block buffer (every rank holds the same copy after the broadcasts)
        for k in range(kb, ke):
            o = k % G
            ck = col(o, k)
            s = slice(k * NB, (k + 1) * NB)
            L, Li, _, info = diag_block(ck[s, :])
            assert info == 0
            ck[s, :] = L
            ck[(k + 1) * NB:, :] = ck[(k + 1) * NB:, :] @ Li.T
            pan = ck[(k + 1) * NB:, :].copy()       # the broadcast
            blk[(k + 1) * NB:, (k - kb) * NB:(k - kb + 1) * NB] = pan
            for r in range(G):                      # immediate updates: owned columns j in (k, ke)
                for j in range(k + 1, ke):
                    if j % G != r:
                        continue
                    pj = pan[(j - k - 1) * NB:(j - k) * NB, :]
                    cj = col(r, j)
                    upd = pan[(j - k - 1) * NB:, :] @ pj.T
                    rows = slice(j * NB, np_)
                    tile = cj[rows, :]
                    tile[NB:, :] -= upd[NB:, :]
                    tile[:NB, :] -= np.tril(upd[:NB, :])
        if ke < T:
            nt = T - ke                             # (the device has one more row tile: the y - m row of the augmented matrix)
            P = blk[ke * NB:, :]
            full = P @ P.T
            for r in range(G):
                jf = ke + ((r - ke) % G + G) % G
                if jf >= T:
                    continue
                ncols = (T - 1 - jf) // G + 1
                C = loc[r][ke * NB:, (jf // G) * NB:]          # rows from the slice origin, local columns from jf's
                for ti, tjg, tjl in oz_cyclic_tiles(nt, ncols, jf - ke, G):
                    rs, cg, cl = slice(ti * NB, (ti + 1) * NB), slice(tjg * 64, (tjg + 1) * 64), slice(tjl * 64, (tjl + 1) * 64)
                    u = full[rs, cg]
                    if ti * NB < (tjg + 1) * 64:               # tile touches the diagonal: rows >= columns only
                        gi = np.arange(ti * NB, (ti + 1) * NB)[:, None]
                        gj = np.arange(tjg * 64, (tjg + 1) * 64)[None, :]
                        u = np.where(gi >= gj, u, 0.0)
                    C[rs, cl] -= u
    out = np.zeros_like(A)
    for r in range(G):
        for lj, j in enumerate(range(r, T, G)):
            out[:, j * NB:(j + 1) * NB] = loc[r][:, lj * NB:(lj + 1) * NB]
    return np.tril(out)
