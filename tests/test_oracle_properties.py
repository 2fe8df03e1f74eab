"""Property tests that pin the CPU oracle (the parity anchor of the GPU tests) against brute-force numpy statements of the same
definitions, on small random inputs including the awkward ones (duplicate points, ties, empty cells, odd sizes)."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle


def _cloud(rng, n, dup):
    pts = rng.normal(0, 3, (n, 3)).astype(np.float32)
    if dup and n > 4:                                   # exact duplicates and a lattice patch: distance ties
        pts[n // 2:n // 2 + 2] = pts[0]
        pts[-3:] = np.round(pts[-3:])
    return pts


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 150), st.integers(1, 32), st.booleans(), st.integers(0, 10**6))
def test_knn_oracle_is_exact_with_index_tie_break(n, k, dup, seed):
    """kd-tree kNN of the oracle == brute force under (squared distance computed as fma(dz,dz,fma(dy,dy,dx*dx)), index) order"""
    rng = np.random.default_rng(seed)
    pts = _cloud(rng, n, dup)
    q = pts if seed % 2 == 0 else _cloud(rng, max(1, n // 3), False)
    idx, d2 = oracle.knn(pts, q, k)
    d = (q[:, None, :] - pts[None, :, :]).astype(np.float32)
    # the oracle's distance: dx*dx rounded, then two float fmas -- reproduce with float64 intermediates of exact products
    dx2 = (d[..., 0].astype(np.float64) * d[..., 0].astype(np.float64)).astype(np.float32)
    t = (d[..., 1].astype(np.float64) * d[..., 1].astype(np.float64) + dx2.astype(np.float64)).astype(np.float32)
    ref = (d[..., 2].astype(np.float64) * d[..., 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    kk = min(k, n)
    for i in range(len(q)):
        order = np.lexsort((np.arange(n), ref[i]))[:kk]
        assert np.array_equal(idx[i, :kk], order)
        assert np.array_equal(d2[i, :kk], ref[i, order])
        assert (idx[i, kk:] == -1).all()


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 4), st.integers(0, 400), st.sampled_from([0.05, 0.2, 0.5, 1.3]), st.integers(0, 10**6))
def test_keyframe_merge_oracle_vs_numpy_statement(nkf, n, leaf, seed):
    """transform (se3 form) + crop + VoxelGrid centroids, restated with plain numpy float32 scalars"""
    rng = np.random.default_rng(seed)
    kfs = [np.hstack([rng.normal(0, 6, (n, 3)), rng.random((n, 1))]).astype(np.float32) for _ in range(nkf)]
    Ts = []
    for _ in range(nkf):
        a = rng.normal(0, 0.3)
        T = np.eye(4, dtype=np.float32)
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        T[:3, 3] = rng.normal(0, 2, 3)
        Ts.append(T)
    lim = 9.0
    got = oracle.merge_nearest_keyframes(kfs, np.stack(Ts), lim, leaf)
    f = np.float32
    pts = []
    for kf, T in zip(kfs, Ts):
        for p in kf:
            x, y, z = f(p[0]), f(p[1]), f(p[2])
            q = [f(T[r, 0] * x) + (f(T[r, 1] * y) + (f(T[r, 2] * z) + T[r, 3])) for r in range(3)]
            q = [f(v) for v in q]
            if all(np.isfinite(q)) and -lim <= q[0] <= lim and -lim <= q[1] <= lim:
                pts.append(q + [f(p[3])])
    if not pts:
        assert len(got) == 0
        return
    P = np.array(pts, np.float32)
    inv = f(1.0) / f(leaf)
    mn, mx = P[:, :3].min(0), P[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    div = np.floor(mx * inv).astype(np.int64) - min_b + 1
    ijk = (np.floor(P[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    exp = []
    for kv in np.unique(key):
        s = np.zeros(4, np.float32)
        members = np.nonzero(key == kv)[0]
        for j in members:                                # ascending input order, float32 running sums
            s = (s + P[j]).astype(np.float32)
        exp.append(s / f(len(members)))
    assert np.array_equal(got, np.array(exp, np.float32))


@settings(max_examples=15, deadline=None)
@given(st.sampled_from([(4, 6), (8, 12), (40, 120), (5, 9)]), st.integers(0, 10**6))
def test_disco_rel_ori_oracle_vs_numpy(shape, seed):
    H, W = shape
    rng = np.random.default_rng(seed)
    ra, ia, rb, ib = (rng.normal(0, 1, (H, W)).astype(np.float32) for _ in range(4))
    ang, loc = oracle.disco_rel_ori(ra, ia, rb, ib)
    cross = (ra * rb + ia * ib).astype(np.float64) + 1j * (ra * ib + rb * ia).astype(np.float64)
    back = (np.fft.ifft2(cross) * H * W).real.astype(np.float32)
    best = back.max()
    assert back.reshape(-1)[loc] >= best - 1e-4 * max(1.0, abs(best))          # same peak (ties to float rounding)
    assert ang == (loc % W) * 3.0


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 2), st.sampled_from([(6, 8), (12, 16), (120, 120)]), st.integers(-7, 7), st.integers(0, 10**6))
def test_solve_translation_row_shifts_vs_numpy_fft(C, shape, row_shift, seed):
    """b[i] = W/2 - argmax(fftshift(sum_c |ifft(fft(q) conj(fft(p)))|)) with the row roll folded in (util.py:404-415)"""
    H, W = shape
    rng = np.random.default_rng(seed)
    q = (rng.random((C, H, W)) * (rng.random((C, H, W)) < 0.4)).astype(np.float32)
    p = (rng.random((C, H, W)) * (rng.random((C, H, W)) < 0.4)).astype(np.float32)
    _, _, _, b = oracle.ring_solve_translation(q, p, 0.3, row_shift)
    prolled = np.roll(p, row_shift, axis=1)
    for i in range(H):
        corr = np.fft.ifft(np.fft.fft(q[:, i, :].astype(np.float64), axis=-1, norm="ortho") *
                           np.conj(np.fft.fft(prolled[:, i, :].astype(np.float64), axis=-1, norm="ortho")), axis=-1, norm="ortho")
        c = np.fft.fftshift(np.abs(corr), axes=-1).sum(0)       # fftshift over the batch axis too is a no-op for the sum
        am = int(np.argmax(c.astype(np.float32)))
        if c[am] - np.partition(c, -2)[-2] > 1e-6 * max(1.0, c[am]):       # unique peak: must agree exactly
            assert b[i] == W // 2 - am
