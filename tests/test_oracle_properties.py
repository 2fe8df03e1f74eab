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


@settings(max_examples=10, deadline=None)
@given(st.integers(40, 300), st.sampled_from([5, 15, 20]), st.integers(0, 10**6))
def test_covariance_oracle_vs_numpy_eigh(n, k, seed):
    """calculate_covariances (fast_gicp_impl.hpp:241-297): neighbour covariance (divisor k) and the PLANE regularisation
    V diag(1, 1, 1e-3) V^T, against numpy.linalg.eigh on the same neighbour sets"""
    rng = np.random.default_rng(seed)
    pts = (rng.normal(0, 1, (n, 3)) * np.array([3.0, 2.0, 0.2])).astype(np.float32)
    idx, _ = oracle.knn(pts, pts, k)
    raw = oracle.covariances(pts, k, oracle.REG_NONE)
    reg = oracle.covariances(pts, k, oracle.REG_PLANE)
    for i in range(0, n, 7):
        nb = pts[idx[i]].astype(np.float64)
        mean = nb.sum(0) / k
        C = (nb - mean).T @ (nb - mean) / k
        assert np.abs(raw[i] - C).max() < 1e-9 * max(1.0, np.abs(C).max())
        w, V = np.linalg.eigh(C)                                   # ascending
        if w[1] - w[0] > 1e-6 * w[2] and w[2] - w[1] > 1e-9 * w[2]:   # a well separated smallest eigenvalue -> unique plane normal
            nrm = V[:, 0]
            exp = np.eye(3) - (1.0 - 1e-3) * np.outer(nrm, nrm)
            assert np.abs(reg[i] - exp).max() < 1e-6


@settings(max_examples=6, deadline=None)
@given(st.integers(0, 10**6))
def test_gicp_oracle_recovers_a_rigid_transform(seed):
    """round trip: a structured cloud and its rigidly moved copy -> align returns the motion (the 1e-4 bar of the GPU parity tests)"""
    from mr_slam_b200 import synth
    rng = np.random.default_rng(seed)
    src = synth.make_scan(synth.pose_matrix(0.0, 0.0, 0.0, 0.0), "vlp16", scan_id=int(rng.integers(0, 1000)))[::4].astype(np.float32)
    T = synth.pose_matrix(*rng.normal(0, 0.15, 3), rng.normal(0, 0.03)).astype(np.float64)
    tgt = (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    g = oracle.FastGICP()
    g.set_max_correspondence_distance(2.0)
    g.set_input_source(src)
    g.set_input_target(tgt)
    got = g.align(np.eye(4, dtype=np.float32))
    assert np.abs(got - T).max() < 2e-3          # float32 clouds: the copy is rounded, so not exactly 1e-4


def test_elevation_oracle_kalman_halves_variance():
    """G_fuse recurrence (gpu_process.cu:477-535): two identical measurements in one cell -> same height, half the variance; a far
    outlier above the cell replaces it, one below is ignored"""
    L, res = 20, 0.5
    m = oracle.ElevationMap(L, res)
    idx = np.array([5 * L + 7] * 2, np.int32)
    z = np.zeros(2, np.int32)
    h = np.array([1.25, 1.25], np.float32)
    v = np.array([0.04, 0.04], np.float32)
    m.fuse(idx, z, z, z, np.zeros(2, np.float32), h, v)
    g = m.get_map()
    assert g["elevation"][5 * L + 7] == np.float32(1.25) and abs(g["variance"][5 * L + 7] - 0.02) < 1e-7
    m.fuse(idx[:1], z[:1], z[:1], z[:1], np.zeros(1, np.float32), np.array([9.0], np.float32), np.array([0.01], np.float32))
    assert m.get_map()["elevation"][5 * L + 7] == np.float32(9.0)           # Mahalanobis > 5 and higher: replaces
    m.fuse(idx[:1], z[:1], z[:1], z[:1], np.zeros(1, np.float32), np.array([-3.0], np.float32), np.array([0.01], np.float32))
    assert m.get_map()["elevation"][5 * L + 7] == np.float32(9.0)           # Mahalanobis > 5 and lower: ignored
    assert (np.delete(m.get_map()["variance"], 5 * L + 7) == np.float32(1e-4)).all()   # every other cell: the 1e-4 floor
