"""CPU tests: the oracle against the reference's own fixtures / known answers / code (tests/golden), and its internal
consistency.  These pin the checker before the CUDA path is compared with it."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, rel_l2
from mr_slam_b200 import synth


def test_kdtree_is_exact():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(3000, 3)).astype(np.float32)
    pts[100:110] = pts[90:100]  # exact duplicates: ties must resolve to the lowest index
    q = np.concatenate([pts[:500], rng.normal(size=(200, 3)).astype(np.float32)])
    for k in (1, 5, 15, 20):
        i1, d1 = oracle.knn(pts, q, k)
        i2, d2 = oracle.knn(pts, q, k, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # fewer points than k
    i1, d1 = oracle.knn(pts[:7], q[:10], 15)
    assert (i1[:, 7:] == -1).all() and np.isinf(d1[:, 7:]).all()


def test_gicp_reference_fixture_tolerances(gicp_golden):
    """FG/src/test/gicp_test.cpp:147-201: forward / backward alignments of the reference's own PCD pair must land within
    t < 0.05 m, r < 1 deg of data/relative.txt and converge.  Library defaults, identity guess."""
    rel = gicp_golden["relative"]
    for name, s, t in (("forward", "source", "target"), ("backward", "target", "source")):
        g = oracle.FastGICP()
        g.set_num_threads(4)
        g.set_input_target(gicp_golden[t])
        g.set_input_source(gicp_golden[s])
        g.align()
        r = g.result()
        T = r["Td"] if name == "forward" else np.linalg.inv(r["Td"])
        te, re = synth.pose_error(T, rel)
        assert te < 0.05 and re < np.deg2rad(1.0) and r["converged"]
        # frozen known answer of the oracle itself (guards against silent drift of the checker)
        assert np.abs(r["Td"] - gicp_golden[name + "_T"]).max() < 1e-9


def test_gicp_readme_known_answer(readme_golden):
    """FG/README.md:110-113: fitness_score of fgicp_mt 0.204412 / fgicp_st 0.204379 on the same pair after
    ApproximateVoxelGrid(0.1)."""
    g = oracle.FastGICP()
    g.set_num_threads(4)
    g.set_input_target(readme_golden["target"])
    g.set_input_source(readme_golden["source"])
    g.align()
    fit = g.get_fitness_score()
    assert abs(fit - 0.204412) < 5e-4
    assert abs(fit - float(readme_golden["fitness"])) < 1e-9


def test_gicp_swap_source_target(gicp_golden):
    """gicp_test.cpp:171-200 "swap and set source/target" cases."""
    rel = gicp_golden["relative"]
    g = oracle.FastGICP()
    g.set_num_threads(4)
    g.set_input_source(gicp_golden["target"])
    g.swap_source_and_target()
    g.set_input_source(gicp_golden["source"])
    g.align()
    te, re = synth.pose_error(g.result()["Td"], rel)
    assert te < 0.05 and re < np.deg2rad(1.0)


def test_polar_oracle_equals_reference_cpu_twin():
    """Oracle vs outputs of the reference's unmodified CPU polar voxeliser on the reference's own 1.bin / 2.bin fixtures."""
    z = np.load(os.path.join(GOLDEN, "polar_fixture.npz"))
    for name in ("disco1", "disco2"):
        pts = z[name + "_points"]
        soa = np.ascontiguousarray(pts.T.reshape(-1))
        for ci, c in enumerate(z["cfgs"]):
            o = oracle.polar(soa, len(pts), *[int(v) for v in c])
            exp = np.zeros_like(o)
            exp[z["%s_cfg%d_idx" % (name, ci)]] = z["%s_cfg%d_val" % (name, ci)]
            assert np.array_equal(o, exp)


def test_polar_cpu_twin_live():
    """Same comparison with the compiled reference twin itself when oracle/_ref is present (authoring container / GPU box)."""
    from oracle import ref
    if not ref.available("polar_cpu"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(4)
    pts = rng.uniform(-0.99, 0.99, (20000, 3)).astype(np.float32)
    soa = np.ascontiguousarray(pts.T.reshape(-1))
    for c in ((1, 1, 40, 120, 20, 1), (1, 1, 40, 120, 4, 3)):
        o, idx, nd = oracle.polar(soa, len(pts), *c, want_idx=True)
        assert nd == 0
        assert np.array_equal(o, ref.polar_cpu(soa, len(pts), *c))


def test_voxelocc_oracle_semantics():
    """manager.cu:64-85 restated in plain python on a small case (sequential scatter in point order)."""
    rng = np.random.default_rng(1)
    n, nx, ny, nh = 3000, 12, 10, 3
    pts = rng.uniform(-1.1, 1.1, (n, 3)).astype(np.float32)
    pts[5] = 0.0
    soa = np.ascontiguousarray(pts.T.reshape(-1))
    out, idx, nd = oracle.voxelocc(soa, n, 1, 1, nx, ny, nh, 1, want_idx=True)
    exp = np.zeros(nx * ny * nh * 3, np.float32)
    maxh = np.zeros(nx * ny)
    gx, gy, gh = (np.float32(2.0 * 1.0 / nx), np.float32(2.0 * 1.0 / ny), np.float32(2.0 * 1.0 / nh))
    for i in range(n):
        v = pts[i].copy()
        v[v == 0] = np.float32(0.0001)
        v[v > 1] = np.float32(0.9999)
        v[v < -1] = np.float32(-0.9999)
        ix, iy, ih = (int(np.floor((np.float64(v[0]) + 1.0) / np.float64(gx))), int(np.floor((np.float64(v[1]) + 1.0) / np.float64(gy))),
                      int(np.floor((np.float64(v[2]) + 1.0) / np.float64(gh))))
        assert (ix, iy, ih) == (idx[i], idx[i + n], idx[i + 2 * n])
        if not (0 <= ix < nx and 0 <= iy < ny and 0 <= ih < nh):
            continue
        slot = 3 * (iy + ix * ny + ih * ny * nx)
        exp[slot], exp[slot + 1] = pts[i, 0], pts[i, 1]
        if maxh[iy + ix * ny] < pts[i, 2]:
            exp[slot + 2] = pts[i, 2]
            maxh[iy + ix * ny] = pts[i, 2]
    assert np.array_equal(out, exp)


def test_elevation_oracle_matches_naive_gfuse():
    """G_fuse restated literally (per cell, loop over ALL points, gpu_process.cu:477-535) on a tiny case."""
    rng = np.random.default_rng(2)
    L, n = 6, 400
    om = oracle.ElevationMap(L, 0.5)
    idx = rng.integers(-1, L * L, n).astype(np.int32)
    h = rng.normal(0, 1.0, n).astype(np.float32)
    h[rng.random(n) < 0.05] = -1.0
    v = rng.uniform(1e-5, 0.05, n).astype(np.float32)
    R = (rng.integers(0, 3, n)).astype(np.int32)
    G = np.ones(n, np.int32)
    B = np.ones(n, np.int32)
    I = rng.random(n).astype(np.float32)
    om.fuse(idx, R, G, B, I, h, v)
    m = om.get_map()
    e = np.full(L * L, -10, np.float32)
    var = np.full(L * L, -10, np.float32)
    inten = np.zeros(L * L, np.float32)
    for c in range(L * L):
        for i in range(n):
            if idx[i] != c or h[i] == -1:
                continue
            col = R[i] != 0 and G[i] != 0 and B[i] != 0 and I[i] != 0
            if e[c] == -10:
                e[c], var[c] = h[i], v[i]
                if col:
                    inten[c] = I[i]
            else:
                md = np.float32(abs(np.float32(h[i] - e[c]))) / np.float32(np.sqrt(var[c]))
                if md > 5:
                    if e[c] < h[i]:
                        e[c], var[c] = h[i], v[i]
                        if col:
                            inten[c] = I[i]
                else:
                    ne = np.float32(np.float32(var[c] * h[i]) + np.float32(v[i] * e[c])) / np.float32(var[c] + v[i])
                    nv = np.float32(v[i] * var[c]) / np.float32(v[i] + var[c])
                    e[c], var[c] = ne, nv
                    if col:
                        inten[c] = I[i]
        if var[c] < 0.0001:
            var[c] = 0.0001
    assert np.allclose(m["elevation"], e, rtol=0, atol=2e-6)  # the naive loop has no FMA contraction: last-bit differences only
    assert np.allclose(m["variance"], var, rtol=1e-6, atol=0)
    assert np.array_equal(m["intensity"], inten)
    assert (m["variance"] >= 0.0001).all()  # empty cells are clamped from -10 to 1e-4, like the reference


def test_radon_oracle_vs_reference_analytic():
    """torch-radon's own test (tests/test_parallel_beam.py:48-70): forward projection of a discretised random phantom vs the
    analytic line integrals of the reference's symbolic.cpp, tolerance 2e-3*(512/A)*(512/D)."""
    from oracle import ref
    if not ref.available("symbolic"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    H = W = D = 128
    A = 128
    f = ref.Symbolic(H, W)
    for _ in range(30):
        w, cx, cy = rng.uniform(0, 1), rng.uniform(-W / 2, W / 2), rng.uniform(-H / 2, H / 2)
        rx, ry = rng.uniform(W / 32, W / 4), rng.uniform(H / 32, H / 4)
        if rng.integers(0, 2) < 0.25:
            f.add_ellipse(w, cx, cy, rx, ry)
        else:
            f.add_gaussian(w, cx, cy, rx, ry)
    img = f.discretize()
    angles = np.linspace(0, 2 * np.pi, A, endpoint=False).astype(np.float32)
    tol = 2e-3 * (512 / A) * (512 / D)
    sino_ref = f.forward(angles, D)
    for wb in (0, 8):
        sino = oracle.radon_forward(img[None], angles, D, 1.0, wb)[0]
        assert rel_l2(sino_ref, sino) < tol


def test_ring_tail_oracle_vs_reference_python():
    """Oracle vs golden vectors produced by the reference's own python (util.py:197-198 and fast_corr, :362-374)."""
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    for i in range(2):
        _, tir = oracle.ring_tiring(z["sino"][i])
        assert np.abs(tir - z["tiring"][i]).max() / np.abs(z["tiring"][i]).max() < 1e-5
    for i in range(2):
        for j in range(2):
            d, a, _ = oracle.ring_fast_corr(z["tiring"][i], z["tiring"][j])
            assert abs(d - z["fast_corr_dist"][i, j]) < 1e-5 and a == z["fast_corr_angle"][i, j]
    # load_pc_infer restatement used by the bench
    got = synth.ring_preprocess(z["pc"])
    assert np.array_equal(got, z["pc_infer"])  # the reference keeps all columns of the surviving rows


def test_solve_translation_oracle_vs_reference_python():
    """solve_translation (util.py:388-423) goldens produced by the reference's own function.  The per-row shifts b are bit-equal in
    all cases.  x, y: the reference multiplies by v.t() where torch.svd's V itself is needed (util.py:506); that is harmless only
    when V happens to be symmetric (cases 0 and 2) -- in case 1 LAPACK returned a rotation and the reference's own answer has a
    residual of 108 against 3.0 for the least-squares solution.  The oracle (and the CUDA path) return the least-squares solution."""
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    for ci in range(3):
        rot, rs = z["st%d_params" % ci]
        x, y, e, b = oracle.ring_solve_translation(z["st%d_query" % ci], z["st%d_positive" % ci], float(rot), int(rs))
        assert np.array_equal(b, z["st%d_b" % ci])
        gx, gy, ge = z["st%d_xye" % ci]
        if ci != 1:
            assert abs(x - gx) < 1e-4 and abs(y - gy) < 1e-4 and abs(e - ge) < 1e-4
        else:
            assert e < 5.0 < ge and abs(x - 11.0) < 0.1 and abs(y + 5.5) < 0.1       # the shifts encode (tx, ty) = (-11, 5.5)


def _merge_case(nkf=3, kind="vlp16"):
    kfs, Ts = [], []
    P0 = synth.pose_matrix(0.5, -0.2, 0.0, 0.1)
    for f in range(nkf):
        pose = synth.pose_matrix(1.5 * f - 1.0, 0.4 * f, 0.0, 0.05 * f)
        pts = synth.make_scan(pose, kind, scan_id=300 + f).astype(np.float32)
        inten = np.random.default_rng(f).random((len(pts), 1)).astype(np.float32)
        kfs.append(np.hstack([pts, inten]))
        Ts.append((np.linalg.inv(P0) @ pose).astype(np.float32))
    return kfs, np.stack(Ts)


def test_keyframe_merge_oracle_properties():
    """mergeNearestKeyframes restated from PCL 1.10's published VoxelGrid / PassThrough / transformPointCloud (PCL is absent: parity
    unpinned by reference artefacts, see oracle/gicp_oracle.cpp).  Properties the algorithm guarantees: one output per occupied
    voxel, each output inside its voxel and inside the crop, ascending voxel index, means of the members, idempotent count."""
    kfs, Ts = _merge_case()
    leaf, lim = 0.4, 20.0
    out = oracle.merge_nearest_keyframes(kfs, Ts, lim, leaf)
    allp = np.concatenate([(k[:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32) for k, T in zip(kfs, Ts)])
    keep = (np.abs(allp[:, 0]) <= lim) & (np.abs(allp[:, 1]) <= lim)
    assert (np.abs(out[:, 0]) <= lim).all() and (np.abs(out[:, 1]) <= lim).all()
    inv = np.float32(1.0) / np.float32(leaf)
    vox_in = np.unique(np.floor(allp[keep] * inv).astype(np.int64), axis=0)
    vox_out = np.floor(out[:, :3] * inv).astype(np.int64)
    assert abs(len(out) - len(vox_in)) <= 2                       # numpy's transform differs from the se3 form in the last bit
    assert len(np.unique(vox_out, axis=0)) >= len(out) - 2        # one centroid per voxel, each inside its own voxel
    again = oracle.merge_nearest_keyframes([out], np.eye(4, dtype=np.float32)[None], lim, leaf)
    assert len(again) == len(out) and np.abs(again - out).max() < 1e-6
    # intensity is averaged like the coordinates
    assert 0.0 <= out[:, 3].min() and out[:, 3].max() <= 1.0
    # leaf too small for int indices -> PCL returns the (cropped) input unchanged
    tiny = oracle.merge_nearest_keyframes(kfs[:1], Ts[:1], lim, 1e-4)
    assert len(tiny) == int(((np.abs((kfs[0][:, :3] @ Ts[0][:3, :3].T + Ts[0][:3, 3])[:, :2]) <= lim).all(axis=1)).sum())
    assert len(oracle.merge_nearest_keyframes([np.zeros((0, 4), np.float32)], np.eye(4, dtype=np.float32)[None], lim, leaf)) == 0


def _nn_close(eig, ent, cov, z):
    e0 = z["nn_eigens"][:, :1]
    assert (np.abs(eig - z["nn_eigens"]) <= 2e-5 * e0 + 1e-4 * np.abs(z["nn_eigens"])).all()
    assert np.abs(cov.reshape(-1, 9) - z["nn_cov"].reshape(-1, 9)).max() <= 2e-5 * e0.max()
    ok = np.isfinite(z["nn_entropy"])
    assert np.abs(ent[ok] - z["nn_entropy"][ok]).max() < 2e-3


def test_ringpp_front_end_oracle_vs_reference_python():
    """RING++ front end: golden vectors from the reference's own build_neighbors_NN (sklearn kd_tree kNN, k = 30, +
    covariation_eigenvalue).  Neighbour lists are identical row for row (also pins this oracle's exact kNN against sklearn)."""
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    idx, cov, ent, eig = oracle.ringpp_build_neighbors(z["nn_pc"], 30)
    assert np.array_equal(idx, z["nn_idx"])
    _nn_close(eig, ent, cov, z)


def _disco_cases(n=6, H=40, W=120, seed=3):
    rng = np.random.default_rng(seed)
    img = (rng.random((H, W)) * (rng.random((H, W)) < 0.3)).astype(np.float32)
    A = np.fft.fft2(img)
    a = (A.real.astype(np.float32), A.imag.astype(np.float32))
    rb, ib, shifts = [], [], []
    for i in range(n):
        sh = int(rng.integers(0, W))
        other = np.roll(img, sh, axis=1) + 0.02 * rng.random((H, W)).astype(np.float32)
        if i == n - 1:
            other = (rng.random((H, W)) * (rng.random((H, W)) < 0.3)).astype(np.float32)     # unrelated scene
        B = np.fft.fft2(other)
        rb.append(B.real.astype(np.float32))
        ib.append(B.imag.astype(np.float32))
        shifts.append(sh)
    return a, (np.stack(rb), np.stack(ib)), shifts


def test_disco_rel_ori_oracle_vs_numpy_fft():
    """calcRelOri (global_manager.cpp:2719-2762): FFTW is an absent third-party dependency; its un-normalised backward 2-D DFT is
    restated directly and checked against numpy's FFT of the same definition."""
    (ra, ia), (rb, ib), shifts = _disco_cases()
    H, W = ra.shape
    for i in range(len(shifts)):
        ang, loc = oracle.disco_rel_ori(ra, ia, rb[i], ib[i])
        cross = (ra * rb[i] + ia * ib[i]).astype(np.float64) + 1j * (ra * ib[i] + rb[i] * ia).astype(np.float64)
        back = (np.fft.ifft2(cross) * H * W).real.astype(np.float32)
        assert loc == int(np.argmax(back)) and ang == (loc % W) * 3.0
        if i < len(shifts) - 1:
            # the reference's imaginary part is Im(a b), not Im(a conj(b)): the real part of the transform is then symmetric up to the
            # noise, so the peak sits at the shift or at its mirror image
            assert ang in (shifts[i] * 3.0, ((W - shifts[i]) % W) * 3.0)

