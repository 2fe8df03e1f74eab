"""GPU parity tests of the Cartesian (voxelocc) and polar (gputransform) voxelisers: bit-exact against the oracle and,
where oracle/_ref is present, against the reference's own GPU classes (kernel + host scatter) compiled unmodified."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from mr_slam_b200 import gputransform, ring, synth, voxelocc

pytestmark = pytest.mark.gpu


def clouds():
    rng = np.random.default_rng(3)
    a = rng.uniform(-1.05, 1.05, (50000, 3)).astype(np.float32)
    a[:50] = 0.0
    a[50:60] = 1.0
    a[60:70] = -1.0
    b = synth.ring_preprocess(synth.make_scan(synth.pose_matrix(1, 2, 0, 0.4), "hdl64", 7) + np.array([0, 0, 1.8], np.float32))
    return {"random": a, "scan": b}


def run(cls, pts, *cfg):
    soa = np.ascontiguousarray(pts.T.reshape(-1))
    t = cls(soa, len(pts), *cfg)
    t.transform()
    return soa, t.retreive()


@pytest.mark.parametrize("nh,el", [(1, 1), (4, 1), (3, 2)])
def test_voxelocc_vs_oracle(gpu, nh, el):
    for name, pts in clouds().items():
        soa, out = run(voxelocc.GPUTransformer, pts, 1, 1, 120, 120, nh, el)
        assert out.dtype == np.float32 and out.shape == (120 * 120 * nh * el * 3,)
        assert np.array_equal(out, oracle.voxelocc(soa, len(pts), 1, 1, 120, 120, nh, el)), name


@pytest.mark.parametrize("nh,el", [(1, 1), (20, 1), (4, 3)])
def test_polar_vs_oracle(gpu, nh, el):
    for name, pts in clouds().items():
        soa, out = run(gputransform.GPUTransformer, pts, 1, 1, 40, 120, nh, el)
        exp = oracle.polar(soa, len(pts), 1, 1, 40, 120, nh, el)
        # CUDA atanf (what the reference GPU kernel uses, and what we use) and glibc atanf may differ in the last ulp:
        # a sector index could flip for a point sitting exactly on a 3-degree boundary.  None does on these inputs.
        assert np.array_equal(out, exp), name


def test_polar_reference_fixture(gpu):
    z = np.load(os.path.join(GOLDEN, "polar_fixture.npz"))
    for name in ("disco1", "disco2"):
        pts = z[name + "_points"]
        for ci, c in enumerate(z["cfgs"]):
            soa, out = run(gputransform.GPUTransformer, pts, *[int(v) for v in c])
            exp = np.zeros_like(out)
            exp[z["%s_cfg%d_idx" % (name, ci)]] = z["%s_cfg%d_val" % (name, ci)]
            assert np.array_equal(out, exp)


def test_against_reference_gpu_classes(gpu):
    from oracle import ref
    if not (ref.available("voxelocc") and ref.available("polar_gpu")):
        pytest.skip("oracle/_ref not present")
    for name, pts in clouds().items():
        pts = pts[(np.abs(pts) < 1.0).all(1)]   # the reference writes out of bounds for |v| == 1 / out-of-range heights
        soa, out = run(voxelocc.GPUTransformer, pts, 1, 1, 120, 120, 1, 1)
        assert np.array_equal(out, ref.voxelocc_gpu(soa, len(pts), 1, 1, 120, 120, 1, 1)), name
        soa, out = run(gputransform.GPUTransformer, pts, 1, 1, 40, 120, 20, 1)
        assert np.array_equal(out, ref.polar_gpu(soa, len(pts), 1, 1, 40, 120, 20, 1)), name


def test_batched_device_bev_equals_per_scan(gpu):
    import torch
    scans = [synth.ring_preprocess(synth.make_scan(synth.pose_matrix(i, -i, 0, 0.2 * i), "vlp16", 40 + i) + np.array([0, 0, 1.8], np.float32))
             for i in range(5)]
    scans.append(np.zeros((0, 3), np.float32))    # an empty scan in the batch
    bev = ring.bev_batch(scans).cpu().numpy()
    for i, p in enumerate(scans):
        soa = np.ascontiguousarray(p.T.reshape(-1)) if len(p) else np.zeros(0, np.float32)
        exp = oracle.voxelocc(soa if len(p) else np.zeros(3, np.float32), len(p), 1, 1, 120, 120, 1, 1).reshape(-1, 3)[:, 2].reshape(120, 120)
        assert np.array_equal(bev[i], exp)


def test_fused_scan_to_bev_equals_load_pc_infer_plus_voxelocc(gpu):
    """mrslam_ring_scans_to_bev_device (raw scans -> BEV in one kernel) == load_pc_infer (restated in synth.ring_preprocess, RING
    util.py:91-112) followed by the voxelocc oracle, bit for bit -- including points on the filter limits, z = 0 and NaNs."""
    import torch
    raw = [synth.make_scan(synth.pose_matrix(i, -i, 0, 0.2 * i), "vlp16", 60 + i) + np.array([0, 0, 1.8], np.float32) for i in range(3)]
    edge = np.array([[70.0, 0, 1], [-70.0, 0, 1], [69.99999, 1, 1], [0, 0, 0], [1, 1, 30.0], [1, 1, 29.999998], [np.nan, 0, 1],
                     [0.0, 0.0, 5.0], [5, -69.99999, 0.5]], np.float32)
    raw[1] = np.concatenate([raw[1][:-len(edge)], edge]).astype(np.float32)
    d = torch.from_numpy(np.stack(raw)).cuda()
    bev = ring.scans_to_bev_device(d).cpu().numpy()
    for i, r in enumerate(raw):
        p = synth.ring_preprocess(r)
        soa = np.ascontiguousarray(p.T.reshape(-1))
        exp = oracle.voxelocc(soa, len(p), 1, 1, 120, 120, 1, 1).reshape(-1, 3)[:, 2].reshape(120, 120)
        assert np.array_equal(bev[i], exp), i
    # ragged batch with 4 floats per point (x y z intensity) and an empty scan
    rag = [np.concatenate([r[: 1000 * (i + 1)], np.ones((1000 * (i + 1), 1), np.float32)], 1) for i, r in enumerate(raw)] + [np.zeros((0, 4), np.float32)]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in rag])]).astype(np.int64)
    bev2 = ring.scans_to_bev_device(torch.from_numpy(np.concatenate(rag)).cuda(), offsets=offs).cpu().numpy()
    for i, r in enumerate(rag):
        p = synth.ring_preprocess(r[:, :3])
        exp = oracle.voxelocc(np.ascontiguousarray(p.T.reshape(-1)) if len(p) else np.zeros(3, np.float32), len(p), 1, 1, 120, 120, 1, 1)
        assert np.array_equal(bev2[i], exp.reshape(-1, 3)[:, 2].reshape(120, 120)), i


def test_empty_and_argument_errors(gpu):
    out = np.zeros(3, np.float32)
    t = voxelocc.GPUTransformer(np.zeros(3, np.float32), 0, 1, 1, 120, 120, 1, 1)
    t.transform()
    assert (t.retreive() == 0).all()
    with pytest.raises(ValueError):
        voxelocc.GPUTransformer(np.zeros(3, np.float64), 1, 1, 1, 120, 120, 1, 1)   # the Cython signature demands float32
    with pytest.raises(ValueError):
        voxelocc.GPUTransformer(np.zeros(3, np.float32), 5, 1, 1, 120, 120, 1, 1)   # buffer shorter than 3*size
    assert out.shape == (3,)


def _feat_inputs(n=6000, k=30, seed=8):
    rng = np.random.default_rng(seed)
    pc = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    nbr = rng.integers(0, n, (n, k)).astype(np.int32)
    e3 = np.sort(rng.uniform(1e-4, 1.0, (n, 3)).astype(np.float32), axis=1)[:, ::-1]
    e2 = np.sort(rng.uniform(1e-4, 1.0, (n, 2)).astype(np.float32), axis=1)[:, ::-1]
    eig = np.ascontiguousarray(np.concatenate([e3, e2], 1), np.float32)
    return pc, nbr, eig


def test_voxelfeat_features(gpu):
    """RING++ point features (BEVF/src/kernel.cu:16-104): 13 features per point, vs the oracle and the reference kernel."""
    from mr_slam_b200 import voxelfeat
    pc, nbr, eig = _feat_inputs()
    n, k = nbr.shape
    fe = voxelfeat.GPUFeatureExtractor(pc.reshape(-1), n, 13, k, nbr.reshape(-1), eig.reshape(-1))
    got = fe.get_features().reshape(n, 13)
    exp = oracle.voxelfeat_features(pc.reshape(-1), n, 13, k, nbr.reshape(-1), eig.reshape(-1)).reshape(n, 13)
    assert np.allclose(got, exp, rtol=2e-5, atol=1e-6)      # logf / pow differ by an ulp or two between glibc and CUDA
    from oracle import ref
    if ref.available("voxelfeat"):
        r = ref.voxelfeat_features(pc.reshape(-1), n, 13, k, nbr.reshape(-1), eig.reshape(-1)).reshape(n, 13)
        assert np.allclose(got, r, rtol=2e-6, atol=1e-7)     # same device math library: last-bit agreement


def test_voxelfeat_bev(gpu):
    """feature BEV (BEVF/src/kernel.cu:106-164): per cell / channel maximum of the positive values."""
    from mr_slam_b200 import voxelfeat
    rng = np.random.default_rng(9)
    n, fs = 40000, 9
    pts = rng.uniform(-1.05, 1.05, (n, fs)).astype(np.float32)
    soa = np.ascontiguousarray(pts.T.reshape(-1))
    t = voxelfeat.GPUTransformer(soa, n, 1, 1, 120, 120, 1, fs)
    t.transform()
    got = t.retreive()
    assert got.shape == (120 * 120 * fs,)
    assert np.array_equal(got, oracle.voxelfeat_bev(soa, n, 1, 1, 120, 120, 1, fs))
    from oracle import ref
    if ref.available("voxelfeat"):
        # the reference update is a non-atomic compare-then-write: compare on an input without collisions (one point per cell)
        gx, gy = np.meshgrid(np.arange(120), np.arange(120), indexing="ij")
        cx = ((gx.reshape(-1) + 0.5) / 60.0 - 1.0).astype(np.float32)
        cy = ((gy.reshape(-1) + 0.5) / 60.0 - 1.0).astype(np.float32)
        m = len(cx)
        p2 = rng.uniform(-0.9, 0.9, (m, fs)).astype(np.float32)
        p2[:, 0], p2[:, 1] = cx, cy
        soa2 = np.ascontiguousarray(p2.T.reshape(-1))
        t2 = voxelfeat.GPUTransformer(soa2, m, 1, 1, 120, 120, 1, fs)
        t2.transform()
        assert np.array_equal(t2.retreive(), ref.voxelfeat_bev(soa2, m, 1, 1, 120, 120, 1, fs))


def test_ringpp_build_neighbors_vs_reference_python(gpu):
    """build_neighbors_NN (CPU sklearn kNN(30) + numpy/LAPACK eigenvalues in the reference) on the GPU: neighbour lists
    identical to the reference's, eigenvalues / covariances / entropy to float32 rounding; its output drives GPUFeatureExtractor."""
    import os
    from conftest import GOLDEN
    from test_oracle_golden import _nn_close
    from mr_slam_b200 import ring
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    idx, cov, ent, eig, vec = ring.build_neighbors_NN(z["nn_pc"], 30)
    assert vec == [] and idx.dtype == np.int32
    assert np.array_equal(idx, z["nn_idx"])
    _nn_close(eig, ent, cov, z)
    # chained with the feature extractor exactly like generate_RINGplusplus (util.py:211-220)
    from mr_slam_b200 import voxelfeat
    pc = z["nn_pc"]
    n = len(pc)
    feat = voxelfeat.GPUFeatureExtractor(pc.flatten(), n, 13, 30, idx.flatten().astype(np.int32), eig.flatten()).get_features()
    exp = oracle.voxelfeat_features(pc.reshape(-1), n, 13, 30, z["nn_idx"].reshape(-1), eig.reshape(-1))
    assert np.allclose(np.asarray(feat).reshape(n, 13), exp.reshape(n, 13), rtol=2e-5, atol=1e-6, equal_nan=True)   # logf/pow: glibc vs CUDA
    with pytest.raises(Exception):
        ring.build_neighbors_NN(pc, 40)          # k <= 32

