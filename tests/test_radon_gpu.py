"""GPU parity tests of the Radon forward projection and the RING descriptor tail / matcher.
Tolerance (north_star): RING descriptor within 1e-4 (relative)."""
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, rel_l2
from mr_slam_b200 import ring
from mr_slam_b200.torch_radon import ParallelBeam

pytestmark = pytest.mark.gpu
TOL = 1e-4
ANG = np.linspace(0, 2 * np.pi, 120).astype(np.float32)   # util.py:192 (endpoint inclusive)


def images(b=4, s=120, seed=5):
    rng = np.random.default_rng(seed)
    return (rng.random((b, s, s)) * (rng.random((b, s, s)) < 0.2)).astype(np.float32)


@pytest.mark.parametrize("interp,wb", [("exact", 0), ("texture", 8)])
def test_radon_vs_oracle(gpu, interp, wb):
    img = images()
    pb = ParallelBeam(120, ANG, interp=interp, fast_trig=False)
    got = pb.forward(torch.from_numpy(img).cuda()).cpu().numpy()
    exp = oracle.radon_forward(img, ANG, 120, 1.0, wb)
    assert got.shape == (4, 120, 120)
    assert np.abs(got - exp).max() <= TOL * np.abs(exp).max()
    assert rel_l2(exp, got) < 1e-6


def test_radon_shapes_and_linearity(gpu):
    pb = ParallelBeam(96, (0.0, np.pi, 50), interp="exact", fast_trig=False)     # (start, end, n) tuple form, endpoint=False
    x = torch.from_numpy(images(6, 80, 1)).cuda().reshape(2, 3, 80, 80)
    y = pb.forward(x)
    assert y.shape == (2, 3, 50, 96) and y.is_cuda
    a, b = x[0, 0], x[1, 2]
    lin = pb.forward(2.0 * a - 0.5 * b)
    assert torch.allclose(lin, 2.0 * pb.forward(a) - 0.5 * pb.forward(b), rtol=1e-5, atol=1e-4)   # the projection is linear
    with pytest.raises(RuntimeError):
        pb.forward(x.cpu())                                                       # no CPU fallback


def test_radon_vs_reference_analytic(gpu):
    """the reference's own acceptance test (torch-radon/tests/test_parallel_beam.py:48-70)"""
    from oracle import ref
    if not ref.available("symbolic"):
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(7)
    H = W = D = A = 128
    f = ref.Symbolic(H, W)
    for _ in range(30):
        args = (rng.uniform(0, 1), rng.uniform(-W / 2, W / 2), rng.uniform(-H / 2, H / 2), rng.uniform(W / 32, W / 4), rng.uniform(H / 32, H / 4))
        (f.add_ellipse if rng.integers(0, 2) < 0.25 else f.add_gaussian)(*args)
    img = f.discretize()
    angles = np.linspace(0, 2 * np.pi, A, endpoint=False).astype(np.float32)
    sino = ParallelBeam(D, angles).forward(torch.from_numpy(img).cuda()[None])[0].cpu().numpy()
    assert rel_l2(f.forward(angles, D), sino) < 2e-3 * (512 / A) * (512 / D)


def test_radon_vs_reference_kernel(gpu):
    """Side by side with the reference's own CUDA kernel (hardware texture filtering, __sinf/__cosf).  Our default mode
    (texture-weight emulation + fast trig) tracks it much closer than exact bilinear does; both deviations are reported in
    DESIGN.md.  Bounds here are loose upper limits of what was measured."""
    from oracle import ref
    if not ref.available("radon"):
        pytest.skip("oracle/_ref not present")
    img = images(4, 120, 2)
    exp = ref.radon_forward(img, ANG, 120, 1.0)
    x = torch.from_numpy(img).cuda()
    emu = ParallelBeam(120, ANG, interp="texture", fast_trig=True).forward(x).cpu().numpy()
    exact = ParallelBeam(120, ANG, interp="exact", fast_trig=False).forward(x).cpu().numpy()
    e_emu, e_exact = rel_l2(exp, emu), rel_l2(exp, exact)
    print("relative L2 vs reference kernel: texture-emulation %.3e, exact bilinear %.3e" % (e_emu, e_exact))
    assert e_emu < 2e-3 and e_exact < 5e-3


def test_radon_mode_gap_on_the_matcher(gpu):
    """What the Radon-mode difference means for the MATCHER (VERDICT r1 task 10): 256 scan pairs, sinograms from (i) the
    reference's own hardware-texture kernel (libradon_ref.so, forward.cu unmodified) and (ii) the product default
    (texture-weight emulation + fast trig) and (iii) exact bilinear; TIRING + fast_corr of scan j against scan 0 in each mode.
    Reported: max |delta dist| and the number of yaw (angle) flips; asserted: the product default never flips the matched
    angle where the correlation peak is unambiguous and moves `dist` by < 2e-3 (the loop threshold is 0.3-0.5 wide)."""
    from oracle import ref
    from mr_slam_b200 import synth
    if not ref.available("radon"):
        pytest.skip("oracle/_ref not present")
    n_base, reps = 16, 16
    base = [synth.make_scan(synth.pose_matrix(3.0 * np.cos(0.7 * i), 2.0 * np.sin(1.3 * i), 0.0, 0.2 * i), "vlp16", scan_id=300 + i) +
            np.array([0, 0, 1.8], np.float32) for i in range(n_base)]
    scans = []
    for j in range(n_base * reps):
        b = base[j % n_base]
        yaw = 2 * np.pi * (j // n_base) / reps + 0.013 * j
        c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
        scans.append(np.stack([c * b[:, 0] - s * b[:, 1], s * b[:, 0] + c * b[:, 1], b[:, 2]], 1).astype(np.float32))
    bev = ring.scans_to_bev_device(torch.from_numpy(np.stack(scans)).cuda())
    bev_h = bev.cpu().numpy()
    sino_ref = ref.radon_forward(bev_h, ANG, 120, 1.0)                                            # (i) reference kernel on this GPU
    sino_emu = ParallelBeam(120, ANG, interp="texture", fast_trig=True).forward(bev).cpu().numpy()   # (ii) product default
    sino_ex = ParallelBeam(120, ANG, interp="exact", fast_trig=False).forward(bev).cpu().numpy()     # (iii) exact bilinear
    res = {}
    for name, sino in (("ref", sino_ref), ("default", sino_emu), ("exact", sino_ex)):
        tir = ring.tiring(torch.from_numpy(np.ascontiguousarray(sino[:, None])).cuda())
        d, a, corr = ring.fast_corr_batch(tir[0], tir, want_corr=True)
        res[name] = (d.cpu().numpy(), a.cpu().numpy(), corr.cpu().numpy())
    d0, a0, c0 = res["ref"]
    top2 = np.sort(c0, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-3 * top2[:, 1]            # pairs whose correlation peak is not a near-tie
    out = {}
    for name in ("default", "exact"):
        d, a, _ = res[name]
        out[name] = (float(np.abs(d - d0).max()), int((a != a0).sum()), int(((a != a0) & clear).sum()))
    print("matcher gap vs the reference Radon kernel over %d pairs: default max|ddist| %.2e, angle flips %d (%d on unambiguous peaks); "
          "exact-bilinear max|ddist| %.2e, angle flips %d (%d unambiguous)" % ((len(scans),) + out["default"] + out["exact"]))
    assert out["default"][0] < 2e-3 and out["default"][2] == 0
    assert out["exact"][0] < 5e-3


def test_tiring_and_fast_corr_vs_reference_python(gpu):
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    tir = ring.tiring(torch.from_numpy(z["sino"]).cuda())
    assert tir.dtype == torch.complex64 and tuple(tir.shape) == (2, 1, 120, 120)
    got = tir.cpu().numpy()
    assert np.abs(got - z["tiring"]).max() <= TOL * np.abs(z["tiring"]).max()
    db = torch.from_numpy(z["tiring"]).cuda()
    for i in range(2):
        dist, ang = ring.fast_corr_batch(db[i], db)
        assert np.abs(dist.cpu().numpy() - z["fast_corr_dist"][i]).max() < 1e-4
        assert np.array_equal(ang.cpu().numpy(), z["fast_corr_angle"][i])
    # yaw recovery: the descriptor of a sinogram rolled by s rows correlates at angle s (reference python gives the same)
    rolled = torch.from_numpy(np.roll(z["sino"][0], int(z["rolled_by"]), axis=-2)[None]).cuda()
    dist, ang = ring.fast_corr_batch(db[0], ring.tiring(rolled))
    assert int(ang[0]) == int(z["rolled_angle"]) and abs(float(dist[0]) - float(z["rolled_dist"])) < 1e-4


def test_tiring_vs_torch_fft_on_gpu(gpu):
    """the reference's two lines (util.py:197-198) run with torch on the GPU, next to our fused kernels"""
    sino = torch.from_numpy(oracle.radon_forward(images(3), ANG, 120, 1.0, 0)).cuda()[:, None]
    got = ring.tiring(sino)
    for i in range(3):
        x = sino[i]
        exp = torch.fft.fft2((x - x.mean()) / x.std(), dim=-2, norm="ortho")
        assert (got[i] - exp).abs().max() <= TOL * exp.abs().max()


def test_generate_ring_pipeline_descriptor_tolerance(gpu):
    """BEV -> Radon -> normalise -> FFT for a batch, against the oracle chain: RING descriptor within 1e-4."""
    from mr_slam_b200 import synth
    scans = [synth.ring_preprocess(synth.make_scan(synth.pose_matrix(i, i, 0, 0.5 * i), "vlp16", 60 + i) + np.array([0, 0, 1.8], np.float32))
             for i in range(4)]
    bev = ring.bev_batch(scans)
    sino, tir = ring.generate_RING_batch(bev, interp="exact", fast_trig=False, want_sino=True)
    for i, p in enumerate(scans):
        soa = np.ascontiguousarray(p.T.reshape(-1))
        bev_o = oracle.voxelocc(soa, len(p), 1, 1, 120, 120, 1, 1).reshape(-1, 3)[:, 2].reshape(1, 120, 120)
        assert np.array_equal(bev[i].cpu().numpy(), bev_o[0])
        sino_o = oracle.radon_forward(bev_o, ANG, 120, 1.0, 0)
        assert np.abs(sino[i].cpu().numpy() - sino_o).max() <= TOL * np.abs(sino_o).max()
        _, tir_o = oracle.ring_tiring(sino_o)
        assert np.abs(tir[i].cpu().numpy() - tir_o).max() <= TOL * np.abs(tir_o).max()
    # batch of 256 (BASELINE config 3 size): every descriptor equals its single-image evaluation (batch independence)
    big = bev.repeat(64, 1, 1)
    _, tb = ring.generate_RING_batch(big, interp="exact", fast_trig=False)
    assert torch.equal(tb[:4], tir) and torch.equal(tb[252:], tir)


def test_solve_translation_vs_reference_python_and_oracle(gpu):
    """the python row loop of solve_translation as one kernel.  Per-row shifts bit-equal to the reference's own
    (golden) and the oracle's; x / y / error within 1e-4 of the oracle (least-squares solution, see the CPU test for the reference's
    v.t() slip)."""
    z = np.load(os.path.join(GOLDEN, "ring_fixture.npz"))
    for ci in range(3):
        rot, rs = z["st%d_params" % ci]
        q = torch.from_numpy(z["st%d_query" % ci]).cuda()
        pos = torch.from_numpy(z["st%d_positive" % ci]).cuda()
        x, y, e, b = ring.solve_translation(q, pos, float(rot), row_shift=int(rs), want_shifts=True)
        ox, oy, oe, ob = oracle.ring_solve_translation(z["st%d_query" % ci], z["st%d_positive" % ci], float(rot), int(rs))
        assert np.array_equal(b.cpu().numpy(), z["st%d_b" % ci]) and np.array_equal(ob, z["st%d_b" % ci])
        assert abs(x - ox) < 1e-4 and abs(y - oy) < 1e-4 and abs(e - oe) < 1e-3
        # the reference's calling convention: roll first, no row_shift argument
        x2, y2, e2 = ring.solve_translation(q, torch.roll(pos, int(rs), dims=1), float(rot), q.device)
        assert (x2, y2, e2) == (x, y, e)
        if ci != 1:
            gx, gy, ge = z["st%d_xye" % ci]
            assert abs(x - gx) < 1e-4 and abs(y - gy) < 1e-4 and abs(e - ge) < 1e-3
    assert ring.calculate_row_shift(17) == -17 and ring.calculate_row_shift(100) == -20


def test_disco_rel_ori_vs_oracle(gpu):
    """calcRelOri batched over the database signatures: flat arg-max and angle identical to the oracle's double DFT"""
    from test_oracle_golden import _disco_cases
    (ra, ia), (rb, ib), shifts = _disco_cases(n=9, seed=8)
    ang, loc = ring.calc_rel_ori((torch.from_numpy(ra).cuda(), torch.from_numpy(ia).cuda()),
                                 (torch.from_numpy(rb).cuda(), torch.from_numpy(ib).cuda()), want_loc=True)
    ang, loc = ang.cpu().numpy(), loc.cpu().numpy()
    for i in range(len(shifts)):
        oa, ol = oracle.disco_rel_ori(ra, ia, rb[i], ib[i])
        assert int(loc[i]) == ol and float(ang[i]) == oa
    for i in range(len(shifts) - 1):
        assert float(ang[i]) in (shifts[i] * 3.0, ((120 - shifts[i]) % 120) * 3.0)      # shift or its mirror image (see the CPU test)

