"""GPU parity tests of the FastGICP path (through the C-ABI) against the CPU oracle and the reference's fixtures.
Tolerance (BASELINE.json north_star): final transform within 1e-4 in R|t; neighbour / correspondence indices bit-exact."""
import numpy as np
import pytest

import oracle
from mr_slam_b200 import pygicp, synth

pytestmark = pytest.mark.gpu
TOL_T = 1e-4


def make(kind_cls, src, tgt, k=15, max_corr=100.0, iters=50, teps=1e-3, reg=None, threads=0):
    r = kind_cls()
    r.set_num_threads(threads)
    r.set_correspondence_randomness(k)
    r.set_max_correspondence_distance(max_corr)
    r.set_maximum_iterations(iters)
    r.set_transformation_epsilon(teps)
    if reg is not None:
        r.set_regularization_method(reg)
    r.set_input_source(src)
    r.set_input_target(tgt)
    return r


@pytest.fixture(scope="module")
def vlp_pair():
    return synth.make_pair(0, "vlp16")


@pytest.mark.parametrize("k", [1, 8, 15, 20, 32])
def test_knn_bit_exact(gpu, vlp_pair, k):
    tgt = vlp_pair[1]
    gi, gd = pygicp.Cloud(tgt).knn(k)
    oi, od = oracle.knn(tgt, tgt, k)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)


def test_knn_duplicates_and_tiny_clouds(gpu):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(700, 3)).astype(np.float32)
    pts[300:340] = pts[100:140]          # exact duplicates -> ties resolved by lowest index
    pts[:, 2] = np.round(pts[:, 2], 1)   # many equal coordinates
    gi, gd = pygicp.Cloud(pts).knn(15)
    oi, od = oracle.knn(pts, pts, 15)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    small = pts[:9]
    gi, gd = pygicp.Cloud(small).knn(15)   # fewer points than k: missing entries are -1 / inf
    oi, od = oracle.knn(small, small, 15)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


@pytest.mark.parametrize("reg", [oracle.REG_PLANE, oracle.REG_NONE, oracle.REG_MIN_EIG, oracle.REG_NORMALIZED_MIN_EIG, oracle.REG_FROBENIUS])
def test_covariances_all_regularisations(gpu, vlp_pair, reg):
    src, tgt = vlp_pair[0], vlp_pair[1]
    g = make(pygicp.FastGICP, src, tgt, k=15, reg=reg)
    cg = g.get_target_covariances(len(tgt))
    co = oracle.covariances(tgt, 15, reg)
    assert np.abs(cg[:, :3, :3] - co).max() <= 1e-12 * max(1.0, np.abs(co).max())
    assert (cg[:, 3, :] == 0).all() and (cg[:, :, 3] == 0).all()   # the reference's Matrix4d has a zero 4th row/column


def test_linearize_stage_parity(gpu, vlp_pair):
    src, tgt, T_gt, T_guess = vlp_pair[:4]
    g = make(pygicp.FastGICP, src, tgt)
    o = make(oracle.FastGICP, src, tgt)
    Tg = T_guess.astype(np.float32)
    cost_g, Hg, bg = g.evaluate_cost(Tg)
    cost_o, Ho, bo, corr_o, sqd_o, mah_o = o.linearize(Tg.astype(np.float64), len(src))
    corr_g, sqd_g, mah_g = g.last_correspondences(len(src))
    assert np.array_equal(corr_g, corr_o)          # correspondence indices bit-exact
    assert np.array_equal(sqd_g, sqd_o)
    assert np.abs(mah_g - mah_o).max() <= 1e-12 * np.abs(mah_o).max()
    assert abs(cost_g - cost_o) <= 1e-11 * abs(cost_o)
    assert np.abs(Hg - Ho).max() <= 1e-11 * np.abs(Ho).max()
    assert np.abs(bg - bo).max() <= 1e-11 * np.abs(bo).max() + 1e-9
    e_g, e_o = g.compute_error(T_gt), o.compute_error(T_gt)
    assert abs(e_g - e_o) <= 1e-11 * abs(e_o)


def test_correspondence_ties_pick_lowest_index(gpu):
    """NN-1 under exact distance ties: every target point exists three times (shuffled original indices) on a coarse lattice and
    the queries sit on the half lattice, so each query has several nearest neighbours at bit-equal distances -- the answer is the
    lowest original index (oracle: gicp_oracle.cpp nn1; reference fast_gicp_impl.hpp:183-214).  Unseeded pass: evaluate_cost();
    seeded pass: the second linearisation inside align(), compared with the oracle's exact search at the same x1."""
    rng = np.random.default_rng(5)
    base = np.round(rng.uniform(-6, 6, size=(5000, 3)) * 4) / 4
    tgt = np.concatenate([base, base, base]).astype(np.float32)
    tgt = tgt[rng.permutation(len(tgt))]
    src = (np.round(rng.uniform(-6, 6, size=(4000, 3)) * 8) / 8).astype(np.float32)
    g = make(pygicp.FastGICP, src, tgt)
    o = make(oracle.FastGICP, src, tgt)
    for T in (np.eye(4), synth.pose_matrix(0.125, 0.0, -0.125, 0.0)):   # lattice-preserving shifts keep the ties exact
        Tg = np.asarray(T, np.float32)
        g.evaluate_cost(Tg)
        corr_o, sqd_o = o.linearize(Tg.astype(np.float64), len(src))[3:5]
        corr_g, sqd_g, _ = g.last_correspondences(len(src))
        assert np.array_equal(sqd_g, sqd_o)
        assert np.array_equal(corr_g, corr_o)
    guess = synth.pose_matrix(0.03, -0.02, 0.01, 0.01)
    g2 = make(pygicp.FastGICP, src, tgt, iters=2)
    g2.align(guess)
    o1 = make(oracle.FastGICP, src, tgt, iters=1)
    o1.align(guess)
    x1 = o1.result()["Td"]
    corr_o, sqd_o = o.linearize(x1, len(src))[3:5]
    corr_g, sqd_g, _ = g2.last_correspondences(len(src))
    assert np.array_equal(sqd_g, sqd_o)
    assert np.array_equal(corr_g, corr_o)      # the three copies of a target point tie exactly: lowest index wins


@pytest.mark.parametrize("pid,k", [(0, 15), (1, 20), (2, 15)])
def test_align_matches_oracle(gpu, pid, k):
    src, tgt, T_gt, T_guess = synth.make_pair(pid, "vlp16")[:4]
    g = make(pygicp.FastGICP, src, tgt, k=k)
    o = make(oracle.FastGICP, src, tgt, k=k)
    Tg = g.align(T_guess.astype(np.float32))
    To = o.align(T_guess.astype(np.float32))
    ro = o.result()
    assert np.abs(g.get_final_transformation_d() - ro["Td"]).max() < TOL_T
    assert np.abs(Tg - To).max() < TOL_T
    assert g.has_converged() == ro["converged"] and g.num_iterations() == ro["iterations"]
    assert g.eval_counts() == (ro["n_linearize"], ro["n_error"])
    assert np.abs(g.get_final_hessian() - ro["H"]).max() <= 1e-9 * np.abs(ro["H"]).max()
    assert abs(g.get_fitness_score(1.0) - o.get_fitness_score(1.0)) < 1e-9
    # `output` cloud of align(): the source moved by final_transformation_
    al = g.get_aligned(len(src))
    exp = (src.astype(np.float64) @ Tg[:3, :3].astype(np.float64).T + Tg[:3, 3]).astype(np.float32)
    assert al.shape == src.shape and np.abs(al - exp).max() < 1e-4


def test_reference_fixture_gtest_cases(gpu, gicp_golden):
    """FG/src/test/gicp_test.cpp:147-201 on the reference's own fixture: forward, backward, swap-and-set-source,
    swap-and-set-target must be within t < 0.05 m / r < 1 deg of relative.txt and converge; and match the oracle to 1e-4."""
    rel = gicp_golden["relative"]
    tgt, src = gicp_golden["target"], gicp_golden["source"]
    t_tol, r_tol = 0.05, np.deg2rad(1.0)
    reg = pygicp.FastGICP()            # library defaults (k=20, 64 iterations), identity guess
    reg.set_input_target(tgt)
    reg.set_input_source(src)
    T = reg.align()
    te, re = synth.pose_error(T, rel)
    assert te < t_tol and re < r_tol and reg.has_converged()
    assert np.abs(reg.get_final_transformation_d() - gicp_golden["forward_T"]).max() < TOL_T
    assert abs(reg.get_fitness_score(1.0) - float(gicp_golden["forward_fitness"])) < 1e-6
    reg.set_input_target(src)          # backward
    reg.set_input_source(tgt)
    T = reg.align()
    te, re = synth.pose_error(np.linalg.inv(T), rel)
    assert te < t_tol and re < r_tol and reg.has_converged()
    assert np.abs(reg.get_final_transformation_d() - gicp_golden["backward_T"]).max() < TOL_T
    reg = pygicp.FastGICP()            # swap and set source
    reg.set_input_source(tgt)
    reg.swap_source_and_target()
    reg.set_input_source(src)
    te, re = synth.pose_error(reg.align(), rel)
    assert te < t_tol and re < r_tol and reg.has_converged()
    reg = pygicp.FastGICP()            # swap and set target
    reg.set_input_target(src)
    reg.swap_source_and_target()
    reg.set_input_target(tgt)
    te, re = synth.pose_error(reg.align(), rel)
    assert te < t_tol and re < r_tol and reg.has_converged()


def test_readme_known_answer(gpu, readme_golden):
    """FG/README.md:112-113 fgicp_mt fitness_score 0.204412 (ApproximateVoxelGrid 0.1 preprocessing)."""
    reg = pygicp.FastGICP()
    reg.set_input_target(readme_golden["target"])
    reg.set_input_source(readme_golden["source"])
    reg.align()
    fit = reg.get_fitness_score()
    assert abs(fit - 0.204412) < 5e-4
    assert abs(fit - float(readme_golden["fitness"])) < 1e-6
    assert np.abs(reg.get_final_transformation_d() - readme_golden["T"]).max() < TOL_T


def test_python_caller_parameters(gpu):
    """RING/main_RING.py:81-104: downsample 0.2, max correspondence distance 1.0/5.0, default k=20, threads 4."""
    src, tgt, T_gt, T_guess = synth.make_pair(4, "vlp16")[:4]
    s, t = pygicp.downsample(src, 0.2), pygicp.downsample(tgt, 0.2)
    for mc in (1.0, 5.0):
        g, o = pygicp.FastGICP(), oracle.FastGICP()
        for r in (g, o):
            r.set_input_target(t)
            r.set_input_source(s)
            r.set_num_threads(4)
            r.set_max_correspondence_distance(mc)
        Tg = g.align(initial_guess=T_guess)
        To = o.align(T_guess)
        assert np.abs(Tg - To).max() < TOL_T
        assert abs(g.get_fitness_score(1.0) - o.get_fitness_score(1.0)) < 1e-9
    T = pygicp.align_points(t, s, k_correspondences=15, initial_guess=T_guess)
    assert T.shape == (4, 4) and T.dtype == np.float64


def test_gauss_newton_and_batch_equal_single(gpu):
    pairs = [synth.make_pair(i, "vlp16")[:4] for i in range(3)]
    singles, regs = [], []
    for src, tgt, T_gt, T_guess in pairs:
        g = make(pygicp.FastGICP, src, tgt)
        singles.append(g.align(T_guess.astype(np.float32)).copy())
        regs.append(make(pygicp.FastGICP, src, tgt))
    Ts = pygicp.align_batch(regs, np.stack([p[3].astype(np.float32) for p in pairs]))
    for a, b in zip(singles, Ts):
        assert np.array_equal(a, b)      # lock-step batching must not change any single result (deterministic reductions)
    fb = pygicp.fitness_batch(regs, 1.0)
    for r, f in zip(regs, fb):
        assert abs(r.get_fitness_score(1.0) - f) < 1e-12
    src, tgt, T_gt, T_guess = pairs[0]
    g, o = make(pygicp.FastGICP, src, tgt), make(oracle.FastGICP, src, tgt)
    g.set_lsq_optimizer(1)
    o.optimizer = 1
    assert np.abs(g.align(T_guess.astype(np.float32)) - o.align(T_guess.astype(np.float32))).max() < TOL_T


def test_device_batch_create_and_prepare_equal_plain_path(gpu):
    """clouds packed from device pointers by one launch + covariances prepared ahead (source first, targets later, as the
    overlapped multi-GPU step does) must give bit-identical transforms to the plain host-buffer path"""
    import torch
    pairs = [synth.make_pair(i, "vlp16")[:4] for i in range(3)]
    guesses = np.stack([p[3].astype(np.float32) for p in pairs])
    plain = pygicp.align_batch([make(pygicp.FastGICP, p[0], p[1]) for p in pairs], guesses)
    # ragged sizes: every cloud keeps its own length
    src_t = [torch.from_numpy(np.ascontiguousarray(p[0], np.float32)).cuda() for p in pairs]
    tgt_t = [torch.from_numpy(np.ascontiguousarray(p[1], np.float32)).cuda() for p in pairs]
    torch.cuda.synchronize()
    regs = []
    for _ in pairs:
        g = pygicp.FastGICP()
        g.set_correspondence_randomness(15)
        g.set_max_correspondence_distance(100.0)
        g.set_maximum_iterations(50)
        g.set_transformation_epsilon(1e-3)
        regs.append(g)
    with pytest.raises(Exception):
        pygicp.prepare_batch(regs, pygicp.SOURCE)     # no source set yet
    sc = pygicp.Cloud.from_device_batch([t.data_ptr() for t in src_t], [t.shape[0] for t in src_t], 3)
    assert [len(c) for c in sc] == [t.shape[0] for t in src_t]
    for g, c in zip(regs, sc):
        g.set_input_source(c)
    pygicp.prepare_batch(regs, pygicp.SOURCE)
    tc = pygicp.Cloud.from_device_batch([t.data_ptr() for t in tgt_t], [t.shape[0] for t in tgt_t], 3)
    for g, c in zip(regs, tc):
        g.set_input_target(c)
    pygicp.prepare_batch(regs, pygicp.TARGET)
    Ts = pygicp.align_batch(regs, guesses)
    assert np.array_equal(Ts, plain)
    covs = regs[0].get_source_covariances(len(sc[0]))
    ref = make(pygicp.FastGICP, pairs[0][0], pairs[0][1]).get_source_covariances(len(sc[0]))
    assert np.array_equal(covs, ref)
    assert pygicp.Cloud.from_device_batch([], 0) == []


def test_edge_cases(gpu):
    src, tgt = synth.make_pair(0, "vlp16")[:2]
    g = pygicp.FastGICP()
    with pytest.raises(Exception):
        g.align()                        # inputs not set -> error status, no crash
    g = make(pygicp.FastGICP, src[:0], tgt)   # empty source: nothing to align, pose stays at the guess
    T = g.align(np.eye(4, dtype=np.float32))
    assert np.isfinite(T).all()
    g = make(pygicp.FastGICP, src, src)       # identical clouds: identity, converged
    T = g.align()
    assert np.abs(T - np.eye(4)).max() < 1e-6 and g.has_converged()
    assert g.get_fitness_score(1.0) == 0.0
    g = make(pygicp.FastGICP, src, tgt, max_corr=1e-3)   # nothing within the correspondence distance
    o = make(oracle.FastGICP, src, tgt, max_corr=1e-3)
    assert np.abs(g.align() - o.align()).max() < TOL_T
    with pytest.raises(Exception):
        g.set_correspondence_randomness(64)   # k limited to 32 (documented)
    g = make(pygicp.FastGICP, src, tgt)       # user-supplied covariances round trip
    c = g.get_source_covariances(len(src))
    g.set_source_covariances(c)
    assert np.array_equal(g.get_source_covariances(len(src)), c)


def test_full_size_120k_properties(gpu):
    """BASELINE config sizes (120 000-pt 64-beam scans): oracle parity of one full align plus size-independent properties:
    inverse consistency (A->B vs B->A), batch determinism and the correspondence nearest-neighbour property."""
    src, tgt, T_gt, T_guess = synth.make_pair(0, "hdl64")[:4]
    assert len(src) == 120000
    g = make(pygicp.FastGICP, src, tgt)
    Tg = g.align(T_guess.astype(np.float32))
    o = make(oracle.FastGICP, src, tgt)
    o.align(T_guess.astype(np.float32))
    assert np.abs(g.get_final_transformation_d() - o.result()["Td"]).max() < TOL_T
    te, re = synth.pose_error(Tg, T_gt)
    assert te < 0.05 and re < np.deg2rad(0.5)
    gb = make(pygicp.FastGICP, tgt, src)
    Tb = gb.align(np.linalg.inv(T_guess).astype(np.float32))
    te, re = synth.pose_error(np.linalg.inv(Tb), Tg)
    assert te < 0.02 and re < np.deg2rad(0.2)            # forward and backward refinements agree
    # every reported correspondence is the true nearest neighbour at the linearisation point: spot-check with the kd-tree
    sel = np.random.default_rng(0).choice(len(src), 2000, replace=False)
    g.evaluate_cost(Tg)
    corr, sqd, _ = g.last_correspondences(len(src))
    Tf = Tg.astype(np.float32)
    moved = (src[sel].astype(np.float64) @ Tf[:3, :3].astype(np.float64).T + Tf[:3, 3]).astype(np.float32)
    bi, bd = oracle.knn(tgt, moved, 1)
    assert (bi[:, 0] == corr[sel]).mean() > 0.995          # (rounding of the re-transformed query can flip exact near-ties)


def test_keyframe_merge_matches_oracle(gpu):
    """mergeNearestKeyframes (transform + concat + PassThrough + VoxelGrid centroids) as the GICP front end:
    host-pointer and device-pointer paths bit-equal to the oracle (same voxel order, same summation order), and the merged cloud
    feeds a registration directly on the device"""
    import torch
    from test_oracle_golden import _merge_case
    for nkf, kind, leaf, lim in [(3, "vlp16", 0.4, 20.0), (5, "hdl64", 0.2, 60.0)]:
        kfs, Ts = _merge_case(nkf, kind)
        exp = oracle.merge_nearest_keyframes(kfs, Ts, lim, leaf)
        got_h = pygicp.merge_nearest_keyframes(kfs, Ts, lim, leaf)
        assert got_h.shape == exp.shape and np.array_equal(got_h, exp)
        dev = [torch.from_numpy(k).cuda() for k in kfs]
        got_d = pygicp.merge_nearest_keyframes(dev, Ts, lim, leaf)
        assert np.array_equal(got_d, exp)
        xyz_only = pygicp.merge_nearest_keyframes([k[:, :3].copy() for k in kfs], Ts, lim, leaf)       # stride 3: intensity = 0
        assert np.array_equal(xyz_only[:, :3], exp[:, :3]) and (xyz_only[:, 3] == 0).all()
    # leaf too small -> unfiltered (PCL warning path); empty input
    kfs, Ts = _merge_case(1, "vlp16")
    assert np.array_equal(pygicp.merge_nearest_keyframes(kfs, Ts, 20.0, 1e-4), oracle.merge_nearest_keyframes(kfs, Ts, 20.0, 1e-4))
    assert len(pygicp.merge_nearest_keyframes([np.zeros((0, 4), np.float32)], np.eye(4, dtype=np.float32)[None])) == 0
    # ICPCheck: both clouds merged on the device, then aligned without leaving it
    kfs, Ts = _merge_case(3, "vlp16")
    T_true = synth.pose_matrix(0.3, -0.2, 0.05, 0.04).astype(np.float32)
    src_t, src_c = pygicp.merge_nearest_keyframes(kfs, Ts, 60.0, 0.4, as_cloud=True)
    tgt_t, tgt_c = pygicp.merge_nearest_keyframes(kfs, np.stack([T_true @ T for T in Ts]), 60.0, 0.4, as_cloud=True)
    g = make(pygicp.FastGICP, src_c, tgt_c, max_corr=2.0)
    T = g.align(np.eye(4, dtype=np.float32))
    o = make(oracle.FastGICP, src_t[:, :3].cpu().numpy(), tgt_t[:, :3].cpu().numpy(), max_corr=2.0)
    To = o.align(np.eye(4, dtype=np.float32))
    assert np.abs(T - To).max() < TOL_T
    assert synth.pose_error(T, T_true)[0] < 0.05


def test_pointcloud2_unpack(gpu):
    """sensor_msgs/PointCloud2 blob -> device x,y,z,intensity (main_RING.py:251-254 does this with a per-point python loop): records of
    odd size with unaligned float fields, NaN records dropped in order; host and device blobs; result usable as a cloud"""
    import torch
    from mr_slam_b200 import msgs
    rng = np.random.default_rng(5)
    n, step = 5000, 22                                           # x@1 y@5 z@9 intensity@14 (unaligned on purpose), 4 bytes of padding
    pts = rng.normal(0, 20, (n, 4)).astype(np.float32)
    pts[rng.random(n) < 0.05, rng.integers(0, 3)] = np.nan
    pts[7, 3] = np.nan
    blob = np.zeros((n, step), np.uint8)
    for j, off in enumerate((1, 5, 9, 14)):
        blob[:, off:off + 4] = pts[:, j:j + 1].copy().view(np.uint8)
    offs = dict(x=1, y=5, z=9, intensity=14)
    exp = pts[~np.isnan(pts).any(axis=1)]
    got = msgs.pointcloud2_unpack(blob.tobytes(), n, step, offs).cpu().numpy()
    assert np.array_equal(got, exp)
    got_d = msgs.pointcloud2_unpack(torch.from_numpy(blob).cuda(), n, step, offs).cpu().numpy()
    assert np.array_equal(got_d, exp)
    keep_all = msgs.pointcloud2_unpack(blob, n, step, offs, skip_nans=False).cpu().numpy()
    assert keep_all.shape == (n, 4) and np.array_equal(np.isnan(keep_all), np.isnan(pts))
    xyz = msgs.pointcloud2_unpack(blob, n, step, dict(x=1, y=5, z=9)).cpu().numpy()          # no intensity field: w = 0
    e3 = pts[~np.isnan(pts[:, :3]).any(axis=1)]
    assert np.array_equal(xyz[:, :3], e3[:, :3]) and (xyz[:, 3] == 0).all()
    t, cloud = msgs.pointcloud2_unpack(blob, n, step, offs, as_cloud=True)
    assert len(cloud) == len(exp)
    idx, _ = cloud.knn(4)
    oi, _ = oracle.knn(exp[:, :3], exp[:, :3], 4)
    assert np.array_equal(idx, oi)
    assert msgs.pointcloud2_unpack(b"", 0, step, offs).shape == (0, 4)

