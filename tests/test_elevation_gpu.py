"""GPU parity tests of the elevation process + fuse path: cell indices bit-exact, heights within 1e-5 (north_star) --
against the CPU oracle and, where oracle/_ref is present, against the reference's own kernels compiled from
gpu_process.cu (with the Eigen shim)."""
import numpy as np
import pytest

import oracle
from mr_slam_b200 import elevation, synth

pytestmark = pytest.mark.gpu


def frame(f, kind):
    pose = synth.pose_matrix(2.0 * f + 0.3, -1.0 * f, 0.1 * f, 0.3 * f)
    pts = synth.make_scan(pose, kind, scan_id=100 + f)
    ei = synth.elevation_inputs(pose)
    args = (ei["T"], -5.0, 5.0, ei["min_r"], ei["beam_a"], ei["beam_c"], ei["sensorJacobian"], ei["rotationVariance"],
            ei["C_SB_transpose"], ei["P_mul_C_BM_transpose"], ei["B_r_BS_skew"])
    x, y, z = (np.ascontiguousarray(pts[:, i]) for i in range(3))
    return pose, x, y, z, args


def colours(n, seed):
    rng = np.random.default_rng(seed)
    R = (rng.integers(1, 255, n) * (rng.random(n) < 0.6)).astype(np.int32)
    G = rng.integers(1, 255, n).astype(np.int32)
    B = rng.integers(1, 255, n).astype(np.int32)
    I = rng.random(n).astype(np.float32)
    return R, G, B, I


@pytest.mark.parametrize("L,res,kind", [(1200, 0.1, "hdl64"), (60, 0.2, "vlp16"), (61, 0.2, "vlp16"), (201, 0.15, "vlp16")])
def test_process_fuse_sequence_vs_oracle(gpu, L, res, kind):
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    o = oracle.ElevationMap(L, res, 2.5, 0.2)
    for f in range(3):
        pose, x, y, z, args = frame(f, kind)
        n = len(x)
        pos = pose[:3, 3].astype(np.float32)
        for a, b in zip(g.Move(pos, res, L), o.move(pos)):
            assert np.array_equal(a, b)
        rg = g.Process_points(x, y, z, args[0], n, *args[1:])
        ro = o.process_points(x, y, z, *args)
        assert np.array_equal(rg[0], ro[0])                    # cell indices: bit-exact
        for a, b in zip(rg[1:], ro[1:]):
            assert np.array_equal(a, b)                        # variance and transformed coordinates: bit-exact too
        assert (rg[0][rg[4] == -1] == -1).all()
        R, G, B, I = colours(n, f)
        nf = g.Fuse(L, n, rg[0], R, G, B, I, rg[4], rg[1])
        assert nf == o.fuse(ro[0], R, G, B, I, ro[4], ro[1])
        mg, mo = g.get_map(), o.get_map()
        assert np.abs(mg["elevation"] - mo["elevation"]).max() <= 1e-5   # heights within 1e-5 (they are in fact identical)
        for k in ("elevation", "variance", "intensity", "R", "G", "B"):
            assert np.array_equal(mg[k], mo[k]), k
        g.Mapvar_update(L, 0.001)
        o.mapvar_update(0.001)
    out_g = g.Map_optmove(np.array([1.234, -0.77], np.float32), 0.05, res, L)
    out_o = o.map_optmove(np.array([1.234, -0.77], np.float32), 0.05)
    assert np.array_equal(out_g, out_o)
    g.Map_closeloop(np.array([3.3, 0.4], np.float32), -0.02, L, res)
    o.map_closeloop(np.array([3.3, 0.4], np.float32), -0.02)
    assert np.array_equal(g.get_map()["elevation"], o.get_map()["elevation"])
    assert all(np.array_equal(a, b) for a, b in zip(g.state(), o.state()))


@pytest.mark.parametrize("jump", [-(2 * 60 + 3), -60, 2 * 60 + 3, -59])
def test_move_large_jump_matches_oracle(gpu, jump):
    """|shift| >= length in either direction is a clear-all (the reference's one-sided test overruns its planes
    for large negative shifts); shifts just below the length still clear strips.  Map state must equal the oracle's."""
    L, res = 60, 0.2
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    o = oracle.ElevationMap(L, res, 2.5, 0.2)
    pose, x, y, z, args = frame(0, "vlp16")
    n = len(x)
    p0 = pose[:3, 3].astype(np.float32)
    for pos in (p0, p0 + np.array([jump * res, 0.0, 0.0], np.float32), p0 + np.array([jump * res, -jump * res, 0.0], np.float32)):
        for a, b in zip(g.Move(pos, res, L), o.move(pos)):
            assert np.array_equal(a, b)
        rg = g.Process_points(x, y, z, args[0], n, *args[1:])
        ro = o.process_points(x, y, z, *args)
        zi = np.zeros(n, np.int32)
        g.Fuse(L, n, rg[0], zi, zi, zi, np.zeros(n, np.float32), rg[4], rg[1])
        o.fuse(ro[0], zi, zi, zi, np.zeros(n, np.float32), ro[4], ro[1])
        mg, mo = g.get_map(), o.get_map()
        for k in ("elevation", "variance", "intensity"):
            assert np.array_equal(mg[k], mo[k]), k


def test_many_points_per_cell_order_dependence(gpu):
    """The Mahalanobis branch makes the result depend on the per-cell point order: feed thousands of points into a handful
    of cells, in an order where sorting by height would give a different answer."""
    rng = np.random.default_rng(3)
    L, n = 8, 20000
    idx = rng.integers(-2, L * L + 3, n).astype(np.int32)          # includes invalid indices on both sides
    h = (rng.normal(0, 0.5, n) + (rng.random(n) < 0.1) * 3.0).astype(np.float32)
    h[rng.random(n) < 0.02] = -1.0
    v = rng.uniform(1e-6, 0.02, n).astype(np.float32)
    R, G, B, I = colours(n, 9)
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, 0.5, 2.5, 0.2)
    o = oracle.ElevationMap(L, 0.5, 2.5, 0.2)
    for rep in range(2):
        assert g.Fuse(L, n, idx, R, G, B, I, h, v) == o.fuse(idx, R, G, B, I, h, v)
        mg, mo = g.get_map(), o.get_map()
        for k in ("elevation", "variance", "intensity", "R", "G", "B"):
            assert np.array_equal(mg[k], mo[k]), k
    perm = rng.permutation(n)   # a different point order gives a different map (the fuse is not commutative) ...
    g2 = elevation.ElevationMap()
    g2.Init_GPU_elevationmap(L, 0.5, 2.5, 0.2)
    g2.Fuse(L, n, idx[perm], R[perm], G[perm], B[perm], I[perm], h[perm], v[perm])
    o2 = oracle.ElevationMap(L, 0.5, 2.5, 0.2)
    o2.fuse(idx[perm], R[perm], G[perm], B[perm], I[perm], h[perm], v[perm])
    assert np.array_equal(g2.get_map()["elevation"], o2.get_map()["elevation"])   # ... and the GPU follows it exactly


def test_device_fused_path_equals_host_path(gpu):
    import torch
    L, res = 1200, 0.1
    pose, x, y, z, args = frame(1, "hdl64")
    n = len(x)
    a = elevation.ElevationMap()
    a.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    b = elevation.ElevationMap()
    b.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    pos = pose[:3, 3].astype(np.float32)
    a.Move(pos, res, L)
    b.Move(pos, res, L)
    r = a.Process_points(x, y, z, args[0], n, *args[1:])
    zi = np.zeros(n, np.int32)
    nf_a = a.Fuse(L, n, r[0], zi, zi, zi, np.zeros(n, np.float32), r[4], r[1])
    d = [torch.from_numpy(t).cuda() for t in (x, y, z)]
    nf_b = b.process_fuse_device(d[0], d[1], d[2], n, *args, near_filter=True)
    assert nf_a == nf_b == int((r[0] >= 0).sum())
    ma, mb = a.get_map(), b.get_map()
    assert np.array_equal(ma["elevation"], mb["elevation"]) and np.array_equal(ma["variance"], mb["variance"])
    # size-independent properties at the full config size: touched cells == distinct valid indices; empty cells are (-10, 1e-4)
    touched = ma["elevation"] != -10
    assert touched.sum() == len(np.unique(r[0][r[0] >= 0]))
    assert (ma["variance"][~touched] == np.float32(0.0001)).all()
    nf_off = b.process_fuse_device(d[0], d[1], d[2], n, *args, near_filter=False)   # throughput mode keeps every in-range point
    assert nf_off > nf_b


def test_pipelined_batch_equals_per_frame_sequence(gpu):
    """mrslam_elev_process_fuse_device_batch (two-stream pipeline over frames) == Move + process_fuse_device per frame, bit for bit
    for elevation / variance and the fused-point count; `lowest` (process kernels, auxiliary stream) touches the same cells"""
    import torch
    L, res = 1200, 0.1
    a = elevation.ElevationMap()
    a.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    b = elevation.ElevationMap()
    b.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    o = oracle.ElevationMap(L, res, 2.5, 0.2)
    frames, keep, total = [], [], 0
    for f in range(5):
        pose, x, y, z, args = frame(f, "hdl64")
        n = len(x) if f != 3 else 0                     # one empty frame in the middle
        pos = pose[:3, 3].astype(np.float32)
        d = [torch.from_numpy(t).cuda() for t in (x, y, z)]
        keep.append(d)
        a.Move(pos, res, L)
        total += a.process_fuse_device(d[0], d[1], d[2], n, *args, near_filter=True)
        frames.append(elevation.ElevationMap.make_frame(d[0], d[1], d[2], n, *args, near_filter=True, position=pos))
        o.move(pos)
        r = o.process_points(x[:n], y[:n], z[:n], *args)
        zi = np.zeros(n, np.int32)
        o.fuse(r[0], zi, zi, zi, np.zeros(n, np.float32), r[4], r[1])
    torch.cuda.synchronize()
    b.fused_total(reset=True)
    b.process_fuse_device_batch(frames, res, L)
    assert b.fused_total(reset=True) == total
    ma, mb, mo = a.get_map(), b.get_map(), o.get_map()
    for key in ("elevation", "variance"):
        assert np.array_equal(ma[key], mb[key]), key
    # `lowest` is updated by a racy read-modify-write inside one frame (in the reference too, gpu_process.cu:434-438): only the
    # set of cells it touched is deterministic
    assert np.array_equal(ma["lowest"] != 100, mb["lowest"] != 100)
    assert np.array_equal(mb["elevation"] != -10, mo["elevation"] != -10)
    assert np.abs(mb["elevation"] - mo["elevation"]).max() <= 1e-5
    b.process_fuse_device_batch([], res, L)             # empty batch is a no-op
    # asynchronous replay (map-writing half on the map's own stream): same map; get_map / fused_total join it
    c = elevation.ElevationMap()
    c.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    c.process_fuse_device_batch(frames, res, L, asynchronous=True)
    mc = c.get_map()
    assert c.fused_total(reset=True) == total
    for key in ("elevation", "variance"):
        assert np.array_equal(ma[key], mc[key]), key
    c.process_fuse_device_batch(frames[:2], res, L, asynchronous=True)      # two replays back to back, then a per-frame call
    c.process_fuse_device_batch(frames[2:], res, L, asynchronous=True)
    a.process_fuse_device_batch(frames, res, L)
    assert np.array_equal(a.get_map()["elevation"], c.get_map()["elevation"])


@pytest.mark.parametrize("L,res,nframes,rows", [(60, 0.2, 70, None), (201, 0.15, 9, None), (120, 0.25, 12, (30, 95))])
def test_onepass_replay_with_moves_equals_per_frame_sequence(gpu, L, res, nframes, rows):
    """One-pass replay of a whole group of frames (one sort + one run walk with in-run frame-end clamps and Move clear masks) ==
    the per-frame sequence Move, process_fuse_device -- with Moves that wipe cells holding points of earlier frames of the same
    batch, a jump of more than a map length in the middle (clear-all), more than 64 frames (two groups), frames without a Move,
    an empty frame, and (third case) a map restricted to a block of storage rows.  Elevation, variance and the count bit for bit."""
    import torch
    a = elevation.ElevationMap()
    a.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    b = elevation.ElevationMap()
    b.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    if rows:
        a.set_owned_rows(*rows)
        b.set_owned_rows(*rows)
    rng = np.random.default_rng(11)
    frames, keep, total = [], [], 0
    base = [frame(f, "vlp16") for f in range(4)]
    pos = np.zeros(3, np.float32)
    for f in range(nframes):
        pose, x, y, z, args = base[f % 4]
        n = len(x) if f % 7 != 5 else 0
        # a random walk of a few cells per frame (wraps the ring buffer several times), one long jump, some frames without a Move
        step = rng.integers(-6, 7, 2) * res
        if f == nframes // 2:
            step = np.array([-(L + 5) * res, 0.0])
        pos = (pos + np.array([step[0], step[1], 0.0])).astype(np.float32)
        T = np.array(args[0], np.float32).copy()
        T[0, 3], T[1, 3] = pos[0] + 0.3, pos[1] - 0.2           # the scan moves with the robot
        args_f = (T,) + tuple(args[1:])
        do_move = f % 5 != 4
        d = [torch.from_numpy(t).cuda() for t in (x, y, z)]
        keep.append(d)
        if do_move:
            a.Move(pos, res, L)
        total += a.process_fuse_device(d[0], d[1], d[2], n, *args_f, near_filter=(f % 3 != 0))
        frames.append(elevation.ElevationMap.make_frame(d[0], d[1], d[2], n, *args_f, near_filter=(f % 3 != 0), position=pos if do_move else None))
    torch.cuda.synchronize()
    b.fused_total(reset=True)
    b.process_fuse_device_batch(frames, res, L)
    assert b.fused_total(reset=True) == total and total > 0
    ma, mb = a.get_map(), b.get_map()
    for key in ("elevation", "variance", "intensity", "R"):
        assert np.array_equal(ma[key], mb[key]), key
    assert all(np.array_equal(u, v) for u, v in zip(a.state(), b.state()))
    assert (ma["elevation"] != -10).sum() > 50


@pytest.mark.parametrize("n,bits", [(0, 21), (1, 21), (2047, 8), (2048, 21), (2049, 24), (100003, 21), (1 << 20, 24), (3000000, 21)])
def test_own_radix_sort_is_a_stable_sort_with_drop(gpu, n, bits):
    """csrc/radix_sort.cuh (the sort of the elevation replay, device-side counts, sentinel drop in the first pass) against numpy's
    stable argsort: same keys, same values in the same order -- i.e. equal keys keep their input order -- and the dropped sentinel."""
    import ctypes as C
    from mr_slam_b200._lib import check, lib, ptr
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, max(hi // 7, 2), n, dtype=np.uint32) if n % 2 else rng.integers(0, hi, n, dtype=np.uint32)   # many / few duplicates
    drop = np.uint32(hi)
    keys[rng.random(n) < 0.3] = drop
    vals = np.arange(n, dtype=np.uint32)
    for use_drop in (1, 0):
        ko, vo, m = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32), C.c_int(-1)
        check(lib().mrslam_sort_pairs_u32(ptr(keys) if n else None, ptr(vals) if n else None, n, bits, int(drop), use_drop, ptr(ko), ptr(vo), C.byref(m)))
        keep = keys != drop if use_drop else np.ones(n, bool)
        order = np.argsort(keys[keep], kind="stable")
        assert m.value == int(keep.sum())
        assert np.array_equal(ko[:m.value], keys[keep][order]) and np.array_equal(vo[:m.value], vals[keep][order])


def test_empty_and_invalid_inputs(gpu):
    g = elevation.ElevationMap()
    with pytest.raises(Exception):
        g.Fuse(10, 0, [], [], [], [], [], [], [])          # Init not called
    g.Init_GPU_elevationmap(10, 0.5, 2.5, 0.2)
    e = np.zeros(0, np.float32)
    assert g.Fuse(10, 0, e.astype(np.int32), e.astype(np.int32), e.astype(np.int32), e.astype(np.int32), e, e, e) == 0
    m = g.get_map()
    assert (m["elevation"] == -10).all() and (m["variance"] == np.float32(0.0001)).all()
    with pytest.raises(Exception):
        g.Fuse(11, 0, e.astype(np.int32), e.astype(np.int32), e.astype(np.int32), e.astype(np.int32), e, e, e)   # wrong length


def test_against_reference_kernels(gpu):
    """The reference's own gpu_process.cu kernels (compiled unmodified against the Eigen shim into oracle/_ref) run next to
    ours on the same inputs: indices bit-exact, heights within 1e-5.  Small map: the reference's G_fuse is O(cells x points)."""
    from oracle import ref
    if not ref.available("elev"):
        pytest.skip("oracle/_ref/libelev_ref.so not present")
    L, res = 100, 0.2
    rk = ref.ElevationRef(L, res)
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    for f in range(3):
        pose, x, y, z, args = frame(f, "vlp16")
        n = len(x)
        pos = pose[:3, 3].astype(np.float32)
        for a, b in zip(g.Move(pos, res, L), rk.move(pos)):
            assert np.array_equal(a, b)
        rg = g.Process_points(x, y, z, args[0], n, *args[1:])
        rr = rk.process_points(x, y, z, *args)
        assert np.array_equal(rg[0], rr[0])                               # cell indices bit-exact vs the reference kernel
        assert np.abs(rg[4] - rr[4]).max() <= 1e-6 and np.abs(rg[2] - rr[2]).max() <= 1e-5
        assert np.abs(rg[1] - rr[1]).max() <= 1e-6 * max(1.0, np.abs(rr[1]).max())
        R, G, B, I = colours(n, f)
        g.Fuse(L, n, rg[0], R, G, B, I, rg[4], rg[1])
        rk.fuse(rr[0], R, G, B, I, rr[4], rr[1])
        mg, mr = g.get_map(), rk.get_map()
        assert np.abs(mg["elevation"] - mr["elevation"]).max() <= 1e-5
        assert np.abs(mg["variance"] - mr["variance"]).max() <= 1e-6
        for k in ("R", "G", "B"):
            assert np.array_equal(mg[k], mr[k])
        g.Mapvar_update(L, 0.002)
        rk.mapvar_update(0.002)


def test_map_feature_and_raytracing_vs_reference_kernels(gpu):
    """the traversability read-out (Map_feature) and the ray-tracing clean-up (Raytracing) that the node runs after
    every Fuse, next to the reference's own kernels (oracle/_ref) over a sequence of frames.  The reference leaves rough / slope /
    traver unwritten where elevation == -10, so those planes are compared on cells with elevation only."""
    from oracle import ref
    if not ref.available("elev"):
        pytest.skip("oracle/_ref/libelev_ref.so not present")
    L, res = 100, 0.2
    rk = ref.ElevationRef(L, res)
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    removed_total, obstacles_total = 0, 0
    for f in range(4):
        pose, x, y, z, args = frame(f, "vlp16")
        n = len(x)
        pos = pose[:3, 3].astype(np.float32)
        g.Move(pos, res, L)
        rk.move(pos)
        rg = g.Process_points(x, y, z, args[0], n, *args[1:])
        rr = rk.process_points(x, y, z, *args)
        assert np.array_equal(rg[0], rr[0])
        R, G, B, I = colours(n, f)
        g.Fuse(L, n, rg[0], R, G, B, I, rg[4], rg[1])
        rk.fuse(rr[0], R, G, B, I, rr[4], rr[1])
        fg, fr = g.Map_feature(L), rk.map_feature()
        has = fr["elevation"] != -10
        assert np.array_equal(fg["elevation"] != -10, has)
        assert np.abs(fg["elevation"] - fr["elevation"]).max() <= 1e-5 and np.abs(fg["var"] - fr["var"]).max() <= 1e-6
        for k in ("colorR", "colorG", "colorB"):
            assert np.array_equal(fg[k], fr[k]), k
        assert np.array_equal(fg["intensity"], fr["intensity"])
        # cells whose heights are bit-identical must give bit-identical features (same arithmetic, same contraction)
        same = has & (fg["elevation"] == fr["elevation"])
        assert same.sum() > 0.9 * has.sum()
        for k in ("rough", "slope", "traver"):
            d = np.abs(fg[k][has] - fr[k][has])
            assert d.max() <= 2e-4, (k, float(d.max()))
            assert (d == 0).mean() > 0.99, (k, float((d == 0).mean()))
        assert (fg["traver"][~has] == -10).all() and (fg["slope"][~has] == 0).all()       # our convention for empty cells
        obstacles_total += int(((fr["traver"] < 0.2) & has).sum())
        g.Raytracing(L)
        rk.raytracing()
        ag, ar = g.Map_feature(L), rk.map_feature()
        assert np.array_equal(ag["elevation"] != -10, ar["elevation"] != -10)             # the same cells were cleaned up
        assert np.abs(ag["elevation"] - ar["elevation"]).max() <= 1e-5
        removed_total += int((has & (ar["elevation"] == -10)).sum())
        assert (g.get_map()["lowest"] == 10).all()                                        # G_Clear_maplowest
    assert obstacles_total > 0            # the ray walk was exercised
    print("raytracing removed %d cells of %d obstacle cells over the sequence" % (removed_total, obstacles_total))


def test_map_feature_and_raytracing_vs_oracle_full_size(gpu):
    """the same post-passes at the benchmark's map size against the CPU twin (CPU libm: agreement to ~1e-6, with a few cells where a
    Jacobi pivot / eigenvalue ordering flips on a rounding difference)"""
    L, res = 1200, 0.1
    g = elevation.ElevationMap()
    g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    o = oracle.ElevationMap(L, res, 2.5, 0.2)
    for f in range(2):
        pose, x, y, z, args = frame(f, "hdl64")
        n = len(x)
        pos = pose[:3, 3].astype(np.float32)
        g.Move(pos, res, L)
        o.move(pos)
        rg = g.Process_points(x, y, z, args[0], n, *args[1:])
        ro = o.process_points(x, y, z, *args)
        zi = np.zeros(n, np.int32)
        g.Fuse(L, n, rg[0], zi, zi, zi, np.zeros(n, np.float32), rg[4], rg[1])
        o.fuse(ro[0], zi, zi, zi, np.zeros(n, np.float32), ro[4], ro[1])
        fg, fo = g.Map_feature(L), o.map_feature()
        if f == 0:
            assert np.array_equal(fg["elevation"], fo["elevation"]) and np.array_equal(fg["var"], fo["var"])
        else:   # a borderline clean-up decision of the previous frame may differ (CPU libm), and its 5x5 neighbourhood with it
            assert (fg["elevation"] != fo["elevation"]).sum() <= 10
        has = fo["elevation"] != -10
        assert ((fg["traver"] == -10) != (fo["traver"] == -10)).sum() <= (0 if f == 0 else 250)
        for k in ("rough", "slope", "traver"):
            d = np.abs(fg[k] - fo[k])
            assert (d <= 1e-4).mean() > 0.999, (k, float((d <= 1e-4).mean()), float(d.max()))
        g.Raytracing(L)
        o.raytracing()
        eg, eo = g.get_map()["elevation"], o.get_map()["elevation"]
        assert ((eg != -10) != (eo != -10)).sum() <= max(3, int(1e-4 * has.sum()))       # borderline cells of the height test only
        assert (g.get_map()["lowest"] == 10).all()


def test_tile_ownership_rows_and_halo(gpu):
    """BASELINE config 5 primitives on the GPU: set_owned_rows (points of other tiles are outside the map), rows copy in both directions,
    and Map_feature on a tile whose halo rows were filled from the neighbouring tile -- against the CPU oracle twins and against the
    whole map: union of two tiles == whole map bit for bit (elevation, variance), features on owned rows identical."""
    import torch
    from mr_slam_b200 import dist as mdist
    L, res = 201, 0.15
    whole = elevation.ElevationMap()
    whole.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    tiles = []
    for r in range(2):
        m = elevation.ElevationMap()
        m.Init_GPU_elevationmap(L, res, 2.5, 0.2)
        tiles.append((m, mdist.TiledElevationMap(m, L, rank=r, world=2, device="cuda")))
    otile = oracle.ElevationMap(L, res, 2.5, 0.2)
    otile.set_owned_rows(tiles[1][1].row_lo, tiles[1][1].row_hi)
    for f in range(3):
        pose, x, y, z, args = frame(f, "vlp16")
        n = len(x)
        pos = pose[:3, 3].astype(np.float32)
        zi = np.zeros(n, np.int32)
        for m in [whole] + [t[0] for t in tiles]:
            m.Move(pos, res, L)
            r = m.Process_points(x, y, z, args[0], n, *args[1:])
            m.Fuse(L, n, r[0], zi, zi, zi, np.zeros(n, np.float32), r[4], r[1])
            if m is tiles[1][0]:
                r_tile = r
        otile.move(pos)
        ro = otile.process_points(x, y, z, *args)
        otile.fuse(ro[0], zi, zi, zi, np.zeros(n, np.float32), ro[4], ro[1])
        assert np.array_equal(r_tile[0], ro[0])                       # ownership filter: same indices as the oracle twin
    W = whole.get_map()
    union_e = np.empty((L, L), np.float32)
    union_v = np.empty((L, L), np.float32)
    for m, t in tiles:
        g = m.get_map()
        e = g["elevation"].reshape(L, L)
        union_e[t.row_lo:t.row_hi] = e[t.row_lo:t.row_hi]
        union_v[t.row_lo:t.row_hi] = g["variance"].reshape(L, L)[t.row_lo:t.row_hi]
        outside = np.ones(L, bool)
        outside[t.row_lo:t.row_hi] = False
        assert (e[outside] == -10).all()
    assert np.array_equal(union_e.reshape(-1), W["elevation"]) and np.array_equal(union_v.reshape(-1), W["variance"])
    assert np.array_equal(tiles[1][0].get_map()["elevation"], otile.get_map()["elevation"])
    # rows copy round trip + manual halo exchange between the two tile instances (what TiledElevationMap does through all_gather)
    h = 2
    (m0, t0), (m1, t1) = tiles
    buf = torch.empty((h, L), dtype=torch.float32, device="cuda")
    for src, dst, src_row, dst_row in [(m1, m0, t1.row_lo, t0.row_hi), (m0, m1, t0.row_hi - h, t1.row_lo - h),
                                       (m1, m0, t1.row_hi - h, (t0.row_lo - h) % L), (m0, m1, t0.row_lo, t1.row_hi % L)]:
        for j in range(h):
            src.rows_to_device("elevation", (src_row + j) % L, 1, buf[j])
            dst.rows_from_device("elevation", (dst_row + j) % L, 1, buf[j])
    fw = whole.Map_feature(L)
    for m, t in tiles:
        ft = m.Map_feature(L)
        for k in ("elevation", "rough", "slope", "traver"):
            assert np.array_equal(ft[k].reshape(L, L)[t.row_lo:t.row_hi], fw[k].reshape(L, L)[t.row_lo:t.row_hi]), k
    with pytest.raises(Exception):
        m0.set_owned_rows(5, L + 1)

