"""Development check on the GPU box: elevation / BEV / Radon / RING parity vs the oracle + rough timings."""
import os, sys, time
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import oracle
from mr_slam_b200 import synth, _lib, elevation, voxelocc, gputransform, ring
from mr_slam_b200.torch_radon import ParallelBeam
_lib.init(0)

def elev_case(L, res, kind, nframes=3):
    print(f"--- elevation L={L} res={res} {kind}")
    g = elevation.ElevationMap(); g.Init_GPU_elevationmap(L, res, 2.5, 0.2)
    o = oracle.ElevationMap(L, res, 2.5, 0.2)
    rng = np.random.default_rng(5)
    for f in range(nframes):
        pose = synth.pose_matrix(2.0 * f + 0.3, -1.0 * f, 0.0, 0.3 * f)
        pts = synth.make_scan(pose, kind, scan_id=100 + f)
        n = len(pts)
        x, y, z = (np.ascontiguousarray(pts[:, i]) for i in range(3))
        ei = synth.elevation_inputs(pose)
        pos = np.array([pose[0, 3], pose[1, 3], pose[2, 3]], np.float32)
        mg = g.Move(pos, res, L); mo = o.move(pos)
        assert all(np.array_equal(a, b) for a, b in zip(mg, mo)), (mg, mo)
        args = (ei["T"], -5.0, 5.0, ei["min_r"], ei["beam_a"], ei["beam_c"], ei["sensorJacobian"], ei["rotationVariance"],
                ei["C_SB_transpose"], ei["P_mul_C_BM_transpose"], ei["B_r_BS_skew"])
        t = time.time(); rg = g.Process_points(x, y, z, args[0], n, *args[1:]); tg = time.time() - t
        ro = o.process_points(x, y, z, *args)
        names = ["map_index", "var", "x_ts", "y_ts", "z_ts"]
        print(f" frame {f}: process exact:", {k: bool(np.array_equal(a, b)) for k, a, b in zip(names, rg, ro)},
              f"valid {int((rg[0] >= 0).sum())}/{n} ({tg*1e3:.2f} ms)")
        R = rng.integers(0, 255, n).astype(np.int32) * (rng.random(n) < 0.5); G = rng.integers(1, 255, n).astype(np.int32); B = G.copy()
        I = rng.random(n).astype(np.float32)
        t = time.time(); nf_g = g.Fuse(L, n, rg[0], R, G, B, I, rg[4], rg[1]); tf = time.time() - t
        nf_o = o.fuse(ro[0], R, G, B, I, ro[4], ro[1])
        mgp, mop = g.get_map(), o.get_map()
        print(f"          fuse n_fused {nf_g}/{nf_o} ({tf*1e3:.2f} ms) map exact:",
              {k: bool(np.array_equal(mgp[k], mop[k])) for k in ("elevation", "variance", "intensity", "R", "G", "B")},
              "max |dh|", float(np.abs(mgp["elevation"] - mop["elevation"]).max()))
        g.Mapvar_update(L, 0.001); o.mapvar_update(0.001)
    # device fused path timing
    d = [torch.from_numpy(np.ascontiguousarray(pts[:, i])).cuda() for i in range(3)]
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        nf = g.process_fuse_device(d[0], d[1], d[2], n, *args, near_filter=True)
        dt = time.time() - t
    print(f"          device process+fuse: {dt*1e6:.0f} us, {nf} cell fuses -> {nf/dt/1e6:.1f} M cell-fuses/s")

elev_case(1200, 0.1, "hdl64")
elev_case(60, 0.2, "vlp16")
elev_case(61, 0.2, "vlp16")

print("--- voxelocc / polar")
rng = np.random.default_rng(3)
for name, pts in (("random", rng.uniform(-1.05, 1.05, (50000, 3)).astype(np.float32)),
                  ("scan", synth.ring_preprocess(synth.make_scan(synth.pose_matrix(1, 2, 0, 0.4), "hdl64", 7) + np.array([0, 0, 1.8], np.float32)))):
    soa = np.ascontiguousarray(pts.T.reshape(-1)); n = len(pts)
    for nh, el in ((1, 1), (4, 1)):
        t = voxelocc.GPUTransformer(soa, n, 1, 1, 120, 120, nh, el); t.transform(); og = t.retreive()
        oo = oracle.voxelocc(soa, n, 1, 1, 120, 120, nh, el)
        print(f" voxelocc {name} n={n} num_height={nh}: exact {np.array_equal(og, oo)} nonzero {int((og != 0).sum())}")
    for nh, el in ((1, 1), (20, 1), (4, 3)):
        mh = 1
        t = gputransform.GPUTransformer(soa, n, 1, mh, 40, 120, nh, el); t.transform(); og = t.retreive()
        oo, idx, nd = oracle.polar(soa, n, 1, mh, 40, 120, nh, el, want_idx=True)
        print(f" polar {name} n={n} num_height={nh} enough_large={el}: exact {np.array_equal(og, oo)} mismatches {int((og != oo).sum())} dropped {nd}")

print("--- radon / ring")
img = rng.random((4, 120, 120)).astype(np.float32) * (rng.random((4, 120, 120)) < 0.2)
ang = np.linspace(0, 2 * np.pi, 120).astype(np.float32)
x = torch.from_numpy(img).cuda()
for interp, wb in (("exact", 0), ("texture", 8)):
    pb = ParallelBeam(120, ang, interp=interp, fast_trig=False)
    sg = pb.forward(x).cpu().numpy()
    so = oracle.radon_forward(img, ang, 120, 1.0, wb)
    print(f" radon {interp}: max abs diff {np.abs(sg - so).max():.3e} rel {np.abs(sg - so).max() / np.abs(so).max():.3e}")
pbf = ParallelBeam(120, ang, interp="exact", fast_trig=True)
print(" radon fast_trig vs accurate rel:", float(np.abs(pbf.forward(x).cpu().numpy() - so_ex).max() / np.abs(so_ex).max()) if (so_ex := oracle.radon_forward(img, ang, 120, 1.0, 0)) is not None else None)
sino = oracle.radon_forward(img, ang, 120, 1.0, 0)
tg = ring.tiring(torch.from_numpy(sino[:, None]).cuda()).cpu().numpy()
to = np.stack([oracle.ring_tiring(sino[i:i + 1])[1] for i in range(4)])
print(" tiring max abs diff", np.abs(tg - to).max(), "scale", np.abs(to).max())
dist, angle, corr = ring.fast_corr_batch(torch.from_numpy(to[0]).cuda(), torch.from_numpy(to).cuda(), want_corr=True)
for j in range(4):
    do, ao, co = oracle.ring_fast_corr(to[0], to[j])
    print(f"  fast_corr[{j}] dist {dist[j].item():.6f}/{do:.6f} angle {angle[j].item()}/{ao} corr max diff {np.abs(corr[j].cpu().numpy() - co).max():.3e}")
# throughput of the batched descriptor pipeline
B = 256
bev = torch.from_numpy((rng.random((B, 120, 120)) * (rng.random((B, 120, 120)) < 0.15)).astype(np.float32)).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    s, tir = ring.generate_RING_batch(bev)
    dt = time.time() - t
print(f" generate_RING_batch {B} images: {dt*1e3:.2f} ms -> {B/dt:.0f} descriptors/s")
