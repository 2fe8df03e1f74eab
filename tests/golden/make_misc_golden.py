"""Generate tests/golden/{polar,ring}_fixture.npz in the authoring container (needs /root/reference and oracle/_ref).

polar: the reference's own fixtures (DISCO .../multi-layer-polar-cpu/cython/{1,2}.bin, BEV/test.bin; 4096x3 float64 already in
       [-1,1]) run through the reference's UNMODIFIED CPU twin of the polar voxeliser (oracle/_ref/libpolar_cpu_ref.so).
ring : the reference's python functions themselves -- LoopDetection/src/RING_ros/util.py is imported (its CUDA-extension and
       plotting imports are stubbed out, nothing else) and `fast_corr`, `fftshift1d`, `load_pc_infer` plus the two tail lines
       of `generate_RING` (util.py:197-198) are executed on CPU tensors.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
REF = "/root/reference/LoopDetection"


def polar():
    from oracle import ref
    out = {}
    files = {"disco1": REF + "/src/disco_ros/tools/multi-layer-polar-cpu/cython/1.bin",
             "disco2": REF + "/src/disco_ros/tools/multi-layer-polar-cpu/cython/2.bin"}
    cfgs = [(1, 1, 40, 120, 20, 1), (1, 1, 40, 120, 1, 1), (1, 1, 40, 120, 4, 3)]
    for name, f in files.items():
        pts = np.fromfile(f, dtype=np.float64).reshape(-1, 3).astype(np.float32)
        out[name + "_points"] = pts
        soa = np.ascontiguousarray(pts.T.reshape(-1))
        for ci, c in enumerate(cfgs):
            o = ref.polar_cpu(soa, len(pts), *c)
            nz = np.nonzero(o)[0].astype(np.int32)
            out["%s_cfg%d_idx" % (name, ci)] = nz
            out["%s_cfg%d_val" % (name, ci)] = o[nz]
    out["cfgs"] = np.array(cfgs, np.int32)
    np.savez_compressed(os.path.join(HERE, "polar_fixture.npz"), **out)
    print("polar fixture:", {k: v.shape for k, v in out.items() if k.endswith("points")})


def ring():
    import torch
    # stub the imports of util.py that need a GPU extension / plotting; everything else is the reference's code
    for name in ("voxelocc", "voxelfeat", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "skimage",
                 "torch_radon", "torchvision", "torchvision.transforms", "torchvision.transforms.functional"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["skimage"].morphology = None
    for n in ("Radon", "ParallelBeam", "RadonFanbeam"):
        setattr(sys.modules["torch_radon"], n, object)
    # torchvision.transforms.functional.normalize on a tensor: tensor.sub_(mean).div_(std) (torchvision/transforms/_functional_tensor.py)
    sys.modules["torchvision.transforms.functional"].normalize = lambda t, mean, std: (t - mean) / std
    sys.path.insert(0, REF + "/src/RING_ros")
    import util  # the reference module

    rng = np.random.default_rng(11)
    sino = (rng.random((2, 1, 120, 120)) * 30 * (rng.random((2, 1, 120, 120)) < 0.3)).astype(np.float32)
    out = {"sino": sino}
    tir = []
    for i in range(2):
        pc_RING = torch.from_numpy(sino[i])
        pc_RING_normalized = util.fn.normalize(pc_RING, mean=pc_RING.mean(), std=pc_RING.std())   # util.py:197
        pc_TIRING = torch.fft.fft2(pc_RING_normalized, dim=-2, norm="ortho")                          # util.py:198
        tir.append(pc_TIRING.numpy())
    out["tiring"] = np.stack(tir)
    d, a = [], []
    for i in range(2):
        for j in range(2):
            dist, angle = util.fast_corr(torch.from_numpy(tir[i]), torch.from_numpy(tir[j]))          # util.py:362-374
            d.append(float(dist))
            a.append(int(angle))
    out["fast_corr_dist"] = np.array(d, np.float32).reshape(2, 2)
    out["fast_corr_angle"] = np.array(a, np.int32).reshape(2, 2)
    # a shifted copy: rolling the sinogram along the angle axis by s must be recovered as the yaw
    rolled = np.roll(sino[0], 17, axis=-2)
    pr = torch.from_numpy(rolled)
    trr = torch.fft.fft2(util.fn.normalize(pr, mean=pr.mean(), std=pr.std()), dim=-2, norm="ortho")
    dist, angle = util.fast_corr(torch.from_numpy(tir[0]), trr)
    out["rolled_by"] = np.int32(17)
    out["rolled_dist"] = np.float32(dist)
    out["rolled_angle"] = np.int32(angle)
    # load_pc_infer on a synthetic cloud
    pc = (rng.random((5000, 4)).astype(np.float32) - 0.5) * np.array([200, 200, 80, 1], np.float32)
    out["pc"] = pc
    out["pc_infer"] = util.load_pc_infer(pc)
    # solve_translation (util.py:388-423) after the row roll of main_RING.py:171-172, run by the reference itself; the per-row
    # shifts b are captured by wrapping util.solve_overdetermined_linear_system (its second argument)
    captured = {}
    orig_lsq = util.solve_overdetermined_linear_system

    def spy(A, b, method="pinv"):
        captured["b"] = b.clone().numpy()
        return orig_lsq(A, b, method=method)
    util.solve_overdetermined_linear_system = spy
    H = W = 120
    cases = []
    for ci, (tx, ty, rot, rs, nch) in enumerate([(7.0, -4.0, 0.0, 0, 1), (-11.0, 5.5, 0.6, 13, 1), (3.0, 9.0, -1.1, -29, 2)]):
        base = (rng.random((nch, H, W)) * (rng.random((nch, H, W)) < 0.15)).astype(np.float32) * 20
        ker = np.array([0.25, 0.5, 0.25], np.float32)
        base = sum(np.roll(base, d - 1, axis=2) * ker[d] for d in range(3))          # a little smoothing along the detector axis
        ang = np.linspace(0, 2 * np.pi, H).astype(np.float32) + np.float32(rot)
        pos = np.stack([np.stack([np.roll(base[c, i], -int(round(tx * np.cos(ang[i]) + ty * np.sin(ang[i])))) for i in range(H)])
                        for c in range(nch)]).astype(np.float32)
        pos += (rng.random(pos.shape).astype(np.float32) - 0.5) * 0.05
        pos_unrolled = np.roll(pos, -rs, axis=1)                                      # what the database holds (RING_matched)
        q_t = torch.from_numpy(base)
        p_t = torch.roll(torch.from_numpy(pos_unrolled), rs, dims=1)                  # main_RING.py:171
        x, y, err = util.solve_translation(q_t, p_t, float(rot), util.device)
        out["st%d_query" % ci] = base
        out["st%d_positive" % ci] = pos_unrolled
        out["st%d_params" % ci] = np.array([rot, rs], np.float32)
        out["st%d_b" % ci] = captured["b"].astype(np.float32)
        out["st%d_xye" % ci] = np.array([float(x), float(y), float(err)], np.float32)
        cases.append((float(x), float(y), float(err)))
    util.solve_overdetermined_linear_system = orig_lsq
    print("solve_translation goldens (x, y, err):", cases)
    # RING++ front end: the reference's build_neighbors_NN (sklearn kd_tree kNN + covariation_eigenvalue), util.py:124-170
    from mr_slam_b200 import synth
    scan = synth.make_scan(synth.pose_matrix(1.0, 2.0, 0.0, 0.3), "vlp16", scan_id=77).astype(np.float32)
    pcn = util.load_pc_infer(np.hstack([scan, np.zeros((len(scan), 1), np.float32)]))[:, :3]
    pcn = np.ascontiguousarray(pcn[:: max(1, len(pcn) // 2500)][:2500], np.float32)
    idx, covs, entropy, eigens, _ = util.build_neighbors_NN(pcn, 30)
    out["nn_pc"] = pcn
    out["nn_idx"] = idx.astype(np.int32)
    out["nn_cov"] = covs.astype(np.float32)
    out["nn_entropy"] = entropy.astype(np.float32)
    out["nn_eigens"] = eigens.astype(np.float32)
    print("build_neighbors_NN golden:", pcn.shape, idx.shape, eigens.shape, covs.dtype, eigens.dtype)
    np.savez_compressed(os.path.join(HERE, "ring_fixture.npz"), **out)
    print("ring fixture: fast_corr dist", out["fast_corr_dist"].tolist(), "angle", out["fast_corr_angle"].tolist(), "rolled ->", int(angle))


if __name__ == "__main__":
    polar()
    ring()
