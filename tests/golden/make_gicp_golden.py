"""Generate tests/golden/gicp_fixture.npz from the reference's own registration fixture.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_gicp_golden.py
Inputs : /root/reference/Mapping/src/fast_gicp/data/{251370668,251371071}.pcd, relative.txt
         (the fixture of FG/src/test/gicp_test.cpp:29-65).
Steps  : read the binary PCDs (FIELDS x y z intensity, 16 B/pt), apply a restatement of
         pcl::VoxelGrid(0.2) (PCL 1.10 voxel_grid.hpp: floor(p/leaf)-min_b linear index, centroid
         per voxel in float, output in ascending voxel index) as gicp_test.cpp:55-65 does, and
         freeze the two downsampled clouds + relative pose.  The oracle's transforms for the
         forward / backward alignments (identity guess, library defaults as in the gtest) are
         stored next to them as the known answers that the CUDA path is compared with.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
REF = "/root/reference/Mapping/src/fast_gicp/data"


def read_pcd(path):
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("POINTS"):
                n = int(line.split()[1])
            if line.startswith("DATA"):
                assert line.split()[1] == "binary"
                break
        raw = np.frombuffer(f.read(n * 16), dtype=np.float32).reshape(n, 4)
    return raw[:, :3].copy()


def voxel_grid(pts, leaf):
    pts = pts[np.isfinite(pts).all(1)].astype(np.float32)
    inv = (np.float32(1.0) / np.float32(leaf)).astype(np.float32)
    mn, mx = pts.min(0), pts.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div_b = max_b - min_b + 1
    mul = np.array([1, div_b[0], div_b[0] * div_b[1]], np.int64)
    ijk = np.floor(pts * inv).astype(np.int64) - min_b
    idx = ijk @ mul
    order = np.argsort(idx, kind="stable")
    idx_s, pts_s = idx[order], pts[order]
    uniq, start, counts = np.unique(idx_s, return_index=True, return_counts=True)
    sums = np.add.reduceat(pts_s.astype(np.float32), start, axis=0).astype(np.float32)
    return (sums / counts[:, None].astype(np.float32)).astype(np.float32)


def approx_voxel_grid(pts, leaf, histsize=512):
    """pcl::ApproximateVoxelGrid (PCL 1.10 approximate_voxel_grid.hpp) restated: 512-entry hash history
    keyed (ix*7171 + iy*3079 + iz*4231) & 511, flushed on collision, remaining entries flushed at the end."""
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts.astype(np.float32) * inv).astype(np.int64)
    h = (ijk[:, 0] * 7171 + ijk[:, 1] * 3079 + ijk[:, 2] * 4231) & (histsize - 1)
    cnt = np.zeros(histsize, np.int64)
    key = np.zeros((histsize, 3), np.int64)
    cen = np.zeros((histsize, 3), np.float32)
    out = []
    for i in range(len(pts)):
        hh = h[i]
        if cnt[hh] and (key[hh] != ijk[i]).any():
            out.append(cen[hh] / np.float32(cnt[hh]))
            cnt[hh] = 0
            cen[hh] = 0
        key[hh] = ijk[i]
        cnt[hh] += 1
        cen[hh] += pts[i]
    for hh in range(histsize):
        if cnt[hh]:
            out.append(cen[hh] / np.float32(cnt[hh]))
    return np.array(out, np.float32)


def readme_fixture(oracle):
    """FG/src/align.cpp:128-149 preprocessing (drop |p|^2 < 1e-3, ApproximateVoxelGrid 0.1) and the
    README known answer: fgicp_mt fitness_score 0.204412, fgicp_st 0.204379 (FG/README.md:110-113)."""
    clouds = []
    for f in ("251370668.pcd", "251371071.pcd"):
        p = read_pcd(os.path.join(REF, f))
        p = p[(p.astype(np.float32) ** 2).sum(1) >= 1e-3]
        clouds.append(approx_voxel_grid(p, 0.1))
    tgt, src = clouds
    print("approx voxel grid 0.1:", tgt.shape, src.shape, "(README: 17249 / 17518)")
    g = oracle.FastGICP()
    g.set_num_threads(4)
    g.set_input_target(tgt)
    g.set_input_source(src)
    g.align()
    fit = g.get_fitness_score()
    print("README fixture: fitness %.6f (README fgicp_mt 0.204412, fgicp_st 0.204379)" % fit)
    assert abs(fit - 0.204412) < 5e-4, "oracle misses the README known answer"
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "gicp_readme_fixture.npz"), target=tgt, source=src,
                        T=g.result()["Td"], fitness=fit, readme_fitness_mt=0.204412, readme_fitness_st=0.204379)


def main():
    import oracle
    from mr_slam_b200 import synth
    tgt = voxel_grid(read_pcd(os.path.join(REF, "251370668.pcd")), 0.2)
    src = voxel_grid(read_pcd(os.path.join(REF, "251371071.pcd")), 0.2)
    rel = np.loadtxt(os.path.join(REF, "relative.txt")).reshape(4, 4)
    print("downsampled:", tgt.shape, src.shape)
    out = dict(target=tgt, source=src, relative=rel)
    for name, (s, t) in dict(forward=(src, tgt), backward=(tgt, src)).items():
        g = oracle.FastGICP()  # library defaults, like gicp_test.cpp:104 (k=20, 64 iters, eps 5e-4/2e-3)
        g.set_num_threads(4)
        g.set_input_target(t)
        g.set_input_source(s)
        g.align()
        r = g.result()
        Tcmp = r["Td"] if name == "forward" else np.linalg.inv(r["Td"])
        te, re = synth.pose_error(Tcmp, rel)
        print(name, "converged", r["converged"], "iters", r["iterations"], "t_err %.4f m  r_err %.4f deg" % (te, np.rad2deg(re)),
              "fitness(1.0) %.6f" % g.get_fitness_score(1.0))
        assert te < 0.05 and re < np.deg2rad(1.0) and r["converged"], "oracle fails the reference gtest tolerances"
        out[name + "_T"] = r["Td"]
        out[name + "_fitness"] = g.get_fitness_score(1.0)
        out[name + "_iters"] = r["iterations"]
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "gicp_fixture.npz"), **out)
    readme_fixture(oracle)


if __name__ == "__main__":
    main()
