"""Small GICP workload for ncu: B fresh aligns of one synthetic 120k-pt pair (covariances recomputed)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from mr_slam_b200 import synth, pygicp, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_lib.init(0)
src, tgt, Tgt, Tg, Pa, Pb = synth.make_pair(0, "hdl64")
Tg = Tg.astype(np.float32)
L = _lib.lib()
for rep in range(reps):
    regs = []
    for i in range(B):
        g = pygicp.FastGICP(); g.set_correspondence_randomness(15); g.set_max_correspondence_distance(100.0)
        g.set_maximum_iterations(50); g.set_transformation_epsilon(1e-3)
        g.set_input_source(src); g.set_input_target(tgt)
        regs.append(g)
    L.mrslam_device_synchronize()
    L.mrslam_gicp_stats_enable(1)
    t = time.time(); Ts = pygicp.align_batch(regs, np.stack([Tg] * B)); dt = time.time() - t
    pairs = L.mrslam_gicp_stats_pairs()
    nl, ne = regs[0].eval_counts()
    print(f"   knn pair evals/query {L.mrslam_gicp_stats_get(0)/(B*2*120000):.0f}, nn1 pair evals/query/pass {L.mrslam_gicp_stats_get(1)/(B*nl*120000):.0f} ({nl} passes), knn insert rounds/query {L.mrslam_gicp_stats_get(2)/(B*2*120000):.1f}, nn1 max tiles in one warp {L.mrslam_gicp_stats_get(3)}, mean tiles per warp-pass {L.mrslam_gicp_stats_get(1)/32/(B*nl*3750):.1f}")
    print(f"   knn tiles on the per-query fallback path {L.mrslam_gicp_stats_get(4)} of {B*2*3750}, mean shared candidates per tile {L.mrslam_gicp_stats_get(5)/(B*2*3750):.1f}")
    g = L.mrslam_gicp_stats_get
    nq, nw = B * 2 * 120000, B * 2 * 3750
    print(f"   knn v2: visits/query {(g(0) - 1024 * nw) / 32 / nq:.2f}, qualifiers/query {g(2) / nq:.1f}, merge rounds/warp {g(6) / nw:.2f}, overflow flushes {g(7)}, "
          f"blocks handed to the general kernel {g(4) >> 32} of {nw // 4}, candidate tiles per block {g(5) / nw:.1f}")
    print(f"rep {rep}: batch {B}: {dt*1e3:.2f} ms, {B/dt:.1f} aligns/s, pair evals {pairs:.3e} ({pairs/B/ (2*120000):.0f} per query-pass avg over passes), err {synth.pose_error(Ts[0], Tgt)}")
    L.mrslam_gicp_stats_enable(0)
    _lib.prof_enable(True)
    regs2 = []
    for i in range(B):
        g = pygicp.FastGICP(); g.set_correspondence_randomness(15); g.set_max_correspondence_distance(100.0)
        g.set_maximum_iterations(50); g.set_transformation_epsilon(1e-3)
        g.set_input_source(src); g.set_input_target(tgt)
        regs2.append(g)
    pygicp.align_batch(regs2, np.stack([Tg] * B))
    pr = _lib.prof_get(); _lib.prof_enable(False)
    print("   per-kernel ms:", {k: round(v[0], 3) for k, v in pr.items() if v[1]}, " => knn us/cloud %.1f, nn1 us/pair-pass %.1f" % (1e3 * pr["knn_search"][0] / (2 * B), 1e3 * pr["nn1_search"][0] / (B * nl)))
