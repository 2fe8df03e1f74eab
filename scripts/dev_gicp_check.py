"""Development check run on the GPU box: stage-by-stage parity of the CUDA GICP path vs the oracle + timing."""
import os, sys, time
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import oracle
from mr_slam_b200 import synth, pygicp, _lib

_lib.init(0)
print(_lib.lib().mrslam_version().decode())

def check_pair(kind, pid, k):
    src, tgt, Tgt, Tg, Pa, Pb = synth.make_pair(pid, kind)
    n = len(src)
    print(f"--- {kind} pair {pid} n={n} k={k}")
    # kNN
    c = pygicp.Cloud(tgt)
    t = time.time(); gi, gd = c.knn(k); tg = time.time() - t
    t = time.time(); oi, od = oracle.knn(tgt, tgt, k); to = time.time() - t
    print(f"knn: idx equal {np.array_equal(gi, oi)}, d2 equal {np.array_equal(gd, od)}  (gpu {tg*1e3:.1f} ms incl copies, oracle {to*1e3:.1f} ms)")
    if not np.array_equal(gi, oi):
        bad = np.where((gi != oi).any(1))[0]
        print("  mismatching rows:", len(bad), bad[:5], gi[bad[0]], oi[bad[0]], gd[bad[0]], od[bad[0]])
    g = pygicp.FastGICP(); g.set_correspondence_randomness(k); g.set_max_correspondence_distance(100.0)
    g.set_maximum_iterations(50); g.set_transformation_epsilon(1e-3)
    g.set_input_source(src); g.set_input_target(tgt)
    o = oracle.FastGICP(); o.set_correspondence_randomness(k); o.set_max_correspondence_distance(100.0)
    o.set_maximum_iterations(50); o.set_transformation_epsilon(1e-3)
    o.set_input_source(src); o.set_input_target(tgt)
    # linearize at the guess
    Tg32 = Tg.astype(np.float32)
    cost_g, Hg, bg = g.evaluate_cost(Tg32)
    cost_o, Ho, bo, corr_o, sqd_o, mah_o = o.linearize(Tg32.astype(np.float64), n)
    cov_g = g.get_target_covariances(len(tgt))[:, :3, :3]
    cov_o = oracle.covariances(tgt, k)
    print("cov max abs diff", np.abs(cov_g - cov_o).max())
    corr_g, sqd_g, mah_g = g.last_correspondences(n)
    print("corr equal", np.array_equal(corr_g, corr_o), "sqd equal", np.array_equal(sqd_g, sqd_o),
          "mahal max rel", np.abs(mah_g - mah_o).max() / np.abs(mah_o).max())
    print("cost rel", abs(cost_g - cost_o) / abs(cost_o), "H rel", np.abs(Hg - Ho).max() / np.abs(Ho).max(), "b rel", np.abs(bg - bo).max() / np.abs(bo).max())
    print("compute_error rel", abs(g.compute_error(Tgt) - o.compute_error(Tgt)) / o.compute_error(Tgt))
    t = time.time(); Tg_ = g.align(Tg32); tga = time.time() - t
    t = time.time(); To_ = o.align(Tg32); toa = time.time() - t
    ro = o.result()
    print(f"align: gpu {tga*1e3:.1f} ms  oracle({oracle.num_procs()} thr) {toa*1e3:.1f} ms; conv {g.has_converged()}/{ro['converged']} iters {g.num_iterations()}/{ro['iterations']} evals {g.eval_counts()}/{(ro['n_linearize'], ro['n_error'])}")
    print("  |T_gpu - T_oracle| max", np.abs(g.get_final_transformation_d() - ro["Td"]).max(), " err vs gt", synth.pose_error(Tg_, Tgt))
    fg, fo = g.get_fitness_score(1.0), o.get_fitness_score(1.0)
    print("  fitness", fg, fo, abs(fg - fo))
    al = g.get_aligned(n)
    ref = (src @ Tg_[:3, :3].T + Tg_[:3, 3]).astype(np.float32)
    print("  aligned max diff", np.abs(al - ref).max())
    return src, tgt, Tg32

check_pair("vlp16", 0, 15)
check_pair("vlp16", 1, 20)
src, tgt, Tg = check_pair("hdl64", 0, 15)

# batch timing: 8 copies of the 120k pair
import ctypes
for nb in (1, 8, 32):
    regs = []
    for i in range(nb):
        g = pygicp.FastGICP(); g.set_correspondence_randomness(15); g.set_max_correspondence_distance(100.0)
        g.set_maximum_iterations(50); g.set_transformation_epsilon(1e-3)
        g.set_input_source(src); g.set_input_target(tgt)
        regs.append(g)
    _lib.lib().mrslam_device_synchronize()
    l0 = _lib.kernel_launch_count()
    t = time.time(); Ts = pygicp.align_batch(regs, np.stack([Tg] * nb)); dt = time.time() - t
    print(f"batch {nb}: {dt*1e3:.1f} ms -> {nb/dt:.1f} aligns/s, launches {_lib.kernel_launch_count()-l0}, iters {[r.num_iterations() for r in regs][:4]}")
    t = time.time(); Ts = pygicp.align_batch(regs, np.stack([Tg] * nb)); dt = time.time() - t
    print(f"batch {nb} (covs cached): {dt*1e3:.1f} ms -> {nb/dt:.1f} aligns/s")
    del regs
