#!/usr/bin/env python
"""bench.py -- benchmarks of the B200-native MR_SLAM hot path (one JSON line per run, the driver's contract).

  --workload cfg2   (default; BASELINE.json configs[1], the configuration the metric is quoted on; at N > 1 it is configs[3])
      one step = per rank: 32 fresh FastGICP refines of synthetic 64-beam 120 000-point scan pairs (clouds rebuilt and both
      covariance sets recomputed every step, parameters of global_manager.cpp:2438-2442: k=15, max-corr 100 m, eps 1e-3,
      50 iters) + the elevation process+fuse of the 32 source scans into a 0.1 m / 1200x1200 map.
      N > 1: rank r = robot r; its 32 target scans are keyframes of robot (r+1) % N and arrive by ONE neighbour-to-neighbour
      NCCL send/recv of the candidate clouds per step (`--pairs 64` is configs[3]'s 512 refines on 8 GPUs).
  --workload ring   (configs[2]) 256 raw 120k-pt scans -> load_pc_infer + max-height BEV -> Radon -> normalise + FFT ->
      fast_corr of every descriptor against scan 0.  metric: RING descriptors/s.  N > 1: scans are batch-sharded, the
      descriptors all-gathered, every rank correlates its first query against the global database.
  --workload tiled  (configs[4]) 84 scans = 10.08 M points fused in scan order into one 0.05 m / 4000x4000 global grid, storage
      rows sharded over the N GPUs (strong scaling), border rows exchanged.  metric: elevation cell-fuses/s.

  value            whole-job throughput, inputs resident in HBM, device-event timed, max over ranks
  e2e              the same metric through the C-ABI / python mirrors with HOST buffers (H2D + D2H inside the timed region)
  roofline         dominant kernel group with the bound that actually limits it (search kernels: warp-instruction issue;
                   streaming kernels: HBM against MEASURED_PEAKS.json); every group is listed under roofline.groups
  cpu_baseline     the CPU oracle (a port of the reference's CPU path) on a bounded sample, all host threads
  --impl reference the reference's CPU implementation of the same workload on the host cores (the reference itself needs
                   PCL/Eigen/ROS and cannot be built here, DESIGN.md: this is the oracle port, kind "port"), same config dict
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PAIRS = 32
N_POINTS = 120000
K_CORR, MAX_CORR, MAX_ITER, TRANS_EPS = 15, 100.0, 50, 1e-3
ELEV_L, ELEV_RES = 1200, 0.1
# development / CPU-test knobs only: the driver runs the defaults (BASELINE.json sizes); a reduced run describes itself in `config`
RING_SCANS, RING_BASE = int(os.environ.get("BENCH_RING_SCANS", 256)), int(os.environ.get("BENCH_RING_BASE", 32))
TILED_L, TILED_RES = int(os.environ.get("BENCH_TILED_L", 4000)), 0.05
TILED_FRAMES, TILED_BASE = int(os.environ.get("BENCH_TILED_FRAMES", 84)), int(os.environ.get("BENCH_TILED_BASE", 28))
SM_COUNT = 148
# NVML sampling cadence inside the timed steps (1 = every step).  Development knob: with several ranks on one host the NVML
# queries of all ranks go through one driver lock.
SAMPLE_EVERY = int(os.environ.get("BENCH_SAMPLE_EVERY", 1))
# measured under ncu (profiles/r2_summary.txt): warp-instructions per query of the two search kernels, DRAM bytes per cloud
KNN_WARP_INST_PER_QUERY = float(os.environ.get("BENCH_KNN_INST", 701.0))
NN1_WARP_INST_PER_QUERY = float(os.environ.get("BENCH_NN1_INST", 160.0))
KNN_DRAM_BYTES_PER_CLOUD = (16.160256e6 + 5.948416e6) / 8


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_constants():
    """ncu-derived constants of the current kernels (written by scripts/ncu_summary.py into profiles/kernel_constants.json)"""
    global KNN_WARP_INST_PER_QUERY, NN1_WARP_INST_PER_QUERY, KNN_DRAM_BYTES_PER_CLOUD
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "kernel_constants.json")))
    except Exception:  # noqa: BLE001
        return {}
    if "BENCH_KNN_INST" not in os.environ and "knn_warp_inst_per_query" in c:
        KNN_WARP_INST_PER_QUERY = float(c["knn_warp_inst_per_query"])
    if "BENCH_NN1_INST" not in os.environ and "nn1_warp_inst_per_query" in c:
        NN1_WARP_INST_PER_QUERY = float(c["nn1_warp_inst_per_query"])
    if "knn_dram_bytes_per_cloud" in c:
        KNN_DRAM_BYTES_PER_CLOUD = float(c["knn_dram_bytes_per_cloud"])
    return c


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    return float(p.get("hbm_gbs", 6650.0)), ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in p else "fallback (B200_PROFILING.md)")


def make_pairs(rank, n_pairs):
    """The pair set of robot `rank`.  By default every robot refines the SAME 32 pairs (synth seeds 0..n_pairs-1): weak scaling is
    defined on equal work per GPU, and the work of a GICP batch depends on its data -- independent pair sets (round 1:
    seeds rank*1000+i) have different Levenberg-Marquardt iteration counts, and the same 32-pair step takes 13.7 ... 17.9 ms on ONE
    GPU depending on the set (profiles/r2_summary.txt), which showed up as a "scaling efficiency" of 0.79 at N=4 although no rank
    waited for anything.  (A rigid copy per rank does not help: the convergence test of lsq_registration_impl.hpp:84-94 is not
    invariant under a change of frame.)  The clouds still travel between the ranks.  BENCH_DISTINCT_DATA=1 restores independent sets."""
    from mr_slam_b200 import synth
    seed0 = rank * 1000 if os.environ.get("BENCH_DISTINCT_DATA") else 0
    out = []
    for i in range(n_pairs):
        src, tgt, T_gt, T_guess, P_a, P_b = synth.make_pair(seed0 + i, "hdl64", N_POINTS)
        out.append(dict(src=src, tgt=tgt, T_gt=T_gt, guess=T_guess.astype(np.float32), pose=P_b, scan=src))
    return out


def config_for(workload, n_pairs, world, near_filter=True):
    """The `config` block: a function of the workload only, so that both arms (ours / --impl reference) print the same dict."""
    if workload == "cfg2":
        return {"workload": "cfg2: %d pairwise GICP refines of 120k-pt synthetic 64-beam scans per rank + elevation process+fuse @0.1 m "
                            "(1200x1200) of the %d source scans, frames in order" % (n_pairs, n_pairs),
                "pairs_per_rank": n_pairs, "points_per_scan": N_POINTS, "k": K_CORR, "max_iter": MAX_ITER, "max_corr_m": MAX_CORR,
                "trans_eps": TRANS_EPS, "elevation_grid": "%dx%d @ %.2f m" % (ELEV_L, ELEV_L, ELEV_RES), "near_filter": bool(near_filter),
                "parallelism": "robot-per-gpu x%d" % world,
                "data": "every robot refines the same %d pairs (equal work per GPU; independent sets differ by up to 30 %% in step time on one GPU)" % n_pairs,
                "exchange": "none (single robot)" if world == 1 else
                            "one neighbour-to-neighbour NCCL send/recv of the %d candidate clouds (%.1f MB) per rank and step"
                            % (n_pairs, n_pairs * N_POINTS * 12 / 1e6),
                "l2": "per-step working set (%d clouds x ~10 MB + map) exceeds the 126 MB L2; clouds are rebuilt every step" % (2 * n_pairs)}
    if workload == "ring":
        return {"workload": "cfg3: %d raw 120k-pt 64-beam scans per rank -> load_pc_infer + 120x120 max-height BEV -> 120-angle Radon -> "
                            "normalise + FFT (TIRING) -> fast_corr of all descriptors vs scan 0" % RING_SCANS,
                "scans_per_rank": RING_SCANS, "points_per_scan": N_POINTS, "bev": "120x120x1", "angles": 120, "radon": "product default (texture emulation)",
                "parallelism": "batch-sharded x%d%s" % (world, "" if world == 1 else " + all_gather of the 115 KB descriptors"),
                "l2": "%d scans x 1.44 MB = %d MB of points per step exceed the 126 MB L2" % (RING_SCANS, RING_SCANS * N_POINTS * 12 // 1000000)}
    if workload == "tiled":
        return {"workload": "cfg5: %d scans (%.2f M points) fused in scan order into one %dx%d @ %.2f m global grid, storage rows sharded "
                            "over the GPUs" % (TILED_FRAMES, TILED_FRAMES * N_POINTS / 1e6, TILED_L, TILED_L, TILED_RES),
                "frames": TILED_FRAMES, "points_per_scan": N_POINTS, "grid": "%dx%d @ %.2f m" % (TILED_L, TILED_L, TILED_RES),
                "near_filter": bool(near_filter), "parallelism": "row tiles x%d%s" % (world, "" if world == 1 else " + 2-row halo exchange"),
                "l2": "map planes 4 x 64 MB + 121 MB of points per step exceed the 126 MB L2"}
    raise ValueError(workload)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (nvidia_ml_py), once inside every timed
    step.  Both alternatives were measured to perturb this launch/sync-heavy workload: the recipe's `nvidia-smi -lms` loop
    slowed the steps 3-4x, and a background NVML polling thread (50 ms) still cost ~25 %."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.err = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = "nvml unavailable: %r" % (e,)
            self.nv = None

    def sample(self):
        nv = self.nv
        if nv is None:
            return
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def reset(self):
        self.samples, self.reasons = [], set()

    def stop(self):
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples),
               "source": "nvml, one sample inside every timed step while a kernel of the step runs on the GPU"}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------------------------------------
# CPU arms (the oracle port of the reference's CPU path): used by cpu_baseline (bounded sample) and --impl reference
# ------------------------------------------------------------------------------------------------------------------------
def oracle_align_sample(pairs, n_sample, threads):
    """Time the CPU oracle (port of the reference's CPU FastGICP) on the first n_sample pairs: fresh kd-trees,
    covariances and the LM loop per pair, like one ICPCheck call each."""
    import oracle
    t0 = time.perf_counter()
    iters = []
    for p in pairs[:n_sample]:
        o = oracle.FastGICP()
        o.set_num_threads(threads)
        o.set_correspondence_randomness(K_CORR)
        o.set_max_correspondence_distance(MAX_CORR)
        o.set_maximum_iterations(MAX_ITER)
        o.set_transformation_epsilon(TRANS_EPS)
        o.set_input_source(p["src"])
        o.set_input_target(p["tgt"])
        o.align(p["guess"])
        iters.append(o.result()["iterations"])
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, iters


def oracle_elev_sample(pairs, n_sample, om=None, near_filter=True):
    import oracle
    from mr_slam_b200 import synth
    if om is None:
        om = oracle.ElevationMap(ELEV_L, ELEV_RES, 2.5, 0.2)
    fused = 0
    t0 = time.perf_counter()
    for p in pairs[:n_sample]:
        ei = synth.elevation_inputs(p["pose"])
        x, y, z = (np.ascontiguousarray(p.get("scan", p["src"])[:, i]) for i in range(3))
        om.move(p["pose"][:3, 3].astype(np.float32))
        r = om.process_points(x, y, z, ei["T"], p["pose"][2, 3] - 5.0, p["pose"][2, 3] + 5.0, ei["min_r"], ei["beam_a"], ei["beam_c"],
                              ei["sensorJacobian"], ei["rotationVariance"], ei["C_SB_transpose"], ei["P_mul_C_BM_transpose"],
                              ei["B_r_BS_skew"])
        zeros = np.zeros(len(x), np.int32)
        fused += om.fuse(r[0], zeros, zeros, zeros, np.zeros(len(x), np.float32), r[4], r[1])
    dt = time.perf_counter() - t0
    return fused / max(dt, 1e-9), dt, fused


def ref_line(args, workload, metric, unit, value, ms_per_step, config, cpu):
    return {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if workload == "tiled" else "weak", "vs_baseline": None,
            "dtype": "f32/f64" if workload == "cfg2" else "f32", "data": "synthetic", "config": config, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference_cfg2(args):
    """All n_pairs refines AND the elevation replay of the n_pairs source scans per step -- the same step as our arm -- on the
    host cores, all threads.  If the host is so slow that warmup+steps would not end within ~4 minutes the per-step sample is cut
    (and said so); a one-off 8-thread rate (the reference's setNumThreads(8), global_manager.cpp:2438) is reported beside it."""
    import oracle
    threads = oracle.num_procs()
    n_pairs = args.pairs
    pairs = make_pairs(0, n_pairs)
    total_steps = args.warmup + args.steps
    # calibration on one pair (not timed into the result)
    _, dt1, _ = oracle_align_sample(pairs, 1, threads)
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", 240.0))
    n_sample = n_pairs
    if dt1 * n_pairs * 1.15 * total_steps > budget_s:
        n_sample = max(1, min(n_pairs, int(budget_s / (dt1 * 1.15 * total_steps))))
    om = oracle.ElevationMap(ELEV_L, ELEV_RES, 2.5, 0.2)
    vals = []
    for s in range(total_steps):
        t0 = time.perf_counter()
        _, dt_a, _ = oracle_align_sample(pairs, n_sample, threads)
        ev, dt_e, fused = oracle_elev_sample(pairs, n_sample, om, near_filter=True)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            vals.append((n_sample / dt, dt, n_sample / dt_a, fused / max(dt_e, 1e-9)))
    value = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    thr8 = min(8, threads)
    v8, dt8, _ = oracle_align_sample(pairs, min(2, n_pairs), thr8)
    full = n_sample == n_pairs
    cpu = {"value": value, "unit": "aligns/s", "cores": threads, "kind": "port",
           "sample": ("the full step: all %d pairs (fresh kd-trees + covariances + LM each) + the elevation oracle on the %d source scans"
                      % (n_pairs, n_pairs)) if full else
                     ("%d of the %d pairs + their %d elevation frames per step (host too slow for the full step inside %.0f s)"
                      % (n_sample, n_pairs, n_sample, budget_s)),
           "gicp_only_aligns_per_s": float(np.mean([v[2] for v in vals])),
           "elevation_cell_fuses_per_s": float(np.mean([v[3] for v in vals])), "elevation_threads": 1,
           "value_8_threads": v8, "value_8_threads_note": "GICP only, %d threads like setNumThreads(8) at global_manager.cpp:2438, %d pairs"
                                                          % (thr8, min(2, n_pairs))}
    line = ref_line(args, "cfg2", "gicp_aligns_per_s", "aligns/s", value, ms, config_for("cfg2", n_pairs, args.gpus), cpu)
    print(json.dumps(line))


def ring_scans(rank, n_scans, n_base):
    """n_base ray-cast 64-beam scans, each seen under n_scans/n_base yaw angles (a yaw-rotated scan is what RING must match)."""
    from mr_slam_b200 import synth
    base = []
    for i in range(n_base):
        pose = synth.pose_matrix(3.0 * np.cos(0.7 * i + rank), 2.0 * np.sin(1.3 * i + rank), 0.0, 0.2 * i)
        base.append(synth.make_scan(pose, "hdl64", scan_id=7000 + 100 * rank + i) + np.array([0, 0, 1.8], np.float32))
    out = np.empty((n_scans, N_POINTS, 3), np.float32)
    rep = n_scans // n_base
    for j in range(n_scans):
        b = base[j % n_base]
        yaw = 2 * np.pi * (j // n_base) / rep
        c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
        out[j, :, 0] = c * b[:, 0] - s * b[:, 1]
        out[j, :, 1] = s * b[:, 0] + c * b[:, 1]
        out[j, :, 2] = b[:, 2]
    return out


def oracle_ring_sample(scans, idx, desc0=None):
    """CPU restatement of generate_RING (util.py:174-200) + fast_corr (:362-374) for the scans `idx`; returns (descriptors, secs)"""
    import oracle
    from mr_slam_b200 import synth
    ang = np.linspace(0, 2 * np.pi, 120).astype(np.float32)
    t0 = time.perf_counter()
    descs = []
    for j in idx:
        p = synth.ring_preprocess(scans[j])
        soa = np.ascontiguousarray(p.T.reshape(-1))
        bev = oracle.voxelocc(soa, len(p), 1, 1, 120, 120, 1, 1).reshape(-1, 3)[:, 2].reshape(1, 120, 120)
        sino = oracle.radon_forward(bev, ang, 120, 1.0, 0)
        _, tir = oracle.ring_tiring(sino)
        descs.append(tir)
    d0 = descs[0] if desc0 is None else desc0
    for t in descs:
        oracle.ring_fast_corr(d0, t)
    return descs, time.perf_counter() - t0


def run_reference_ring(args):
    import oracle
    threads = oracle.num_procs()
    scans = ring_scans(0, RING_SCANS, RING_BASE)
    total_steps = args.warmup + args.steps
    _, dt1 = oracle_ring_sample(scans, [0])
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", 240.0))
    n_sample = max(1, min(RING_SCANS, int(budget_s / (max(dt1, 1e-4) * 1.15 * total_steps))))
    vals = []
    for s in range(total_steps):
        idx = [(s * n_sample + j) % RING_SCANS for j in range(n_sample)]
        _, dt = oracle_ring_sample(scans, idx)
        if s >= args.warmup:
            vals.append((n_sample / dt, dt))
    value = float(np.mean([v[0] for v in vals]))
    cpu = {"value": value, "unit": "descriptors/s", "cores": threads, "kind": "port",
           "sample": "%d of the %d scans per step: load_pc_infer (numpy) + voxeliser + exact-bilinear Radon (OpenMP) + normalise/FFT + fast_corr vs the "
                     "first descriptor of the sample" % (n_sample, RING_SCANS)}
    line = ref_line(args, "ring", "ring_descriptors_per_s", "descriptors/s", value, float(np.mean([v[1] for v in vals])) * 1e3 * RING_SCANS / n_sample,
                    config_for("ring", 0, args.gpus), cpu)
    line["ms_per_step_note"] = "scaled from the %d-scan sample to the %d-scan step" % (n_sample, RING_SCANS) if n_sample < RING_SCANS else "full step"
    print(json.dumps(line))


def tiled_frames_host(n_frames, n_base):
    """cfg5 input: n_base ray-cast scans of the (80 m x 60 m) scene placed along a Lissajous path that spans the 200 m x 200 m
    grid (the scene is replicated along the path: frame f = base scan f % n_base shifted to the path point f).  Returns
    (list of (x, y, z SoA float32, pose 4x4))."""
    from mr_slam_b200 import synth
    base = []
    for i in range(n_base):
        pose = synth.pose_matrix(6.0 * np.cos(0.9 * i), 5.0 * np.sin(1.1 * i), 0.0, 0.37 * i)
        base.append((synth.make_scan(pose, "hdl64", scan_id=9000 + i), pose))
    frames = []
    for f in range(n_frames):
        pts, pose = base[f % n_base]
        t = 2 * np.pi * f / n_frames
        P = pose.copy()
        P[0, 3] += 58.0 * np.sin(2 * t + 0.4)
        P[1, 3] += 68.0 * np.sin(3 * t)
        frames.append((np.ascontiguousarray(pts.T), P))
    return frames


def tiled_elev_args(P):
    from mr_slam_b200 import synth
    e = synth.elevation_inputs(P)
    z0 = float(P[2, 3])
    return (e["T"], z0 - 5.0, z0 + 5.0, e["min_r"], e["beam_a"], e["beam_c"], e["sensorJacobian"], e["rotationVariance"],
            e["C_SB_transpose"], e["P_mul_C_BM_transpose"], e["B_r_BS_skew"])


def run_reference_tiled(args):
    import oracle
    frames = tiled_frames_host(TILED_FRAMES, TILED_BASE)
    om = oracle.ElevationMap(TILED_L, TILED_RES, 2.5, 0.2)
    total_steps = args.warmup + args.steps

    def run(idx):
        fused = 0
        t0 = time.perf_counter()
        for f in idx:
            soa, P = frames[f]
            a = tiled_elev_args(P)
            r = om.process_points(soa[0], soa[1], soa[2], *a)
            zeros = np.zeros(N_POINTS, np.int32)
            fused += om.fuse(r[0], zeros, zeros, zeros, np.zeros(N_POINTS, np.float32), r[4], r[1])
        return fused, time.perf_counter() - t0

    _, dt1 = run([0])
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", 240.0))
    n_sample = max(1, min(TILED_FRAMES, int(budget_s / (max(dt1, 1e-4) * 1.15 * total_steps))))
    vals = []
    for s in range(total_steps):
        idx = [(s * n_sample + j) % TILED_FRAMES for j in range(n_sample)]
        fused, dt = run(idx)
        if s >= args.warmup:
            vals.append((fused / dt, dt * TILED_FRAMES / n_sample))
    value = float(np.mean([v[0] for v in vals]))
    cpu = {"value": value, "unit": "cell-fuses/s", "cores": 1, "kind": "port",
           "sample": "%d of the %d frames per step, sequential in-order fuse (one thread; the reference's O(cells x points) G_fuse loop is "
                     "infeasible at 4000^2)" % (n_sample, TILED_FRAMES)}
    line = ref_line(args, "tiled", "elevation_cell_fuses_per_s", "cell-fuses/s", value, float(np.mean([v[1] for v in vals])) * 1e3,
                    config_for("tiled", 0, args.gpus), cpu)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------------
# GPU arms
# ------------------------------------------------------------------------------------------------------------------------
class Ctx:
    """process-wide set-up shared by the three GPU workloads"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from mr_slam_b200 import _lib
        from mr_slam_b200 import dist as mdist
        self.torch, self.dist, self.mdist, self._lib = torch, dist, mdist, _lib
        self.rank, self.world, self.local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback")
        self.cores = mdist.bind_to_gpu_numa(self.local_rank) if self.world > 1 and not os.environ.get("BENCH_NO_BIND") else None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL_DEBUG=VERSION (the image default on some boxes) prints a banner on stdout, which must carry the JSON line only
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        _lib.init(self.local_rank)
        self.L = _lib.lib()
        self.lib_stream = torch.cuda.ExternalStream(self.L.mrslam_stream(), device=self.dev)
        self.sampler = ClockSampler(self.local_rank)
        self.sampler.start()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def record(self, e):
        with self.torch.cuda.stream(self.lib_stream):
            e.record()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def gather_floats(self, vals):
        """[world][len(vals)] of per-rank floats"""
        if self.world == 1:
            return [list(vals)]
        tt = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        allt = [self.torch.zeros_like(tt) for _ in range(self.world)]
        self.dist.all_gather(allt, tt)
        return [[float(v) for v in x] for x in allt]

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_cfg2(args):
    C = Ctx(args)
    torch, dist, mdist, _lib, L = C.torch, C.dist, C.mdist, C._lib, C.L
    from mr_slam_b200 import elevation, pygicp, synth
    rank, world, dev, lib_stream = C.rank, C.world, C.dev, C.lib_stream
    near_filter = not args.near_filter_off

    n_pairs = args.pairs
    # BENCH_DATA_RANK (development): run another rank's pair set on this GPU -- the LM iteration counts, and with them the step
    # time, depend on the pairs, so ranks with different data do different amounts of work
    data_rank = env_int("BENCH_DATA_RANK", rank) if world == 1 else rank
    pairs = make_pairs(data_rank, n_pairs)
    guesses = np.stack([p["guess"] for p in pairs])
    # device-resident inputs: this robot's source scans and the candidate clouds it offers to its neighbour
    src_dev = torch.from_numpy(np.stack([p["src"] for p in pairs])).to(dev)           # [P, N, 3]
    # rank q holds (and offers) the keyframes that robot (q-1) % world matches against: the targets of that rank's pairs
    offer_pairs = pairs if (world == 1 or not os.environ.get("BENCH_DISTINCT_DATA")) else make_pairs((rank - 1) % world, n_pairs)
    offer_dev = torch.from_numpy(np.stack([p["tgt"] for p in offer_pairs])).to(dev)
    recv_dev = torch.empty_like(offer_dev) if world > 1 else None
    # pinned host copies for the e2e path
    src_host = [torch.from_numpy(p["src"]).pin_memory() for p in pairs]
    tgt_host = [torch.from_numpy(p["tgt"]).pin_memory() for p in pairs]
    soa_dev = [torch.from_numpy(np.ascontiguousarray(p["scan"].T)).to(dev) for p in pairs]  # elevation inputs x,y,z SoA (sensor frame)
    soa_host = [torch.from_numpy(np.ascontiguousarray(p["scan"].T)).pin_memory().numpy() for p in pairs]  # page-locked x,y,z
    pin = lambda dt: torch.zeros(N_POINTS, dtype=dt).pin_memory().numpy()
    elev_out = (pin(torch.int32), pin(torch.float32), pin(torch.float32), pin(torch.float32), pin(torch.float32))
    zeros_i_pin, zeros_f_pin = pin(torch.int32), pin(torch.float32)
    elev_in = [synth.elevation_inputs(p["pose"]) for p in pairs]

    emap = elevation.ElevationMap()
    emap.Init_GPU_elevationmap(ELEV_L, ELEV_RES, 2.5, 0.2)

    def new_regs():
        regs = []
        for _ in range(n_pairs):
            g = pygicp.FastGICP()
            g.set_num_threads(8)
            g.set_correspondence_randomness(K_CORR)
            g.set_max_correspondence_distance(MAX_CORR)
            g.set_maximum_iterations(MAX_ITER)
            g.set_transformation_epsilon(TRANS_EPS)
            regs.append(g)
        return regs

    def elev_args(i):
        e = elev_in[i]
        z0 = float(pairs[i]["pose"][2, 3])
        return (e["T"], z0 - 5.0, z0 + 5.0, e["min_r"], e["beam_a"], e["beam_c"], e["sensorJacobian"], e["rotationVariance"],
                e["C_SB_transpose"], e["P_mul_C_BM_transpose"], e["B_r_BS_skew"])

    elev_frames = [elevation.ElevationMap.make_frame(soa_dev[i][0], soa_dev[i][1], soa_dev[i][2], N_POINTS, *elev_args(i), near_filter=near_filter,
                                                     position=pairs[i]["pose"][:3, 3].astype(np.float32)) for i in range(n_pairs)]
    ev = C.ev
    exch_ms = [0.0]
    step_no = [0]
    sampler = C.sampler

    def step_device(overlap=True, sample_clocks=False, local_targets=None):
        """inputs resident in HBM; returns (t_gicp_ms, t_elev_ms, n_fused, results) measured with events on the library stream.
        Order inside a step: start the exchange (NCCL stream) -> local work that does not need it (source tiles + covariances,
        elevation replay of the robot's own scans) -> wait -> target tiles + covariances + LM.
        overlap=True (the timed steps): the elevation replay is enqueued on the map's own streams and runs concurrently with
        the rest of the GICP batch; the step ends when both are done, and only the total is meaningful (returned as t_gicp,
        t_elev = 0).  overlap=False: the replay runs in line on the library stream, which gives the per-part times.
        local_targets: skip the exchange and use this tensor as the targets (the N > 1 parity check)."""
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        C.record(e0)
        # exchange step: candidate clouds of the neighbouring robot, exactly the bytes this rank needs (NCCL send/recv over NVLink)
        works = []
        if local_targets is not None:
            tgt_dev = local_targets
        elif world > 1:
            tgt_dev, works = mdist.neighbor_exchange(offer_dev, send_to=(rank - 1) % world, recv_from=(rank + 1) % world, out=recv_dev, async_op=True)
        else:
            tgt_dev = offer_dev
        regs = new_regs()
        src_clouds = pygicp.Cloud.from_device_batch([src_dev[i].data_ptr() for i in range(n_pairs)], N_POINTS, 3)
        for g, cs in zip(regs, src_clouds):
            g.set_input_source(cs)
        pygicp.prepare_batch(regs, pygicp.SOURCE)
        if sample_clocks and (len(sampler.samples) < 1 or SAMPLE_EVERY <= 1 or step_no[0] % SAMPLE_EVERY == 0):
            sampler.sample()      # the source kNN (ALU-bound work) is running on the GPU right now
        step_no[0] += 1
        C.record(e1)
        # the robot's scans replayed into its map by one call (frames pipelined inside the library)
        emap.process_fuse_device_batch(elev_frames, ELEV_RES, ELEV_L, asynchronous=overlap)
        C.record(e2)
        tx0 = time.perf_counter()
        if works:
            mdist.wait_all(works)             # orders torch's current stream after the transfer (no host block)
            done = torch.cuda.Event()
            done.record()
            lib_stream.wait_event(done)       # ... and the library stream after that
        exch_ms[0] += 1e3 * (time.perf_counter() - tx0)
        tgt_clouds = pygicp.Cloud.from_device_batch([tgt_dev[i].data_ptr() for i in range(n_pairs)], N_POINTS, 3)
        for g, ct in zip(regs, tgt_clouds):
            g.set_input_target(ct)
        Ts = pygicp.align_batch(regs, guesses)
        fused = emap.fused_total(reset=True) if overlap else None   # joins the asynchronous replay into the library stream
        C.record(e3)
        e3.synchronize()
        if fused is None:
            fused = emap.fused_total(reset=True)
        t_gicp = e0.elapsed_time(e1) + e2.elapsed_time(e3)
        t_elev = e1.elapsed_time(e2)
        if overlap:
            t_gicp, t_elev = e0.elapsed_time(e3), 0.0
        counts = [r.eval_counts() for r in regs]
        # the registrations (and their clouds) are released here, like ICPCheck's locals: the library recycles their device
        # buffers for the next batch -- holding them across steps would make every other step allocate 64 fresh clouds
        return t_gicp, t_elev, fused, Ts, (sum(c[0] for c in counts), sum(c[1] for c in counts))

    # device staging of the e2e elevation uploads (x, y, z SoA per frame) and the frames that point at it
    stage_soa = [torch.empty_like(s_) for s_ in soa_dev]
    stage_frames = [elevation.ElevationMap.make_frame(stage_soa[i][0], stage_soa[i][1], stage_soa[i][2], N_POINTS, *elev_args(i), near_filter=near_filter,
                                                      position=pairs[i]["pose"][:3, 3].astype(np.float32)) for i in range(n_pairs)]
    soa_pin_t = [torch.from_numpy(h) for h in soa_host]
    copy_stream = torch.cuda.Stream(device=dev)

    stage_soa_b = [torch.empty_like(s_) for s_ in soa_dev]      # second staging set: batch k + 1 goes up while batch k is fused
    stage_frames_b = [elevation.ElevationMap.make_frame(stage_soa_b[i][0], stage_soa_b[i][1], stage_soa_b[i][2], N_POINTS, *elev_args(i), near_filter=near_filter,
                                                        position=pairs[i]["pose"][:3, 3].astype(np.float32)) for i in range(n_pairs)]
    stage_sets = ((stage_soa, stage_frames), (stage_soa_b, stage_frames_b))

    def e2e_submit(slot):
        """Hand one batch of HOST inputs (pinned) to the public API: every cloud through set_input_source / set_input_target, every
        scan's x, y, z into its device staging.  Only uploads are enqueued (the library's and torch's copy streams); the library
        stream is not touched, so a batch submitted while the previous one is being aligned goes up under its kernels."""
        regs = new_regs()
        for i, g in enumerate(regs):
            g.set_input_source(src_host[i].numpy())       # first in the PCIe queue: the source kNN is the first consumer
        with torch.cuda.stream(copy_stream):
            for i in range(n_pairs):
                stage_sets[slot][0][i].copy_(soa_pin_t[i], non_blocking=True)
            up = torch.cuda.Event()
            up.record()
        for i, g in enumerate(regs):
            g.set_input_target(tgt_host[i].numpy())
        return regs, up

    def e2e_compute_head(regs, up, slot):
        """first, asynchronous part of the step on a submitted batch: source tiles + covariances, elevation replay"""
        pygicp.prepare_batch(regs, pygicp.SOURCE)
        lib_stream.wait_event(up)
        emap.process_fuse_device_batch(stage_sets[slot][1], ELEV_RES, ELEV_L, asynchronous=True)

    def e2e_compute_tail(regs):
        """target tiles + covariances + LM; results (32 transforms, the fused count) back on the host"""
        Ts = pygicp.align_batch(regs, guesses)  # reads back 32 x 4x4
        fused = emap.fused_total(reset=True)
        return fused, Ts

    def step_e2e():
        """One isolated step with HOST inputs through the public API: H2D of every cloud (GICP) and of every scan's x, y, z
        (elevation) and the results back, nothing overlapped with a neighbouring step (the latency of a single batch).  The
        elevation scans go up on a copy stream while the source clouds are searched, and are fused on the device by the replay
        entry point (ElevationMap::add-style: the per-point outputs of Process_points never leave the GPU)."""
        e0, e1 = ev(), ev()
        C.record(e0)
        regs, up = e2e_submit(0)
        e2e_compute_head(regs, up, 0)
        fused, Ts = e2e_compute_tail(regs)
        C.record(e1)
        e1.synchronize()
        return e0.elapsed_time(e1), fused, Ts

    def run_e2e_pipelined(steps):
        """`steps` batches back to back, batch k + 1 SUBMITTED (its uploads enqueued) before batch k is computed -- what a node
        that receives scans continuously does.  Every byte of every batch is copied inside the timed region; the copies of batch
        k + 1 overlap the kernels of batch k.  Returns (ms for all steps, fused count of the last step, transforms of the last)."""
        e0, e1 = ev(), ev()
        C.record(e0)
        nxt = e2e_submit(0)
        fused = Ts = None
        for k_ in range(steps):
            cur = nxt
            e2e_compute_head(cur[0], cur[1], k_ & 1)
            if k_ + 1 < steps:
                nxt = e2e_submit((k_ + 1) & 1)   # host work of the submission runs while the GPU searches batch k's sources
            fused, Ts = e2e_compute_tail(cur[0])
            cur = None
        C.record(e1)
        e1.synchronize()
        return e0.elapsed_time(e1), fused, Ts

    def step_e2e_compat():
        """the reference-shaped calls with HOST buffers: per frame Move + Process_points (5 output arrays back to the host) +
        Fuse (7 input arrays up again), as the elevation_mapping node drives libgpu.so"""
        e0, e1 = ev(), ev()
        C.record(e0)
        regs = new_regs()
        for i, g in enumerate(regs):
            g.set_input_source(src_host[i].numpy())
            g.set_input_target(tgt_host[i].numpy())
        Ts = pygicp.align_batch(regs, guesses)
        zeros_i, zeros_f = zeros_i_pin, zeros_f_pin
        fused = 0
        for i in range(n_pairs):
            emap.Move(pairs[i]["pose"][:3, 3].astype(np.float32), ELEV_RES, ELEV_L)
            x, y, z = soa_host[i]
            r = emap.Process_points(x, y, z, elev_args(i)[0], N_POINTS, *elev_args(i)[1:], out=elev_out)
            fused += emap.Fuse(ELEV_L, N_POINTS, r[0], zeros_i, zeros_i, zeros_i, zeros_f, r[4], r[1])
        C.record(e1)
        e1.synchronize()
        return e0.elapsed_time(e1), fused, Ts

    # ---- warm-up (also grows the memory pools; NVML's lazy initialisation happens here, outside the timed region) ----
    for _ in range(args.warmup):
        step_device(sample_clocks=True)
    sampler.reset()
    C.barrier()

    # ---- timed region: device-resident ----
    launches0 = _lib.kernel_launch_count()
    C.barrier()
    t_g, t_e, fused_tot, n_lin, n_err = 0.0, 0.0, 0, 0, 0
    exch_ms[0] = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        tg, te, fused, Ts, (a, b) = step_device(sample_clocks=True)
        t_g += tg
        t_e += te
        fused_tot += fused
        n_lin += a
        n_err += b
    C.barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.kernel_launch_count() - launches0
    exchange_ms_per_step = exch_ms[0] / args.steps
    clocks = sampler.stop()
    last_Ts = np.array(Ts, copy=True)

    # ---- N > 1: the transforms computed from clouds that arrived through NCCL must equal, bit for bit, the transforms of the
    # same pairs aligned with locally generated targets (outside the timed region) ----
    parity = None
    if world > 1:
        local_tgt = torch.from_numpy(np.stack([p["tgt"] for p in pairs])).to(dev)
        same_bytes = bool(torch.equal(local_tgt, recv_dev))
        _tg, _te, _f, Ts_local, _c = step_device(local_targets=local_tgt)
        same_T = bool(np.array_equal(np.asarray(Ts_local).view(np.uint32), last_Ts.view(np.uint32)))
        ok_all = C.reduce(1.0 if (same_bytes and same_T) else 0.0, "min")
        parity = {"multi_gpu_transforms_bitwise_equal_local": bool(ok_all == 1.0), "this_rank": {"clouds": same_bytes, "transforms": same_T},
                  "checked": "%d pairs per rank on all %d ranks, exchanged vs locally generated targets" % (n_pairs, world)}
        del local_tgt
        assert ok_all == 1.0, "multi-GPU parity violated: %r" % (parity,)

    # per-part device times: the same step with the elevation replay in line
    SEQ_STEPS = 3
    seq_g = seq_e = 0.0
    seq_fused = 0
    for _ in range(SEQ_STEPS):
        tg_, te_, f_, _Ts, _c = step_device(overlap=False)
        seq_g += tg_
        seq_e += te_
        seq_fused += f_
    # per-kernel device times for the roofline block: CUDA events around every launch group on the library stream (kept out
    # of the timed region: ~300 extra event records per step would be timed with it)
    PROF_STEPS = 2
    _lib.prof_enable(True)
    p_lin = p_err = 0
    tgp = 0.0
    for _ in range(PROF_STEPS):
        tg_, _te, _f, _Ts, (a, b) = step_device(overlap=False)
        tgp += tg_ + _te
        p_lin += a
        p_err += b
    prof = _lib.prof_get()
    _lib.prof_enable(False)
    C.barrier()
    per_rank = C.gather_floats([(t_g + t_e) / args.steps, float(clocks.get("sm_mhz") or 0.0), prof["knn_search"][0] / PROF_STEPS, exchange_ms_per_step])
    t_total_ms = C.reduce(t_g + t_e, "max")
    seq_g_ms, seq_e_ms = C.reduce(seq_g, "max"), C.reduce(seq_e, "max")
    aligns_total = n_pairs * args.steps * world
    fused_all = C.reduce(float(fused_tot), "sum")
    seq_fused_all = C.reduce(float(seq_fused), "sum")
    value = aligns_total / (t_total_ms * 1e-3)

    # accuracy of the last step against ground truth (sanity, not the parity gate)
    errs = [synth.pose_error(last_Ts[i], pairs[i]["T_gt"]) for i in range(n_pairs)]

    # ---- timed region: end to end with host buffers ----
    run_e2e_pipelined(3)      # warm-up: grows the cloud pool to two batches
    C.barrier()
    wall_e2e0 = time.perf_counter()
    te2e, fused_e2e, Ts_e2e = run_e2e_pipelined(args.steps)
    wall_e2e = time.perf_counter() - wall_e2e0
    C.barrier()
    te2e_ms = C.reduce(te2e, "max")
    e2e_value = aligns_total / (te2e_ms * 1e-3)
    if os.environ.get("BENCH_E2E_DIAG") and rank == 0:   # development: where does the e2e step spend its host time?
        tm = {"head": 0.0, "submit": 0.0, "tail": 0.0}
        _lib.lib().mrslam_build_diag_dump(b"device-resident steps")
        nxt = e2e_submit(0)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        ND = 6
        for k_ in range(ND):
            cur = nxt
            a0 = time.perf_counter()
            e2e_compute_head(cur[0], cur[1], k_ & 1)
            a1 = time.perf_counter()
            nxt = e2e_submit((k_ + 1) & 1)
            a2 = time.perf_counter()
            e2e_compute_tail(cur[0])
            cur = None
            a3 = time.perf_counter()
            tm["head"] += a1 - a0
            tm["submit"] += a2 - a1
            tm["tail"] += a3 - a2
        torch.cuda.synchronize()
        wd = time.perf_counter() - w0
        nxt = None
        _lib.lib().mrslam_build_diag_dump(b"e2e pipelined steps")
        _lib.prof_enable(True)
        run_e2e_pipelined(3)
        pr = _lib.prof_get()
        _lib.prof_enable(False)
        sys.stderr.write("e2e diag: wall %.2f ms/step; host ms/step: %s; kernel groups with profiling on (ms/step): %s\n" % (
            1e3 * wd / ND, {k: round(1e3 * v / ND, 3) for k, v in tm.items()}, {k: round(v[0] / 3, 3) for k, v in pr.items() if v[1]}))
    # the same batch in isolation (no neighbouring step to hide its uploads behind): single-batch latency
    NSINGLE = max(3, args.steps // 4)
    step_e2e()
    tsingle = 0.0
    for _ in range(NSINGLE):
        t, _f, _T = step_e2e()
        tsingle += t
    C.barrier()
    tsingle_ms = C.reduce(tsingle, "max") / NSINGLE
    e2e_same = bool(np.array_equal(np.asarray(Ts_e2e).view(np.uint32), last_Ts.view(np.uint32)))
    h2d = n_pairs * 2 * N_POINTS * 12 + n_pairs * 3 * N_POINTS * 4
    d2h = n_pairs * 64 + 8
    # the reference-shaped (stateless Process_points / Fuse) host path, fewer repetitions
    step_e2e_compat()
    C.barrier()
    tcomp = 0.0
    NCOMP = max(2, args.steps // 3)
    for _ in range(NCOMP):
        t, _, _ = step_e2e_compat()
        tcomp += t
    C.barrier()
    tcomp_ms = C.reduce(tcomp, "max") / NCOMP
    h2d_c = n_pairs * 2 * N_POINTS * 12 + n_pairs * (3 * N_POINTS * 4 + 7 * N_POINTS * 4)
    d2h_c = n_pairs * 64 + n_pairs * 5 * N_POINTS * 4

    # ---- CPU baseline (rank 0, bounded sample) ----
    cpu = None
    if rank == 0:
        import oracle
        threads = oracle.num_procs()
        v, dt, _ = oracle_align_sample(pairs, 2, threads)
        ev_, edt, _f = oracle_elev_sample(pairs, 4, near_filter=near_filter)
        cpu = {"value": v, "unit": "aligns/s", "cores": threads, "kind": "port",
               "sample": "2 of the %d pairs (fresh kd-trees + covariances + LM, %.2f s); elevation oracle on 4 frames: %.3g cell-fuses/s single thread"
                         % (n_pairs, dt, ev_), "elevation_cell_fuses_per_s": ev_}

    if rank == 0:
        hbm_peak, peak_src = peaks()
        f_ghz = (clocks.get("sm_mhz") or 1965.0) * 1e-3
        issue_peak = SM_COUNT * 4 * f_ghz           # G warp-instructions / s
        clouds_prof = 2 * n_pairs * PROF_STEPS
        q_knn = float(N_POINTS) * clouds_prof
        q_nn1 = float(N_POINTS) * p_lin

        def grp(ms_key, launches_key, bound, work, unit, peak, hbm_bytes=None, note=None):
            ms = sum(prof[k][0] for k in ms_key)
            d = {"ms_per_step": round(ms / PROF_STEPS, 4), "launches_per_step": prof[launches_key][1] / PROF_STEPS, "bound": bound,
                 "achieved": (work / (ms * 1e-3) / 1e9) if ms > 0 else 0.0, "peak": peak, "unit": unit}
            d["frac"] = d["achieved"] / peak if peak else None
            if hbm_bytes is not None and ms > 0:
                d["hbm_gbs"] = hbm_bytes / (ms * 1e-3) / 1e9
                d["hbm_frac"] = d["hbm_gbs"] / hbm_peak
            if note:
                d["note"] = note
            return d

        groups = {
            "knn_search": grp(["knn_search"], "knn_search", "issue", KNN_WARP_INST_PER_QUERY * q_knn, "Gwarp-inst/s", issue_peak,
                              hbm_bytes=(16.0 + 4.0 * K_CORR) * q_knn,
                              note="%.0f warp-instructions per query (ncu smsp__inst_executed.sum / queries, profiles/); peak = 148 SM x 4 schedulers x SM clock"
                                   % KNN_WARP_INST_PER_QUERY),
            "cov": grp(["cov"], "cov", "hbm", 36.0 * q_knn, "GB/s", hbm_peak, note="36 B/pt/cloud (SURVEY 8d); FP64 eigen-solve dominated, see profiles/"),
            "nn1_search": grp(["nn1_search"], "nn1_search", "issue", NN1_WARP_INST_PER_QUERY * q_nn1, "Gwarp-inst/s", issue_peak,
                              hbm_bytes=24.0 * q_nn1, note="%.0f warp-instructions per query; about a third of them are packed FADD2/FMUL2/FFMA2, which issue once but occupy the FP32 pipe for two cycles, so the issue fraction understates the pipe load of the tile scan" % NN1_WARP_INST_PER_QUERY),
            "linearize": grp(["linearize"], "linearize", "hbm", 100.0 * q_nn1, "GB/s", hbm_peak, note="100 B/pt/pass (SURVEY 8d)"),
            "error": grp(["error"], "error", "hbm", 52.0 * N_POINTS * p_err, "GB/s", hbm_peak, note="52 B/pt/eval"),
            "cloud_build": grp(["cloud_build"], "cloud_build", "hbm", (12.0 + 40.0 + 6 * 24.0) * q_knn, "GB/s", hbm_peak,
                               note="12 B/pt in + 40 B/pt out + 6 radix passes x 24 B/pt"),
            "elev_process": grp(["elev_process"], "elev_process", "hbm", 24.0 * N_POINTS * n_pairs * PROF_STEPS, "GB/s", hbm_peak, note="12 B/pt in + 12 B/pt out"),
            "elev_fuse": grp(["elev_fuse"], "elev_fuse", "hbm", (16.0 * seq_fused / max(SEQ_STEPS, 1) + 16.0 * N_POINTS * n_pairs) * PROF_STEPS, "GB/s", hbm_peak,
                             note="16 B per fused point (sorted h,var stream + cell read/write) + sort traffic"),
        }
        dom = max(groups, key=lambda k: groups[k]["ms_per_step"])
        gd = groups[dom]
        roofline = {"bound": gd["bound"], "kernel": dom, "achieved": gd["achieved"], "peak": gd["peak"], "unit": gd["unit"], "frac": gd["frac"],
                    "traffic": (KNN_DRAM_BYTES_PER_CLOUD * 2 * n_pairs / max(gd["launches_per_step"], 1)) if dom == "knn_search" else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of k_knn_search (ncu --set full, profiles/), scaled to the clouds of one launch",
                    "peak_source": ("148 x 4 x measured SM clock (issue slots)" if gd["bound"] == "issue" else peak_src),
                    "hbm": {"achieved_gbs": gd.get("hbm_gbs"), "peak_gbs": hbm_peak, "frac": gd.get("hbm_frac"), "peak_source": peak_src,
                            "note": "algorithmic bytes of the same kernel against the measured HBM peak (SURVEY 8d); not the limiting resource"},
                    "ms_per_launch": gd["ms_per_step"] / max(gd["launches_per_step"], 1), "launches": gd["launches_per_step"],
                    "share_of_step": gd["ms_per_step"] * PROF_STEPS / tgp if tgp > 0 else None,
                    "timed_over": "%d extra steps of the same workload with per-kernel CUDA events on the launch stream" % PROF_STEPS,
                    "groups": groups}
        clocks["per_rank"] = {"ms_per_step": [round(r[0], 3) for r in per_rank], "sm_mhz": [r[1] for r in per_rank],
                              "knn_ms_per_step": [round(r[2], 3) for r in per_rank], "exchange_host_wait_ms": [round(r[3], 3) for r in per_rank],
                              "exchange_bytes_received_per_rank": 0 if world == 1 else n_pairs * N_POINTS * 12}
        line = {
            "metric": "gicp_aligns_per_s", "value": value, "unit": "aligns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64",
            "data": "synthetic",
            "config": config_for("cfg2", n_pairs, world, near_filter),
            "overlap": "timed steps: the elevation replay runs on the map's own streams concurrently with the GICP batch (the step ends "
                       "when both are done); the gicp / elevation blocks below are device times of %d further steps with the replay in "
                       "line on the library stream (%.2f ms per step in that mode)" % (SEQ_STEPS, (seq_g_ms + seq_e_ms) / SEQ_STEPS),
            "gicp": {"aligns_per_s": n_pairs * SEQ_STEPS * world / (seq_g_ms * 1e-3), "ms_per_step": seq_g_ms / SEQ_STEPS,
                     "mean_linearize_per_align": n_lin / float(n_pairs * args.steps), "mean_error_evals_per_align": n_err / float(n_pairs * args.steps),
                     "median_t_err_m": float(np.median([e[0] for e in errs])), "median_r_err_rad": float(np.median([e[1] for e in errs])),
                     "parity_note": "R|t to 1e-4 is pinned by the repo's oracle only (the reference's own tests pin 0.05 m / 1 deg)"},
            "elevation": {"metric": "elevation_cell_fuses_per_s", "value": seq_fused_all / (seq_e_ms * 1e-3), "unit": "cell-fuses/s",
                          "cell_fuses_per_step": fused_all / args.steps / world, "ms_per_step": seq_e_ms / SEQ_STEPS,
                          "grid": "%dx%d @ %.2f m" % (ELEV_L, ELEV_L, ELEV_RES), "near_filter": near_filter},
            "e2e": {"value": e2e_value, "unit": "aligns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": te2e_ms / args.steps, "transforms_bitwise_equal_device_path": e2e_same,
                    "path": "host clouds -> set_input_* / align_batch; host scans -> device staging (copy stream) -> process_fuse_device_batch; "
                            "32 transforms + fused count back",
                    "pipelined": "the %d steps are timed as ONE region (CUDA events on the library stream; host wall clock of the same "
                                 "region %.2f ms); batch k+1 is submitted (uploads enqueued on the copy streams) before batch k is "
                                 "computed, so its H2D copies run under batch k's kernels -- all copies of all steps are inside the region"
                                 % (args.steps, 1e3 * wall_e2e),
                    "single_batch": {"value": n_pairs * world / (tsingle_ms * 1e-3), "unit": "aligns/s", "ms_per_step": tsingle_ms,
                                     "note": "one batch in isolation, its own uploads exposed (%d repetitions)" % NSINGLE},
                    "compat": {"value": n_pairs * world / (tcomp_ms * 1e-3), "unit": "aligns/s", "ms_per_step": tcomp_ms,
                               "h2d_bytes_per_step": h2d_c, "d2h_bytes_per_step": d2h_c,
                               "path": "reference-shaped stateless calls per frame: Move + Process_points (5 arrays D2H) + Fuse (7 arrays H2D)"}},
            "parity": parity,
            "cpu_affinity": ("rank 0 bound to the %d cores NVML lists for its GPU" % len(C.cores)) if C.cores else "unbound",
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line))
    C.finish()


def run_ring(args):
    C = Ctx(args)
    torch, dist, mdist, _lib, L = C.torch, C.dist, C.mdist, C._lib, C.L
    from mr_slam_b200 import ring
    rank, world, dev = C.rank, C.world, C.dev
    scans = ring_scans(rank, RING_SCANS, RING_BASE)                       # [256, 120000, 3] float32
    host = torch.from_numpy(scans).pin_memory()
    d_scans = host.to(dev)
    ang = ring.ring_angles(dev)
    n_total = RING_SCANS * world
    ev = C.ev

    def step(d_in, sample_clocks=False):
        """scans -> BEV -> Radon -> TIRING (-> all_gather) -> fast_corr vs the first descriptor of this rank; returns dist, angle on the host"""
        bev = ring.scans_to_bev_device(d_in)
        if sample_clocks:
            C.sampler.sample()
        _, tir = ring.generate_RING_batch(bev, ang)
        db = tir
        if world > 1:
            db = torch.view_as_complex(mdist.allgather_descriptors(torch.view_as_real(tir).contiguous(), n_total))
        d, a = ring.fast_corr_batch(tir[0], db)
        return d.cpu().numpy(), a.cpu().numpy(), tir

    for _ in range(args.warmup):
        step(d_scans, True)
    C.sampler.reset()
    C.barrier()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = ev(), ev()
    C.record(e0)
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        dist_v, ang_v, tir = step(d_scans, True)
    C.record(e1)
    e1.synchronize()
    C.barrier()
    wall = time.perf_counter() - wall0
    t_ms = C.reduce(e0.elapsed_time(e1), "max")
    launches = _lib.kernel_launch_count() - launches0
    clocks = C.sampler.stop()
    value = n_total * args.steps / (t_ms * 1e-3)

    # e2e: host scans (pinned) -> device inside the timed region, dist/angle back on the host
    def step_e2e():
        d_in = host.to(dev, non_blocking=True)
        return step(d_in)
    for _ in range(2):
        step_e2e()
    C.barrier()
    C.record(e0)
    for _ in range(args.steps):
        step_e2e()
    C.record(e1)
    e1.synchronize()
    C.barrier()
    te2e = C.reduce(e0.elapsed_time(e1), "max")

    # per-stage device times (in-line, events on the library stream; a few further steps)
    stage = {"scan_to_bev": 0.0, "radon+tiring": 0.0, "fast_corr": 0.0}
    REP = 5
    for _ in range(REP):
        a0, a1, a2, a3 = ev(), ev(), ev(), ev()
        C.record(a0)
        bev = ring.scans_to_bev_device(d_scans)
        C.record(a1)
        _, tir = ring.generate_RING_batch(bev, ang)
        C.record(a2)
        ring.fast_corr_batch(tir[0], tir)
        C.record(a3)
        a3.synchronize()
        stage["scan_to_bev"] += a0.elapsed_time(a1) / REP
        stage["radon+tiring"] += a1.elapsed_time(a2) / REP
        stage["fast_corr"] += a2.elapsed_time(a3) / REP

    # parity of the timed outputs against the CPU oracle on a bounded sample (exact-bilinear oracle: RING tolerance 1e-4 applies to
    # interp="exact"; the product default emulates the reference's texture unit, so dist is compared at 2e-3 and the angle exactly)
    cpu = None
    parity = None
    if rank == 0:
        import oracle
        ns = 4
        t0 = time.perf_counter()
        descs, dt = oracle_ring_sample(scans, list(range(ns)))
        ok_ang, max_dd = True, 0.0
        for j in range(ns):
            d_o, a_o, _ = oracle.ring_fast_corr(descs[0], descs[j])
            max_dd = max(max_dd, abs(float(d_o) - float(dist_v[j])))
            ok_ang = ok_ang and int(a_o) == int(ang_v[j])
        parity = {"fast_corr_vs_oracle_first_scans": ns, "max_abs_dist_diff": max_dd, "angles_equal": bool(ok_ang)}
        cpu = {"value": ns / dt, "unit": "descriptors/s", "cores": oracle.num_procs(), "kind": "port",
               "sample": "%d of the %d scans: load_pc_infer (numpy) + voxeliser + exact-bilinear Radon (OpenMP) + normalise/FFT + fast_corr" % (ns, RING_SCANS)}
    if rank == 0:
        hbm_peak, peak_src = peaks()
        bytes_floor = (12.0 * N_POINTS + 230400.0) * RING_SCANS
        dom = max(stage, key=lambda k: stage[k])
        alg = {"scan_to_bev": (12.0 * N_POINTS + 57600.0) * RING_SCANS, "radon+tiring": (57600.0 + 57600.0 + 115200.0) * RING_SCANS,
               "fast_corr": 115200.0 * (RING_SCANS + 1)}
        groups = {k: {"ms_per_step": round(v, 4), "bound": "hbm" if k != "radon+tiring" else "on-chip (shared-memory gather; HBM floor reported)",
                      "hbm_gbs": alg[k] / (v * 1e-3) / 1e9 if v > 0 else 0.0, "hbm_frac": alg[k] / (v * 1e-3) / 1e9 / hbm_peak if v > 0 else 0.0,
                      "algorithmic_bytes": alg[k]} for k, v in stage.items()}
        gd = groups[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": gd["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": gd["hbm_frac"], "traffic": None,
                    "peak_source": peak_src, "groups": groups,
                    "whole_step": {"algorithmic_bytes": bytes_floor, "hbm_gbs": bytes_floor * world / (t_ms / args.steps * 1e-3) / 1e9 / world,
                                   "note": "12 B/pt + 230 400 B per scan (SURVEY 8d) over the whole device-resident step"}}
        line = {"metric": "ring_descriptors_per_s", "value": value, "unit": "descriptors/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_for("ring", 0, world),
                "pair_correlations_per_s": world * n_total * args.steps / (t_ms * 1e-3),
                "e2e": {"value": n_total * args.steps / (te2e * 1e-3), "unit": "descriptors/s", "h2d_bytes_per_step": RING_SCANS * N_POINTS * 12,
                        "d2h_bytes_per_step": n_total * 8, "ms_per_step": te2e / args.steps},
                "parity": parity, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "wall_s_timed_region": wall}
        print(json.dumps(line))
    C.finish()


def run_tiled(args):
    C = Ctx(args)
    torch, dist, mdist, _lib, L = C.torch, C.dist, C.mdist, C._lib, C.L
    from mr_slam_b200 import elevation
    rank, world, dev = C.rank, C.world, C.dev
    near_filter = not args.near_filter_off
    frames_h = tiled_frames_host(TILED_FRAMES, TILED_BASE)
    soa_dev = [torch.from_numpy(f[0]).to(dev) for f in frames_h]
    soa_pin = [torch.from_numpy(f[0]).pin_memory() for f in frames_h]

    def make_map(tiled):
        m = elevation.ElevationMap()
        m.Init_GPU_elevationmap(TILED_L, TILED_RES, 2.5, 0.2)
        t = mdist.TiledElevationMap(m, TILED_L, rank=rank, world=world, device="cuda") if tiled else None
        return m, t

    emap, tiles = make_map(world > 1)

    def frames_for(soa_list, nf):
        return [elevation.ElevationMap.make_frame(soa_list[i][0], soa_list[i][1], soa_list[i][2], N_POINTS, *tiled_elev_args(frames_h[i][1]),
                                                  near_filter=nf, position=None) for i in range(TILED_FRAMES)]

    fr_dev = frames_for(soa_dev, near_filter)
    fr_dev_alt = frames_for(soa_dev, not near_filter)
    ev = C.ev

    def step(fr, sample_clocks=False):
        emap.process_fuse_device_batch(fr, TILED_RES, TILED_L, asynchronous=False)
        if sample_clocks:
            C.sampler.sample()
        if tiles is not None:
            tiles.exchange_halo("elevation")      # the border exchange of configs[4]: 2 storage rows from each neighbour
        return emap.fused_total(reset=True)

    for _ in range(args.warmup):
        step(fr_dev, True)
    C.sampler.reset()
    C.barrier()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = ev(), ev()
    C.record(e0)
    wall0 = time.perf_counter()
    fused = 0
    for _ in range(args.steps):
        fused += step(fr_dev, True)
    C.record(e1)
    e1.synchronize()
    C.barrier()
    wall = time.perf_counter() - wall0
    t_ms = C.reduce(e0.elapsed_time(e1), "max")
    fused_all = C.reduce(float(fused), "sum")
    per_rank = C.gather_floats([e0.elapsed_time(e1) / args.steps, float(fused) / args.steps])
    launches = _lib.kernel_launch_count() - launches0
    clocks = C.sampler.stop()
    value = fused_all / (t_ms * 1e-3)
    # the other near-filter setting (SURVEY 8d asks for both)
    C.barrier()
    C.record(e0)
    fused_alt = 0
    ALT = 3
    for _ in range(ALT):
        fused_alt += step(fr_dev_alt)
    C.record(e1)
    e1.synchronize()
    t_alt = C.reduce(e0.elapsed_time(e1), "max")
    fused_alt_all = C.reduce(float(fused_alt), "sum")

    # e2e: host SoA scans (pinned) -> device inside the timed region
    stage_dev = [torch.empty_like(s) for s in soa_dev]
    fr_stage = frames_for(stage_dev, near_filter)

    def step_e2e():
        for i in range(TILED_FRAMES):
            stage_dev[i].copy_(soa_pin[i], non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        C.lib_stream.wait_event(done)
        return step(fr_stage)
    step_e2e()
    C.barrier()
    C.record(e0)
    for _ in range(max(2, args.steps // 2)):
        step_e2e()
    C.record(e1)
    e1.synchronize()
    te2e = C.reduce(e0.elapsed_time(e1), "max") / max(2, args.steps // 2)

    # parity outside the timed region: (1) union of the tiles == one whole map, bit for bit (N > 1); (2) a fresh map fed the same
    # frames == the CPU oracle on the cells of a band of rows (rank 0; bounded: 6 frames)
    parity = {}
    if world > 1:
        whole, _ = make_map(False)
        t2_map, t2 = make_map(True)
        for m in (whole, t2_map):
            m.process_fuse_device_batch(fr_dev[:12], TILED_RES, TILED_L, asynchronous=False)
        got = t2.gather_full(("elevation", "variance"))
        ref = whole.get_map()
        ok = bool(np.array_equal(got["elevation"].cpu().numpy(), ref["elevation"]) and np.array_equal(got["variance"].cpu().numpy(), ref["variance"]))
        okall = C.reduce(1.0 if ok else 0.0, "min")
        parity["union_of_tiles_equals_whole_map_bitwise"] = bool(okall == 1.0)
        assert okall == 1.0, "tiled map differs from the whole map"
        del whole, t2_map, t2
    cpu = None
    if rank == 0:
        import oracle
        om = oracle.ElevationMap(TILED_L, TILED_RES, 2.5, 0.2)
        gm, _ = make_map(False)
        gm.process_fuse_device_batch(fr_dev[:6], TILED_RES, TILED_L, asynchronous=False)
        t0 = time.perf_counter()
        fo = 0
        for i in range(6):
            soa, P = frames_h[i]
            a = tiled_elev_args(P)
            r = om.process_points(soa[0], soa[1], soa[2], *a)
            zeros = np.zeros(N_POINTS, np.int32)
            fo += om.fuse(r[0], zeros, zeros, zeros, np.zeros(N_POINTS, np.float32), r[4], r[1])
        dt = time.perf_counter() - t0
        mg, mo = gm.get_map(), om.get_map()
        parity["vs_oracle_6_frames"] = {"elevation_bitwise": bool(np.array_equal(mg["elevation"], mo["elevation"])),
                                        "variance_bitwise": bool(np.array_equal(mg["variance"], mo["variance"]))}
        cpu = {"value": fo / dt, "unit": "cell-fuses/s", "cores": 1, "kind": "port",
               "sample": "6 of the %d frames, sequential in-order fuse on one thread (%.2f s)" % (TILED_FRAMES, dt)}
        del gm
    if rank == 0:
        hbm_peak, peak_src = peaks()
        pts = float(TILED_FRAMES * N_POINTS)
        alg_bytes = 12.0 * pts * world + 16.0 * fused_all / args.steps    # every rank reads all points; 16 B per fused point
        ms_step = t_ms / args.steps
        roofline = {"bound": "hbm", "kernel": "elevation process + sort + fuse (whole step)", "achieved": alg_bytes / world / (ms_step * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / world / (ms_step * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_gpu": alg_bytes / world,
                    "note": "12 B/pt in (every tile reads all points) + 16 B per fused point; per GPU"}
        clocks["per_rank"] = {"ms_per_step": [round(r[0], 3) for r in per_rank], "cell_fuses_per_step": [r[1] for r in per_rank]}
        line = {"metric": "elevation_cell_fuses_per_s", "value": value, "unit": "cell-fuses/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_for("tiled", 0, world, near_filter),
                "cell_fuses_per_step": fused_all / args.steps, "points_per_step": pts,
                "other_near_filter_setting": {"near_filter": not near_filter, "value": fused_alt_all / (t_alt * 1e-3), "unit": "cell-fuses/s",
                                              "cell_fuses_per_step": fused_alt_all / ALT, "ms_per_step": t_alt / ALT},
                "e2e": {"value": (fused_all / args.steps) / (te2e * 1e-3), "unit": "cell-fuses/s", "h2d_bytes_per_step": TILED_FRAMES * N_POINTS * 12,
                        "d2h_bytes_per_step": 8, "ms_per_step": te2e},
                "parity": parity, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "wall_s_timed_region": wall}
        print(json.dumps(line))
    C.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "cfg2"), choices=["cfg2", "ring", "tiled"])
    ap.add_argument("--pairs", type=int, default=N_PAIRS, help="cfg2: pairs per rank (default 32; 64 = BASELINE configs[3] on 8 GPUs)")
    ap.add_argument("--near-filter-off", action="store_true", help="elevation: disable the reference's hard-coded near-field / y>0 filter")
    args = ap.parse_args()
    load_constants()
    rank = env_int("RANK", 0)
    if args.impl == "reference":
        if rank != 0:
            return          # rank 0 alone runs and prints the CPU arm
        {"cfg2": run_reference_cfg2, "ring": run_reference_ring, "tiled": run_reference_tiled}[args.workload](args)
        return
    if args.warmup < 3:
        args.warmup = 3
    {"cfg2": run_cfg2, "ring": run_ring, "tiled": run_tiled}[args.workload](args)


if __name__ == "__main__":
    main()
