"""Deterministic synthetic lidar data for tests and bench.py.

Scene (seed 20240901): ground plane z=-1.8 m with +-0.05 m sinusoidal relief, room walls at
x=+-40 m / y=+-30 m (6 m high), 24 axis-aligned boxes and 12 vertical cylinders inside +-35 m.
Scans are ray-cast from a sensor pose; range noise N(0, 0.02 m).  No network, no datasets.
"""
import numpy as np

SCENE_SEED = 20240901
GROUND_Z = -1.8
WALL_X, WALL_Y, WALL_H = 40.0, 30.0, 6.0


class Scene:
    def __init__(self, seed=SCENE_SEED):
        rng = np.random.default_rng(seed)
        nb, nc = 24, 12
        c = rng.uniform(-35, 35, size=(nb, 2))
        s = rng.uniform(0.5, 4.0, size=(nb, 3))
        self.box_min = np.stack([c[:, 0] - s[:, 0] / 2, c[:, 1] - s[:, 1] / 2, np.full(nb, GROUND_Z)], 1)
        self.box_max = np.stack([c[:, 0] + s[:, 0] / 2, c[:, 1] + s[:, 1] / 2, GROUND_Z + s[:, 2]], 1)
        self.cyl_c = rng.uniform(-35, 35, size=(nc, 2))
        self.cyl_r = rng.uniform(0.2, 0.6, size=nc)
        self.cyl_h = rng.uniform(2.0, 5.0, size=nc)

    @staticmethod
    def relief(x, y):
        return 0.05 * np.sin(0.35 * x) * np.cos(0.27 * y)

    def raycast(self, o, d):
        """o: (3,) origin, d: (N,3) unit directions (world). returns range (N,), inf for misses."""
        N = d.shape[0]
        best = np.full(N, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            # ground
            t = (GROUND_Z - o[2]) / d[:, 2]
            t = np.where((t > 0) & np.isfinite(t), t, np.inf)
            hx, hy = o[0] + t * d[:, 0], o[1] + t * d[:, 1]
            ok = (np.abs(hx) <= WALL_X) & (np.abs(hy) <= WALL_Y)
            # relief: displace along the ray so that the hit lies on z = GROUND_Z + relief(x, y) (first order)
            dz = np.where(ok, self.relief(np.where(ok, hx, 0), np.where(ok, hy, 0)), 0.0)
            t = np.where(ok, t + dz / d[:, 2], np.inf)
            best = np.minimum(best, np.where(t > 0, t, np.inf))
            # walls
            for axis, lim, other, olim in ((0, WALL_X, 1, WALL_Y), (1, WALL_Y, 0, WALL_X)):
                for sgn in (-1.0, 1.0):
                    t = (sgn * lim - o[axis]) / d[:, axis]
                    t = np.where((t > 0) & np.isfinite(t), t, np.inf)
                    ho = o[other] + t * d[:, other]
                    hz = o[2] + t * d[:, 2]
                    ok = (np.abs(ho) <= olim) & (hz >= GROUND_Z) & (hz <= GROUND_Z + WALL_H)
                    best = np.minimum(best, np.where(ok, t, np.inf))
            # boxes (slab method)
            inv = 1.0 / d
            for bmin, bmax in zip(self.box_min, self.box_max):
                t1 = (bmin - o) * inv
                t2 = (bmax - o) * inv
                tn = np.nanmax(np.minimum(t1, t2), axis=1)
                tf = np.nanmin(np.maximum(t1, t2), axis=1)
                ok = (tf >= tn) & (tn > 0)
                best = np.minimum(best, np.where(ok, tn, np.inf))
            # vertical cylinders
            for c, r, h in zip(self.cyl_c, self.cyl_r, self.cyl_h):
                ox, oy = o[0] - c[0], o[1] - c[1]
                a = d[:, 0] ** 2 + d[:, 1] ** 2
                b = 2 * (ox * d[:, 0] + oy * d[:, 1])
                cc = ox * ox + oy * oy - r * r
                disc = b * b - 4 * a * cc
                t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * a)
                hz = o[2] + t * d[:, 2]
                ok = (disc > 0) & (t > 0) & (hz >= GROUND_Z) & (hz <= GROUND_Z + h)
                best = np.minimum(best, np.where(ok, t, np.inf))
        return best


def beam_directions(kind):
    """Sensor-frame unit ray directions, azimuth-major ordering (like a spinning lidar driver)."""
    if kind == "vlp16":
        elev = np.deg2rad(np.arange(-15.0, 15.0 + 1e-9, 2.0))  # 16 beams
        n_az = 1024
    elif kind == "hdl64":
        elev = np.deg2rad(np.linspace(-24.8, 2.0, 64))
        n_az = 1875
    else:
        raise ValueError(kind)
    az = np.arange(n_az) * (2 * np.pi / n_az)
    A, E = np.meshgrid(az, elev, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    return d


def pose_matrix(x, y, z, yaw, pitch=0.0, roll=0.0):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [x, y, z]
    return T


_SCENE = None


def scene():
    global _SCENE
    if _SCENE is None:
        _SCENE = Scene()
    return _SCENE


def make_scan(pose, kind="hdl64", scan_id=0, n_points=None, noise=0.02, max_range=100.0):
    """Ray-cast one scan from `pose` (4x4 sensor->world). Returns float32 (N,3) in the SENSOR frame.

    Misses are dropped and the cloud is padded back to exactly `n_points` (default: the ray count)
    by re-sampling hits (seeded), as SURVEY.md 8d prescribes."""
    d_s = beam_directions(kind)
    N = d_s.shape[0]
    n_points = n_points or N
    rng = np.random.default_rng(1000 + scan_id)
    d_w = d_s @ pose[:3, :3].T
    r = scene().raycast(pose[:3, 3], d_w)
    r = r + rng.normal(0.0, noise, size=N)
    ok = np.isfinite(r) & (r > 0.5) & (r < max_range)
    pts = (d_s[ok] * r[ok, None])
    if pts.shape[0] < n_points:
        extra = rng.integers(0, pts.shape[0], size=n_points - pts.shape[0])
        jitter = rng.normal(0.0, 0.005, size=(len(extra), 3))
        pts = np.concatenate([pts, pts[extra] + jitter], 0)
    elif pts.shape[0] > n_points:
        pts = pts[:n_points]
    return np.ascontiguousarray(pts, dtype=np.float32)


def make_pair(pair_id, kind="hdl64", n_points=None):
    """(source, target, T_gt, T_guess): aligning source onto target gives T_gt (4x4 float64).

    target = scan from keyframe pose P_a; source = scan from P_b = P_a @ T_gt;
    T_guess = T_gt perturbed by yaw +-2 deg, t +-0.3 m (mimics the RING initial pose)."""
    rng = np.random.default_rng(2000 + pair_id)
    P_a = pose_matrix(rng.uniform(-20, 20), rng.uniform(-15, 15), 0.0, rng.uniform(-np.pi, np.pi))
    T_gt = pose_matrix(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.1, 0.1), np.deg2rad(rng.uniform(-5, 5)))
    P_b = P_a @ T_gt
    tgt = make_scan(P_a, kind, scan_id=2 * pair_id, n_points=n_points)
    src = make_scan(P_b, kind, scan_id=2 * pair_id + 1, n_points=n_points)
    dT = pose_matrix(rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.03, 0.03), np.deg2rad(rng.uniform(-2, 2)))
    return src, tgt, T_gt, dT @ T_gt, P_a, P_b


def pose_error(T_est, T_gt):
    """(translation error [m], rotation error [rad]) as FG/src/test/gicp_test.cpp:70-75 computes it."""
    d = np.linalg.inv(np.asarray(T_gt, np.float64)) @ np.asarray(T_est, np.float64)
    t_err = float(np.linalg.norm(d[:3, 3]))
    c = (np.trace(d[:3, :3]) - 1.0) / 2.0
    return t_err, float(np.arccos(np.clip(c, -1.0, 1.0)))


# --- elevation mapping sensor constants (EM/config/sensor_processors/velodyne_VLP16.yaml) ---
VLP16_SENSOR = dict(min_radius=0.018, beam_angle=0.0006, beam_constant=0.0015)


def elevation_inputs(pose):
    """Arguments of Process_points for a sensor whose frame == base frame (cfg 2):
    rotationVariance = 0, C_SB = I, B_r_BS = 0; sensorJacobian = third row of C_BM^T-projected C_SB^T."""
    R = np.asarray(pose[:3, :3], np.float64)
    P = np.array([0.0, 0.0, 1.0])
    C_BM_T = R  # map<-base rotation transposed appropriately: C_BM = R^T, so C_BM^T = R
    P_mul_C_BM_T = P @ C_BM_T
    C_SB_T = np.eye(3)
    sensorJacobian = P_mul_C_BM_T @ C_SB_T
    return dict(T=np.asarray(pose, np.float32), sensorJacobian=sensorJacobian.astype(np.float32),
                rotationVariance=np.zeros((3, 3), np.float32), C_SB_transpose=C_SB_T.astype(np.float32),
                P_mul_C_BM_transpose=P_mul_C_BM_T.astype(np.float32), B_r_BS_skew=np.zeros((3, 3), np.float32),
                min_r=VLP16_SENSOR["min_radius"], beam_a=VLP16_SENSOR["beam_angle"], beam_c=VLP16_SENSOR["beam_constant"])


def ring_preprocess(pts):
    """load_pc_infer (RING/util.py:91-112): keep |x|,|y|<70, 0<z<30; scale to [-1,1]; returns SoA float32[3N]."""
    pc = np.asarray(pts, np.float32)
    keep = (np.abs(pc[:, 0]) < 70.0) & (np.abs(pc[:, 1]) < 70.0) & (pc[:, 2] < 30.0) & (pc[:, 2] > 0.0)
    hits = pc[keep].copy()
    hits[:, 0] = hits[:, 0] / 70.0
    hits[:, 1] = hits[:, 1] / 70.0
    hits[:, 2] = hits[:, 2] / 30.0
    return hits
