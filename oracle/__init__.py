"""ctypes loader for the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package `mr_slam_b200` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = 0, 1, 2, 3, 4


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("gicp_oracle.cpp", "elev_oracle.cpp", "bev_oracle.cpp", "radon_oracle.cpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _decl(_LIB)
    return _LIB


def _p(dtype):
    return np.ctypeslib.ndpointer(dtype=dtype, flags="C_CONTIGUOUS")


def _decl(L):
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    L.orc_gicp_create.restype = vp
    L.orc_gicp_destroy.argtypes = [vp]
    L.orc_gicp_set_num_threads.argtypes = [vp, i]
    L.orc_gicp_set_params.argtypes = [vp, i, d, i, i, d, d, i, d]
    L.orc_gicp_set_source.argtypes = [vp, _p(np.float32), i]
    L.orc_gicp_set_target.argtypes = [vp, _p(np.float32), i]
    L.orc_gicp_swap.argtypes = [vp]
    L.orc_gicp_align.argtypes = [vp, _p(np.float32)]
    L.orc_gicp_result.argtypes = [vp, _p(np.float32), _p(np.float64), _p(np.float64)] + [C.POINTER(i)] * 4
    L.orc_gicp_fitness.argtypes = [vp, d]
    L.orc_gicp_fitness.restype = d
    L.orc_knn.argtypes = [_p(np.float32), i, _p(np.float32), i, i, _p(np.int32), _p(np.float32), i]
    L.orc_knn_bruteforce.argtypes = [_p(np.float32), i, _p(np.float32), i, i, _p(np.int32), _p(np.float32)]
    L.orc_covariances.argtypes = [_p(np.float32), i, i, i, _p(np.float64), i]
    L.orc_gicp_linearize.argtypes = [vp, _p(np.float64), _p(np.float64), _p(np.float64), _p(np.int32), _p(np.float32), _p(np.float64)]
    L.orc_gicp_linearize.restype = d
    L.orc_gicp_compute_error.argtypes = [vp, _p(np.float64)]
    L.orc_gicp_compute_error.restype = d
    L.orc_gicp_get_covariances.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.orc_num_procs.restype = i

    L.orc_elev_create.restype = vp
    L.orc_elev_destroy.argtypes = [vp]
    L.orc_elev_init.argtypes = [vp, i, f, f, f]
    L.orc_elev_move.argtypes = [vp, _p(np.float32), f, i, _p(np.float32), _p(np.int32), _p(np.float32)]
    L.orc_elev_process_points.argtypes = ([vp, _p(np.int32)] + [_p(np.float32)] * 7 + [_p(np.float32), i, d, d, f, f, f] +
                                          [_p(np.float32)] * 5)
    L.orc_elev_fuse.argtypes = [vp, i, i] + [_p(np.int32)] * 4 + [_p(np.float32)] * 3
    L.orc_elev_fuse.restype = C.c_long
    L.orc_elev_mapvar_update.argtypes = [vp, i, f]
    L.orc_elev_map_optmove.argtypes = [vp, _p(np.float32), f, f, i, _p(np.float32)]
    L.orc_elev_map_closeloop.argtypes = [vp, _p(np.float32), f, i, f]
    L.orc_elev_get_map.argtypes = [vp] + [C.c_void_p] * 7
    L.orc_elev_map_feature.argtypes = [vp] + [C.c_void_p] * 9
    L.orc_elev_raytracing.argtypes = [vp]
    L.orc_elev_set_owned_rows.argtypes = [vp, i, i]
    L.orc_elev_rows_copy.argtypes = [vp, i, i, i, C.c_void_p, i]
    L.orc_elev_get_state.argtypes = [vp, _p(np.float32), _p(np.int32)]

    L.orc_voxelocc.argtypes = [_p(np.float32), i, i, i, i, i, i, i, _p(np.float32), C.c_void_p, C.POINTER(i)]
    L.orc_polar.argtypes = [_p(np.float32), i, i, i, i, i, i, i, _p(np.float32), C.c_void_p, C.POINTER(i)]
    L.orc_voxelfeat_features.argtypes = [_p(np.float32), i, i, i, _p(np.int32), _p(np.float32), _p(np.float32)]
    L.orc_voxelfeat_bev.argtypes = [_p(np.float32), i, i, i, i, i, i, i, _p(np.float32), C.POINTER(i)]
    L.orc_radon_forward.argtypes = [_p(np.float32), i, i, i, _p(np.float32), i, i, f, i, _p(np.float32)]
    L.orc_ring_tiring.argtypes = [_p(np.float32), i, i, i, C.c_void_p, _p(np.float32)]
    L.orc_ring_fast_corr.argtypes = [_p(np.float32), _p(np.float32), i, i, i, _p(np.float32), C.POINTER(f), C.POINTER(i)]
    L.orc_ring_solve_translation.argtypes = [_p(np.float32), _p(np.float32), i, i, i, f, i, _p(np.float32), _p(np.float32)]


def num_procs():
    return int(lib().orc_num_procs())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class FastGICP:
    """Oracle twin of fast_gicp::FastGICP (same setter names as the pygicp binding)."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_gicp_create())
        self.k = 20
        self.max_corr = float(np.sqrt(np.finfo(np.float64).max))
        self.reg = REG_PLANE
        self.max_iter = 64
        self.rot_eps = 2e-3
        self.trans_eps = 5e-4
        self.optimizer = 0
        self.lm_init_lambda_factor = 1e-9
        self.set_num_threads(0)

    def __del__(self):
        try:
            self._L.orc_gicp_destroy(self._h)
        except Exception:
            pass

    def _push(self):
        self._L.orc_gicp_set_params(self._h, self.k, self.max_corr, self.reg, self.max_iter, self.rot_eps, self.trans_eps,
                                    self.optimizer, self.lm_init_lambda_factor)

    def set_num_threads(self, n):
        self._L.orc_gicp_set_num_threads(self._h, int(n))

    def set_correspondence_randomness(self, k):
        self.k = int(k)

    def set_max_correspondence_distance(self, d):
        self.max_corr = float(d)

    def set_regularization_method(self, r):
        self.reg = int(r)

    def set_maximum_iterations(self, n):
        self.max_iter = int(n)

    def set_transformation_epsilon(self, e):
        self.trans_eps = float(e)

    def set_rotation_epsilon(self, e):
        self.rot_eps = float(e)

    def set_input_source(self, pts):
        p = _f32(pts)[:, :3].copy()
        self._L.orc_gicp_set_source(self._h, p, p.shape[0])

    def set_input_target(self, pts):
        p = _f32(pts)[:, :3].copy()
        self._L.orc_gicp_set_target(self._h, p, p.shape[0])

    def swap_source_and_target(self):
        self._L.orc_gicp_swap(self._h)

    def align(self, initial_guess=None):
        self._push()
        g = _f32(np.eye(4) if initial_guess is None else initial_guess).reshape(16).copy()
        self._L.orc_gicp_align(self._h, g)
        return self.get_final_transformation()

    def result(self):
        T = np.zeros(16, np.float32)
        Td = np.zeros(16, np.float64)
        H = np.zeros(36, np.float64)
        c, it, nl, ne = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._L.orc_gicp_result(self._h, T, Td, H, C.byref(c), C.byref(it), C.byref(nl), C.byref(ne))
        return dict(T=T.reshape(4, 4), Td=Td.reshape(4, 4), H=H.reshape(6, 6), converged=bool(c.value), iterations=it.value,
                    n_linearize=nl.value, n_error=ne.value)

    def get_final_transformation(self):
        return self.result()["T"]

    def get_final_hessian(self):
        return self.result()["H"]

    def has_converged(self):
        return self.result()["converged"]

    def get_fitness_score(self, max_range=float(np.finfo(np.float64).max)):
        return float(self._L.orc_gicp_fitness(self._h, float(max_range)))

    def linearize(self, T, n_src):
        self._push()
        Td = np.ascontiguousarray(T, np.float64).reshape(16)
        H = np.zeros(36)
        b = np.zeros(6)
        corr = np.zeros(n_src, np.int32)
        sqd = np.zeros(n_src, np.float32)
        mah = np.zeros((n_src, 9))
        cost = self._L.orc_gicp_linearize(self._h, Td, H, b, corr, sqd, mah)
        return cost, H.reshape(6, 6), b, corr, sqd, mah

    def compute_error(self, T):
        return float(self._L.orc_gicp_compute_error(self._h, np.ascontiguousarray(T, np.float64).reshape(16)))


def knn(pts, queries, k, threads=0, brute=False):
    pts, queries = _f32(pts), _f32(queries)
    idx = np.zeros((queries.shape[0], k), np.int32)
    d2 = np.zeros((queries.shape[0], k), np.float32)
    if brute:
        lib().orc_knn_bruteforce(pts, pts.shape[0], queries, queries.shape[0], k, idx, d2)
    else:
        lib().orc_knn(pts, pts.shape[0], queries, queries.shape[0], k, idx, d2, threads or num_procs())
    return idx, d2


def covariances(pts, k, reg=REG_PLANE, threads=0):
    pts = _f32(pts)
    out = np.zeros((pts.shape[0], 9))
    lib().orc_covariances(pts, pts.shape[0], k, reg, out, threads or num_procs())
    return out.reshape(-1, 3, 3)


class ElevationMap:
    """Oracle twin of the reference's hidden device-side map + the free functions of gpu_process.cu."""

    def __init__(self, length, resolution, mahalanobis=2.5, obstacle=0.2):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_elev_create())
        self.length, self.resolution = int(length), float(resolution)
        self._L.orc_elev_init(self._h, self.length, self.resolution, mahalanobis, obstacle)

    def __del__(self):
        try:
            self._L.orc_elev_destroy(self._h)
        except Exception:
            pass

    def move(self, position3):
        c, s, a = np.zeros(2, np.float32), np.zeros(2, np.int32), np.zeros(2, np.float32)
        self._L.orc_elev_move(self._h, _f32(position3), self.resolution, self.length, c, s, a)
        return c, s, a

    def process_points(self, x, y, z, T, lower, upper, min_r, beam_a, beam_c, sensorJacobian, rotationVariance, C_SB_transpose,
                       P_mul_C_BM_transpose, B_r_BS_skew):
        n = len(x)
        mi = np.zeros(n, np.int32)
        var, xt, yt, zt = (np.zeros(n, np.float32) for _ in range(4))
        self._L.orc_elev_process_points(self._h, mi, _f32(x), _f32(y), _f32(z), var, xt, yt, zt, _f32(T).reshape(16), n,
                                        float(lower), float(upper), min_r, beam_a, beam_c, _f32(sensorJacobian).reshape(3),
                                        _f32(rotationVariance).reshape(9), _f32(C_SB_transpose).reshape(9),
                                        _f32(P_mul_C_BM_transpose).reshape(3), _f32(B_r_BS_skew).reshape(9))
        return mi, var, xt, yt, zt

    def fuse(self, idx, R, G, B, intensity, height, var):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        return int(self._L.orc_elev_fuse(self._h, self.length, len(idx), i32(idx), i32(R), i32(G), i32(B), _f32(intensity),
                                         _f32(height), _f32(var)))

    def mapvar_update(self, v):
        self._L.orc_elev_mapvar_update(self._h, self.length, float(v))

    def map_optmove(self, opt_p, height_update):
        out = np.zeros(2, np.float32)
        self._L.orc_elev_map_optmove(self._h, _f32(opt_p), float(height_update), self.resolution, self.length, out)
        return out

    def map_closeloop(self, update_position, height_update):
        self._L.orc_elev_map_closeloop(self._h, _f32(update_position), float(height_update), self.length, self.resolution)

    def map_feature(self):
        n = self.length * self.length
        out = dict(elevation=np.zeros(n, np.float32), var=np.zeros(n, np.float32), colorR=np.zeros(n, np.int32), colorG=np.zeros(n, np.int32),
                   colorB=np.zeros(n, np.int32), rough=np.zeros(n, np.float32), slope=np.zeros(n, np.float32), traver=np.zeros(n, np.float32),
                   intensity=np.zeros(n, np.float32))
        self._L.orc_elev_map_feature(self._h, *[out[k].ctypes.data for k in ("elevation", "var", "colorR", "colorG", "colorB", "rough", "slope",
                                                                              "traver", "intensity")])
        return out

    def raytracing(self):
        self._L.orc_elev_raytracing(self._h)

    PLANES = dict(elevation=0, variance=1, intensity=2, lowest=3, traver=4, R=5, G=6, B=7)

    def set_owned_rows(self, row_lo, row_hi):
        self._L.orc_elev_set_owned_rows(self._h, int(row_lo), int(row_hi))

    def rows_to_device(self, plane, storage_row, n_rows, out):
        """same name as the product class; `out` is a CPU torch tensor / numpy array here"""
        buf = out.numpy() if hasattr(out, "numpy") else out
        self._L.orc_elev_rows_copy(self._h, self.PLANES[plane], int(storage_row), int(n_rows), buf.ctypes.data, 0)

    def rows_from_device(self, plane, storage_row, n_rows, src):
        buf = src.contiguous().numpy() if hasattr(src, "numpy") else np.ascontiguousarray(src)
        self._L.orc_elev_rows_copy(self._h, self.PLANES[plane], int(storage_row), int(n_rows), buf.ctypes.data, 1)

    def get_map(self):
        n = self.length * self.length
        out = dict(elevation=np.zeros(n, np.float32), variance=np.zeros(n, np.float32), intensity=np.zeros(n, np.float32),
                   R=np.zeros(n, np.int32), G=np.zeros(n, np.int32), B=np.zeros(n, np.int32), lowest=np.zeros(n, np.float32))
        self._L.orc_elev_get_map(self._h, *[out[k].ctypes.data for k in ("elevation", "variance", "intensity", "R", "G", "B", "lowest")])
        return out

    def state(self):
        c, s = np.zeros(2, np.float32), np.zeros(2, np.int32)
        self._L.orc_elev_get_state(self._h, c, s)
        return c, s


def voxelocc(point_soa, size, max_length, max_height, num_x, num_y, num_height, enough_large=1, want_idx=False):
    p = _f32(point_soa)
    out = np.zeros(num_x * num_y * num_height * enough_large * 3, np.float32)
    idx = np.zeros(3 * size, np.int32) if want_idx else None
    nd = C.c_int()
    lib().orc_voxelocc(p, size, max_length, max_height, num_x, num_y, num_height, enough_large, out,
                       idx.ctypes.data if want_idx else None, C.byref(nd))
    return (out, idx, nd.value) if want_idx else out


def polar(point_soa, size, max_length, max_height, num_ring, num_sector, num_height, enough_large=1, want_idx=False):
    p = _f32(point_soa)
    out = np.zeros(num_ring * num_sector * num_height * enough_large * 3, np.float32)
    idx = np.zeros(3 * size, np.int32) if want_idx else None
    nd = C.c_int()
    lib().orc_polar(p, size, max_length, max_height, num_ring, num_sector, num_height, enough_large, out,
                    idx.ctypes.data if want_idx else None, C.byref(nd))
    return (out, idx, nd.value) if want_idx else out


def voxelfeat_features(pc_aos, size, featsize, k, neighbors, eigens):
    out = np.zeros(size * featsize, np.float32)
    lib().orc_voxelfeat_features(_f32(pc_aos), size, featsize, k, np.ascontiguousarray(neighbors, np.int32), _f32(eigens), out)
    return out


def voxelfeat_bev(point_soa, size, max_length, max_height, num_x, num_y, num_height, featsize):
    out = np.zeros(num_x * num_y * num_height * featsize, np.float32)
    nd = C.c_int()
    lib().orc_voxelfeat_bev(_f32(point_soa), size, max_length, max_height, num_x, num_y, num_height, featsize, out, C.byref(nd))
    return out


def radon_forward(img, angles, det_count=None, det_spacing=1.0, weight_bits=0):
    img = _f32(img)
    shape = img.shape
    im3 = img.reshape(-1, shape[-2], shape[-1])
    angles = _f32(angles)
    det_count = det_count or shape[-1]
    out = np.zeros((im3.shape[0], len(angles), det_count), np.float32)
    lib().orc_radon_forward(im3, im3.shape[0], shape[-2], shape[-1], angles, len(angles), det_count, det_spacing, weight_bits, out)
    return out.reshape(shape[:-2] + (len(angles), det_count))


def ring_tiring(sino):
    s = _f32(sino)
    Cc, A, D = s.shape
    out = np.zeros((Cc, A, D, 2), np.float32)
    norm = np.zeros((Cc, A, D), np.float32)
    lib().orc_ring_tiring(s, Cc, A, D, norm.ctypes.data, out.reshape(-1))
    return norm, out[..., 0] + 1j * out[..., 1]


def ring_fast_corr(a, b):
    def il(x):
        x = np.asarray(x, np.complex64)
        return np.ascontiguousarray(np.stack([x.real, x.imag], -1).astype(np.float32)).reshape(-1)
    Cc, A, D = np.asarray(a).shape
    corr = np.zeros(A, np.float32)
    dist, ang = C.c_float(), C.c_int()
    lib().orc_ring_fast_corr(il(a), il(b), Cc, A, D, corr, C.byref(dist), C.byref(ang))
    return dist.value, ang.value, corr


def ring_solve_translation(query, positive, rot_angle, row_shift=0):
    """solve_translation (RING/util.py:388-423) on float32 [C,H,W] arrays; returns (x, y, error, b[H])."""
    q, p = _f32(query), _f32(positive)
    Cc, H, W = q.shape
    b = np.zeros(H, np.float32)
    out = np.zeros(3, np.float32)
    lib().orc_ring_solve_translation(q.reshape(-1), p.reshape(-1), Cc, H, W, float(rot_angle), int(row_shift), b, out)
    return float(out[0]), float(out[1]), float(out[2]), b


def merge_nearest_keyframes(keyframes, transforms, xy_limit=60.0, leaf=0.2):
    """mergeNearestKeyframes (global_manager.cpp:1894-1939): list of (n_i, 3|4) float32 arrays + (k,4,4) transforms -> (m,4) xyzi."""
    k = len(keyframes)
    kfs = [np.ascontiguousarray(a, np.float32) for a in keyframes]
    T = np.ascontiguousarray(np.asarray(transforms, np.float32).reshape(k, 16))
    ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in kfs])
    ns = (C.c_int * k)(*[a.shape[0] for a in kfs])
    total = max(sum(a.shape[0] for a in kfs), 1)
    out = np.zeros((total, 4), np.float32)
    fn = lib().orc_keyframe_merge
    fn.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int]
    m = fn(ptrs, ns, kfs[0].shape[1] if k else 3, k, T.ctypes.data, float(xy_limit), float(leaf), out.ctypes.data, total)
    assert m >= 0
    return out[:m].copy()


def ringpp_build_neighbors(pointcloud, k):
    """build_neighbors_NN (RING/util.py:163-170): exact kNN (this oracle's kd-tree; sklearn's kd_tree in the reference) followed by
    covariation_eigenvalue (util.py:124-160) restated with the same numpy float32 operations and LAPACK eigvalsh."""
    pc = np.ascontiguousarray(np.asarray(pointcloud)[:, :3], np.float32)
    idx, _ = knn(pc, pc, k)
    nb = pc[idx]                                             # [n, k, 3]
    ex = np.average(nb, axis=1)
    P = nb - ex[:, None, :]
    cov = np.matmul(P.transpose((0, 2, 1)), P) / (nb.shape[1] - 1)
    e3 = np.sort(np.linalg.eigvalsh(cov), axis=1)[:, ::-1]
    L_, P_, S_ = (e3[:, 0] - e3[:, 1]) / e3[:, 0], (e3[:, 1] - e3[:, 2]) / e3[:, 0], e3[:, 2] / e3[:, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        entropy = -L_ * np.log(L_) - P_ * np.log(P_) - S_ * np.log(S_)
    e2 = np.sort(np.linalg.eigvalsh(cov[:, :2, :2]), axis=1)[:, ::-1]
    return idx, cov, entropy, np.append(e3, e2, axis=1)


def disco_rel_ori(real_a, imag_a, real_b, imag_b):
    """GlobalManager::calcRelOri (global_manager.cpp:2719-2762) for one pair of [H, W] float32 planes -> (angle_deg, flat arg-max)."""
    ra, ia, rb, ib = (_f32(x) for x in (real_a, imag_a, real_b, imag_b))
    H, W = ra.shape
    fn = lib().orc_disco_rel_ori
    fn.restype = C.c_float
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.POINTER(C.c_int)]
    loc = C.c_int()
    ang = fn(ra.ctypes.data, ia.ctypes.data, rb.ctypes.data, ib.ctypes.data, H, W, C.byref(loc))
    return float(ang), loc.value

