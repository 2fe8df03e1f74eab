"""`pygicp`-compatible host mirror over the C-ABI (reference: Mapping/src/fast_gicp/src/python/main.cpp:152-190).

Same class and method names, argument meaning and defaults as the reference's pybind module, so the
RING nodes' `fast_gicp()` helper (LoopDetection/src/RING_ros/main_RING.py:81-104) runs unchanged:

    gicp = pygicp.FastGICP()
    gicp.set_input_target(target); gicp.set_input_source(source)
    gicp.set_num_threads(4); gicp.set_max_correspondence_distance(1.0)
    T = gicp.align(initial_guess=init_pose); fitness = gicp.get_fitness_score(1.0)

All compute runs in libmrslam_b200.so on the GPU; numpy arrays are copied on entry like `eigen2pcl`.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, ptr

DBL_MAX = float(np.finfo(np.float64).max)

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = 0, 1, 2, 3, 4


def _xyz32(points):
    p = np.asarray(points)
    if p.ndim != 2 or p.shape[1] < 3:
        raise ValueError("expected an (N,3) array")
    return np.ascontiguousarray(p[:, :3], dtype=np.float32)  # eigen2pcl: cast<float>()


class Cloud:
    """Device-resident cloud (with cached covariances) that several registrations can share."""

    def __init__(self, points=None, device_ptr=None, n=None, stride=3):
        L = lib()
        if device_ptr is not None:
            self._h = L.mrslam_cloud_create_device(device_ptr, int(n), int(stride))
        else:
            p = _xyz32(points)
            self._h = L.mrslam_cloud_create_host(ptr(p), p.shape[0], 3)
        if not self._h:
            raise _lib.MrSlamError(L.mrslam_last_error().decode())

    @classmethod
    def from_device_batch(cls, device_ptrs, n, stride=3):
        """One launch packs `len(device_ptrs)` clouds of n points each (mrslam_cloud_create_device_batch)."""
        cnt = len(device_ptrs)
        ptrs = (C.c_void_p * cnt)(*[int(p) for p in device_ptrs])
        ns = (C.c_int * cnt)(*([int(n)] * cnt if np.isscalar(n) else [int(v) for v in n]))
        out = (C.c_void_p * cnt)()
        check(lib().mrslam_cloud_create_device_batch(ptrs, ns, int(stride), cnt, out))
        clouds = []
        for h in out:
            c = cls.__new__(cls)
            c._h = h
            clouds.append(c)
        return clouds

    def __len__(self):
        return lib().mrslam_cloud_size(self._h)

    def knn(self, k):
        n = len(self)
        idx = np.zeros((n, k), np.int32)
        d2 = np.zeros((n, k), np.float32)
        check(lib().mrslam_cloud_knn(self._h, k, ptr(idx), ptr(d2)))
        return idx, d2

    def __del__(self):
        try:
            if self._h:
                lib().mrslam_cloud_destroy(self._h)
        except Exception:
            pass


class LsqRegistration:
    def __init__(self):
        self._L = lib()
        self._h = self._L.mrslam_gicp_create()
        if not self._h:
            raise _lib.MrSlamError(self._L.mrslam_last_error().decode())
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                self._L.mrslam_gicp_destroy(self._h)
        except Exception:
            pass

    # --- pygicp surface (main.cpp:169-181) ---
    def set_input_target(self, points):
        if isinstance(points, Cloud):
            check(self._L.mrslam_gicp_set_input_target_cloud(self._h, points._h))
        else:
            p = _xyz32(points)
            check(self._L.mrslam_gicp_set_input_target(self._h, ptr(p), p.shape[0], 3))

    def set_input_source(self, points):
        if isinstance(points, Cloud):
            check(self._L.mrslam_gicp_set_input_source_cloud(self._h, points._h))
        else:
            p = _xyz32(points)
            check(self._L.mrslam_gicp_set_input_source(self._h, ptr(p), p.shape[0], 3))

    def swap_source_and_target(self):
        check(self._L.mrslam_gicp_swap_source_and_target(self._h))

    def get_final_hessian(self):
        H = np.zeros((6, 6))
        check(self._L.mrslam_gicp_get_final_hessian(self._h, ptr(H)))
        return H

    def get_final_transformation(self):
        T = np.zeros((4, 4), np.float32)
        check(self._L.mrslam_gicp_get_final_transformation(self._h, ptr(T)))
        return T

    def get_fitness_score(self, max_range=DBL_MAX):
        v = self._L.mrslam_gicp_get_fitness_score(self._h, float(max_range))
        if v < 0:
            raise _lib.MrSlamError(self._L.mrslam_last_error().decode())
        return v

    def align(self, initial_guess=None):
        g = np.ascontiguousarray(np.eye(4) if initial_guess is None else initial_guess, dtype=np.float32)
        s = self._L.mrslam_gicp_align(self._h, ptr(g))
        if s not in (0, -4):  # -4: "lm not converged!!" -- the reference prints and returns the pose reached
            check(s)
        return self.get_final_transformation()

    # --- C++-only surface, exposed for tests ---
    def get_final_transformation_d(self):
        T = np.zeros((4, 4))
        check(self._L.mrslam_gicp_get_final_transformation_d(self._h, ptr(T)))
        return T

    def has_converged(self):
        return bool(self._L.mrslam_gicp_has_converged(self._h))

    def num_iterations(self):
        return int(self._L.mrslam_gicp_get_num_iterations(self._h))

    def eval_counts(self):
        a, b = C.c_int(), C.c_int()
        check(self._L.mrslam_gicp_get_eval_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_maximum_iterations(self, n):
        check(self._L.mrslam_gicp_set_maximum_iterations(self._h, int(n)))

    def set_transformation_epsilon(self, e):
        check(self._L.mrslam_gicp_set_transformation_epsilon(self._h, float(e)))

    def set_rotation_epsilon(self, e):
        check(self._L.mrslam_gicp_set_rotation_epsilon(self._h, float(e)))

    def set_initial_lambda_factor(self, f):
        check(self._L.mrslam_gicp_set_initial_lambda_factor(self._h, float(f)))

    def set_lsq_optimizer(self, t):
        check(self._L.mrslam_gicp_set_lsq_optimizer(self._h, int(t)))

    def set_debug_print(self, on):
        check(self._L.mrslam_gicp_set_debug_print(self._h, int(bool(on))))

    def evaluate_cost(self, T):
        H, b, c = np.zeros((6, 6)), np.zeros(6), C.c_double()
        T = np.ascontiguousarray(T, np.float32)
        check(self._L.mrslam_gicp_evaluate_cost(self._h, ptr(T), ptr(H), ptr(b), C.byref(c)))
        return c.value, H, b

    def compute_error(self, T):
        c = C.c_double()
        T = np.ascontiguousarray(T, np.float64)
        check(self._L.mrslam_gicp_compute_error(self._h, ptr(T), C.byref(c)))
        return c.value

    def last_correspondences(self, n_src):
        corr = np.zeros(n_src, np.int32)
        sqd = np.zeros(n_src, np.float32)
        mah = np.zeros((n_src, 9))
        check(self._L.mrslam_gicp_debug_last_correspondences(self._h, ptr(corr), ptr(sqd), ptr(mah)))
        return corr, sqd, mah

    def get_aligned(self, n_src):
        out = np.zeros((n_src, 3), np.float32)
        check(self._L.mrslam_gicp_get_aligned(self._h, ptr(out), 3))
        return out

    def get_source_covariances(self, n):
        out = np.zeros((n, 4, 4))
        check(self._L.mrslam_gicp_get_source_covariances(self._h, ptr(out), n))
        return out

    def get_target_covariances(self, n):
        out = np.zeros((n, 4, 4))
        check(self._L.mrslam_gicp_get_target_covariances(self._h, ptr(out), n))
        return out

    def set_source_covariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64)
        check(self._L.mrslam_gicp_set_source_covariances(self._h, ptr(c), c.shape[0]))

    def set_target_covariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64)
        check(self._L.mrslam_gicp_set_target_covariances(self._h, ptr(c), c.shape[0]))


class FastGICP(LsqRegistration):
    # main.cpp:183-188
    def set_num_threads(self, n):
        check(self._L.mrslam_gicp_set_num_threads(self._h, int(n)))

    def set_correspondence_randomness(self, k):
        check(self._L.mrslam_gicp_set_correspondence_randomness(self._h, int(k)))

    def set_max_correspondence_distance(self, d):
        check(self._L.mrslam_gicp_set_max_correspondence_distance(self._h, float(d)))

    def set_regularization_method(self, m):
        check(self._L.mrslam_gicp_set_regularization_method(self._h, int(m)))


def align_batch(regs, initial_guesses=None):
    """Advance many registrations in lock step (mrslam_gicp_align_batch); returns (n,4,4) float32."""
    n = len(regs)
    L = lib()
    arr = (C.c_void_p * n)(*[r._h for r in regs])
    g = None
    if initial_guesses is not None:
        g = np.ascontiguousarray(initial_guesses, np.float32).reshape(n, 16)
    s = L.mrslam_gicp_align_batch(arr, n, ptr(g))
    if s not in (0, -4):
        check(s)
    return np.stack([r.get_final_transformation() for r in regs])


SOURCE, TARGET = 1, 2


def prepare_batch(regs, which=SOURCE | TARGET):
    """Build the spatial tiles and covariances of the source and/or target clouds ahead of align_batch."""
    n = len(regs)
    arr = (C.c_void_p * n)(*[r._h for r in regs])
    check(lib().mrslam_gicp_prepare_batch(arr, n, int(which)))


def fitness_batch(regs, max_range=DBL_MAX):
    n = len(regs)
    arr = (C.c_void_p * n)(*[r._h for r in regs])
    out = np.zeros(n)
    check(lib().mrslam_gicp_get_fitness_score_batch(arr, n, float(max_range), ptr(out)))
    return out


def align_points(target, source, method="GICP", downsample_resolution=-1.0, k_correspondences=15,
                 max_correspondence_distance=DBL_MAX, voxel_resolution=1.0, num_threads=0, neighbor_search_method="DIRECT1",
                 neighbor_search_radius=1.5, initial_guess=None):
    """main.cpp:62-150; only method == "GICP" is on the hot path (the others print an error and return identity
    exactly like a reference build without BUILD_VGICP_CUDA does for the CUDA methods)."""
    if method != "GICP":
        import sys
        print("error: registration method %s is not provided by mr_slam_b200 (GICP only)" % method, file=sys.stderr)
        return np.eye(4)
    if downsample_resolution > 0.0:
        target = downsample(target, downsample_resolution)
        source = downsample(source, downsample_resolution)
    reg = FastGICP()
    reg.set_max_correspondence_distance(np.sqrt(DBL_MAX) if max_correspondence_distance >= DBL_MAX else max_correspondence_distance)
    reg.set_correspondence_randomness(k_correspondences)
    reg.set_num_threads(num_threads)
    reg.set_input_target(target)
    reg.set_input_source(source)
    return reg.align(initial_guess).astype(np.float64)


def downsample(points, downsample_resolution):
    """pcl::ApproximateVoxelGrid as main.cpp:46-60 uses it (PCL 1.10 approximate_voxel_grid.hpp): a sequential,
    order-dependent 512-entry hash history -- host code in the reference and here (not a data-parallel kernel)."""
    pts = np.asarray(points, np.float64)[:, :3].astype(np.float32)
    inv = np.float32(1.0) / np.float32(downsample_resolution)
    ijk = np.floor(pts * inv).astype(np.int64)
    h = (ijk[:, 0] * 7171 + ijk[:, 1] * 3079 + ijk[:, 2] * 4231) & 511
    cnt = np.zeros(512, np.int64)
    key = np.zeros((512, 3), np.int64)
    cen = np.zeros((512, 3), np.float32)
    out = []
    for i in range(len(pts)):
        hh = h[i]
        if cnt[hh] and (key[hh, 0] != ijk[i, 0] or key[hh, 1] != ijk[i, 1] or key[hh, 2] != ijk[i, 2]):
            out.append(cen[hh] / np.float32(cnt[hh]))
            cnt[hh] = 0
            cen[hh] = 0
        key[hh] = ijk[i]
        cnt[hh] += 1
        cen[hh] += pts[i]
    for hh in range(512):
        if cnt[hh]:
            out.append(cen[hh] / np.float32(cnt[hh]))
    return np.array(out, np.float32).astype(np.float64).reshape(-1, 3)


def merge_nearest_keyframes(keyframes, transforms, xy_limit=60.0, leaf=0.2, as_cloud=False):
    """GlobalManager::mergeNearestKeyframes (global_manager.cpp:1894-1939) on the GPU: keyframes = list of (n_i, 3|4) float32
    arrays (numpy -> host path) or CUDA torch tensors (device path), transforms = (k,4,4) float32 (currPose^-1 * nearPose).
    Returns the filtered (m,4) x,y,z,intensity array (numpy), or (torch CUDA tensor [m,4], Cloud) when as_cloud=True."""
    import torch
    k = len(keyframes)
    T = np.ascontiguousarray(np.asarray(transforms, np.float32).reshape(k, 16))
    on_dev = k > 0 and hasattr(keyframes[0], "data_ptr")
    # one row stride for the whole pointer table: a list mixing (n,3) and (n,4) keyframes would be read with the wrong stride
    widths = {int(a.shape[1]) if len(a.shape) == 2 else -1 for a in keyframes}
    if k and (len(widths) != 1 or min(widths) < 3):
        raise ValueError("merge_nearest_keyframes: every keyframe must be (n_i, c) with the same c >= 3, got widths %s" % sorted(widths))
    if k and any(hasattr(a, "data_ptr") != on_dev for a in keyframes):
        raise ValueError("merge_nearest_keyframes: keyframes must be all CUDA tensors or all host arrays")
    if on_dev:
        kfs = [t.contiguous().float() for t in keyframes]
        ptrs = (C.c_void_p * k)(*[t.data_ptr() for t in kfs])
        ns = [int(t.shape[0]) for t in kfs]
        stride = int(kfs[0].shape[1]) if k else 3
    else:
        kfs = [np.ascontiguousarray(a, np.float32) for a in keyframes]
        ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in kfs])
        ns = [int(a.shape[0]) for a in kfs]
        stride = int(kfs[0].shape[1]) if k else 3
    total = max(sum(ns), 1)
    out = torch.empty((total, 4), dtype=torch.float32, device="cuda")
    n_out = C.c_int(0)
    torch.cuda.current_stream().synchronize()
    check(lib().mrslam_keyframe_merge(ptrs, (C.c_int * k)(*ns), stride, k, ptr(T), float(xy_limit), float(leaf), 1 if on_dev else 0,
                                      out.data_ptr(), total, C.byref(n_out)))
    m = n_out.value
    if as_cloud:
        return out[:m], Cloud(device_ptr=out.data_ptr(), n=m, stride=4)
    lib().mrslam_device_synchronize()
    return out[:m].cpu().numpy()

