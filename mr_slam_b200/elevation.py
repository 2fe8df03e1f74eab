"""Host mirror of the elevation-mapping free functions of the reference's libgpu.so
(Mapping/src/elevation_mapping_periodical/elevation_mapping/cuda/gpu_process.cu:938-1252, declared at
src/ElevationMapping.cpp:44-50, SensorProcessorBase.cpp:34, RobotMotionMapUpdater.cpp:18).

Same function names and argument order as the reference; the map the reference keeps in `__device__`
globals is a module-level handle here (one map per process, like the reference).  Eigen arguments are
numpy arrays (row-major).  All compute runs on the GPU through the C-ABI.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MrSlamError, check, lib, ptr


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class ElevationMap:
    """One hidden device-resident map (Init_GPU_elevationmap state)."""

    def __init__(self):
        self._L = lib()
        self._h = self._L.mrslam_elev_create()
        if not self._h:
            raise MrSlamError(self._L.mrslam_last_error().decode())
        self.length = 0
        self.resolution = 0.0

    def __del__(self):
        try:
            if self._h:
                self._L.mrslam_elev_destroy(self._h)
        except Exception:
            pass

    def Init_GPU_elevationmap(self, length, resolution, mahalanobisThr, obstacle_threshold):
        check(self._L.mrslam_elev_init(self._h, int(length), float(resolution), float(mahalanobisThr), float(obstacle_threshold)))
        self.length, self.resolution = int(length), float(resolution)

    def Move(self, current_Position, resolution, length):
        c, s, a = np.zeros(2, np.float32), np.zeros(2, np.int32), np.zeros(2, np.float32)
        check(self._L.mrslam_elev_move(self._h, ptr(_f32(current_Position)), float(resolution), int(length), ptr(c), ptr(s), ptr(a)))
        return c, s, a

    def Process_points(self, x, y, z, T, n, lower, upper, min_r, beam_a, beam_c, sensorJacobian, rotationVariance, C_SB_transpose,
                       P_mul_C_BM_transpose, B_r_BS_skew, out=None):
        """`out` (optional): caller-owned (map_index int32, var, x_ts, y_ts, z_ts float32) arrays of length n -- the reference's
        calling convention (caller-owned output arrays); pass page-locked arrays to get DMA copies."""
        x, y, z = _f32(x), _f32(y), _f32(z)
        if out is not None:
            mi, var, xt, yt, zt = out
        else:
            mi = np.zeros(n, np.int32)
            var, xt, yt, zt = (np.zeros(n, np.float32) for _ in range(4))
        args = [_f32(T).reshape(16), _f32(sensorJacobian).reshape(3), _f32(rotationVariance).reshape(9), _f32(C_SB_transpose).reshape(9),
                _f32(P_mul_C_BM_transpose).reshape(3), _f32(B_r_BS_skew).reshape(9)]
        check(self._L.mrslam_elev_process_points(self._h, ptr(mi), ptr(x), ptr(y), ptr(z), ptr(var), ptr(xt), ptr(yt), ptr(zt),
                                                 ptr(args[0]), int(n), float(lower), float(upper), float(min_r), float(beam_a),
                                                 float(beam_c), *[ptr(a) for a in args[1:]]))
        return mi, var, xt, yt, zt

    def Fuse(self, length, n, idx, R, G, B, intensity, height, var):
        nf = C.c_longlong()
        idx, R, G, B = _i32(idx), _i32(R), _i32(G), _i32(B)
        intensity, height, var = _f32(intensity), _f32(height), _f32(var)
        check(self._L.mrslam_elev_fuse(self._h, int(length), int(n), ptr(idx), ptr(R), ptr(G), ptr(B), ptr(intensity), ptr(height),
                                       ptr(var), C.byref(nf)))
        return nf.value

    def process_fuse_device(self, d_x, d_y, d_z, n, T, lower, upper, min_r, beam_a, beam_c, sensorJacobian, rotationVariance,
                            C_SB_transpose, P_mul_C_BM_transpose, B_r_BS_skew, near_filter=True, want_count=True):
        """Device-resident fast path: d_x/d_y/d_z are torch CUDA float32 tensors (or raw device addresses)."""
        nf = C.c_longlong()
        args = [_f32(T).reshape(16), _f32(sensorJacobian).reshape(3), _f32(rotationVariance).reshape(9), _f32(C_SB_transpose).reshape(9),
                _f32(P_mul_C_BM_transpose).reshape(3), _f32(B_r_BS_skew).reshape(9)]
        dp = lambda a: a if isinstance(a, int) else ptr(a)
        check(self._L.mrslam_elev_process_fuse_device(self._h, dp(d_x), dp(d_y), dp(d_z), int(n), ptr(args[0]), float(lower),
                                                      float(upper), float(min_r), float(beam_a), float(beam_c),
                                                      *[ptr(a) for a in args[1:]], int(bool(near_filter)),
                                                      C.byref(nf) if want_count else None))
        return nf.value if want_count else None

    @staticmethod
    def make_frame(d_x, d_y, d_z, n, T, lower, upper, min_r, beam_a, beam_c, sensorJacobian, rotationVariance, C_SB_transpose,
                   P_mul_C_BM_transpose, B_r_BS_skew, near_filter=True, position=None):
        """One mrslam_elev_frame_t for process_fuse_device_batch (position != None: Move(position) first)."""
        dp = lambda a: a if isinstance(a, int) else a.data_ptr() if hasattr(a, "data_ptr") else ptr(a)
        F = _lib.ElevFrame()
        F.d_x, F.d_y, F.d_z, F.n = dp(d_x), dp(d_y), dp(d_z), int(n)
        F.do_move = 0 if position is None else 1
        if position is not None:
            F.position[:] = [float(v) for v in np.asarray(position).reshape(3)]
        F.T16[:] = _f32(T).reshape(16).tolist()
        F.lower, F.upper = float(lower), float(upper)
        F.min_r, F.beam_a, F.beam_c = float(min_r), float(beam_a), float(beam_c)
        F.sJ[:] = _f32(sensorJacobian).reshape(3).tolist()
        F.rotVar[:] = _f32(rotationVariance).reshape(9).tolist()
        F.CSBT[:] = _f32(C_SB_transpose).reshape(9).tolist()
        F.PCBM[:] = _f32(P_mul_C_BM_transpose).reshape(3).tolist()
        F.BrBS[:] = _f32(B_r_BS_skew).reshape(9).tolist()
        F.near_filter = int(bool(near_filter))
        return F

    def process_fuse_device_batch(self, frames, resolution, length, asynchronous=False):
        """frames: list of make_frame(...) results, fused in order (mrslam_elev_process_fuse_device_batch); no host sync.
        asynchronous=True: the replay runs on the map's own streams, concurrently with later work on the library stream; any
        later call on this map (e.g. fused_total, get_map) joins it."""
        arr = (_lib.ElevFrame * len(frames))(*frames)
        fn = self._L.mrslam_elev_process_fuse_device_batch_async if asynchronous else self._L.mrslam_elev_process_fuse_device_batch
        check(fn(self._h, arr, len(frames), float(resolution), int(length)))

    def fused_total(self, reset=True):
        """cell fuses accumulated on the device by process_fuse_device(..., want_count=False) calls"""
        t = C.c_longlong()
        check(self._L.mrslam_elev_fused_total(self._h, int(bool(reset)), C.byref(t)))
        return t.value

    def Mapvar_update(self, length, var_update):
        check(self._L.mrslam_elev_mapvar_update(self._h, int(length), float(var_update)))

    def Map_optmove(self, opt_p, height_update, resolution, length):
        out = np.zeros(2, np.float32)
        check(self._L.mrslam_elev_map_optmove(self._h, ptr(_f32(opt_p)), float(height_update), float(resolution), int(length), ptr(out)))
        return out

    def Map_closeloop(self, update_position, height_update, length, resolution):
        check(self._L.mrslam_elev_map_closeloop(self._h, ptr(_f32(update_position)), float(height_update), int(length), float(resolution)))

    def Map_feature(self, length):
        """Map_feature (gpu_process.cu:1254): returns dict(elevation, var, colorR, colorG, colorB, rough, slope, traver, intensity),
        each of length*length in storage order."""
        n = int(length) * int(length)
        out = dict(elevation=np.zeros(n, np.float32), var=np.zeros(n, np.float32), colorR=np.zeros(n, np.int32), colorG=np.zeros(n, np.int32),
                   colorB=np.zeros(n, np.int32), rough=np.zeros(n, np.float32), slope=np.zeros(n, np.float32), traver=np.zeros(n, np.float32),
                   intensity=np.zeros(n, np.float32))
        check(self._L.mrslam_elev_map_feature(self._h, int(length), *[ptr(out[k]) for k in ("elevation", "var", "colorR", "colorG", "colorB", "rough",
                                                                                         "slope", "traver", "intensity")]))
        return out

    def Raytracing(self, length):
        check(self._L.mrslam_elev_raytracing(self._h, int(length)))

    PLANES = dict(elevation=0, variance=1, intensity=2, lowest=3, traver=4, R=5, G=6, B=7)

    def set_owned_rows(self, row_lo, row_hi):
        """tile-sharded global grid: this instance only fuses points that fall into storage rows [row_lo, row_hi)"""
        check(self._L.mrslam_elev_set_owned_rows(self._h, int(row_lo), int(row_hi)))

    def rows_to_device(self, plane, storage_row, n_rows, out):
        """copy n_rows storage rows of a plane into the CUDA tensor `out` (stream-ordered on the library stream)"""
        check(self._L.mrslam_elev_rows_copy(self._h, self.PLANES[plane], int(storage_row), int(n_rows), out.data_ptr(), 0))

    def rows_from_device(self, plane, storage_row, n_rows, src):
        check(self._L.mrslam_elev_rows_copy(self._h, self.PLANES[plane], int(storage_row), int(n_rows), src.data_ptr(), 1))

    def get_map(self):
        n = self.length * self.length
        out = dict(elevation=np.zeros(n, np.float32), variance=np.zeros(n, np.float32), intensity=np.zeros(n, np.float32),
                   R=np.zeros(n, np.int32), G=np.zeros(n, np.int32), B=np.zeros(n, np.int32), lowest=np.zeros(n, np.float32))
        check(self._L.mrslam_elev_get_map(self._h, *[out[k].ctypes.data for k in ("elevation", "variance", "intensity", "R", "G", "B", "lowest")]))
        return out

    def state(self):
        c, s = np.zeros(2, np.float32), np.zeros(2, np.int32)
        check(self._L.mrslam_elev_get_state(self._h, ptr(c), ptr(s)))
        return c, s


# ---- module-level free functions with the reference's names (one hidden map per process) ----
_MAP = None


def _map():
    global _MAP
    if _MAP is None:
        _MAP = ElevationMap()
    return _MAP


def Init_GPU_elevationmap(length, resolution, mahalanobisThr, obstacle_threshold):
    _map().Init_GPU_elevationmap(length, resolution, mahalanobisThr, obstacle_threshold)


def Move(current_Position, resolution, length):
    return _map().Move(current_Position, resolution, length)


def Process_points(*args, **kw):
    return _map().Process_points(*args, **kw)


def Fuse(*args, **kw):
    return _map().Fuse(*args, **kw)


def Map_feature(length):
    return _map().Map_feature(length)


def Raytracing(length):
    return _map().Raytracing(length)


def Mapvar_update(length, var_update):
    _map().Mapvar_update(length, var_update)


def Map_optmove(opt_p, height_update, resolution, length):
    return _map().Map_optmove(opt_p, height_update, resolution, length)


def Map_closeloop(update_position, height_update, length, resolution):
    _map().Map_closeloop(update_position, height_update, length, resolution)
