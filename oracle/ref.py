"""ctypes access to the pieces of the REAL reference compiled from their own sources into oracle/_ref/
(see oracle/build_ref.sh).  TEST INFRASTRUCTURE: imported by tests/ only.  The .so files are built in the
authoring container (where /root/reference exists) and travel to the GPU box; nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = {}


def available(name):
    return os.path.exists(os.path.join(_DIR, "lib%s_ref.so" % name))


def _lib(name):
    if name not in _cache:
        _cache[name] = C.CDLL(os.path.join(_DIR, "lib%s_ref.so" % name))
    return _cache[name]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _vox(libname, fn, soa, size, max_length, max_height, a, b, num_height, enough_large):
    L = _lib(libname)
    f = getattr(L, fn)
    f.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    p = _f32(soa).copy()
    out = np.zeros(a * b * num_height * enough_large * 3, np.float32)
    f(p.ctypes.data, size, max_length, max_height, a, b, num_height, enough_large, out.ctypes.data)
    return out


def polar_cpu(soa, size, max_length, max_height, num_ring, num_sector, num_height, enough_large):
    return _vox("polar_cpu", "ref_polar_cpu", soa, size, max_length, max_height, num_ring, num_sector, num_height, enough_large)


def polar_gpu(soa, size, max_length, max_height, num_ring, num_sector, num_height, enough_large):
    return _vox("polar_gpu", "ref_polar_gpu", soa, size, max_length, max_height, num_ring, num_sector, num_height, enough_large)


def voxelocc_gpu(soa, size, max_length, max_height, num_x, num_y, num_height, enough_large):
    return _vox("voxelocc", "ref_voxelocc_gpu", soa, size, max_length, max_height, num_x, num_y, num_height, enough_large)


def voxelfeat_features(pc_aos, size, featsize, k, nbr, eig):
    L = _lib("voxelfeat")
    L.ref_voxelfeat_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    pc, eig = _f32(pc_aos).copy(), _f32(eig).copy()
    nbr = np.ascontiguousarray(nbr, np.int32).copy()
    out = np.zeros(size * featsize, np.float32)
    L.ref_voxelfeat_features(pc.ctypes.data, size, featsize, k, nbr.ctypes.data, eig.ctypes.data, out.ctypes.data)
    return out


def voxelfeat_bev(soa, size, max_length, max_height, num_x, num_y, num_height, featsize):
    L = _lib("voxelfeat")
    L.ref_voxelfeat_bev.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    p = _f32(soa).copy()
    out = np.zeros(num_x * num_y * num_height * featsize, np.float32)
    L.ref_voxelfeat_bev(p.ctypes.data, size, max_length, max_height, num_x, num_y, num_height, featsize, out.ctypes.data)
    return out


def radon_forward(img, angles, det_count, det_spacing=1.0):
    L = _lib("radon")
    L.ref_radon_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    img = _f32(img)
    angles = _f32(angles)
    B, H, W = img.shape
    out = np.zeros((B, len(angles), det_count), np.float32)
    rc = L.ref_radon_forward(img.ctypes.data, B, H, W, angles.ctypes.data, len(angles), det_count, det_spacing, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("reference radon kernel failed with CUDA error %d" % rc)
    return out


class Symbolic:
    """torch-radon's analytic phantom (tests/utils.py: random_symbolic_function / symbolic_discretize / symbolic_forward)."""

    def __init__(self, h, w):
        L = self.L = _lib("symbolic")
        L.ref_sym_create.restype = C.c_void_p
        L.ref_sym_create.argtypes = [C.c_float, C.c_float]
        L.ref_sym_destroy.argtypes = [C.c_void_p]
        L.ref_sym_add_gaussian.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.ref_sym_add_ellipse.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.ref_sym_discretize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_sym_forward.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        self.h, self.w = h, w
        self._f = C.c_void_p(L.ref_sym_create(float(h), float(w)))

    def __del__(self):
        try:
            self.L.ref_sym_destroy(self._f)
        except Exception:
            pass

    def add_gaussian(self, k, cx, cy, a, b):
        self.L.ref_sym_add_gaussian(self._f, k, cx, cy, a, b)

    def add_ellipse(self, k, cx, cy, r, a):
        self.L.ref_sym_add_ellipse(self._f, k, cx, cy, r, a)

    def discretize(self):
        out = np.zeros((self.h, self.w), np.float32)
        self.L.ref_sym_discretize(self._f, out.ctypes.data, self.h, self.w)
        return out

    def forward(self, angles, det_count, det_spacing=1.0):
        angles = _f32(angles)
        out = np.zeros((len(angles), det_count), np.float32)
        self.L.ref_sym_forward(self._f, det_count, det_spacing, angles.ctypes.data, len(angles), out.ctypes.data)
        return out


class ElevationRef:
    """The reference's own elevation kernels (one hidden map per process, as in the reference)."""

    def __init__(self, length, resolution, mahal=2.5, obst=0.2):
        L = self.L = _lib("elev")
        L.ref_elev_init.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_elev_move.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_elev_process_points.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_double, C.c_double, C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 5
        L.ref_elev_fuse.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7
        L.ref_elev_mapvar_update.argtypes = [C.c_int, C.c_float]
        L.ref_elev_get_map.argtypes = [C.c_int] + [C.c_void_p] * 6
        L.ref_elev_map_feature.argtypes = [C.c_int] + [C.c_void_p] * 9
        L.ref_elev_raytracing.argtypes = [C.c_int]
        self.length, self.res = int(length), float(resolution)
        L.ref_elev_init(self.length, self.res, mahal, obst)

    def move(self, pos3):
        p = _f32(pos3).copy()
        c, s, a = np.zeros(2, np.float32), np.zeros(2, np.int32), np.zeros(2, np.float32)
        self.L.ref_elev_move(p.ctypes.data, self.res, self.length, c.ctypes.data, s.ctypes.data, a.ctypes.data)
        return c, s, a

    def process_points(self, x, y, z, T, lower, upper, min_r, beam_a, beam_c, sJ, rotVar, CSBT, PCBM, BrBS):
        n = len(x)
        x, y, z = _f32(x).copy(), _f32(y).copy(), _f32(z).copy()
        mi = np.zeros(n, np.int32)
        var, xt, yt, zt = (np.zeros(n, np.float32) for _ in range(4))
        mats = [_f32(T).reshape(16), _f32(sJ).reshape(3), _f32(rotVar).reshape(9), _f32(CSBT).reshape(9), _f32(PCBM).reshape(3),
                _f32(BrBS).reshape(9)]
        self.L.ref_elev_process_points(mi.ctypes.data, x.ctypes.data, y.ctypes.data, z.ctypes.data, var.ctypes.data, xt.ctypes.data,
                                       yt.ctypes.data, zt.ctypes.data, mats[0].ctypes.data, n, float(lower), float(upper), min_r, beam_a,
                                       beam_c, *[m.ctypes.data for m in mats[1:]])
        return mi, var, xt, yt, zt

    def fuse(self, idx, R, G, B, inten, h, v):
        i32 = lambda a: np.ascontiguousarray(a, np.int32).copy()
        idx, R, G, B = i32(idx), i32(R), i32(G), i32(B)
        inten, h, v = _f32(inten).copy(), _f32(h).copy(), _f32(v).copy()
        self.L.ref_elev_fuse(self.length, len(idx), idx.ctypes.data, R.ctypes.data, G.ctypes.data, B.ctypes.data, inten.ctypes.data,
                             h.ctypes.data, v.ctypes.data)

    def mapvar_update(self, upd):
        self.L.ref_elev_mapvar_update(self.length, float(upd))

    def map_feature(self):
        """the reference's Map_feature: NOTE rough / slope / traver are only defined where elevation != -10, and the call updates
        the resident traversability plane (which Raytracing reads)"""
        n = self.length * self.length
        out = dict(elevation=np.zeros(n, np.float32), var=np.zeros(n, np.float32), colorR=np.zeros(n, np.int32), colorG=np.zeros(n, np.int32),
                   colorB=np.zeros(n, np.int32), rough=np.zeros(n, np.float32), slope=np.zeros(n, np.float32), traver=np.zeros(n, np.float32),
                   intensity=np.zeros(n, np.float32))
        self.L.ref_elev_map_feature(self.length, *[out[k].ctypes.data for k in ("elevation", "var", "colorR", "colorG", "colorB", "rough", "slope",
                                                                              "traver", "intensity")])
        return out

    def raytracing(self):
        self.L.ref_elev_raytracing(self.length)

    def get_map(self):
        n = self.length * self.length
        out = dict(elevation=np.zeros(n, np.float32), variance=np.zeros(n, np.float32), R=np.zeros(n, np.int32), G=np.zeros(n, np.int32),
                   B=np.zeros(n, np.int32), intensity=np.zeros(n, np.float32))
        self.L.ref_elev_get_map(self.length, *[out[k].ctypes.data for k in ("elevation", "variance", "R", "G", "B", "intensity")])
        return out
