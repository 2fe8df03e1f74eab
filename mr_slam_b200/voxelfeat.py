"""`voxelfeat` mirror (RING++): the Cython classes of LoopDetection/generate_bev_pointfeat_cython/wrapper.pyx:17-60.

    fe = voxelfeat.GPUFeatureExtractor(pc_flatten, size, 13, 30, k_indices, k_eigens); feat = fe.get_features()
    t  = voxelfeat.GPUTransformer(point_soa, size, max_length, max_height, num_x, num_y, num_height, featsize)
    t.transform(); bev = t.retreive()

The reference's feature-BEV kernel updates a running maximum with a non-atomic compare-then-write (racy); this
implementation returns the race-free result it intends: per cell and channel the maximum of the positive values."""
import numpy as np

from ._lib import check, lib, ptr


def _chk(a, dtype, name):
    a = np.asarray(a)
    if a.dtype != dtype or a.ndim != 1 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("Buffer dtype mismatch for %s, expected 1-d C-contiguous %s" % (name, np.dtype(dtype).name))
    return a


class GPUFeatureExtractor:
    def __init__(self, point, size, featsize, k, neighbors_indices, eigens):
        self.point = _chk(point, np.float32, "point")
        self.nbr = _chk(neighbors_indices, np.int32, "neighbors_indices")
        self.eig = _chk(eigens, np.float32, "eigens")
        self.size, self.featsize, self.k = int(size), int(featsize), int(k)
        if len(self.point) < 3 * self.size or len(self.nbr) < self.size * self.k or len(self.eig) < 5 * self.size:
            raise ValueError("input buffers shorter than size implies")

    def get_features(self):
        out = np.zeros(self.featsize * self.size, dtype=np.float32)
        check(lib().mrslam_voxelfeat_features(ptr(self.point), self.size, self.featsize, self.k, ptr(self.nbr), ptr(self.eig), ptr(out)))
        return out


class GPUTransformer:
    def __init__(self, point, size, max_length, max_height, num_x, num_y, num_height, featsize):
        self.point = _chk(point, np.float32, "point")
        self.size = int(size)
        self.cfg = (int(max_length), int(max_height), int(num_x), int(num_y), int(num_height), int(featsize))
        if len(self.point) < self.size * int(featsize):
            raise ValueError("point buffer shorter than featsize*size")
        self.grid_size = num_x * num_y * num_height * featsize
        self._out = None

    def transform(self):
        out = np.zeros(self.grid_size, dtype=np.float32)
        check(lib().mrslam_voxelfeat_transform(ptr(self.point), self.size, *self.cfg, ptr(out)))
        self._out = out

    def retreive(self):
        if self._out is None:
            self.transform()
        return self._out
