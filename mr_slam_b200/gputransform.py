"""`gputransform` mirror: DiSCO's polar voxeliser class
(LoopDetection/src/disco_ros/tools/multi-layer-polar-gpu/cython/gputransform.pyx:13-39); same surface as voxelocc."""
import numpy as np

from ._lib import check, lib, ptr


class GPUTransformer:
    def __init__(self, point, size, max_length, max_height, num_ring, num_sector, num_height, enough_large):
        point = np.asarray(point)
        if point.dtype != np.float32 or point.ndim != 1 or not point.flags["C_CONTIGUOUS"]:
            raise ValueError("Buffer dtype mismatch, expected 'float' 1-d C-contiguous")
        if point.shape[0] < 3 * size:
            raise ValueError("point buffer shorter than 3*size")
        self.point = point
        self.size = int(size)
        self.cfg = (int(max_length), int(max_height), int(num_ring), int(num_sector), int(num_height), int(enough_large))
        self.grid_size = num_ring * num_sector * num_height * enough_large
        self._out = None

    def transform(self):
        out = np.zeros(self.grid_size * 3, dtype=np.float32)
        check(lib().mrslam_polar_transform(ptr(self.point), self.size, *self.cfg, ptr(out)))
        self._out = out

    def retreive(self):
        if self._out is None:
            self.transform()
        return self._out
