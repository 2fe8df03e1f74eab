"""`voxelocc` mirror: the Cython class of LoopDetection/generate_bev_cython_binary/wrapper.pyx:13-39.

    t = voxelocc.GPUTransformer(point, size, max_length, max_height, num_x, num_y, num_height, enough_large)
    t.transform(); out = t.retreive()        # (sic) -- same method names as the reference

`point` is a C-contiguous float32 SoA array (all x, all y, all z).  The reference computes indices on the GPU and
scatters on the host; here both steps run on the GPU inside `transform()`, `retreive()` returns the host buffer."""
import numpy as np

from ._lib import check, lib, ptr


class GPUTransformer:
    def __init__(self, point, size, max_length, max_height, num_x, num_y, num_height, enough_large):
        point = np.asarray(point)
        if point.dtype != np.float32 or point.ndim != 1 or not point.flags["C_CONTIGUOUS"]:
            raise ValueError("Buffer dtype mismatch, expected 'float' 1-d C-contiguous")  # what the Cython signature enforces
        if point.shape[0] < 3 * size:
            raise ValueError("point buffer shorter than 3*size")
        self.point = point
        self.size = int(size)
        self.cfg = (int(max_length), int(max_height), int(num_x), int(num_y), int(num_height), int(enough_large))
        self.grid_size = num_x * num_y * num_height * enough_large
        self._out = None

    def transform(self):
        out = np.zeros(self.grid_size * 3, dtype=np.float32)
        check(lib().mrslam_voxelocc_transform(ptr(self.point), self.size, *self.cfg, ptr(out)))
        self._out = out

    def retreive(self):
        if self._out is None:
            self.transform()
        return self._out
