"""`torch_radon.ParallelBeam` mirror (LoopDetection/torch-radon/torch_radon/radon.py:62-87,139-167): forward only,
parallel beam, float32, 2-D -- the part RING uses (RING_ros/util.py:192-195).

    radon = ParallelBeam(det_count, angles)          # angles: tensor / ndarray / (start, end, n) tuple
    sino = radon.forward(x)                          # x: CUDA float32 [..., H, W] -> [..., n_angles, det_count]

`interp`: "texture" (default) emulates the texture unit's 8-bit interpolation weights like the reference's hardware
path, "exact" is exact fp32 bilinear.  `fast_trig=True` uses __sinf/__cosf like the reference kernel."""
import numpy as np
import torch

from ._lib import check, lib, ptr


class ParallelBeam:
    def __init__(self, det_count, angles, det_spacing=1.0, volume=None, interp="texture", fast_trig=True):
        if isinstance(angles, tuple) and len(angles) == 3:
            s, e, n = angles
            angles = np.linspace(s, e, n, endpoint=False)  # radon.py:30-32
        if not isinstance(angles, torch.Tensor):
            angles = torch.FloatTensor(np.asarray(angles, dtype=np.float32))
        if volume is not None:
            raise NotImplementedError("only the default centred Volume2D() is on the hot path")
        self.angles = angles.float()
        self.det_count = int(det_count)
        self.det_spacing = float(det_spacing)
        self.interp = 1 if interp == "texture" else 0
        self.fast_trig = 1 if fast_trig else 0

    def forward(self, x, angles=None, exec_cfg=None):
        if not x.is_cuda:
            raise RuntimeError("ParallelBeam.forward expects a CUDA tensor (there is no CPU fallback)")
        if x.dtype != torch.float32:
            raise NotImplementedError("only float32 is on the hot path")
        x = x.contiguous()
        ang = (angles if angles is not None else self.angles).to(x.device).contiguous()
        self.angles = ang if angles is None else self.angles
        shape = x.shape
        H, W = shape[-2], shape[-1]
        batch = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        out = torch.empty(tuple(shape[:-2]) + (ang.numel(), self.det_count), dtype=torch.float32, device=x.device)
        torch.cuda.current_stream(x.device).synchronize()
        check(lib().mrslam_radon_forward(ptr(x), batch, H, W, ptr(ang), ang.numel(), self.det_count, self.det_spacing, self.interp,
                                         self.fast_trig, ptr(out)))
        check(lib().mrslam_device_synchronize())
        return out
