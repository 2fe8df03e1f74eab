import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from mr_slam_b200 import synth, pygicp, _lib
_lib.init(0)
L = _lib.lib()
src, tgt, Tgt, Tg, Pa, Pb = synth.make_pair(0, "hdl64")
B = 32
src_dev = torch.from_numpy(np.stack([src] * B)).cuda(); tgt_dev = torch.from_numpy(np.stack([tgt] * B)).cuda()
guesses = np.stack([Tg.astype(np.float32)] * B)
def sync(): L.mrslam_device_synchronize()
for prof in (False, True, False):
    _lib.prof_enable(prof)
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        regs = []
        for i in range(B):
            g = pygicp.FastGICP(); g.set_correspondence_randomness(15); g.set_max_correspondence_distance(100.0)
            g.set_maximum_iterations(50); g.set_transformation_epsilon(1e-3); regs.append(g)
        t1 = time.perf_counter()
        clouds = []
        for i, g in enumerate(regs):
            cs = pygicp.Cloud(device_ptr=src_dev[i].data_ptr(), n=120000, stride=3)
            ct = pygicp.Cloud(device_ptr=tgt_dev[i].data_ptr(), n=120000, stride=3)
            g.set_input_source(cs); g.set_input_target(ct); clouds += [cs, ct]
        t2 = time.perf_counter(); sync(); t3 = time.perf_counter()
        Ts = pygicp.align_batch(regs, guesses)
        t4 = time.perf_counter()
        del regs, clouds
        sync(); t5 = time.perf_counter()
        print(f"prof={prof} rep{rep}: handles {1e3*(t1-t0):.2f} ms, clouds {1e3*(t2-t1):.2f} (+sync {1e3*(t3-t2):.2f}), align_batch {1e3*(t4-t3):.2f}, destroy {1e3*(t5-t4):.2f}")
    if prof: print({k: (round(v[0],2), v[1]) for k, v in _lib.prof_get().items()})
