"""Wire formats at the seams of the hot path: the key packing of dislam_msgs/Loop and the PointCloud2 blob of
dislam_msgs/SubMap -- without ROS (the messages are described by plain fields).

Reference: Mapping/src/dislam_msgs/msg/{SubMap,Loop,Loops}.msg; key packing Mapping/src/global_manager/src/global_manager.cpp:2587-2609
and LoopDetection/src/RING_ros/util.py:253-260; PointCloud2 decoding LoopDetection/src/RING_ros/main_RING.py:251-254."""
import ctypes as C

import numpy as np

from ._lib import check, lib


def robotid_to_key(robotid):
    """util.py:253-260 / GlobalManager::robotID2Key: the key of a robot's frame 0 is robotid_to_key(r) + 1"""
    return int(lib().mrslam_key_pack(int(robotid), -1))


def key_pack(robotid, keyframe_index):
    return int(lib().mrslam_key_pack(int(robotid), int(keyframe_index)))


def key_unpack(key):
    """-> (robot id, keyframe index) as ICPCheck decodes them (global_manager.cpp:1948-1951)"""
    L = lib()
    return int(L.mrslam_key_robot(int(key))), int(L.mrslam_key_index(int(key)))


def pointcloud2_unpack(data, n_points, point_step, offsets, skip_nans=True, as_cloud=False):
    """data: the PointCloud2 `data` blob (bytes / uint8 numpy array, or a CUDA uint8 torch tensor); offsets = dict(x=, y=, z=[, intensity=])
    byte offsets of the float32 fields inside a record.  Returns a CUDA float32 tensor [m, 4] (x, y, z, intensity) in record order with
    NaN records dropped (and a pygicp.Cloud over it when as_cloud=True)."""
    import torch
    on_dev = hasattr(data, "data_ptr")
    if on_dev:
        buf, p = data, data.data_ptr()
    else:
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        p = buf.ctypes.data
    out = torch.empty((max(int(n_points), 1), 4), dtype=torch.float32, device="cuda")
    m = C.c_int(0)
    torch.cuda.current_stream().synchronize()
    check(lib().mrslam_pointcloud2_unpack(p, 1 if on_dev else 0, int(n_points), int(point_step), int(offsets["x"]), int(offsets["y"]),
                                          int(offsets["z"]), int(offsets.get("intensity", -1)), 1 if skip_nans else 0, out.data_ptr(),
                                          C.byref(m)))
    if as_cloud:
        from . import pygicp
        return out[:m.value], pygicp.Cloud(device_ptr=out.data_ptr(), n=m.value, stride=4)
    return out[:m.value]
