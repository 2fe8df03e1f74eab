"""ctypes binding of libmrslam_b200.so (the C-ABI declared in include/mrslam_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no CPU
fallback: if the library is missing, or no B200 is present when a compute entry point is called,
the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRSLAM_LIB", os.path.join(_HERE, "libmrslam_b200.so"))  # MRSLAM_LIB: development variants only
_LIB = None


class MrSlamError(RuntimeError):
    pass


def _p(dtype):
    return np.ctypeslib.ndpointer(dtype=dtype, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MrSlamError(
                "libmrslam_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "there is no CPU fallback for the CUDA path")
        _LIB = C.CDLL(LIB_PATH)
        _declare(_LIB)
    return _LIB


def check(status):
    if status != 0:
        raise MrSlamError("mrslam error %d: %s" % (status, lib().mrslam_last_error().decode()))
    return status


class ElevFrame(C.Structure):
    """mrslam_elev_frame_t (include/mrslam_b200.h)"""
    _fields_ = [("d_x", C.c_void_p), ("d_y", C.c_void_p), ("d_z", C.c_void_p), ("n", C.c_int), ("do_move", C.c_int),
                ("position", C.c_float * 3), ("T16", C.c_float * 16), ("lower", C.c_double), ("upper", C.c_double),
                ("min_r", C.c_float), ("beam_a", C.c_float), ("beam_c", C.c_float), ("sJ", C.c_float * 3), ("rotVar", C.c_float * 9),
                ("CSBT", C.c_float * 9), ("PCBM", C.c_float * 3), ("BrBS", C.c_float * 9), ("near_filter", C.c_int)]


def _declare(L):
    vp, i, f, d, ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong
    fp, ip, dp = C.c_void_p, C.c_void_p, C.c_void_p  # raw pointers (host or device) passed as integers
    L.mrslam_init.argtypes = [i]
    L.mrslam_last_error.restype = C.c_char_p
    L.mrslam_version.restype = C.c_char_p
    L.mrslam_stream.restype = vp
    L.mrslam_kernel_launch_count.restype = C.c_ulonglong
    L.mrslam_prof_enable.argtypes = [i]
    L.mrslam_prof_get.argtypes = [i, C.POINTER(d), C.POINTER(ll)]

    L.mrslam_cloud_create_host.argtypes = [fp, i, i]
    L.mrslam_cloud_create_host.restype = vp
    L.mrslam_cloud_create_device.argtypes = [fp, i, i]
    L.mrslam_cloud_create_device.restype = vp
    L.mrslam_cloud_create_device_batch.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i, i, C.POINTER(vp)]
    L.mrslam_key_pack.argtypes = [i, ll]
    L.mrslam_key_pack.restype = C.c_ulonglong
    L.mrslam_key_robot.argtypes = [C.c_ulonglong]
    L.mrslam_key_index.argtypes = [C.c_ulonglong]
    L.mrslam_key_index.restype = ll
    L.mrslam_pointcloud2_unpack.argtypes = [fp, i, i, i, i, i, i, i, i, fp, C.POINTER(C.c_int)]
    L.mrslam_ringpp_build_neighbors.argtypes = [fp, i, i, i, ip, fp, fp, fp]
    L.mrslam_keyframe_merge.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i, i, fp, f, f, i, fp, i, C.POINTER(C.c_int)]
    L.mrslam_cloud_destroy.argtypes = [vp]
    L.mrslam_cloud_size.argtypes = [vp]
    L.mrslam_cloud_knn.argtypes = [vp, i, ip, fp]

    L.mrslam_gicp_create.restype = vp
    L.mrslam_gicp_destroy.argtypes = [vp]
    for name in ("set_num_threads", "set_correspondence_randomness", "set_regularization_method", "set_maximum_iterations",
                 "set_lsq_optimizer", "set_debug_print"):
        getattr(L, "mrslam_gicp_" + name).argtypes = [vp, i]
    for name in ("set_max_correspondence_distance", "set_transformation_epsilon", "set_rotation_epsilon",
                 "set_initial_lambda_factor"):
        getattr(L, "mrslam_gicp_" + name).argtypes = [vp, d]
    L.mrslam_gicp_set_input_source.argtypes = [vp, fp, i, i]
    L.mrslam_gicp_set_input_target.argtypes = [vp, fp, i, i]
    L.mrslam_gicp_set_input_source_cloud.argtypes = [vp, vp]
    L.mrslam_gicp_set_input_target_cloud.argtypes = [vp, vp]
    for name in ("swap_source_and_target", "clear_source", "clear_target"):
        getattr(L, "mrslam_gicp_" + name).argtypes = [vp]
    for name in ("set_source_covariances", "set_target_covariances", "get_source_covariances", "get_target_covariances"):
        getattr(L, "mrslam_gicp_" + name).argtypes = [vp, dp, i]
    L.mrslam_gicp_align.argtypes = [vp, fp]
    L.mrslam_gicp_align_batch.argtypes = [C.POINTER(vp), i, fp]
    L.mrslam_gicp_prepare_batch.argtypes = [C.POINTER(vp), i, i]
    L.mrslam_gicp_get_final_transformation.argtypes = [vp, fp]
    L.mrslam_gicp_get_final_transformation_d.argtypes = [vp, dp]
    L.mrslam_gicp_has_converged.argtypes = [vp]
    L.mrslam_gicp_get_num_iterations.argtypes = [vp]
    L.mrslam_gicp_get_final_hessian.argtypes = [vp, dp]
    L.mrslam_gicp_get_eval_counts.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.mrslam_gicp_get_fitness_score.argtypes = [vp, d]
    L.mrslam_gicp_get_fitness_score.restype = d
    L.mrslam_gicp_get_fitness_score_batch.argtypes = [C.POINTER(vp), i, d, dp]
    L.mrslam_gicp_get_aligned.argtypes = [vp, fp, i]
    L.mrslam_gicp_evaluate_cost.argtypes = [vp, fp, dp, dp, C.POINTER(d)]
    L.mrslam_gicp_debug_last_correspondences.argtypes = [vp, ip, fp, dp]
    L.mrslam_gicp_compute_error.argtypes = [vp, dp, C.POINTER(d)]
    L.mrslam_gicp_stats_enable.argtypes = [i]
    L.mrslam_gicp_stats_pairs.restype = C.c_ulonglong
    L.mrslam_gicp_stats_get.argtypes = [i]
    L.mrslam_gicp_stats_get.restype = C.c_ulonglong

    if hasattr(L, "mrslam_elev_create"):
        L.mrslam_elev_create.restype = vp
        L.mrslam_elev_destroy.argtypes = [vp]
        L.mrslam_elev_init.argtypes = [vp, i, f, f, f]
        L.mrslam_elev_move.argtypes = [vp, fp, f, i, fp, ip, fp]
        L.mrslam_elev_process_points.argtypes = [vp, ip] + [fp] * 7 + [fp, i, d, d, f, f, f] + [fp] * 5
        L.mrslam_elev_fuse.argtypes = [vp, i, i, ip, ip, ip, ip, fp, fp, fp, C.POINTER(ll)]
        L.mrslam_elev_process_fuse_device.argtypes = [vp, fp, fp, fp, i, fp, d, d, f, f, f] + [fp] * 5 + [i, C.POINTER(ll)]
        L.mrslam_elev_mapvar_update.argtypes = [vp, i, f]
        L.mrslam_elev_process_fuse_device_batch.argtypes = [vp, C.POINTER(ElevFrame), i, f, i]
        L.mrslam_elev_process_fuse_device_batch_async.argtypes = [vp, C.POINTER(ElevFrame), i, f, i]
        L.mrslam_elev_fused_total.argtypes = [vp, i, C.POINTER(ll)]
        L.mrslam_elev_map_optmove.argtypes = [vp, fp, f, f, i, fp]
        L.mrslam_elev_map_closeloop.argtypes = [vp, fp, f, i, f]
        L.mrslam_elev_map_feature.argtypes = [vp, i] + [fp] * 9
        L.mrslam_elev_raytracing.argtypes = [vp, i]
        L.mrslam_elev_set_owned_rows.argtypes = [vp, i, i]
        L.mrslam_elev_rows_copy.argtypes = [vp, i, i, i, fp, i]
        L.mrslam_elev_get_map.argtypes = [vp] + [C.c_void_p] * 7
        L.mrslam_elev_get_state.argtypes = [vp, fp, ip]
    if hasattr(L, "mrslam_voxelocc_transform"):
        L.mrslam_voxelocc_transform.argtypes = [fp, i, i, i, i, i, i, i, fp]
        L.mrslam_polar_transform.argtypes = [fp, i, i, i, i, i, i, i, fp]
        L.mrslam_voxelfeat_features.argtypes = [fp, i, i, i, ip, fp, fp]
        L.mrslam_voxelfeat_transform.argtypes = [fp, i, i, i, i, i, i, i, fp]
        L.mrslam_voxelocc_bev_device.argtypes = [fp, C.c_void_p, i, i, i, i, i, fp]
        L.mrslam_sort_pairs_u32.argtypes = [fp, fp, i, i, C.c_uint, i, fp, fp, C.POINTER(C.c_int)]
        L.mrslam_ring_scans_to_bev_device.argtypes = [fp, C.c_void_p, i, i, f, f, i, i, i, i, fp]
    if hasattr(L, "mrslam_radon_forward"):
        L.mrslam_radon_forward.argtypes = [fp, i, i, i, fp, i, i, f, i, i, fp]
        L.mrslam_ring_tiring.argtypes = [fp, i, i, i, i, fp]
        L.mrslam_ring_fast_corr.argtypes = [fp, fp, i, i, i, i, fp, ip, fp]
        L.mrslam_ring_generate.argtypes = [fp, i, i, fp, i, i, i, fp, fp]
        L.mrslam_ring_solve_translation.argtypes = [fp, fp, i, i, i, f, i, fp, fp]
        L.mrslam_disco_rel_ori.argtypes = [fp, fp, fp, fp, i, i, i, fp, ip]


def ptr(a):
    """raw address of a numpy array (host) or a torch tensor (device/host)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def init(device=-1):
    check(lib().mrslam_init(int(device)))


PROF_IDS = ["knn_search", "cov", "nn1_search", "linearize", "error", "cloud_build", "elev_process", "elev_fuse"]


def prof_enable(on=True):
    check(lib().mrslam_prof_enable(1 if on else 0))


def prof_get():
    """{name: (total_ms, launches)} of the instrumented kernels since prof_enable()"""
    out = {}
    for k, name in enumerate(PROF_IDS):
        ms, n = C.c_double(), C.c_longlong()
        check(lib().mrslam_prof_get(k, C.byref(ms), C.byref(n)))
        out[name] = (ms.value, n.value)
    return out


def kernel_launch_count():
    return int(lib().mrslam_kernel_launch_count())
