"""The C++ side of the drop-in boundary:
 * header-only mirrors (include/fast_gicp_b200/fast_gicp.hpp, include/elevation_b200/gpu_process.hpp) compile with a plain g++;
 * libgpu.so exports the nine elevation free functions with the reference's C++-mangled names: a caller that only has the
   reference's forward declarations (EM/src/ElevationMapping.cpp:44-50, SensorProcessorBase.cpp:34, RobotMotionMapUpdater.cpp:18)
   links against it with --no-undefined, and `nm -C` shows the expected signatures;
 * libfast_gicp.so holds the explicit instantiations of fast_gicp::FastGICP / LsqRegistration for PointXYZ / PointXYZI /
   PointNormal (FG/src/fast_gicp/gicp/fast_gicp.cpp:4-6); a caller written like GlobalManager::ICPCheck links against it without
   seeing the implementation header.
Eigen / PCL are not installed in this image: the signature-compatible shims under tests/cpp/abi_shim stand in (see their headers).
On the GPU box the programs also run."""
import os
import subprocess

import pytest

from conftest import ROOT

LIBDIR = os.path.join(ROOT, "mr_slam_b200")
SHIM = os.path.join(ROOT, "tests", "cpp", "abi_shim")


def _gxx(tmp_path, name, src, libs, extra_inc=()):
    exe = str(tmp_path / name)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include")]
    for i in extra_inc:
        cmd += ["-I", i]
    cmd += [os.path.join(ROOT, "tests", "cpp", src), "-o", exe, "-L", LIBDIR] + ["-l" + l for l in libs] + ["-Wl,-rpath," + LIBDIR, "-Wl,--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _build(tmp_path):
    return _gxx(tmp_path, "test_adapters", "test_adapters.cpp", ["mrslam_b200"])


def _run(exe, ok_word, need_gpu):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    if need_gpu:
        assert ok_word in r.stdout, r.stdout + r.stderr
    else:
        assert "NO_GPU" in r.stdout or ok_word in r.stdout, r.stdout + r.stderr


def test_adapters_compile_and_link(tmp_path):
    _run(_build(tmp_path), "ADAPTERS_OK", False)


def test_libgpu_exports_the_reference_symbols(tmp_path):
    out = subprocess.run(["nm", "-C", "--defined-only", os.path.join(LIBDIR, "libgpu.so")], capture_output=True, text=True).stdout
    want = [
        "Init_GPU_elevationmap(int, float, float, float)",                                       # gpu_process.cu:938
        "Move(float*, float, int, float*, int*, float*)",                                        # :1002
        "Process_points(int*, float*, float*, float*, float*, float*, float*, float*, Eigen::Matrix<float, 4, 4, 0, 4, 4>, int, double, double, "
        "float, float, float, Eigen::Matrix<float, 1, 3, 1, 1, 3>, Eigen::Matrix<float, 3, 3, 0, 3, 3>, Eigen::Matrix<float, 3, 3, 0, 3, 3>, "
        "Eigen::Matrix<float, 1, 3, 1, 1, 3>, Eigen::Matrix<float, 3, 3, 0, 3, 3>)",             # :1083
        "Mapvar_update(int, float)",                                                             # :1144
        "Fuse(int, int, int*, int*, int*, int*, float*, float*, float*)",                        # :1152
        "Map_optmove(float*, float, float, int, float*)",                                        # :1213
        "Map_closeloop(float*, float, int, float)",                                              # :1233
        "Map_feature(int, float*, float*, int*, int*, int*, float*, float*, float*, float*)",    # :1254
        "Raytracing(int)",                                                                       # :1302
    ]
    for w in want:
        assert (" T " + w) in out, w
    # a caller with the reference's forward declarations only
    _run(_gxx(tmp_path, "libgpu_caller", "test_libgpu_caller.cpp", ["gpu", "mrslam_b200"], [SHIM]), "LIBGPU_OK", False)


def test_libfast_gicp_has_the_explicit_instantiations(tmp_path):
    out = subprocess.run(["nm", "-C", "--defined-only", os.path.join(LIBDIR, "libfast_gicp.so")], capture_output=True, text=True).stdout
    for pt in ("pcl::PointXYZ", "pcl::PointXYZI", "pcl::PointNormal"):
        for member in ("FastGICP<%s, %s>::FastGICP()", "FastGICP<%s, %s>::setCorrespondenceRandomness(int)", "FastGICP<%s, %s>::swapSourceAndTarget()",
                       "FastGICP<%s, %s>::setSourceCovariances(", "FastGICP<%s, %s>::computeTransformation(", "LsqRegistration<%s, %s>::setRotationEpsilon(double)",
                       "LsqRegistration<%s, %s>::getFinalHessian() const", "LsqRegistration<%s, %s>::evaluateCost("):
            assert ("fast_gicp::" + member % (pt, pt)) in out, (pt, member)
        assert ("vtable for fast_gicp::FastGICP<%s, %s>" % (pt, pt)) in out
    _run(_gxx(tmp_path, "dropin", "test_fast_gicp_dropin.cpp", ["fast_gicp", "mrslam_b200"], [SHIM]), "DROPIN_OK", False)


@pytest.mark.gpu
def test_adapters_run_on_gpu(tmp_path, gpu):
    _run(_build(tmp_path), "ADAPTERS_OK", True)


@pytest.mark.gpu
def test_libgpu_caller_runs_on_gpu(tmp_path, gpu):
    _run(_gxx(tmp_path, "libgpu_caller", "test_libgpu_caller.cpp", ["gpu", "mrslam_b200"], [SHIM]), "LIBGPU_OK", True)


@pytest.mark.gpu
def test_pcl_dropin_align_runs_on_gpu(tmp_path, gpu):
    """align(out, guess) through a pcl::Registration<...>::Ptr to the drop-in fast_gicp::FastGICP, as ICPCheck drives it"""
    _run(_gxx(tmp_path, "dropin", "test_fast_gicp_dropin.cpp", ["fast_gicp", "mrslam_b200"], [SHIM]), "DROPIN_OK", True)
