// Compile-and-link check of the C++ host mirrors (runs a tiny registration when a GPU is present; otherwise only verifies
// that the library refuses to work without one).  Built and run by tests/test_cpp_adapters.py.
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "elevation_b200/gpu_process.hpp"
#include "fast_gicp_b200/fast_gicp.hpp"

struct PointXYZI {  // pcl::PointXYZI layout: x y z pad | intensity pad pad pad
  float x, y, z, pad0, intensity, pad1, pad2, pad3;
};

int main() {
  if (mrslam_init(0) != MRSLAM_OK) {
    std::printf("NO_GPU %s\n", mrslam_last_error());
    return 0;
  }
  using Reg = fast_gicp_b200::FastGICP<PointXYZI>;
  auto tgt = std::make_shared<Reg::Cloud>();
  auto src = std::make_shared<Reg::Cloud>();
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  for (int i = 0; i < 4000; i++) {
    float x = rnd() * 20 - 10, y = rnd() * 20 - 10, z = 0.3f * std::sin(0.5f * x) + 0.2f * std::cos(0.7f * y);
    tgt->push_back(PointXYZI{x, y, z, 1, (float)i, 0, 0, 0});
    src->push_back(PointXYZI{x - 0.10f, y + 0.05f, z - 0.02f, 1, (float)i, 0, 0, 0});  // target = source + (0.10, -0.05, 0.02)
  }
  Reg reg;
  reg.setNumThreads(8);
  reg.setTransformationEpsilon(1e-3);
  reg.setMaximumIterations(50);
  reg.setMaxCorrespondenceDistance(100.0);
  reg.setCorrespondenceRandomness(15);
  reg.setInputSource(src);
  reg.setInputTarget(tgt);
  Reg::Cloud out;
  reg.align(out);
  auto T = reg.getFinalTransformation();
  std::printf("converged %d t = %.4f %.4f %.4f fitness %.6f out %zu intensity %.0f\n", (int)reg.hasConverged(), T[3], T[7], T[11],
              reg.getFitnessScore(1.0), out.size(), out[7].intensity);
  bool ok = reg.hasConverged() && std::fabs(T[3] - 0.10f) < 0.02f && std::fabs(T[7] + 0.05f) < 0.02f && out.size() == src->size() &&
            out[7].intensity == 7.0f;
  // elevation free functions (POD overload)
  Init_GPU_elevationmap(60, 0.2f, 2.5f, 0.2f);
  float pos[3] = {0, 0, 0}, cc[2], sh[2];
  int si[2];
  Move(pos, 0.2f, 60, cc, si, sh);
  const int n = 1000;
  std::vector<float> x(n), y(n), z(n), var(n), xt(n), yt(n), zt(n), inten(n, 0.f);
  std::vector<int> mi(n), zero(n, 0);
  for (int i = 0; i < n; i++) { x[i] = rnd() * 10 - 5; y[i] = -1.5f - rnd() * 4; z[i] = -1.0f + 0.1f * rnd(); }
  const float T16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, sJ[3] = {0, 0, 1}, I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z9[9] = {0};
  Process_points(mi.data(), x.data(), y.data(), z.data(), var.data(), xt.data(), yt.data(), zt.data(), T16, n, -5.0, 5.0, 0.018f, 0.0006f,
                 0.0015f, sJ, Z9, I3, sJ, Z9);
  Fuse(60, n, mi.data(), zero.data(), zero.data(), zero.data(), inten.data(), zt.data(), var.data());
  int valid = 0;
  for (int i = 0; i < n; i++) valid += mi[i] >= 0;
  std::printf("elevation valid %d\n", valid);
  ok = ok && valid > 500;
  std::printf(ok ? "ADAPTERS_OK\n" : "ADAPTERS_FAIL\n");
  return ok ? 0 : 1;
}
