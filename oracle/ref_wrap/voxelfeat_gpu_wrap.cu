// extern "C" shim over the reference's voxelfeat classes (generate_bev_pointfeat_cython/src/manager.cu + kernel.cu,
// compiled UNMODIFIED).  TEST INFRASTRUCTURE: GPU box only.
#include <manager.hh>
#include <vector>
extern "C" int ref_voxelfeat_features(float* point_aos, int size, int featsize, int k, int* nbr, float* eig, float* out) {
  GPUFeatureExtractor fe(point_aos, size, featsize, k, nbr, eig);
  fe.get_features(out);
  return 0;
}
extern "C" int ref_voxelfeat_bev(float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_height, int featsize,
                                 float* out) {
  std::vector<int> a(size, 0), b(size, 0), c(size, 0);
  GPUTransformer t(point, size, a.data(), b.data(), c.data(), max_length, max_height, num_x, num_y, num_height, featsize);
  t.transform();
  t.retreive(out);
  return 0;
}
