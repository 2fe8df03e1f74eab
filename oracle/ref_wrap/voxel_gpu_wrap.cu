// extern "C" shim over the reference's GPU voxeliser class (voxelocc: generate_bev_cython_binary/src/manager.cu, or the
// DiSCO polar one: multi-layer-polar-gpu/cython/src/manager.cu -- selected by the include path), compiled UNMODIFIED.
// TEST INFRASTRUCTURE: runs only on the GPU box inside `-m gpu` tests.
#include <manager.hh>
#include <vector>
extern "C" int REF_FN(float* point, int size, int max_length, int max_height, int num_a, int num_b, int num_height, int enough_large,
                      float* out /* zeroed by the caller */) {
  std::vector<int> a(size, 0), b(size, 0), c(size, 0);  // wrapper.pyx passes zero-initialised index arrays
  GPUTransformer t(point, size, a.data(), b.data(), c.data(), max_length, max_height, num_a, num_b, num_height, enough_large);
  t.transform();
  t.retreive(out);
  return 0;
}
