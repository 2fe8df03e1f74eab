// extern "C" shim over the reference's CPU twin of the DiSCO polar voxeliser, compiled UNMODIFIED from
// /root/reference/LoopDetection/src/disco_ros/tools/multi-layer-polar-cpu/cython/src/{kernel.cpp,manager.cpp}
// (the reference sources are #included from where they lie; nothing is copied).  TEST INFRASTRUCTURE.
#include <manager.hh>
extern "C" int ref_polar_cpu(float* point, int size, int max_length, int max_height, int num_ring, int num_sector, int num_height,
                             int enough_large, float* out /* zeroed by the caller */) {
  int* dummy = nullptr;
  GPUTransformer t(point, size, dummy, dummy, dummy, max_length, max_height, num_ring, num_sector, num_height, enough_large);
  t.transform();
  t.retreive(out);
  return 0;
}
