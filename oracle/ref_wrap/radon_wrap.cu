// extern "C" shim over the reference's torch-radon forward projection (torch-radon/src/forward.cu + texture.cu +
// parameter_classes.cu, compiled UNMODIFIED for sm_100).  TEST INFRASTRUCTURE: GPU box only.
#include <cuda_runtime.h>
#include "parameter_classes.h"
#include "texture.h"
#include "forward.h"
extern "C" int ref_radon_forward(const float* h_img, int batch, int H, int W, const float* h_angles, int n_angles, int det_count,
                                 float det_spacing, float* h_out) {
  float *d_img, *d_ang, *d_out;
  cudaMalloc(&d_img, sizeof(float) * batch * H * W);
  cudaMalloc(&d_ang, sizeof(float) * n_angles);
  cudaMalloc(&d_out, sizeof(float) * batch * n_angles * det_count);
  cudaMemcpy(d_img, h_img, sizeof(float) * batch * H * W, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ang, h_angles, sizeof(float) * n_angles, cudaMemcpyHostToDevice);
  {
    TextureCache cache(8);
    // Volume2D().to_cfg(): VolumeCfg(0, h, w, 0,0,0, 1,1,1, false); Projection.parallel_beam: ProjectionCfg(det_count, det_spacing)
    VolumeCfg vol(0, H, W, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f, false);
    ProjectionCfg proj(det_count, det_spacing);
    proj.n_angles = n_angles;
    ExecCfg exec(16, 16, 1, 4);  // radon.py:24
    radon_forward_cuda<float>(d_img, d_ang, d_out, cache, vol, proj, exec, batch, 0);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h_out, d_out, sizeof(float) * batch * n_angles * det_count, cudaMemcpyDeviceToHost);
  cudaFree(d_img);
  cudaFree(d_ang);
  cudaFree(d_out);
  return (int)cudaGetLastError();
}
