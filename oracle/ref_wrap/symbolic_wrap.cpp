// extern "C" shim over torch-radon's analytic test oracle (torch-radon/src/symbolic.cpp + parameter_classes, compiled
// UNMODIFIED): random Gaussians/ellipses, their discretisation and exact line integrals -- what the reference's own
// tests/test_parallel_beam.py compares the forward projection with.  TEST INFRASTRUCTURE.
#include "symbolic.h"
extern "C" {
void* ref_sym_create(float h, float w) { return new SymbolicFunction(h, w); }
void ref_sym_destroy(void* f) { delete (SymbolicFunction*)f; }
void ref_sym_add_gaussian(void* f, float k, float cx, float cy, float a, float b) { ((SymbolicFunction*)f)->add_gaussian(k, cx, cy, a, b); }
void ref_sym_add_ellipse(void* f, float k, float cx, float cy, float r, float a) { ((SymbolicFunction*)f)->add_ellipse(k, cx, cy, r, a); }
void ref_sym_move(void* f, float dx, float dy) { ((SymbolicFunction*)f)->move(dx, dy); }
void ref_sym_discretize(void* f, float* data, int h, int w) { ((SymbolicFunction*)f)->discretize(data, h, w); }
void ref_sym_forward(void* f, int det_count, float det_spacing, const float* angles, int n_angles, float* sino) {
  ProjectionCfg proj(det_count, det_spacing);
  proj.n_angles = n_angles;
  symbolic_forward(*(SymbolicFunction*)f, proj, angles, n_angles, sino);
}
}
