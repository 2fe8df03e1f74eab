// fast_gicp_b200/fast_gicp.hpp -- C++ host mirror of fast_gicp::FastGICP over the C-ABI (include/mrslam_b200.h).
//
// Two layers:
//  (1) fast_gicp_b200::FastGICP<PointT>: PCL-free class with the reference's method names and argument meaning
//      (FG/include/fast_gicp/gicp/fast_gicp.hpp:19-100, lsq_registration.hpp:15-85).  PointT is any POD whose first three
//      floats are x,y,z (pcl::PointXYZ: 4 floats, pcl::PointXYZI / PointNormal: 8 floats).  Used by tests/cpp and by ports
//      that do not link PCL.
//  (2) fast_gicp::LsqRegistration / fast_gicp::FastGICP<PointSource,PointTarget> deriving from
//      pcl::Registration<PointSource,PointTarget,float> -- same class names, namespace and virtual surface as the reference --
//      are in pcl_dropin.hpp (libfast_gicp.so), so Mapping/src/global_manager/src/global_manager.cpp:2435-2443 and
//      FG/src/python/main.cpp compile unchanged against <fast_gicp/gicp/fast_gicp.hpp> of this repository.
//
// Eigen matrices are column-major, the C-ABI is row-major: the adapters transpose.
#pragma once
#include <array>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../mrslam_b200.h"

namespace fast_gicp_b200 {

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };  // gicp_settings.hpp:6
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };                         // lsq_registration.hpp:13

using Matrix4f = std::array<float, 16>;   // row-major
using Matrix6d = std::array<double, 36>;  // row-major

template <typename PointT>
class FastGICP {
public:
  using Cloud = std::vector<PointT>;
  using CloudConstPtr = std::shared_ptr<const Cloud>;

  FastGICP() : h_(mrslam_gicp_create()) {
    if (!h_) throw std::runtime_error(std::string("mrslam_gicp_create: ") + mrslam_last_error());
  }
  ~FastGICP() { mrslam_gicp_destroy(h_); }
  FastGICP(const FastGICP&) = delete;
  FastGICP& operator=(const FastGICP&) = delete;

  void setNumThreads(int n) { mrslam_gicp_set_num_threads(h_, n); }
  void setCorrespondenceRandomness(int k) { check(mrslam_gicp_set_correspondence_randomness(h_, k)); }
  void setRegularizationMethod(RegularizationMethod m) { check(mrslam_gicp_set_regularization_method(h_, static_cast<int>(m))); }
  void setMaxCorrespondenceDistance(double d) { mrslam_gicp_set_max_correspondence_distance(h_, d); }
  void setMaximumIterations(int n) { mrslam_gicp_set_maximum_iterations(h_, n); }
  void setTransformationEpsilon(double e) { mrslam_gicp_set_transformation_epsilon(h_, e); }
  void setRotationEpsilon(double e) { mrslam_gicp_set_rotation_epsilon(h_, e); }
  void setInitialLambdaFactor(double f) { mrslam_gicp_set_initial_lambda_factor(h_, f); }
  void setDebugPrint(bool on) { mrslam_gicp_set_debug_print(h_, on ? 1 : 0); }
  void setLSQType(LSQ_OPTIMIZER_TYPE t) { mrslam_gicp_set_lsq_optimizer(h_, t == LSQ_OPTIMIZER_TYPE::GaussNewton ? MRSLAM_OPT_GN : MRSLAM_OPT_LM); }

  // fast_gicp_impl.hpp:72-90: early-out when the same cloud object is passed again (covariances are kept)
  void setInputSource(const CloudConstPtr& cloud) {
    if (cloud == source_) return;
    source_ = cloud;
    check(mrslam_gicp_set_input_source(h_, floats(cloud), static_cast<int>(cloud->size()), stride()));
  }
  void setInputTarget(const CloudConstPtr& cloud) {
    if (cloud == target_) return;
    target_ = cloud;
    check(mrslam_gicp_set_input_target(h_, floats(cloud), static_cast<int>(cloud->size()), stride()));
  }
  void swapSourceAndTarget() {
    source_.swap(target_);
    check(mrslam_gicp_swap_source_and_target(h_));
  }
  void clearSource() {
    source_.reset();
    mrslam_gicp_clear_source(h_);
  }
  void clearTarget() {
    target_.reset();
    mrslam_gicp_clear_target(h_);
  }

  // pcl::Registration::align(output, guess): failure is reported through hasConverged(), never thrown (lsq_registration_impl.hpp:69-72)
  void align(Cloud& output, const Matrix4f& guess = identity()) {
    int s = mrslam_gicp_align(h_, guess.data());
    if (s != MRSLAM_OK && s != MRSLAM_ERR_NOTCONV) std::fprintf(stderr, "FastGICP::align: %s\n", mrslam_last_error());
    output.resize(source_ ? source_->size() : 0);
    if (!output.empty()) {
      if (source_.get() != &output) output = *source_;  // keep the non-xyz fields like pcl::transformPointCloud
      mrslam_gicp_get_aligned(h_, reinterpret_cast<float*>(output.data()), stride());
    }
  }
  Matrix4f getFinalTransformation() const {
    Matrix4f T;
    mrslam_gicp_get_final_transformation(h_, T.data());
    return T;
  }
  bool hasConverged() const { return mrslam_gicp_has_converged(h_) != 0; }
  double getFitnessScore(double max_range = 1.7976931348623157e308) { return mrslam_gicp_get_fitness_score(h_, max_range); }
  Matrix6d getFinalHessian() const {
    Matrix6d H;
    mrslam_gicp_get_final_hessian(h_, H.data());
    return H;
  }
  double evaluateCost(const Matrix4f& relative_pose, Matrix6d* H = nullptr, std::array<double, 6>* b = nullptr) {
    double c = 0;
    check(mrslam_gicp_evaluate_cost(h_, relative_pose.data(), H ? H->data() : nullptr, b ? b->data() : nullptr, &c));
    return c;
  }
  // n x (row-major 4x4) like std::vector<Eigen::Matrix4d>
  std::vector<double> getSourceCovariances() {
    std::vector<double> c(source_ ? source_->size() * 16 : 0);
    if (!c.empty()) check(mrslam_gicp_get_source_covariances(h_, c.data(), static_cast<int>(source_->size())));
    return c;
  }
  std::vector<double> getTargetCovariances() {
    std::vector<double> c(target_ ? target_->size() * 16 : 0);
    if (!c.empty()) check(mrslam_gicp_get_target_covariances(h_, c.data(), static_cast<int>(target_->size())));
    return c;
  }
  void setSourceCovariances(const std::vector<double>& c) { check(mrslam_gicp_set_source_covariances(h_, c.data(), static_cast<int>(c.size() / 16))); }
  void setTargetCovariances(const std::vector<double>& c) { check(mrslam_gicp_set_target_covariances(h_, c.data(), static_cast<int>(c.size() / 16))); }

  mrslam_gicp_t* handle() { return h_; }
  static Matrix4f identity() { return Matrix4f{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

private:
  static_assert(sizeof(PointT) % sizeof(float) == 0 && sizeof(PointT) >= 3 * sizeof(float), "PointT must start with float x,y,z");
  static int stride() { return static_cast<int>(sizeof(PointT) / sizeof(float)); }
  static const float* floats(const CloudConstPtr& c) { return c->empty() ? nullptr : reinterpret_cast<const float*>(c->data()); }
  static void check(int s) {
    if (s != MRSLAM_OK) std::fprintf(stderr, "fast_gicp_b200: %s\n", mrslam_last_error());
  }
  mrslam_gicp_t* h_;
  CloudConstPtr source_, target_;
};

}  // namespace fast_gicp_b200

// The pcl::Registration-derived drop-in classes (fast_gicp::LsqRegistration / fast_gicp::FastGICP, the ones the ROS nodes
// instantiate) live in pcl_dropin.hpp + pcl_dropin_impl.hpp and are built into libfast_gicp.so; include <fast_gicp/gicp/fast_gicp.hpp>
// from this repository's include/ directory to get them.
