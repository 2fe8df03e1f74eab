// fast_gicp_b200/pcl_dropin_impl.hpp -- out-of-line members of the drop-in classes (see pcl_dropin.hpp).  Logic-free by
// intent: parameter plumbing and row-major <-> Eigen conversions around the C-ABI calls.
#pragma once
#include <cmath>
#include <cstdio>

#include "pcl_dropin.hpp"

namespace fast_gicp {
namespace detail {
inline void warn(int status, const char* what) {
  if (status != MRSLAM_OK && status != MRSLAM_ERR_NOTCONV) std::fprintf(stderr, "fast_gicp (mrslam_b200) %s: %s\n", what, mrslam_last_error());
}
template <typename M>
inline void to_rows16(const M& m, float* out) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out[4 * i + j] = static_cast<float>(m(i, j));
}
template <typename PointT>
inline int floats_per_point() {
  static_assert(sizeof(PointT) % sizeof(float) == 0 && sizeof(PointT) >= 3 * sizeof(float), "point type must start with float x, y, z");
  return static_cast<int>(sizeof(PointT) / sizeof(float));
}
inline void covs_to_rows(const CovarianceList& covs, std::vector<double>& rows) {
  rows.resize(covs.size() * 16);
  for (std::size_t p = 0; p < covs.size(); p++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) rows[p * 16 + 4 * i + j] = covs[p](i, j);
}
inline void rows_to_covs(const std::vector<double>& rows, CovarianceList& covs) {
  covs.resize(rows.size() / 16);
  for (std::size_t p = 0; p < covs.size(); p++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) covs[p](i, j) = rows[p * 16 + 4 * i + j];
}
}  // namespace detail

// ---- LsqRegistration (lsq_registration_impl.hpp:9-22 for the defaults) -------------------------------------------------
template <typename PointSource, typename PointTarget>
LsqRegistration<PointSource, PointTarget>::LsqRegistration()
    : rotation_epsilon_(2e-3), lsq_optimizer_type_(LSQ_OPTIMIZER_TYPE::LevenbergMarquardt), lm_max_iterations_(10),
      lm_init_lambda_factor_(1e-9), lm_lambda_(-1.0), lm_debug_print_(false), h_(mrslam_gicp_create()) {
  this->reg_name_ = "LsqRegistration";
  max_iterations_ = 64;
  transformation_epsilon_ = 5e-4;
  final_hessian_.setIdentity();
  if (!h_) std::fprintf(stderr, "fast_gicp (mrslam_b200): %s\n", mrslam_last_error());
}
template <typename PointSource, typename PointTarget>
LsqRegistration<PointSource, PointTarget>::~LsqRegistration() {
  if (h_) mrslam_gicp_destroy(h_);
}
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::setRotationEpsilon(double eps) {
  rotation_epsilon_ = eps;
}
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::setInitialLambdaFactor(double f) {
  lm_init_lambda_factor_ = f;
}
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::setDebugPrint(bool on) {
  lm_debug_print_ = on;
}
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::setLSQType(LSQ_OPTIMIZER_TYPE type) {
  lsq_optimizer_type_ = type;
}
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::push_parameters() {
  mrslam_gicp_set_maximum_iterations(h_, max_iterations_);
  mrslam_gicp_set_transformation_epsilon(h_, transformation_epsilon_);
  mrslam_gicp_set_rotation_epsilon(h_, rotation_epsilon_);
  mrslam_gicp_set_initial_lambda_factor(h_, lm_init_lambda_factor_);
  mrslam_gicp_set_debug_print(h_, lm_debug_print_ ? 1 : 0);
  mrslam_gicp_set_lsq_optimizer(h_, lsq_optimizer_type_ == LSQ_OPTIMIZER_TYPE::GaussNewton ? MRSLAM_OPT_GN : MRSLAM_OPT_LM);
}
template <typename PointSource, typename PointTarget>
const Eigen::Matrix<double, 6, 6>& LsqRegistration<PointSource, PointTarget>::getFinalHessian() const {
  double H[36];
  if (mrslam_gicp_get_final_hessian(h_, H) == MRSLAM_OK)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) final_hessian_(i, j) = H[6 * i + j];
  return final_hessian_;
}
template <typename PointSource, typename PointTarget>
double LsqRegistration<PointSource, PointTarget>::evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H,
                                                               Eigen::Matrix<double, 6, 1>* b) {
  push_parameters();
  float T[16];
  detail::to_rows16(relative_pose, T);
  double H36[36], b6[6], cost = 0.0;
  detail::warn(mrslam_gicp_evaluate_cost(h_, T, H ? H36 : nullptr, b ? b6 : nullptr, &cost), "evaluateCost");
  if (H)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) (*H)(i, j) = H36[6 * i + j];
  if (b)
    for (int i = 0; i < 6; i++) (*b)(i) = b6[i];
  return cost;
}
// lsq_registration_impl.hpp:84-94: element-wise test on R - I and t
template <typename PointSource, typename PointTarget>
bool LsqRegistration<PointSource, PointTarget>::is_converged(const Eigen::Isometry3d& delta) const {
  double worst = 0.0;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) worst = std::fmax(worst, std::fabs(delta(i, j) - (i == j ? 1.0 : 0.0)) / rotation_epsilon_);
    worst = std::fmax(worst, std::fabs(delta(i, 3)) / transformation_epsilon_);
  }
  return worst < 1.0;
}
// pcl::Registration::align() is non-virtual and lands here; the outer Gauss-Newton / Levenberg-Marquardt loop
// (lsq_registration_impl.hpp:53-82) runs inside the library in lock step with the device kernels
template <typename PointSource, typename PointTarget>
void LsqRegistration<PointSource, PointTarget>::computeTransformation(PointCloudSource& output, const Matrix4& guess) {
  push_parameters();
  float g16[16], T16[16];
  detail::to_rows16(guess, g16);
  detail::warn(mrslam_gicp_align(h_, g16), "align");
  if (mrslam_gicp_get_final_transformation(h_, T16) == MRSLAM_OK)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) final_transformation_(i, j) = T16[4 * i + j];
  converged_ = mrslam_gicp_has_converged(h_) != 0;
  nr_iterations_ = mrslam_gicp_get_num_iterations(h_);
  if (input_ && !input_->points.empty()) {  // pcl::transformPointCloud(*input_, output, final_transformation_): xyz moved, other fields kept
    if (&output != input_.get()) output = *input_;
    mrslam_gicp_get_aligned(h_, reinterpret_cast<float*>(output.points.data()), detail::floats_per_point<PointSource>());
  }
}

// ---- FastGICP (fast_gicp_impl.hpp:15-100 for the defaults and the input bookkeeping) ------------------------------------
template <typename PointSource, typename PointTarget>
FastGICP<PointSource, PointTarget>::FastGICP()
    : num_threads_(0), k_correspondences_(20), regularization_method_(RegularizationMethod::PLANE), source_covs_fetched_(false),
      target_covs_fetched_(false) {
  reg_name_ = "FastGICP";
  corr_dist_threshold_ = std::numeric_limits<float>::max();
}
template <typename PointSource, typename PointTarget>
FastGICP<PointSource, PointTarget>::~FastGICP() {}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setNumThreads(int n) {
  num_threads_ = n;
  mrslam_gicp_set_num_threads(h_, n);
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setCorrespondenceRandomness(int k) {
  k_correspondences_ = k;
  detail::warn(mrslam_gicp_set_correspondence_randomness(h_, k), "setCorrespondenceRandomness");
  source_covs_fetched_ = target_covs_fetched_ = false;
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setRegularizationMethod(RegularizationMethod method) {
  regularization_method_ = method;
  detail::warn(mrslam_gicp_set_regularization_method(h_, static_cast<int>(method)), "setRegularizationMethod");
  source_covs_fetched_ = target_covs_fetched_ = false;
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::swapSourceAndTarget() {
  input_.swap(target_);
  source_covs_.swap(target_covs_);
  std::swap(source_covs_fetched_, target_covs_fetched_);
  detail::warn(mrslam_gicp_swap_source_and_target(h_), "swapSourceAndTarget");
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::clearSource() {
  input_.reset();
  source_covs_.clear();
  source_covs_fetched_ = false;
  mrslam_gicp_clear_source(h_);
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::clearTarget() {
  target_.reset();
  target_covs_.clear();
  target_covs_fetched_ = false;
  mrslam_gicp_clear_target(h_);
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setInputSource(const PointCloudSourceConstPtr& cloud) {
  if (input_ == cloud) return;  // same object again: structures and covariances are kept
  Base::setInputSource(cloud);
  source_covs_.clear();
  source_covs_fetched_ = false;
  detail::warn(mrslam_gicp_set_input_source(h_, cloud->points.empty() ? nullptr : reinterpret_cast<const float*>(cloud->points.data()),
                                            static_cast<int>(cloud->points.size()), detail::floats_per_point<PointSource>()),
               "setInputSource");
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setInputTarget(const PointCloudTargetConstPtr& cloud) {
  if (target_ == cloud) return;
  Base::setInputTarget(cloud);
  target_covs_.clear();
  target_covs_fetched_ = false;
  detail::warn(mrslam_gicp_set_input_target(h_, cloud->points.empty() ? nullptr : reinterpret_cast<const float*>(cloud->points.data()),
                                            static_cast<int>(cloud->points.size()), detail::floats_per_point<PointTarget>()),
               "setInputTarget");
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setSourceCovariances(const CovarianceList& covs) {
  std::vector<double> rows;
  detail::covs_to_rows(covs, rows);
  detail::warn(mrslam_gicp_set_source_covariances(h_, rows.data(), static_cast<int>(covs.size())), "setSourceCovariances");
  source_covs_ = covs;
  source_covs_fetched_ = true;
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::setTargetCovariances(const CovarianceList& covs) {
  std::vector<double> rows;
  detail::covs_to_rows(covs, rows);
  detail::warn(mrslam_gicp_set_target_covariances(h_, rows.data(), static_cast<int>(covs.size())), "setTargetCovariances");
  target_covs_ = covs;
  target_covs_fetched_ = true;
}
template <typename PointSource, typename PointTarget>
const CovarianceList& FastGICP<PointSource, PointTarget>::getSourceCovariances() const {
  if (!source_covs_fetched_ && input_ && !input_->points.empty()) {
    std::vector<double> rows(input_->points.size() * 16);
    if (mrslam_gicp_get_source_covariances(h_, rows.data(), static_cast<int>(input_->points.size())) == MRSLAM_OK) {
      detail::rows_to_covs(rows, source_covs_);
      source_covs_fetched_ = true;
    }
  }
  return source_covs_;
}
template <typename PointSource, typename PointTarget>
const CovarianceList& FastGICP<PointSource, PointTarget>::getTargetCovariances() const {
  if (!target_covs_fetched_ && target_ && !target_->points.empty()) {
    std::vector<double> rows(target_->points.size() * 16);
    if (mrslam_gicp_get_target_covariances(h_, rows.data(), static_cast<int>(target_->points.size())) == MRSLAM_OK) {
      detail::rows_to_covs(rows, target_covs_);
      target_covs_fetched_ = true;
    }
  }
  return target_covs_;
}
template <typename PointSource, typename PointTarget>
double FastGICP<PointSource, PointTarget>::getFitnessScore(double max_range) {
  return mrslam_gicp_get_fitness_score(h_, max_range);
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::push_parameters() {
  Lsq::push_parameters();
  mrslam_gicp_set_max_correspondence_distance(h_, corr_dist_threshold_);
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::computeTransformation(PointCloudSource& output, const Matrix4& guess) {
  Lsq::computeTransformation(output, guess);  // covariances are (re)computed by the library when they are missing
  source_covs_fetched_ = source_covs_fetched_ && !source_covs_.empty();
  target_covs_fetched_ = target_covs_fetched_ && !target_covs_.empty();
}
template <typename PointSource, typename PointTarget>
void FastGICP<PointSource, PointTarget>::update_correspondences(const Eigen::Isometry3d&) {}  // part of the device linearisation pass
template <typename PointSource, typename PointTarget>
double FastGICP<PointSource, PointTarget>::linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) {
  Eigen::Matrix4f T;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T(i, j) = static_cast<float>(trans(i, j));
  return this->evaluateCost(T, H, b);
}
template <typename PointSource, typename PointTarget>
double FastGICP<PointSource, PointTarget>::compute_error(const Eigen::Isometry3d& trans) {
  double T[16], c = 0.0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T[4 * i + j] = trans(i, j);
  detail::warn(mrslam_gicp_compute_error(h_, T, &c), "compute_error");
  return c;
}
}  // namespace fast_gicp
