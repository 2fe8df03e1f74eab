// fast_gicp_b200/pcl_dropin.hpp -- the classes a ROS node sees when it includes <fast_gicp/gicp/fast_gicp.hpp>:
//   fast_gicp::LsqRegistration<PointSource, PointTarget> : pcl::Registration<PointSource, PointTarget, float>
//   fast_gicp::FastGICP<PointSource, PointTarget>        : LsqRegistration<PointSource, PointTarget>
// with the reference's public and protected-virtual surface (FG/include/fast_gicp/gicp/lsq_registration.hpp:15-85,
// FG/include/fast_gicp/gicp/fast_gicp.hpp:19-100) -- but no algorithm: every member forwards to the C-ABI of
// libmrslam_b200.so (include/mrslam_b200.h), where the kNN / covariance / correspondence / linearisation kernels and the
// Levenberg-Marquardt driver live.  Members are defined out of line in pcl_dropin_impl.hpp and explicitly instantiated for
// PointXYZ / PointXYZI / PointNormal in mr_slam_b200/csrc/adapters/fast_gicp_inst.cpp -> libfast_gicp.so
// (FG/src/fast_gicp/gicp/fast_gicp.cpp:4-6, lsq_registration.cpp:4-6), the library name the catkin packages link.
//
// What differs from the reference, by design:
//  * the whole optimisation runs inside mrslam_gicp_align(); step_optimize / step_gn / step_lm do not exist as separate host
//    functions, and a subclass overriding linearize() / compute_error() does not change what align() does (the two virtuals
//    are still callable and return the device results);
//  * getFitnessScore(max_range) is offered by FastGICP itself (one more NN pass on the GPU); through a
//    pcl::Registration::Ptr the caller reaches PCL's own non-virtual CPU implementation instead (global_manager.cpp:2058);
//  * k_correspondences <= 32.
// Eigen matrices are column-major, the C-ABI is row-major: the members go through element access, never through data().
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>

#include <memory>
#include <vector>

#include "../mrslam_b200.h"

namespace fast_gicp {

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };  // gicp_settings.hpp:6
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };                         // lsq_registration.hpp:13
using CovarianceList = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;

template <typename PointSource, typename PointTarget>
class LsqRegistration : public pcl::Registration<PointSource, PointTarget, float> {
public:
  using Scalar = float;
  using Base = pcl::Registration<PointSource, PointTarget, Scalar>;
  using Matrix4 = typename Base::Matrix4;
  using PointCloudSource = typename Base::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Base::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = pcl::shared_ptr<LsqRegistration<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const LsqRegistration<PointSource, PointTarget>>;
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  LsqRegistration();
  virtual ~LsqRegistration();
  LsqRegistration(const LsqRegistration&) = delete;
  LsqRegistration& operator=(const LsqRegistration&) = delete;

  void setRotationEpsilon(double eps);
  void setInitialLambdaFactor(double init_lambda_factor);
  void setDebugPrint(bool lm_debug_print);
  void setLSQType(LSQ_OPTIMIZER_TYPE type);  // upstream fast_gicp's setter for lsq_optimizer_type_ (a protected member in this reference)
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const;
  double evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr);

  virtual void swapSourceAndTarget() {}
  virtual void clearSource() {}
  virtual void clearTarget() {}

  mrslam_gicp_t* handle() const { return h_; }  // the device-side registration object

protected:
  using Base::converged_;
  using Base::final_transformation_;
  using Base::input_;
  using Base::max_iterations_;
  using Base::nr_iterations_;
  using Base::transformation_epsilon_;

  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) override;
  bool is_converged(const Eigen::Isometry3d& delta) const;
  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) = 0;
  virtual double compute_error(const Eigen::Isometry3d& trans) = 0;
  virtual void push_parameters();  // PCL-side members (max_iterations_, epsilons, correspondence distance) -> device object

  double rotation_epsilon_;
  LSQ_OPTIMIZER_TYPE lsq_optimizer_type_;
  int lm_max_iterations_;
  double lm_init_lambda_factor_;
  double lm_lambda_;
  bool lm_debug_print_;
  mutable Eigen::Matrix<double, 6, 6> final_hessian_;
  mrslam_gicp_t* h_;
};

template <typename PointSource, typename PointTarget>
class FastGICP : public LsqRegistration<PointSource, PointTarget> {
public:
  using Scalar = float;
  using Lsq = LsqRegistration<PointSource, PointTarget>;
  using Matrix4 = typename Lsq::Matrix4;
  using PointCloudSource = typename Lsq::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Lsq::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = pcl::shared_ptr<FastGICP<PointSource, PointTarget>>;
  using ConstPtr = pcl::shared_ptr<const FastGICP<PointSource, PointTarget>>;

  FastGICP();
  virtual ~FastGICP() override;

  void setNumThreads(int n);                 // accepted and ignored: the work runs on the GPU
  void setCorrespondenceRandomness(int k);   // 1..32
  void setRegularizationMethod(RegularizationMethod method);

  virtual void swapSourceAndTarget() override;
  virtual void clearSource() override;
  virtual void clearTarget() override;
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) override;
  virtual void setSourceCovariances(const CovarianceList& covs);
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) override;
  virtual void setTargetCovariances(const CovarianceList& covs);
  const CovarianceList& getSourceCovariances() const;  // computed on the device on first use, then cached here
  const CovarianceList& getTargetCovariances() const;
  double getFitnessScore(double max_range = 1.7976931348623157e308);  // hides PCL's CPU version for callers holding a FastGICP

protected:
  using Lsq::h_;
  using typename Lsq::Base;
  using Base::corr_dist_threshold_;
  using Base::input_;
  using Base::reg_name_;
  using Base::target_;

  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) override;
  virtual void update_correspondences(const Eigen::Isometry3d& trans);
  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) override;
  virtual double compute_error(const Eigen::Isometry3d& trans) override;
  virtual void push_parameters() override;

  int num_threads_;
  int k_correspondences_;
  RegularizationMethod regularization_method_;
  mutable CovarianceList source_covs_, target_covs_;
  mutable bool source_covs_fetched_, target_covs_fetched_;
};
}  // namespace fast_gicp
