// TEST INFRASTRUCTURE ONLY -- the part of pcl::Registration<PointSource, PointTarget, Scalar> (PCL 1.10
// registration/registration.h + impl/registration.hpp) that a derived registration method and its callers use: the input
// setters, the parameters, align() -> computeTransformation(), the result getters.  getFitnessScore is deliberately NOT
// restated (PCL runs its own CPU kd-tree there; the drop-in class offers a GPU getFitnessScore of its own).
#pragma once
#include <Eigen/Core>
#include <cfloat>
#include <string>

#include "../point_cloud.h"
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = shared_ptr<Registration<PointSource, PointTarget, Scalar>>;

  Registration()
      : nr_iterations_(0), max_iterations_(10), final_transformation_(Matrix4::Identity()), transformation_(Matrix4::Identity()),
        previous_transformation_(Matrix4::Identity()), transformation_epsilon_(0.0), transformation_rotation_epsilon_(0.0),
        euclidean_fitness_epsilon_(-DBL_MAX), corr_dist_threshold_(DBL_MAX), converged_(false) {}
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  PointCloudSourceConstPtr const getInputSource() { return input_; }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  int getMaximumIterations() { return max_iterations_; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  double getTransformationEpsilon() { return transformation_epsilon_; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  double getMaxCorrespondenceDistance() { return corr_dist_threshold_; }
  void setEuclideanFitnessEpsilon(double e) { euclidean_fitness_epsilon_ = e; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {  // impl/registration.hpp: reset the state, run the method
    nr_iterations_ = 0;
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    output.points.resize(input_ ? input_->points.size() : 0);
    computeTransformation(output, guess);
  }

protected:
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_, max_iterations_;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_, transformation_rotation_epsilon_, euclidean_fitness_epsilon_, corr_dist_threshold_;
  bool converged_;
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
};
}  // namespace pcl
