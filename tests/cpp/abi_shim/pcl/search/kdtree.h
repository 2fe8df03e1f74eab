#pragma once
namespace pcl { namespace search { template <typename PointT> class KdTree {}; } }
