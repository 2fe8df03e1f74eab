// <fast_gicp/gicp/gicp_settings.hpp> of the B200 drop-in: RegularizationMethod lives in pcl_dropin.hpp
#pragma once
#include "../../fast_gicp_b200/pcl_dropin.hpp"
