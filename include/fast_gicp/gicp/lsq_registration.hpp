// <fast_gicp/gicp/lsq_registration.hpp> of the B200 drop-in (see ../../fast_gicp_b200/pcl_dropin.hpp)
#pragma once
#include "../../fast_gicp_b200/pcl_dropin.hpp"
