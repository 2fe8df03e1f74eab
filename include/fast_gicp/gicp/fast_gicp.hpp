// <fast_gicp/gicp/fast_gicp.hpp> of the B200 drop-in: the include path the reference's callers use
// (Mapping/src/global_manager/include/global_manager/global_manager.h, FG/src/python/main.cpp).
#pragma once
#include "../../fast_gicp_b200/pcl_dropin.hpp"
