#include <gtsam_points/types/point_cloud.hpp>
