// stand-in for gtsam_points/config.hpp: no TBB, OpenMP as the default parallelism (the reference's #pragma omp branch)
#pragma once
