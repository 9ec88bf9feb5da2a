/* OpenCL disabled: only the includes the cmake-generated header would pull in */
#include "opencv2/core/ocl.hpp"
#include "opencv2/core/ocl_genbase.hpp"
#include "opencv2/core/opencl/ocl_defs.hpp"
