// stub of <cuda_runtime.h> for the host emulation build (tests/emu)
#pragma once
#include <cstddef>
typedef void* cudaStream_t;
typedef int cudaError_t;
