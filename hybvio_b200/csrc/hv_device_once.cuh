// hybvio_b200/csrc/hv_device_once.cuh -- included by the files that launch kernels (not by kernel bodies: the host emulator does not see it)
#pragma once
#include <cuda_runtime.h>

// Function attributes (dynamic shared memory limit, non-portable cluster size) belong to the DEVICE's context, not to the process: a
// process that drives several GPUs (or whose adapters sit on another GPU than its session) must set them once per device. Returns true
// the first time it is called for the current device with this flag array (64 entries, zero-initialised).
static inline bool hv_first_use_on_device(bool* seen)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (seen[dev]) return false;
    seen[dev] = true;
    return true;
}
