// hybvio_b200/host/cuda_context.hpp -- the ONE hv_ctx (one CUDA stream) that all adapters of a process share.
//
// The reference drives tracker and EKF of a Session from a single thread (src/api/api.cpp:425-428), so one context per
// process is enough; sharing it between cuda_tracker_backends.cpp and cuda_ekf.cpp keeps LK and the EKF on one stream
// (round 1 had one private context per adapter file: two streams, two pinned result areas). HV_DEVICE selects the GPU.
#ifndef HYBVIO_B200_HOST_CUDA_CONTEXT_HPP_
#define HYBVIO_B200_HOST_CUDA_CONTEXT_HPP_
#include "../../include/hybvio_b200.h"
#include <cstdio>
#include <cstdlib>

namespace hybvio_b200 {
[[noreturn]] inline void hvFail(const char* what) {
    std::fprintf(stderr, "hybvio_b200: %s failed: %s\n", what, hv_last_error());
    std::abort();
}
// inline + function-local static: one instance per program, whichever translation unit asks first
inline hv_ctx* sharedContext() {
    static hv_ctx* ctx = [] {
        hv_ctx* c = nullptr;
        const char* dev = std::getenv("HV_DEVICE");
        if (hv_ctx_create(dev ? std::atoi(dev) : 0, &c) != HV_OK) hvFail("hv_ctx_create");   // no CPU fallback
        return c;
    }();
    return ctx;
}
} // namespace hybvio_b200
#endif
