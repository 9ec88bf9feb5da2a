"""hybvio_b200: B200 (sm_100a) implementation of HybVIO's per-frame hot path -- optical-flow pyramid, pyramidal
Lucas-Kanade tracker and the EKF covariance propagate/update -- behind a C ABI (include/hybvio_b200.h).

  csrc/   CUDA kernels + C ABI  -> libhybvio_b200.so
  host/   C++ adapters implementing the reference's tracker::ImagePyramid / OpticalFlow / odometry::EKF interfaces
  capi.py ctypes binding used by tests/ and bench.py
  synth.py deterministic EuRoC-shaped synthetic streams
"""
