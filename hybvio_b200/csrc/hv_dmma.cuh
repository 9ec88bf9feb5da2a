// hybvio_b200/csrc/hv_dmma.cuh -- fp64 tensor-core tile product D(8x8) += A(8x4) B(4x8) (mma.sync.m8n8k4.f64) for the small
// dense fp64 products of the EKF kernels. Measured on B200 (tools/probe2.cu): 64 FMA/clk/SM with 16 warps and a 26-cycle
// dependent latency, against ~46 FMA/clk/SM for scalar DFMA code that is additionally bound by shared-memory loads
// (two 8-byte loads per FMA without register tiling); one DMMA needs ONE operand load per lane for 8 FMAs per lane.
//
// Fragment layout (PTX ISA, m8n8k4 .f64; cute::SM80_8x4 / SM80_8x8_Row): with g = lane >> 2, t = lane & 3
//   a = A[g][t]        b = B[t][g]        c0 = C[g][2t], c1 = C[g][2t + 1]
#pragma once
#ifdef HV_EMU
inline void hv_dmma(double& c0, double& c1, double a, double b) { emu_dmma(c0, c1, a, b); }
#else
__device__ __forceinline__ void hv_dmma(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif
