// tools/tma_probe.cu -- which way of handing a u8 2-D tensor map to cp.async.bulk.tensor.2d works on this B200 / driver?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe MODE X Y      MODE 0: __grid_constant__ CUtensorMap parameter; 1: inside a 4.6 KB __grid_constant__ struct;
//                                 2: descriptor in global memory; 3: inside a 1 KB __grid_constant__ struct
// One mode per process (an illegal instruction kills the context).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

struct alignas(64) Tmap { unsigned long long q[16]; };
struct Big { const void* table; int n; unsigned short idx[32]; const unsigned char* src[32]; int pitch[32]; unsigned char use[32]; Tmap tmap[32]; };
struct Small { const void* table; int n; unsigned short idx[4]; const unsigned char* src[4]; int pitch[4]; unsigned char use[4]; Tmap tmap[4]; };

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ void body(const void* map, int x, int y, int boxW, int boxH, unsigned char* out)
{
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"((unsigned)(boxW * boxH)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(s32(buf)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    }
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[(size_t)blockIdx.x * boxW * boxH + i] = buf[i];
}
__global__ void k_param(const __grid_constant__ Tmap m, int x, int y, int bw, int bh, unsigned char* out) { body(&m, x, y, bw, bh, out); }
__global__ void k_big(const __grid_constant__ Big b, int x, int y, int bw, int bh, unsigned char* out) { body(&b.tmap[blockIdx.x & 1], x, y, bw, bh, out); }
__global__ void k_small(const __grid_constant__ Small b, int x, int y, int bw, int bh, unsigned char* out) { body(&b.tmap[blockIdx.x & 1], x, y, bw, bh, out); }
__global__ void k_global(const Tmap* m, int x, int y, int bw, int bh, unsigned char* out) { body(m, x, y, bw, bh, out); }

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0, x = argc > 2 ? atoi(argv[2]) : 0, y = argc > 3 ? atoi(argv[3]) : 0;
    const int w = 752, h = 480, pitch = 752, bw = argc > 4 ? atoi(argv[4]) : 112, bh = argc > 5 ? atoi(argv[5]) : 108;
    std::vector<unsigned char> img((size_t)pitch * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *out;
    cudaMalloc(&d, img.size()); cudaMalloc(&out, 2 * bw * bh);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 2; }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, estr[2] = {1, 1};
    CUresult r = ((EncodeFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("mode %d: encode failed %d\n", mode, (int)r); return 3; }
    Tmap t; memcpy(&t, &m, sizeof(t));
    const size_t smem = (size_t)bw * bh + 128;
    cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (mode == 0) k_param<<<2, 256, smem>>>(t, x, y, bw, bh, out);
    else if (mode == 1) { Big b; memset(&b, 0, sizeof(b)); b.tmap[0] = t; b.tmap[1] = t; k_big<<<2, 256, smem>>>(b, x, y, bw, bh, out); }
    else if (mode == 3) { Small b; memset(&b, 0, sizeof(b)); b.tmap[0] = t; b.tmap[1] = t; k_small<<<2, 256, smem>>>(b, x, y, bw, bh, out); }
    else { Tmap* dm; cudaMalloc(&dm, sizeof(t)); cudaMemcpy(dm, &t, sizeof(t), cudaMemcpyHostToDevice); k_global<<<2, 256, smem>>>(dm, x, y, bw, bh, out); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d x %d y %d box %dx%d: FAILED: %s\n", mode, x, y, bw, bh, cudaGetErrorString(e)); return 1; }
    std::vector<unsigned char> res((size_t)bw * bh);
    cudaMemcpy(res.data(), out, res.size(), cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) {
        const int gx = x + i, gy = y + j;
        const unsigned char want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[(size_t)gy * pitch + gx] : 0;
        if (res[(size_t)j * bw + i] != want) bad++;
    }
    printf("mode %d x %d y %d box %dx%d: ok, %ld bytes differ (sizeof Big %zu, Small %zu)\n", mode, x, y, bw, bh, bad, sizeof(Big), sizeof(Small));
    return bad ? 4 : 0;
}
