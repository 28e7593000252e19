// p2p_probe.cu -- what SM-issued copies reach over NVLink between two GPUs of a box (development aid).
// One process, devices 0 and 1 with peer access: a grid-stride copy kernel of 16 bytes per thread runs on device 0 and
//   local : reads device 0, writes device 0
//   pull  : reads device 1 (remote loads),  writes device 0      -- what slab_order.cu's gather / scatter do
//   push  : reads device 0, writes device 1 (posted remote stores)
// for a few grid sizes, and both devices at once in opposite directions ("both"), as in the symbol exchange where every
// rank pulls at the same time.     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/p2p_probe tools/p2p_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void __launch_bounds__(256) copy16(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// two loads in flight per thread
__global__ void __launch_bounds__(256) copy16x2(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
        const uint4 a = src[i];
        uint4 b = make_uint4(0, 0, 0, 0);
        if (i + stride < n) b = src[i + stride];
        dst[i] = a;
        if (i + stride < n) dst[i + stride] = b;
    }
}

int main(int argc, char** argv)
{
    const size_t bytes = (argc > 1 ? (size_t)atol(argv[1]) : 400) << 20;     // MB per copy (default: one slab's three layers)
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs two GPUs\n"); return 1; }
    void *a0, *b0, *a1, *b1;
    cudaStream_t s0, s1;
    cudaEvent_t e0, e1, f0, f1;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&a0, bytes)); CK(cudaMalloc(&b0, bytes));
    CK(cudaMemset(a0, 1, bytes)); CK(cudaStreamCreate(&s0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&a1, bytes)); CK(cudaMalloc(&b1, bytes));
    CK(cudaMemset(a1, 2, bytes)); CK(cudaStreamCreate(&s1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    CK(cudaDeviceSynchronize());
    const size_t n = bytes / 16;
    struct Case { const char* name; const void* src0; void* dst0; const void* src1; void* dst1; };
    const Case cases[] = {
        {"local            ", a0, b0, nullptr, nullptr},
        {"pull  (one GPU)  ", a1, b0, nullptr, nullptr},
        {"push  (one GPU)  ", a0, b1, nullptr, nullptr},
        {"pull  (both GPUs)", a1, b0, a0, b1},
        {"push  (both GPUs)", a0, b1, a1, b0},
    };
    const int grids[] = {148 * 2, 148 * 8, 148 * 32};
    for (int variant = 0; variant < 2; variant++)
        for (const Case& c : cases)
            for (int g : grids) {
                float best = 1e30f;
                for (int rep = 0; rep < 4; rep++) {
                    CK(cudaSetDevice(0)); CK(cudaEventRecord(e0, s0));
                    if (variant == 0) copy16<<<g, 256, 0, s0>>>((const uint4*)c.src0, (uint4*)c.dst0, n);
                    else copy16x2<<<g, 256, 0, s0>>>((const uint4*)c.src0, (uint4*)c.dst0, n);
                    CK(cudaEventRecord(e1, s0));
                    if (c.src1) {
                        CK(cudaSetDevice(1)); CK(cudaEventRecord(f0, s1));
                        if (variant == 0) copy16<<<g, 256, 0, s1>>>((const uint4*)c.src1, (uint4*)c.dst1, n);
                        else copy16x2<<<g, 256, 0, s1>>>((const uint4*)c.src1, (uint4*)c.dst1, n);
                        CK(cudaEventRecord(f1, s1));
                        CK(cudaStreamSynchronize(s1));
                    }
                    CK(cudaSetDevice(0)); CK(cudaStreamSynchronize(s0));
                    float ms = 0, ms1 = 0;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (c.src1) { CK(cudaSetDevice(1)); CK(cudaEventElapsedTime(&ms1, f0, f1)); if (ms1 > ms) ms = ms1; }
                    if (rep > 0 && ms < best) best = ms;
                }
                printf("%s %s grid %5d x 256: %7.3f ms  %7.1f GB/s per GPU\n", variant ? "2 loads in flight" : "1 load in flight ", c.name, g, best,
                       bytes / (best * 1e-3) / 1e9);
            }
    return 0;
}
