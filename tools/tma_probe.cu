// tma_probe.cu -- minimal cp.async.bulk.tensor probe (development aid): loads one box of a 3-D tensor into shared memory
// and prints it.  usage: tma_probe dtype(0=f64,1=f32) boxx boxy cx cy cz variant
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <class T>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int cx, int cy, int cz, int nelem, T* out, int variant)
{
    extern __shared__ __align__(128) unsigned char raw[];
    T* tile = reinterpret_cast<T*>(raw);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1));
        if (variant & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(nelem * (int)sizeof(T)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(tile)), "l"(&tm), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(bar)) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) out[i] = tile[i];
}

template <class T>
int run(CUtensorMapDataType dt, int bx, int by, int cx, int cy, int cz, int variant)
{
    const int n0 = 64, n1 = 48, n2 = 6;
    T* d; T* out;
    cudaMalloc(&d, sizeof(T) * n0 * n1 * n2);
    cudaMalloc(&out, 65536);
    T* h = (T*)malloc(sizeof(T) * n0 * n1 * n2);
    for (int i = 0; i < n0 * n1 * n2; i++) h[i] = (T)i;
    cudaMemcpy(d, h, sizeof(T) * n0 * n1 * n2, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 2; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2};
    const cuuint64_t strides[2] = {(cuuint64_t)n0 * sizeof(T), (cuuint64_t)n0 * n1 * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tm, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    (variant & 2) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 3;
    cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    probe<T><<<1, 128, 40000>>>(tm, cx, cy, cz, bx * by, out, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 4;
    T* ho = (T*)malloc(sizeof(T) * bx * by);
    cudaMemcpy(ho, out, sizeof(T) * bx * by, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r_ = 0; r_ < by; r_++) for (int c = 0; c < bx; c++) {
        const int gx = cx + c, gy = cy + r_;
        const double want = (gx >= 0 && gx < n0 && gy >= 0 && gy < n1 && cz >= 0 && cz < n2) ? (double)(gx + n0 * (gy + n1 * cz)) : 0.0;
        if ((double)ho[r_ * bx + c] != want) bad++;
    }
    printf("mismatches %d of %d; first row: %g %g %g %g\n", bad, bx * by, (double)ho[0], (double)ho[1], (double)ho[2], (double)ho[3]);
    return bad ? 5 : 0;
}

int main(int argc, char** argv)
{
    const int dt = atoi(argv[1]), bx = atoi(argv[2]), by = atoi(argv[3]), cx = atoi(argv[4]), cy = atoi(argv[5]), cz = atoi(argv[6]);
    const int variant = argc > 7 ? atoi(argv[7]) : 1;
    printf("dtype %s box %dx%d at (%d,%d,%d) variant %d\n", dt ? "f32" : "f64", bx, by, cx, cy, cz, variant);
    return dt ? run<float>(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, bx, by, cx, cy, cz, variant)
              : run<double>(CU_TENSOR_MAP_DATA_TYPE_FLOAT64, bx, by, cx, cy, cz, variant);
}
