// wr_common.cuh -- shared device-side definitions for the WaveRange B200 codec.
//
// All floating-point arithmetic on the numeric path is compiled with -fmad=false so that
// every + and * is individually rounded, exactly like the ISO evaluation of the reference
// sources.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace wrb {

// ---- stream-format constants (reference src/core/defs.h:34-50) ----------------------------
constexpr uint32_t kBlock   = 60000;   // BLOCKSIZE: symbols per coder block
constexpr int      kNLayMax = 8;       // NLAYMAX
constexpr int      kWavLvl  = 4;       // WAV_LVL
#define WRB_ACC_COEF 1.75              /* WAV_ACC_COEF */

// ---- lifting constants (reference src/waveletcdf97_3d/waveletcdf97_3d.c:41-58) -------------
// Hex literals are the exact doubles the reference's decimal constants parse to; E0..E2 are the
// values its run-time expressions for ext[] evaluate to under ISO arithmetic.
#define WRB_LA   (-0x1.960ce676401a2p+0)   /* lfc[0] = -1.5861343420693648 */
#define WRB_LB   (-0x1.b2035c9357a96p-5)   /* lfc[1] = -0.0529801185718856 */
#define WRB_LC   ( 0x1.c40ceba5738p-1)     /* lfc[2] =  0.8829110755411875 */
#define WRB_LD   ( 0x1.c626a904721eep-2)   /* lfc[3] =  0.4435068520511142 */
#define WRB_SCL  ( 0x1.264c795071464p+0)   /* scl    =  1.1496043988602418 */
#define WRB_PSCL ( 0x1.bd5edf975ce17p-1)   /* 1.0/scl */
#define WRB_E0   (-0x1.4f43b88aa31b3p-3)
#define WRB_E1   ( 0x1.a6be82e3706b1p-4)
#define WRB_E2   ( 0x1.0f7c8ee31b63bp+0)

// ---- device-resident codec state -----------------------------------------------------------
// min/max reductions go through order-preserving uint64 keys so that atomicMin/atomicMax on
// integers implement fmin/fmax on doubles (no NaNs assumed, as in the reference).
struct DevState {
    unsigned long long fmin_key, fmax_key;                     // raw field extrema (wrappers.cpp:244-250)
    unsigned long long rmin_key[kNLayMax + 1], rmax_key[kNLayMax + 1];  // residual extrema before layer l
    double tolabs, midval, halfspan;
    double deps[kNLayMax], minval[kNLayMax], aopt[kNLayMax], bopt[kNLayMax];
    double span[kNLayMax];     // max - min of the residual before layer l (the local-cutoff test, wrappers.cpp:365)
    int    active[kNLayMax];   // layer l is part of the stream
    int    nlay;
    int    done;               // brflag seen (wrappers.cpp:326-333)
    int    trivial;            // halfspan <= 2*DBL_MIN (wrappers.cpp:257)
    int    error;
    unsigned long long ntot_enc;
    unsigned long long len_enc[kNLayMax];
    unsigned long long lay_off[kNLayMax + 1];   // byte offset of every layer inside the blob
    int    nseek_keep;         // seek points per chunk that went into the container (<= those recorded)
};

__host__ __device__ inline unsigned long long dkey(double x)
{
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b; memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double dunkey(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x; memcpy(&x, &b, 8); return x;
#endif
}
// min/max by plain comparison: fmin()/fmax() on doubles expand to ~10 instructions each (NaN
// propagation); the reference assumes NaN-free data and so do these (result for a < b is a)
__device__ __forceinline__ double dmin2(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax2(double a, double b) { return (b > a) ? b : a; }

// DRAM -> L2 prefetch: the bytes a streaming kernel needs in flight to cover HBM latency (~50-100 KB per
// SM) are parked in L2 instead of registers; the register prefetch then only has to cover L2 latency
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

constexpr unsigned long long kKeyMinInit = ~0ull;   // identity for min
constexpr unsigned long long kKeyMaxInit = 0ull;    // identity for max

// block-wide min/max of keys, then one atomic pair per CTA
__device__ inline void block_minmax_commit(unsigned long long kmin, unsigned long long kmax,
                                           unsigned long long* gmin, unsigned long long* gmax)
{
    __shared__ unsigned long long s_min[32], s_max[32];
    for (int o = 16; o; o >>= 1) {
        unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o);
        unsigned long long b = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    int nthr = blockDim.x * blockDim.y * blockDim.z;
    int w = tid >> 5, l = tid & 31, nw = (nthr + 31) >> 5;
    if (l == 0) { s_min[w] = kmin; s_max[w] = kmax; }
    __syncthreads();
    if (w == 0) {
        kmin = l < nw ? s_min[l] : kKeyMinInit;
        kmax = l < nw ? s_max[l] : kKeyMaxInit;
        for (int o = 16; o; o >>= 1) {
            unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o);
            unsigned long long b = __shfl_xor_sync(0xffffffffu, kmax, o);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if (l == 0) {
            if (kmin != kKeyMinInit) atomicMin(gmin, kmin);
            if (kmax != kKeyMaxInit) atomicMax(gmax, kmax);
        }
    }
    __syncthreads();
}

struct Strides { long long y, z; };   // x stride is always 1

// z-segment length (in output pairs) of the one-pass-per-level kernels.  A CTA restarts its z pipeline `lead`
// pairs before its segment, and CTAs run in waves of `slots` (resident CTAs on the machine): pick the segment
// count that minimises  waves * (segment + lead),  i.e. the length of the critical path in plane pairs.
inline int pick_zpairs(long long tiles, int m2, int slots, int lead, int min_zp)
{
    int best = m2;
    long long best_cost = -1;
    for (int nseg = 1; nseg <= m2; nseg++) {
        const int zp = (m2 + nseg - 1) / nseg;
        if (zp < min_zp && nseg > 1) break;
        const long long ctas = tiles * ((m2 + zp - 1) / zp);
        const long long waves = (ctas + slots - 1) / slots;
        const long long cost = waves * (zp + lead);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = zp; }
    }
    return best;
}

}  // namespace wrb
