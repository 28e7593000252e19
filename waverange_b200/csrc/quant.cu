// quant.cu -- byte-plane quantiser of the wavelet coefficients.
//
// Replaces the layer loop of encoding_wrap() (reference src/core/wrappers.cpp:305-441) and the
// accumulate loop of decoding_wrap() (:480-515).
//
// The reference keeps the residual in place and sweeps it ~5 times per layer.  Here the
// coefficient array is never rewritten: the pass for layer l re-derives the residual from the
// coefficient by replaying the l earlier (deps, minval) pairs in registers with bit-identical
// operations, emits the 1-byte symbol, reduces min/max of the new residual (the only global
// dependency between layers, wrappers.cpp:308-314) and accumulates the 256-bin histogram of the
// coder block the symbol belongs to.  Layer parameters live in device memory (DevState), so the
// host never waits between layers; layers after the terminating one exit immediately.
#include <cfloat>
#include <type_traits>
#include "wr_common.cuh"
#include "wr_kernels.h"

namespace wrb {

ChunkGeom make_geom(unsigned long long ntot, unsigned long long chunk_len, unsigned int nseek_req)
{
    ChunkGeom g{};
    g.ntot = ntot;
    g.chunk_len = (chunk_len == 0 || chunk_len > ntot) ? ntot : chunk_len;
    g.pitch = (g.chunk_len + 15ull) & ~15ull;
    g.nchunks = (unsigned int)((ntot + g.chunk_len - 1) / g.chunk_len);
    g.blocks_per_chunk = (unsigned int)(g.chunk_len / kBlock + 1);
    unsigned long long last = ntot - (unsigned long long)(g.nchunks - 1) * g.chunk_len;
    g.nblocks = (g.nchunks - 1) * g.blocks_per_chunk + (unsigned int)(last / kBlock + 1);
    g.nseek = 0;
    g.sub_len = 0;
    if (nseek_req >= 7) nseek_req = 7; else if (nseek_req >= 3) nseek_req = 3; else if (nseek_req >= 1) nseek_req = 1;   // 2, 4 or 8 lanes per chunk
    // Seek grids are nested: the finest one (7 points) has spacing base = ceil16(ceil(chunk_len / 8)), the grids of 3 and
    // of 1 point are every 2nd / 4th point of it.  An encoder can so record the finest grid and decide afterwards how
    // many points the container keeps.  The largest grid <= the request whose last sub-range is not empty is granted.
    if (nseek_req > 0 && g.blocks_per_chunk == 1 && g.chunk_len >= 512) {
        const unsigned long long base = (((g.chunk_len + 7ull) / 8ull) + 15ull) & ~15ull;
        for (unsigned int n = nseek_req; n >= 1; n = (n - 1) / 2) {
            const unsigned long long sub = base * (8u / (n + 1));
            if ((unsigned long long)n * sub < g.chunk_len) { g.nseek = n; g.sub_len = (unsigned int)sub; break; }
        }
    }
    return g;
}

__global__ void state_init_kernel(DevState* st)
{
    st->fmin_key = kKeyMinInit; st->fmax_key = kKeyMaxInit;
    for (int l = 0; l <= kNLayMax; l++) { st->rmin_key[l] = kKeyMinInit; st->rmax_key[l] = kKeyMaxInit; }
    for (int l = 0; l < kNLayMax; l++) {
        st->active[l] = 0; st->deps[l] = 0; st->minval[l] = 0; st->aopt[l] = 0; st->bopt[l] = 0; st->len_enc[l] = 0;
        st->span[l] = 0;
    }
    st->tolabs = 0; st->midval = 0; st->halfspan = 0;
    st->nlay = 0; st->done = 0; st->trivial = 0; st->error = 0; st->ntot_enc = 0; st->nseek_keep = 0;
}
void state_init(DevState* st, cudaStream_t s) { state_init_kernel<<<1, 1, 0, s>>>(st); note_launch(1); }

// reference wrappers.cpp:253-266 (mid / half-span, trivial exit) and :292-299 (tolerance)
__global__ void state_prepare_kernel(DevState* st, double tolrel)
{
    double mn = dunkey(st->fmin_key), mx = dunkey(st->fmax_key);
    double half = (mx - mn) / 2;
    st->halfspan = half;
    st->midval = mn + half;
    if (half <= 2 * DBL_MIN) { st->trivial = 1; st->tolabs = 0; return; }
    double tolabs = tolrel * fmax(fabs(mn), fabs(mx));
    tolabs /= WRB_ACC_COEF;
    st->tolabs = tolabs;
}
void state_prepare(DevState* st, double tolrel, cudaStream_t s) { state_prepare_kernel<<<1, 1, 0, s>>>(st, tolrel); note_launch(1); }

// reference wrappers.cpp:316-340
__global__ void layer_params_kernel(DevState* st, int l)
{
    if (st->trivial || st->done) { st->active[l] = 0; return; }
    double mn = dunkey(st->rmin_key[l]), mx = dunkey(st->rmax_key[l]);
    st->minval[l] = mn;
    st->span[l] = mx - mn;
    double deps = (mx - mn) / 255.0;
    int last = 0;
    if (deps < st->tolabs) { deps = st->tolabs; last = 1; }
    if (l >= kNLayMax - 1) last = 1;
    st->deps[l] = deps;
    double aopt = 1.0 / deps;
    st->aopt[l] = aopt;
    st->bopt[l] = -mn * aopt + 0.5;
    st->active[l] = 1;
    st->nlay = l + 1;
    if (last) st->done = 1;
}
void layer_params(DevState* st, int layer, cudaStream_t s) { layer_params_kernel<<<1, 1, 0, s>>>(st, layer); note_launch(1); }

// One CTA per coder block (<= 60000 symbols).  reference wrappers.cpp:384-398 + rangecod.c:389-397
//
// The kernel has to move 9 bytes per coefficient and nothing else should cost: the first version issued
// ~80 instructions per coefficient (per-element bounds branches, __match_any_sync + shared atomics for the
// histogram, f64<->int conversions) and was issue-bound at 2.5 TB/s.  This one issues ~25-35:
//   * floor by magic number: t = fq + 2^52 rounded DOWN holds floor(fq) in its low mantissa bits and
//     t - 2^52 is that integer as a double -- what (unsigned char)fq followed by the int -> double
//     conversion give for 0 <= fq < 256 (fq lies in [0.5, 255.5] by construction of aopt/bopt; only the
//     low byte is kept, like the reference's cast) -- no conversion instructions;
//   * histogram without atomics: every thread owns a private 256 x 1-byte counter set in shared memory
//     (a thread sees at most ceil(60000/256) = 235 symbols, so a byte cannot overflow); thread t's counter
//     for bin q is byte t>>6 of word q*64 + (t&63), so the 32 lanes of a warp always hit 32 different banks;
//   * full tiles of 2048 coefficients run without predicates, eight 8-byte loads in flight per thread
//     (HBM latency needs ~35 KB in flight per SM at three resident CTAs);
//   * extrema are tracked as doubles (fmin/fmax are order-independent, as in the reference) and turned
//     into ordered keys once per thread.
constexpr int kQThreads = 256;
constexpr int kQLoads = 8;                               // coefficients per thread and tile
constexpr int kQTile = kQThreads * kQLoads;              // 2048
constexpr int kQAhead = 3;                               // L2 prefetch distance in tiles
constexpr int kQHistBytes = 256 * kQThreads;             // private byte counters: 64 KiB
// CTAs per coder block.  The per-CTA costs (clearing 64 KiB of counters, adding them up, the first exposed load, the
// extrema commit) outweigh the tail of the last wave: measured at 512^3 (2237 blocks, 444 CTA slots) the three layers
// take 0.71 ms with one CTA per block, 0.73 with two, 0.77 with three, 0.81 with four.  Partial histograms are added
// with global atomics, so any split works.
constexpr unsigned int kQSubLen = 30 * kQTile;           // >= 60000: one CTA per block
constexpr unsigned int kQSub = (kBlock + kQSubLen - 1) / kQSubLen;

__device__ __forceinline__ double floor_magic(double fq)
{
    return __dadd_rd(fq, 4503599627370496.0);            // 2^52: low mantissa bits = floor(fq)
}

// residual after layer `nl` of coefficient r, symbol of layer nl in the low byte of q
template <int LAYER>
__device__ __forceinline__ double quantise_one(double r, const double* s_a, const double* s_b, const double* s_d,
                                               const double* s_m, int layer, unsigned int& q)
{
    const int nl = (LAYER >= 0) ? LAYER : layer;
#pragma unroll
    for (int m = 0; m < (LAYER >= 0 ? LAYER : kNLayMax - 1); m++) {          // replay earlier layers (:387-398)
        if (LAYER < 0 && m >= nl) break;
        const double t = floor_magic(s_a[m] * r + s_b[m]);
        const double qd = t - 4503599627370496.0;                           // exact: the integer as a double
        r = r - (qd * s_d[m] + s_m[m]);
    }
    const double t = floor_magic(s_a[nl] * r + s_b[nl]);
    q = (unsigned int)__double2loint(t) & 0xFFu;
    return r - ((t - 4503599627370496.0) * s_d[nl] + s_m[nl]);
}

template <int LAYER>
__global__ void __launch_bounds__(kQThreads, (LAYER <= 2) ? 3 : 2) quantise_kernel(const double* __restrict__ coef, ChunkGeom g, int layer,
                                                               DevState* st, uint8_t* __restrict__ sym,
                                                               uint32_t* __restrict__ hist)
{
    if (!st->active[layer]) return;
    const int tid = threadIdx.x;
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    const unsigned int sb0 = blockIdx.y * kQSubLen;       // this CTA's part of the coder block
    if (sb0 >= bs) return;
    const unsigned int n = (bs - sb0 < kQSubLen) ? bs - sb0 : kQSubLen;
    extern __shared__ __align__(16) uint32_t s_cnt[];     // [256 bins][64 words], see above
    __shared__ double s_a[kNLayMax], s_b[kNLayMax], s_d[kNLayMax], s_m[kNLayMax];
    const double* __restrict__ in = coef + cstart + boff + sb0 + tid;
    uint8_t* __restrict__ out = sym + (unsigned long long)c * g.pitch + boff + sb0 + tid;
    // the first tile's loads and the L2 prefetch of the next ones go out before the counters are cleared
    double r[kQLoads], rn[kQLoads];
    const unsigned int nfull = n / kQTile;
    if (nfull > 0) {
#pragma unroll
        for (int k = 0; k < kQLoads; k++) r[k] = in[k * kQThreads];
        for (unsigned int t = 1; t <= kQAhead && t < nfull; t++) {
#pragma unroll
            for (int k = 0; k < kQLoads; k++) prefetch_l2(in + t * kQTile + k * kQThreads);
        }
    }
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        for (int i = tid; i < kQHistBytes / 16; i += kQThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid <= layer) { s_a[tid] = st->aopt[tid]; s_b[tid] = st->bopt[tid]; s_d[tid] = st->deps[tid]; s_m[tid] = st->minval[tid]; }
    __syncthreads();
    uint8_t* const cnt = reinterpret_cast<uint8_t*>(s_cnt);
    const unsigned int mine = (tid & 63) * 4 + (tid >> 6);      // < 256: byte 0 of a counter's offset, the bin is byte 1
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    double rmin = kInf, rmax = -kInf;
    // The residual extrema feed the NEXT layer's parameters: the terminating layer (st->done is set by its layer_params)
    // has no successor, so its pass skips them -- six instructions per coefficient of a pass that issues ~40.  Two
    // copies of the loop (a generic lambda instantiated for both cases): the choice costs nothing per element.
    const bool track = st->done == 0;
    auto sweep = [&](auto track_tag) {
    constexpr bool TRACK = decltype(track_tag)::value;
    auto one = [&](double r, uint8_t* dst) {
        unsigned int q;
        const double res = quantise_one<LAYER>(r, s_a, s_b, s_d, s_m, layer, q);
        *dst = (uint8_t)q;
        cnt[__byte_perm(q, mine, 0x6504)] += 1;                  // offset = bin * 256 + mine in one PRMT
        if (TRACK) {
            rmin = dmin2(rmin, res);
            rmax = dmax2(rmax, res);
        }
    };
    // full tiles: no predicates; the loads of tile t+1 are issued before tile t is processed (two register
    // sets used alternately); running pointers keep every access at base + immediate
    for (unsigned int t = 0; t < nfull; t += 2) {
        if (t + kQAhead + 1 < nfull) {                    // two tiles per iteration: tiles t+kQAhead, t+kQAhead+1
#pragma unroll
            for (int k = 0; k < 2 * kQLoads; k++) prefetch_l2(in + kQAhead * kQTile + k * kQThreads);
        }
        if (t + 1 < nfull) {
#pragma unroll
            for (int k = 0; k < kQLoads; k++) rn[k] = in[kQTile + k * kQThreads];
        }
#pragma unroll
        for (int k = 0; k < kQLoads; k++) one(r[k], out + k * kQThreads);
        in += kQTile; out += kQTile;
        if (t + 1 < nfull) {
            if (t + 2 < nfull) {
#pragma unroll
                for (int k = 0; k < kQLoads; k++) r[k] = in[kQTile + k * kQThreads];
            }
#pragma unroll
            for (int k = 0; k < kQLoads; k++) one(rn[k], out + k * kQThreads);
            in += kQTile; out += kQTile;
        }
    }
    const unsigned int i0 = nfull * kQTile;
    if (i0 < n) {                                         // ragged tail
#pragma unroll
        for (int k = 0; k < kQLoads; k++) r[k] = (i0 + k * kQThreads + tid < n) ? in[k * kQThreads] : 0.0;
#pragma unroll
        for (int k = 0; k < kQLoads; k++)
            if (i0 + k * kQThreads + tid < n) one(r[k], out + k * kQThreads);
    }
    };
    if (track) sweep(std::true_type{}); else sweep(std::false_type{});
    __syncthreads();
    {   // thread = bin: add up the 256 private counters of the bin (64 words, rotated start -> no bank conflicts)
        const uint4* row = reinterpret_cast<const uint4*>(s_cnt + tid * 64);
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < 16; w++) {
            const uint4 x = row[(w + tid) & 15];
            tot = __dp4a(x.x, 0x01010101u, tot);          // sum of the four byte counters of a word (UNSIGNED bytes:
            tot = __dp4a(x.y, 0x01010101u, tot);          // a counter can exceed 127)
            tot = __dp4a(x.z, 0x01010101u, tot);
            tot = __dp4a(x.w, 0x01010101u, tot);
        }
        if (tot) atomicAdd(&hist[(unsigned long long)b * 256 + tid], tot);      // hist is zeroed per encode
    }
    const bool any = track && (unsigned int)tid < n;
    if (track)                                            // uniform over the CTA (it contains barriers)
        block_minmax_commit(any ? dkey(rmin) : kKeyMinInit, any ? dkey(rmax) : kKeyMaxInit, &st->rmin_key[layer + 1],
                            &st->rmax_key[layer + 1]);
}

template <int LAYER>
static void launch_quantise(const double* coef, const ChunkGeom& g, int layer, DevState* st, uint8_t* sym, uint32_t* hist,
                            cudaStream_t s)
{
    static DeviceOnce once;                               // 64 KiB dynamic shared memory: above the 48 KiB default (per device)
    once.run([] { cudaFuncSetAttribute(quantise_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQHistBytes); });
    quantise_kernel<LAYER><<<dim3(g.nblocks, kQSub, 1), kQThreads, kQHistBytes, s>>>(coef, g, layer, st, sym, hist);
}

void quantise_layer(const double* coef, const ChunkGeom& g, int layer, DevState* st, uint8_t* sym, uint32_t* hist,
                    cudaStream_t s)
{
    switch (layer) {
    case 0: launch_quantise<0>(coef, g, layer, st, sym, hist, s); break;
    case 1: launch_quantise<1>(coef, g, layer, st, sym, hist, s); break;
    case 2: launch_quantise<2>(coef, g, layer, st, sym, hist, s); break;
    case 3: launch_quantise<3>(coef, g, layer, st, sym, hist, s); break;
    case 4: launch_quantise<4>(coef, g, layer, st, sym, hist, s); break;
    case 5: launch_quantise<5>(coef, g, layer, st, sym, hist, s); break;
    case 6: launch_quantise<6>(coef, g, layer, st, sym, hist, s); break;
    default: launch_quantise<7>(coef, g, layer, st, sym, hist, s); break;
    }
    note_launch(1);
}

// ------------------------------------------------------------------------------------------
// Local (spatially varying) cutoff: the mx*my*mz > 1 branch of encoding_wrap() (wrappers.cpp:343-379).
// For every point the reference takes the wavelet-space index and level from ind_p2w_3d(), a precision
//   precmask = (level <= LOC_CUTOFF_LVL) ? tolabs / tolrel * cutoffvec[block of the PHYSICAL point] : tolabs,
// and, in every layer, codes symbol 0 and leaves residual 0 (fld = minval; fld -= 0*deps + minval) where the
// layer's residual span max - min is below precmask.  ind_p2w_3d() reports level == lvlin for every point (its
// `chlvl` flag is never cleared, waveletcdf97_3d.c:487,535), so with the transform on (lvlin = 4 > 1) precmask
// is tolabs everywhere, and with the transform off (lvlin = 0) the wavelet index IS the physical index and every
// point gets its block's cutoff.  Both cases are reproduced as they are.  Rarely used: a plain kernel, one CTA per
// coder block, shared-memory atomics for the histogram.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quantise_masked_kernel(const double* __restrict__ coef, ChunkGeom g, int layer,
                                                              DevState* st, uint8_t* __restrict__ sym,
                                                              uint32_t* __restrict__ hist, LocalCutoff lc)
{
    if (!st->active[layer]) return;
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    const double tolabs = st->tolabs;
    const double scale = tolabs / lc.tolrel;                         // tolabs/tolrel * lcl_prec(...), left to right
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    double rmin = kInf, rmax = -kInf;
    for (unsigned int i = tid; i < bs; i += blockDim.x) {
        const unsigned long long jw = cstart + boff + i;
        double pm = tolabs;
        if (lc.per_point) {                                          // level 0: physical index == array index
            const int jx = (int)(jw % (unsigned long long)lc.nx);
            const int jy = (int)((jw / (unsigned long long)lc.nx) % (unsigned long long)lc.ny);
            const int jz = (int)(jw / ((unsigned long long)lc.nx * lc.ny));
            const int kx = (int)((double)jx / (double)lc.nx * (double)lc.mx);     // lcl_prec, wrappers.cpp:55-64
            const int ky = (int)((double)jy / (double)lc.ny * (double)lc.my);
            const int kz = (int)((double)jz / (double)lc.nz * (double)lc.mz);
            pm = scale * lc.cut[kx + lc.mx * ky + lc.mx * lc.my * kz];
        }
        double r = coef[jw];
        unsigned int q = 0;
        for (int m = 0; m <= layer; m++) {
            if (st->span[m] < pm) {                                  // :365-371
                q = 0;
                r = 0.0;                                             // minval - (0*deps + minval)
            } else {
                const double fq = st->aopt[m] * r + st->bopt[m];
                q = (unsigned int)(unsigned char)__double2int_rz(fq);
                r = r - ((double)q * st->deps[m] + st->minval[m]);
            }
        }
        sym[(unsigned long long)c * g.pitch + boff + i] = (uint8_t)q;
        atomicAdd(&s_hist[q], 1u);
        rmin = dmin2(rmin, r);
        rmax = dmax2(rmax, r);
    }
    __syncthreads();
    if (s_hist[tid]) atomicAdd(&hist[(unsigned long long)b * 256 + tid], s_hist[tid]);
    block_minmax_commit(rmin <= rmax ? dkey(rmin) : kKeyMinInit, rmin <= rmax ? dkey(rmax) : kKeyMaxInit, &st->rmin_key[layer + 1],
                        &st->rmax_key[layer + 1]);
}

void quantise_layer_masked(const double* coef, const ChunkGeom& g, int layer, DevState* st, uint8_t* sym, uint32_t* hist,
                           const LocalCutoff& lc, cudaStream_t s)
{
    quantise_masked_kernel<<<g.nblocks, 256, 0, s>>>(coef, g, layer, st, sym, hist, lc);
    note_launch(1);
}

struct DequantParams { double deps[kNLayMax], minval[kNLayMax]; };

// reference wrappers.cpp:480 and :513-514: fld = 0; fld = fld + (q*deps + minval) per layer
__global__ void __launch_bounds__(256) dequantise_kernel(const uint8_t* __restrict__ sym, unsigned long long layer_stride,
                                                         ChunkGeom g, int nlay, DequantParams p,
                                                         double* __restrict__ coef)
{
    __shared__ double s_d[kNLayMax], s_m[kNLayMax];
    if (threadIdx.x < kNLayMax) { s_d[threadIdx.x] = p.deps[threadIdx.x]; s_m[threadIdx.x] = p.minval[threadIdx.x]; }
    __syncthreads();
    const unsigned int c = blockIdx.x;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned int clen = (unsigned int)((g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len);
    const uint8_t* __restrict__ in = sym + (unsigned long long)c * g.pitch;
    double* __restrict__ out = coef + cstart;
    for (unsigned long long i0 = blockIdx.y * 1024ull; i0 < clen; i0 += (unsigned long long)gridDim.y * 1024ull) {
        double f[4] = {0, 0, 0, 0};
        for (int l = 0; l < nlay; l++) {
            const uint8_t* __restrict__ il = in + (unsigned long long)l * layer_stride + i0 + threadIdx.x;
            unsigned int q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = (i0 + k * 256 + threadIdx.x < clen) ? il[k * 256] : 0;
#pragma unroll
            for (int k = 0; k < 4; k++) f[k] = f[k] + ((double)q[k] * s_d[l] + s_m[l]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k * 256 + threadIdx.x < clen) out[i0 + k * 256 + threadIdx.x] = f[k];
    }
}

void dequantise(const uint8_t* sym, unsigned long long layer_stride, const ChunkGeom& g, int nlay, const double* deps,
                const double* minval, double* coef, cudaStream_t s)
{
    DequantParams p{};
    for (int l = 0; l < nlay; l++) { p.deps[l] = deps[l]; p.minval[l] = minval[l]; }
    unsigned long long per = (g.chunk_len + 1023) / 1024;
    unsigned int gy = (unsigned int)(per < 8 ? (per ? per : 1) : 8);
    if (g.nchunks < 148 * 4) {
        unsigned long long want = (148ull * 16 + g.nchunks - 1) / g.nchunks;
        gy = (unsigned int)(per < want ? (per ? per : 1) : want);
    }
    dim3 grid(g.nchunks, gy, 1);
    dequantise_kernel<<<grid, 256, 0, s>>>(sym, layer_stride, g, nlay, p, coef);
    note_launch(1);
}

}  // namespace wrb
