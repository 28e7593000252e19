// slab_order.cu -- the symbol exchange that puts a z-slab partitioned field into the GLOBAL wavelet-space order.
//
// The reference codes the symbols of a layer in the array order of the whole coefficient array
// (wrappers.cpp:384-412), and that order is fixed by the in-box de-interleave of every level
// (waveletcdf97_3d.c:128-135, 256-263): along z, plane w of the global array holds
//   * for (x, y) that left the low box at level k (region k): the level-k' z-high band if
//     nz/2^k' <= w < nz/2^(k'-1) for some k' <= k, else the level-k z-low details (w < nz/2^k);
//   * for (x, y) inside the coarsest low box (region levels+1): the same with k = levels, the planes below nz/2^levels
//     being the approximation.
// Under the z-slab partition rank r owns the pairs [z0/2^k', (z0+nzl)/2^k') of every level, stored in a rank-local
// array laid out like the transform of an independent (nx, ny, nzl) field.  So every (global plane, region) has one
// (owner rank, local plane), in closed form (order_source / order_global_plane below) -- no tables.
//
// Rank r codes a contiguous run of whole chunks of the global sequence (make_order_geom).  Its run is GATHERED straight
// from the peers' symbol planes through peer pointers (CUDA IPC over NVLink, or plain pointers when the ranks are
// emulated in one process), written in the coder's chunk-major layout with the coder blocks' histograms on the way:
// one kernel, no staging buffers, every byte crosses NVLink once.  Decoding mirrors it: every rank decodes its run,
// then gathers its local symbol planes from the peers' runs.
#include "wr_common.cuh"
#include "wr_kernels.h"
#include "slab_comm.h"

namespace wrb {

OrderGeom make_order_geom(int nx, int ny, int nz, int nranks, int levels, unsigned long long chunk_len)
{
    OrderGeom og{};
    og.nx = nx; og.ny = ny; og.nz = nz; og.nzl = nz / nranks; og.levels = levels; og.nranks = nranks;
    og.ntot = (unsigned long long)nx * ny * nz;
    og.chunk_len = (chunk_len == 0 || chunk_len > og.ntot) ? og.ntot : chunk_len;
    og.nchunks = (og.ntot + og.chunk_len - 1) / og.chunk_len;
    for (int r = 0; r <= nranks; r++) {
        og.cb[r] = (unsigned long long)r * og.nchunks / (unsigned long long)nranks;
        const unsigned long long j = og.cb[r] * og.chunk_len;
        og.j0[r] = j < og.ntot ? j : og.ntot;
    }
    return og;
}

struct OrderDev {
    int nx, ny, nz, nzl, levels, nranks;
    unsigned long long plane;                       // nx * ny
    unsigned long long j0[kMaxRanks + 1];
    int mx[8], my[8];                               // low extents in x, y after k levels (k = 0: the full extents)
};

static OrderDev make_dev(const OrderGeom& og)
{
    OrderDev o{};
    o.nx = og.nx; o.ny = og.ny; o.nz = og.nz; o.nzl = og.nzl; o.levels = og.levels; o.nranks = og.nranks;
    o.plane = (unsigned long long)og.nx * og.ny;
    for (int r = 0; r <= og.nranks; r++) o.j0[r] = og.j0[r];
    for (int k = 0; k < 8; k++) {
        o.mx[k] = (int)(((long long)og.nx + (1ll << k) - 1) >> k);
        o.my[k] = (int)(((long long)og.ny + (1ll << k) - 1) >> k);
    }
    return o;
}

// region of an (x, y) position: the level at which it leaves the low box, levels + 1 inside the coarsest box
__host__ __device__ inline int order_region(const OrderDev& o, int x, int y)
{
    for (int k = 1; k <= o.levels; k++)
        if (x >= o.mx[k] || y >= o.my[k]) return k;
    return o.levels + 1;
}

// (owner rank, local plane) of global wavelet-space plane w for a position of region reg
__host__ __device__ inline void order_source(const OrderDev& o, int w, int reg, int& owner, int& p)
{
    const int kmax = reg < o.levels ? reg : o.levels;
    for (int k = 1; k <= kmax; k++) {
        const int low = o.nz >> k;
        if (w >= low) {                                  // z-high band of level k (w < nz >> (k-1) here)
            const int i = w - low, nl = o.nzl >> k;
            owner = i / nl;
            p = nl + (i - owner * nl);
            return;
        }
    }
    const int nl = o.nzl >> kmax;                        // z-low details of level kmax / the approximation
    owner = w / nl;
    p = w - owner * nl;
}

// global wavelet-space plane of local plane p of `rank` for a position of region reg
__host__ __device__ inline int order_plane(const OrderDev& o, int rank, int p, int reg)
{
    const int kmax = reg < o.levels ? reg : o.levels;
    const int z0 = rank * o.nzl;
    for (int k = 1; k <= kmax; k++) {
        const int nl = o.nzl >> k;
        if (p >= nl) return (o.nz >> k) + (z0 >> k) + (p - nl);      // p < 2 nl here
    }
    return (z0 >> kmax) + p;
}

int order_global_plane(const OrderGeom& og, int rank, int p, int reg)
{
    const OrderDev o = make_dev(og);
    return order_plane(o, rank, p, reg);
}

// four / sixteen bytes at an arbitrary address through aligned word loads (the buffers have >= 16 bytes of slack)
__device__ __forceinline__ uint32_t load4_unaligned(const uint8_t* src)
{
    const unsigned long long a = (unsigned long long)src;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~3ull);
    const uint32_t sh = (uint32_t)(a & 3ull) * 8u;
    const uint32_t w0 = wp[0], w1 = wp[1];
    return __funnelshift_r(w0, w1, sh);
}
// Sixteen bytes at an arbitrary address for every lane of a warp, when consecutive lanes mostly read consecutive 16-byte
// items of one contiguous source run.  Loads that cross NVLink are not kept in L1, so five overlapping word loads per
// lane (the obvious way) move every sector five times; here a lane loads ONE aligned 16-byte vector, takes the next one
// from its neighbour by shuffle (or loads it itself where the run ends) and shifts the pair in registers.
// All 32 lanes must call it; `on` = this lane wants data.
// Two phases, so that a thread can have the vectors of several items in flight before it touches the first.  A lane at the
// end of a contiguous run (no neighbour to take the next vector from) loads that vector itself, and it does so in the
// SAME phase as its first one: whether it is a run end follows from the addresses alone.  (Issued after the shuffle of
// the data, as a first version did, the second load cost every warp a second trip over NVLink -- with sub-band rows of
// 256 symbols two lanes of every warp are run ends; tools/p2p_probe.cu: a plain copy kernel pulls 650-750 GB/s where the
// exchange kernels reached 450.)  All 32 lanes must call both phases.
struct Load16 { unsigned long long a; uint4 v, n2; bool on, own; };
__device__ __forceinline__ Load16 load16_issue(const uint8_t* src, bool on)
{
    Load16 l;
    l.a = on ? (unsigned long long)src : 0ull;
    l.on = on;
    const unsigned long long base = l.a & ~15ull;
    // the neighbour's vector is my next one iff its aligned address is mine + 16
    const unsigned long long an = __shfl_down_sync(0xffffffffu, base, 1);
    l.own = on && (l.a & 15ull) != 0 && ((threadIdx.x & 31) == 31 || an != base + 16ull);
    const uint4* vp = reinterpret_cast<const uint4*>(base);
    l.v = make_uint4(0, 0, 0, 0);
    l.n2 = make_uint4(0, 0, 0, 0);
    if (on) l.v = vp[0];
    if (l.own) l.n2 = vp[1];
    return l;
}
__device__ __forceinline__ uint4 load16_finish(const Load16& l)
{
    const unsigned long long a = l.a;
    const uint4 v = l.v;
    uint4 n;
    n.x = __shfl_down_sync(0xffffffffu, v.x, 1); n.y = __shfl_down_sync(0xffffffffu, v.y, 1);
    n.z = __shfl_down_sync(0xffffffffu, v.z, 1); n.w = __shfl_down_sync(0xffffffffu, v.w, 1);
    if (l.own) n = l.n2;
    const unsigned int s = (unsigned int)(a & 15ull);
    // bytes [s, s + 16) of (v, n)
    const unsigned int ws = s >> 2, bs = (s & 3u) * 8u;
    const uint32_t W[8] = {v.x, v.y, v.z, v.w, n.x, n.y, n.z, n.w};
    uint32_t o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint32_t x = W[k];
        x = (ws == 1) ? W[k + 1] : x;
        x = (ws == 2) ? W[k + 2] : x;
        x = (ws == 3) ? W[(k + 3) & 7] : x;
        o[k] = x;
    }
    return make_uint4(__funnelshift_r(o[0], o[1], bs), __funnelshift_r(o[1], o[2], bs), __funnelshift_r(o[2], o[3], bs),
                      __funnelshift_r(o[3], o[4], bs));
}
__device__ __forceinline__ uint4 warp_load16_unaligned(const uint8_t* src, bool on) { return load16_finish(load16_issue(src, on)); }

// ---- encode side -----------------------------------------------------------------------------------------------
// Generic version (any shape; used for rows shorter than 256 symbols).  Pure copy kernel, grid (coder blocks of my run,
// layers), no shared memory: what limits it is the latency of loads
// that cross NVLink (~2-3 us), so it wants as many bytes in flight as the SM can hold -- 2048 threads x 20 bytes.  (A
// first version read 4 bytes per thread and iteration and histogrammed on the way: 180 GB/s.)  The coder blocks'
// histograms are taken afterwards from the local copy (hist_blocks_kernel).
__global__ void __launch_bounds__(256) gather_run_small_kernel(OrderDev o, int rank, PeerPtrs peer, unsigned long long peer_stride,
                                                         const int* __restrict__ active, ChunkGeom g,
                                                         uint8_t* __restrict__ sym, unsigned long long lstride)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
    const int tid = threadIdx.x;
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    // global position of the block's first symbol
    const unsigned long long jb = o.j0[rank] + cstart + boff;
    const unsigned long long wb = jb / o.plane;
    const unsigned int rem = (unsigned int)(jb - wb * o.plane);
    const unsigned int yb = rem / (unsigned int)o.nx, xb = rem - yb * (unsigned int)o.nx;
    uint8_t* __restrict__ out = sym + (unsigned long long)layer * lstride + (unsigned long long)c * g.pitch + boff;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    // position (x, y, w) of symbol q of the block
    auto locate = [&](unsigned int q, int& x, int& y, int& w) {
        const unsigned int t = xb + q;
        const unsigned int dy = t / (unsigned int)o.nx;
        x = (int)(t - dy * (unsigned int)o.nx);
        const unsigned int yy = yb + dy;
        const unsigned int dw = yy / (unsigned int)o.ny;
        y = (int)(yy - dw * (unsigned int)o.ny);
        w = (int)(wb + dw);
    };
    auto source = [&](int x, int y, int w, int reg) -> const uint8_t* {
        int owner, p;
        order_source(o, w, reg, owner, p);
        return peer.p[owner] + loff + ((unsigned long long)p * o.ny + y) * o.nx + x;
    };
    for (unsigned int q0 = 0; q0 < bs; q0 += 16u * 256u) {            // warp-uniform trip count: the lanes shuffle
        const unsigned int q = q0 + 16u * tid;
        int x = 0, y = 0, w = 0, reg = 0;
        bool fast = false;
        const uint8_t* src = nullptr;
        if (q < bs) {
            locate(q, x, y, w);
            reg = order_region(o, x, y);
            fast = q + 16 <= bs && x + 15 < o.nx && order_region(o, x + 15, y) == reg;   // 16 symbols of one row and region
            if (fast) src = source(x, y, w, reg);
        }
        const uint4 v = warp_load16_unaligned(src, fast);
        if (fast) { *reinterpret_cast<uint4*>(out + q) = v; continue; }
        if (q >= bs) continue;
        for (unsigned int q4 = q; q4 < q + 16 && q4 < bs; q4 += 4) {                  // row end / region border / tail
            locate(q4, x, y, w);
            const int r4 = order_region(o, x, y);
            if (q4 + 4 <= bs && x + 3 < o.nx && order_region(o, x + 3, y) == r4) {
                *reinterpret_cast<uint32_t*>(out + q4) = load4_unaligned(source(x, y, w, r4));
            } else {
                for (unsigned int e = q4; e < q4 + 4 && e < bs; e++) {
                    locate(e, x, y, w);
                    out[e] = *source(x, y, w, order_region(o, x, y));
                }
            }
        }
    }
}

// Fast version for rows of >= 256 symbols (every real field).  The generic kernel above spends ~80 instructions per
// byte on index arithmetic (two divisions, the region and source searches, per item and again per byte where an item
// straddles a region border: 1.2 ms for a 512 x 512 x 512 slab's three layers).  Here the CTA first writes, per row its
// coder block touches (<= 60000 / 256 + 2), the row's segments -- runs of one region, i.e. of one (owner, local plane) --
// with their source pointers into shared memory; the copy loop then finds an item's row by one multiply-shift, its
// segment by comparisons and its source by one shared load.
constexpr int kGRowsMax = 240;
struct FastDiv { uint32_t mul, sh; };
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv d) { return __umulhi(n, d.mul) >> d.sh; }   // exact for n <= 2^31, divisor >= 2

__global__ void __launch_bounds__(256) gather_run_kernel(OrderDev o, FastDiv dnx, int rank, PeerPtrs peer, unsigned long long peer_stride,
                                                         const int* __restrict__ active, ChunkGeom g,
                                                         uint8_t* __restrict__ sym, unsigned long long lstride)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
    __shared__ const uint8_t* s_ptr[kGRowsMax * 5];   // source of x = 0 of every (row, segment): add x
    __shared__ uint8_t s_ky[kGRowsMax];               // segments of the row (= region of its y alone)
    const int tid = threadIdx.x;
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    const unsigned long long jb = o.j0[rank] + cstart + boff;
    const unsigned long long wb = jb / o.plane;
    const unsigned int rem = (unsigned int)(jb - wb * o.plane);
    const unsigned int yb = rem / (unsigned int)o.nx, xb = rem - yb * (unsigned int)o.nx;
    uint8_t* __restrict__ out = sym + (unsigned long long)layer * lstride + (unsigned long long)c * g.pitch + boff;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    const int L = o.levels;
    const unsigned int nrows = bs ? (xb + bs - 1) / (unsigned int)o.nx + 1 : 0;
    // ---- the rows' segments: segment s of a row with ky segments covers x in [B_s, B_s+1), B_0 = 0, B_j = mx[ky - j], B_ky = nx,
    //      and has region ky for s == 0, ky - s otherwise ----
    for (unsigned int t = tid; t < nrows * 5u; t += 256u) {
        const unsigned int r = t / 5u, sgm = t - r * 5u;
        const unsigned int yy = yb + r;
        const unsigned int dw = yy / (unsigned int)o.ny;
        const int y = (int)(yy - dw * (unsigned int)o.ny);
        const int w = (int)(wb + dw);
        int ky = L + 1;
        for (int k = L; k >= 1; k--) if (y >= o.my[k]) ky = k;
        if (sgm == 0) s_ky[r] = (uint8_t)ky;
        if ((int)sgm < ky) {
            int owner, p;
            order_source(o, w, ky - (int)sgm, owner, p);
            s_ptr[t] = peer.p[owner] + loff + ((unsigned long long)p * o.ny + y) * o.nx;
        }
    }
    __syncthreads();
    const int m1 = o.mx[1], m2 = o.mx[2], m3 = o.mx[3], m4 = o.mx[4];
    // segment of x in a row with ky segments, and the end of that segment
    auto segment = [&](int x, int ky, int& xe) -> int {
        // boundaries B_j = mx[ky - j], j = 1 .. ky-1: count those <= x
        int s = 0;
        xe = o.nx;
        if (ky > 4) { if (x >= m4) s++; else { xe = m4; return s; } }      // ky = 5: B_1 = mx[4]
        if (ky > 3) { if (x >= m3) s++; else { xe = m3; return s; } }
        if (ky > 2) { if (x >= m2) s++; else { xe = m2; return s; } }
        if (ky > 1) { if (x >= m1) s++; else { xe = m1; return s; } }
        return s;
    };
    // (two items per thread and iteration, both loads in flight before the first is used, were measured on 4 B200: 10 % slower)
    for (unsigned int q0 = 0; q0 < bs; q0 += 16u * 256u) {            // warp-uniform trip count: the lanes shuffle
        const unsigned int q = q0 + 16u * tid;
        bool fast = false;
        const uint8_t* src = nullptr;
        unsigned int r = 0;
        int x = 0, xe = 0, ky = 1, sgm = 0;
        if (q < bs) {
            const unsigned int t = xb + q;
            r = fdiv(t, dnx);
            x = (int)(t - r * (unsigned int)o.nx);
            ky = s_ky[r];
            sgm = segment(x, ky, xe);
            fast = q + 16 <= bs && x + 16 <= xe;
            src = s_ptr[r * 5 + sgm] + x;
        }
        const uint4 v = warp_load16_unaligned(src, fast);
        if (fast) { *reinterpret_cast<uint4*>(out + q) = v; continue; }
        if (q >= bs) continue;
        for (unsigned int e = q; e < q + 16 && e < bs; e++) {          // the item crosses a segment or row end (or is the tail)
            if (x >= xe) {
                if (x >= o.nx) { x = 0; r++; ky = s_ky[r]; }
                sgm = segment(x, ky, xe);
                src = s_ptr[r * 5 + sgm] + x;
            }
            out[e] = *src;
            src++; x++;
        }
    }
}

// 256-bin histogram of every coder block of the (local, chunk-major) symbols: one CTA per block, private one-byte
// counters per thread as in the quantiser (quant.cu) -- no atomics, so a layer that is almost one symbol costs the same
constexpr int kHThreads = 256;
constexpr int kHHistBytes = 256 * kHThreads;             // 64 KiB
__global__ void __launch_bounds__(kHThreads) hist_blocks_kernel(const uint8_t* __restrict__ sym, unsigned long long lstride,
                                                                const int* __restrict__ active, ChunkGeom g,
                                                                uint32_t* __restrict__ hist, unsigned long long hstride)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
    extern __shared__ __align__(16) uint32_t s_cnt[];     // [256 bins][64 words]: thread t's counter of bin q is byte t>>6 of word q*64 + (t&63)
    const int tid = threadIdx.x;
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        for (int i = tid; i < kHHistBytes / 16; i += kHThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    uint8_t* const cnt = reinterpret_cast<uint8_t*>(s_cnt);
    const unsigned int mine = (tid & 63) * 4 + (tid >> 6);
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    const uint8_t* __restrict__ in = sym + (unsigned long long)layer * lstride + (unsigned long long)c * g.pitch + boff;
    // (the next vector is in flight while the sixteen counter updates of this one run: with 64 KiB of counters only three
    //  CTAs share an SM, too few to hide the load otherwise -- r3 profile: 7.6 stall cycles per issue on it)
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (16u * tid < bs) nxt = *reinterpret_cast<const uint4*>(in + 16u * tid);      // chunks start 16-byte aligned, pitch slack covers the tail
    for (unsigned int q = 16u * tid; q < bs; q += 16u * kHThreads) {                  // <= 15 vectors = 240 symbols per thread
        const uint4 v4 = nxt;
        if (q + 16u * kHThreads < bs) nxt = *reinterpret_cast<const uint4*>(in + q + 16u * kHThreads);
        const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
        const unsigned int n = bs - q;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v = vv[k];
            if (n > 4u * k) cnt[__byte_perm(v, mine, 0x6504)] += 1;                  // offset = bin * 256 + mine
            if (n > 4u * k + 1) cnt[__byte_perm(v >> 8, mine, 0x6504)] += 1;
            if (n > 4u * k + 2) cnt[__byte_perm(v >> 16, mine, 0x6504)] += 1;
            if (n > 4u * k + 3) cnt[__byte_perm(v >> 24, mine, 0x6504)] += 1;
        }
    }
    __syncthreads();
    const uint4* row = reinterpret_cast<const uint4*>(s_cnt + tid * 64);
    unsigned int tot = 0;
#pragma unroll
    for (int wd = 0; wd < 16; wd++) {
        const uint4 xx = row[(wd + tid) & 15];
        tot = __dp4a(xx.x, 0x01010101u, tot);
        tot = __dp4a(xx.y, 0x01010101u, tot);
        tot = __dp4a(xx.z, 0x01010101u, tot);
        tot = __dp4a(xx.w, 0x01010101u, tot);
    }
    hist[(unsigned long long)layer * hstride + (unsigned long long)b * 256 + tid] = tot;
}

void gather_global_run(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlayers,
                       const int* active, const ChunkGeom& g, uint8_t* sym, unsigned long long sym_layer_stride,
                       uint32_t* hist, unsigned long long hist_layer_stride, cudaStream_t s, cudaEvent_t after_gather)
{
    if (g.nblocks == 0 || nlayers <= 0) { if (after_gather) cudaEventRecord(after_gather, s); return; }
    const OrderDev o = make_dev(og);
    static DeviceOnce once;
    once.run([] { cudaFuncSetAttribute(hist_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHHistBytes); });
    dim3 grid(g.nblocks, nlayers, 1);
    if (og.nx >= 256 && og.levels <= 4) {
        uint32_t l2 = 0;
        while ((1u << l2) < (uint32_t)og.nx) l2++;                       // ceil(log2 nx)
        FastDiv d{(uint32_t)(((1ull << (31 + l2)) + (unsigned long long)og.nx - 1) / (unsigned long long)og.nx), l2 - 1};
        gather_run_kernel<<<grid, 256, 0, s>>>(o, d, rank, peer, peer_stride, active, g, sym, sym_layer_stride);
    } else {
        gather_run_small_kernel<<<grid, 256, 0, s>>>(o, rank, peer, peer_stride, active, g, sym, sym_layer_stride);
    }
    if (after_gather) cudaEventRecord(after_gather, s);
    hist_blocks_kernel<<<grid, kHThreads, kHHistBytes, s>>>(sym, sym_layer_stride, active, g, hist, hist_layer_stride);
    note_launch(2);
}

// segment of x in a row with ky segments (see gather_run_kernel), and the end of that segment
__device__ __forceinline__ int row_segment(int x, int ky, int m1, int m2, int m3, int m4, int nx, int& xe)
{
    int s = 0;
    xe = nx;
    if (ky > 4) { if (x >= m4) s++; else { xe = m4; return s; } }
    if (ky > 3) { if (x >= m3) s++; else { xe = m3; return s; } }
    if (ky > 2) { if (x >= m2) s++; else { xe = m2; return s; } }
    if (ky > 1) { if (x >= m1) s++; else { xe = m1; return s; } }
    return s;
}

// ---- decode side -----------------------------------------------------------------------------------------------
// grid (ceil(ny / kSRows), nzl, layers): one CTA per group of kSRows rows of one local plane of one layer; a thread moves
// 16 symbols per iteration (same reasoning as gather_run_kernel: the loads cross NVLink)
constexpr int kSRows = 16;

__global__ void __launch_bounds__(256) scatter_local_kernel(OrderDev o, int rank, PeerPtrs peer, unsigned long long peer_stride,
                                                            uint8_t* __restrict__ out, unsigned long long out_stride)
{
    const int p = blockIdx.y, layer = blockIdx.z;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    // global plane of this local plane for every region (1 .. levels + 1)
    int wreg[8];
#pragma unroll
    for (int k = 1; k < 8; k++) wreg[k] = (k <= o.levels + 1) ? order_plane(o, rank, p, k) : 0;
    auto wof = [&](int reg) -> int {                                  // register select instead of a local-memory array
        int w = wreg[1];
#pragma unroll
        for (int k = 2; k < 8; k++) w = (reg == k) ? wreg[k] : w;
        return w;
    };
    auto source = [&](unsigned long long j, unsigned long long& left) -> const uint8_t* {
        int d = 0;
        while (d + 1 < o.nranks && j >= o.j0[d + 1]) d++;
        left = o.j0[d + 1] - j;                                       // symbols of this rank's run from j on
        return peer.p[d] + loff + (j - o.j0[d]);
    };
    const int ipr = (o.nx + 15) >> 4;                                 // 16-symbol items per row
    const int y0 = blockIdx.x * kSRows;
    const int nrows = min(kSRows, o.ny - y0);
    for (int it0 = 0; it0 < nrows * ipr; it0 += 256) {                // warp-uniform trip count: the lanes shuffle
        const int it = it0 + threadIdx.x;
        const bool live = it < nrows * ipr;
        const int ry = live ? it / ipr : 0, x = live ? 16 * (it - ry * ipr) : 0;
        const int y = y0 + ry;
        uint8_t* __restrict__ row = out + (unsigned long long)layer * out_stride + ((unsigned long long)p * o.ny + y) * o.nx;
        unsigned long long left = 0;
        bool fast = false;
        const uint8_t* src = nullptr;
        if (live) {
            const int reg = order_region(o, x, y);
            if (x + 15 < o.nx && ((((unsigned long long)row) + x) & 15ull) == 0 && order_region(o, x + 15, y) == reg) {
                src = source(((unsigned long long)wof(reg) * o.ny + y) * o.nx + x, left);
                fast = left >= 16;
            }
        }
        const uint4 v = warp_load16_unaligned(src, fast);
        if (fast) { *reinterpret_cast<uint4*>(row + x) = v; continue; }
        if (!live) continue;
        for (int x4 = x; x4 < x + 16 && x4 < o.nx; x4 += 4) {
            const int r4 = order_region(o, x4, y);
            if (x4 + 3 < o.nx && ((((unsigned long long)row) + x4) & 3ull) == 0 && order_region(o, x4 + 3, y) == r4) {
                const uint8_t* s4 = source(((unsigned long long)wof(r4) * o.ny + y) * o.nx + x4, left);
                if (left >= 4) { *reinterpret_cast<uint32_t*>(row + x4) = load4_unaligned(s4); continue; }
            }
            for (int e = x4; e < x4 + 4 && e < o.nx; e++)
                row[e] = *source(((unsigned long long)wof(order_region(o, e, y)) * o.ny + y) * o.nx + e, left);
        }
    }
}

// Fast version for rows of >= 256 symbols, as on the encode side: the generic kernel above runs ~350 instructions per
// 16-byte item (region and owner searches, 64-bit index arithmetic; r3 profile: 281 M warp instructions, issue-bound at
// 84 %).  Here the CTA first writes the source pointer of every (row, segment) it covers into shared memory -- a segment
// is a run of one region, i.e. of one global plane, and almost always lies inside one rank's run -- and an item then costs
// one multiply-shift, a few comparisons and one shared load.  Segments that straddle the end of a run keep the generic path.
constexpr int kSSeg = 5;
__global__ void __launch_bounds__(256) scatter_local_fast_kernel(OrderDev o, FastDiv dipr, int ipr, int rank, PeerPtrs peer,
                                                                 unsigned long long peer_stride, uint8_t* __restrict__ out,
                                                                 unsigned long long out_stride)
{
    __shared__ const uint8_t* s_src[kSRows * kSSeg];   // source of x = 0 of the row's segment (add x); null: generic path
    __shared__ uint8_t s_ky[kSRows];
    const int p = blockIdx.y, layer = blockIdx.z, tid = threadIdx.x;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    const int y0 = blockIdx.x * kSRows;
    const int nrows = min(kSRows, o.ny - y0);
    const int L = o.levels;
    if (tid < nrows * kSSeg) {
        const int r = tid / kSSeg, sgm = tid - r * kSSeg;
        const int y = y0 + r;
        int ky = L + 1;
        for (int k = L; k >= 1; k--) if (y >= o.my[k]) ky = k;
        if (sgm == 0) s_ky[r] = (uint8_t)ky;
        const uint8_t* src = nullptr;
        if (sgm < ky) {
            const int reg = (sgm == 0) ? ky : ky - sgm;
            const int xb = (sgm == 0) ? 0 : o.mx[ky - sgm];
            const int xe = (sgm + 1 < ky) ? o.mx[ky - sgm - 1] : o.nx;
            const int w = order_plane(o, rank, p, reg);
            const unsigned long long j = ((unsigned long long)w * o.ny + y) * o.nx + xb;
            int d = 0;
            while (d + 1 < o.nranks && j >= o.j0[d + 1]) d++;
            if (o.j0[d + 1] - j >= (unsigned long long)(xe - xb)) src = peer.p[d] + loff + (j - o.j0[d]) - xb;
        }
        s_src[tid] = src;
    }
    __syncthreads();
    const int m1 = o.mx[1], m2 = o.mx[2], m3 = o.mx[3], m4 = o.mx[4];
    uint8_t* const obase = out + (unsigned long long)layer * out_stride + ((unsigned long long)p * o.ny + y0) * o.nx;
    const int nitems = nrows * ipr;
    struct Item { int ry, x; bool live, fast; const uint8_t* src; };
    auto locate_item = [&](int it) -> Item {
        Item m{0, 0, it < nitems, false, nullptr};
        if (m.live) {
            m.ry = (int)fdiv((uint32_t)it, dipr);
            m.x = 16 * (it - m.ry * ipr);
            int xe;
            const int sgm = row_segment(m.x, s_ky[m.ry], m1, m2, m3, m4, o.nx, xe);
            const uint8_t* s0 = s_src[m.ry * kSSeg + sgm];
            m.fast = s0 != nullptr && m.x + 16 <= xe;
            m.src = s0 + m.x;
        }
        return m;
    };
    auto finish_item = [&](const Item& m, const Load16& ld) {
        const uint4 v = load16_finish(ld);
        uint8_t* __restrict__ row = obase + (unsigned long long)m.ry * o.nx;
        if (m.fast) { *reinterpret_cast<uint4*>(row + m.x) = v; return; }
        if (!m.live) return;
        const int y = y0 + m.ry;                                      // generic path, symbol by symbol
        for (int e = m.x; e < m.x + 16 && e < o.nx; e++) {
            const int reg = order_region(o, e, y);
            const unsigned long long j = ((unsigned long long)order_plane(o, rank, p, reg) * o.ny + y) * o.nx + e;
            int d = 0;
            while (d + 1 < o.nranks && j >= o.j0[d + 1]) d++;
            row[e] = *(peer.p[d] + loff + (j - o.j0[d]));
        }
    };
    for (int it0 = 0; it0 < nitems; it0 += 512) {                     // warp-uniform trip count: the lanes shuffle; two items in flight
        const Item a = locate_item(it0 + tid), b = locate_item(it0 + 256 + tid);
        const Load16 la = load16_issue(a.src, a.fast), lb = load16_issue(b.src, b.fast);
        finish_item(a, la);
        finish_item(b, lb);
    }
}

void scatter_local_planes(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlay,
                          uint8_t* out, unsigned long long out_stride, cudaStream_t s)
{
    if (nlay <= 0) return;
    const OrderDev o = make_dev(og);
    dim3 grid((og.ny + kSRows - 1) / kSRows, og.nzl, nlay);
    const bool aligned = og.nx % 16 == 0 && (reinterpret_cast<size_t>(out) % 16) == 0 && out_stride % 16 == 0;
    if (og.nx >= 256 && og.levels <= 4 && aligned) {
        const uint32_t ipr = (uint32_t)og.nx / 16u;
        uint32_t l2 = 0;
        while ((1u << l2) < ipr) l2++;                                   // ceil(log2 ipr)
        const FastDiv d{(uint32_t)(((1ull << (31 + l2)) + ipr - 1) / ipr), l2 - 1};
        scatter_local_fast_kernel<<<grid, 256, 0, s>>>(o, d, (int)ipr, rank, peer, peer_stride, out, out_stride);
    } else {
        scatter_local_kernel<<<grid, 256, 0, s>>>(o, rank, peer, peer_stride, out, out_stride);
    }
    note_launch(1);
}

}  // namespace wrb
