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

// four bytes at an arbitrary address through two aligned word loads (the buffers have >= 8 bytes of slack)
__device__ __forceinline__ uint32_t load4_unaligned(const uint8_t* src)
{
    const unsigned long long a = (unsigned long long)src;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~3ull);
    const uint32_t sh = (uint32_t)(a & 3ull) * 8u;
    const uint32_t w0 = wp[0], w1 = wp[1];
    return __funnelshift_r(w0, w1, sh);
}

// ---- encode side -----------------------------------------------------------------------------------------------
// grid (coder blocks of my run, layers): one CTA per coder block (<= 60000 symbols), as the quantiser has it
__global__ void __launch_bounds__(256) gather_run_kernel(OrderDev o, int rank, PeerPtrs peer, unsigned long long peer_stride,
                                                         const int* __restrict__ active, ChunkGeom g,
                                                         uint8_t* __restrict__ sym, unsigned long long lstride,
                                                         uint32_t* __restrict__ hist, unsigned long long hstride)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
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
    // global position of the block's first symbol
    const unsigned long long jb = o.j0[rank] + cstart + boff;
    const unsigned long long wb = jb / o.plane;
    const unsigned int rem = (unsigned int)(jb - wb * o.plane);
    const unsigned int yb = rem / (unsigned int)o.nx, xb = rem - yb * (unsigned int)o.nx;
    uint8_t* __restrict__ out = sym + (unsigned long long)layer * lstride + (unsigned long long)c * g.pitch + boff;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    for (unsigned int q = 4u * tid; q < bs; q += 4u * 256u) {
        const unsigned int t = xb + q;
        const unsigned int dy = t / (unsigned int)o.nx;
        const int x = (int)(t - dy * (unsigned int)o.nx);
        const unsigned int yy = yb + dy;
        const unsigned int dw = yy / (unsigned int)o.ny;
        const int y = (int)(yy - dw * (unsigned int)o.ny);
        const int w = (int)(wb + dw);
        const int reg = order_region(o, x, y);
        if (q + 4 <= bs && x + 3 < o.nx && order_region(o, x + 3, y) == reg) {      // the quad stays in one row and region
            int owner, p;
            order_source(o, w, reg, owner, p);
            const uint8_t* src = peer.p[owner] + loff + ((unsigned long long)p * o.ny + y) * o.nx + x;
            const uint32_t v = load4_unaligned(src);
            *reinterpret_cast<uint32_t*>(out + q) = v;
            atomicAdd(&s_hist[v & 0xFFu], 1u);
            atomicAdd(&s_hist[(v >> 8) & 0xFFu], 1u);
            atomicAdd(&s_hist[(v >> 16) & 0xFFu], 1u);
            atomicAdd(&s_hist[v >> 24], 1u);
        } else {
            for (unsigned int e = 0; e < 4 && q + e < bs; e++) {
                const unsigned int t1 = xb + q + e;
                const unsigned int dy1 = t1 / (unsigned int)o.nx;
                const int x1 = (int)(t1 - dy1 * (unsigned int)o.nx);
                const unsigned int yy1 = yb + dy1;
                const unsigned int dw1 = yy1 / (unsigned int)o.ny;
                const int y1 = (int)(yy1 - dw1 * (unsigned int)o.ny);
                const int w1 = (int)(wb + dw1);
                int owner, p;
                order_source(o, w1, order_region(o, x1, y1), owner, p);
                const uint8_t v = peer.p[owner][loff + ((unsigned long long)p * o.ny + y1) * o.nx + x1];
                out[q + e] = v;
                atomicAdd(&s_hist[v], 1u);
            }
        }
    }
    __syncthreads();
    hist[(unsigned long long)layer * hstride + (unsigned long long)b * 256 + tid] = s_hist[tid];
}

void gather_global_run(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlayers,
                       const int* active, const ChunkGeom& g, uint8_t* sym, unsigned long long sym_layer_stride,
                       uint32_t* hist, unsigned long long hist_layer_stride, cudaStream_t s)
{
    if (g.nblocks == 0 || nlayers <= 0) return;
    const OrderDev o = make_dev(og);
    dim3 grid(g.nblocks, nlayers, 1);
    gather_run_kernel<<<grid, 256, 0, s>>>(o, rank, peer, peer_stride, active, g, sym, sym_layer_stride, hist, hist_layer_stride);
    note_launch(1);
}

// ---- decode side -----------------------------------------------------------------------------------------------
// grid (ny, nzl, layers): one CTA per row (y, local plane p) of one layer
__global__ void __launch_bounds__(128) scatter_local_kernel(OrderDev o, int rank, PeerPtrs peer, unsigned long long peer_stride,
                                                            uint8_t* __restrict__ out, unsigned long long out_stride)
{
    const int y = blockIdx.x, p = blockIdx.y, layer = blockIdx.z;
    const unsigned long long loff = (unsigned long long)layer * peer_stride;
    uint8_t* __restrict__ row = out + (unsigned long long)layer * out_stride + ((unsigned long long)p * o.ny + y) * o.nx;
    const bool vec = (o.nx & 3) == 0 && (((unsigned long long)row) & 3ull) == 0;
    auto source = [&](unsigned long long j) -> const uint8_t* {
        int d = 0;
        while (d + 1 < o.nranks && j >= o.j0[d + 1]) d++;
        return peer.p[d] + loff + (j - o.j0[d]);
    };
    for (int x = 4 * threadIdx.x; x < o.nx; x += 4 * 128) {
        const int reg = order_region(o, x, y);
        const int x3 = (x + 3 < o.nx) ? x + 3 : o.nx - 1;
        bool fast = vec && x + 3 < o.nx && order_region(o, x3, y) == reg;
        if (fast) {
            const int w = order_plane(o, rank, p, reg);
            const unsigned long long j = ((unsigned long long)w * o.ny + y) * o.nx + x;
            int d = 0;
            while (d + 1 < o.nranks && j >= o.j0[d + 1]) d++;
            if (j + 3 < o.j0[d + 1]) {                               // the quad lies in one rank's run
                *reinterpret_cast<uint32_t*>(row + x) = load4_unaligned(peer.p[d] + loff + (j - o.j0[d]));
                continue;
            }
        }
        for (int e = 0; e < 4 && x + e < o.nx; e++) {
            const int w = order_plane(o, rank, p, order_region(o, x + e, y));
            row[x + e] = *source(((unsigned long long)w * o.ny + y) * o.nx + (x + e));
        }
    }
}

void scatter_local_planes(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlay,
                          uint8_t* out, unsigned long long out_stride, cudaStream_t s)
{
    if (nlay <= 0) return;
    const OrderDev o = make_dev(og);
    dim3 grid(og.ny, og.nzl, nlay);
    scatter_local_kernel<<<grid, 128, 0, s>>>(o, rank, peer, peer_stride, out, out_stride);
    note_launch(1);
}

}  // namespace wrb
