// wavelet_slab.cu -- z-slab partitioned transform (one large field across the GPUs of a box).
//
// Rank g owns the physical planes [z0, z0 + nzl) of the field.  x- and y-lifting are slab-local (the
// general pass kernels of wavelet.cu on the local box); only the z-lifting reaches across slabs:
//   forward : outputs for the owned pairs need 4 planes from below and 3 from above of the
//             (x,y)-transformed level input            (waveletcdf97_3d.c:112-125)
//   inverse : the owned samples need 2 low-band + 2 high-band planes from either side
// Those planes come from the neighbours through the halo callback (NCCL send/recv in production,
// gloo in the CPU tests).  Every rank keeps its coefficients in a *rank-local* wavelet-space array
// laid out exactly like the transform of an independent (nx, ny, nzl) field, so quantiser and coder
// run unchanged on it; the coefficient VALUES are those of the global transform, bit for bit,
// because the kernels below evaluate the same register windows (fwd_pairs / inv_pairs) with global
// line indices and only remap where a plane is stored.
#include <cstdlib>
#include "wr_common.cuh"
#include "wr_kernels.h"
#include "wavelet_pairs.cuh"

namespace wrb {

constexpr int kSR = 2;          // pairs per thread (local pair counts are even)
constexpr int kHaloLo = 4;      // forward: planes needed from below
constexpr int kHaloHi = 3;      // forward: planes needed from above
constexpr int kHaloInv = 2;     // inverse: planes per band and side

struct FwdSlabArgs {
    const double* src; long long ssy, ssz;   // (x,y)-transformed level input incl. halo; local plane p = global z - zoff
    double* dst; long long dsy, dsz;         // rank-local coefficient array
    double* lll; long long lsy, lsz;         // rank-local low-low-low scratch (or null on the last level)
    int n0, n1, N;                           // box extents in x, y; GLOBAL extent in z
    int zoff;                                // global z of local plane 0 of src
    int pair_lo, pair_hi;                    // owned global pairs [pair_lo, pair_hi)
    int m0, m1;                              // low extents in x, y (lll routing)
    unsigned long long* out_min; unsigned long long* out_max;
};

__global__ void __launch_bounds__(128) fwd_zslab_kernel(FwdSlabArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int i0 = a.pair_lo + blockIdx.z * kSR;
    const int M = (a.N + 1) >> 1, nl = a.pair_hi - a.pair_lo;
    unsigned long long omin = kKeyMinInit, omax = kKeyMaxInit;
    if (x < a.n0 && i0 < a.pair_hi) {
        const long long sbase = x + (long long)y * a.ssy;
        auto ld = [&](int j) -> double { return a.src[sbase + (long long)(j - a.zoff) * a.ssz]; };
        double so[kSR], dd[kSR];
        fwd_pairs<kSR>(ld, a.N, i0, so, dd);
#pragma unroll
        for (int t = 0; t < kSR; t++) {
            const int j = i0 + t;
            if (j >= a.pair_hi || j >= M) continue;
            const int pl = j - a.pair_lo;                       // local plane of the low output
            if (a.lll != nullptr && x < a.m0 && y < a.m1) {
                a.lll[x + (long long)y * a.lsy + (long long)pl * a.lsz] = so[t];
            } else {
                a.dst[x + (long long)y * a.dsy + (long long)pl * a.dsz] = so[t];
                const unsigned long long k = dkey(so[t]); omin = k < omin ? k : omin; omax = k > omax ? k : omax;
            }
            if (2 * j + 1 < a.N) {
                a.dst[x + (long long)y * a.dsy + (long long)(nl + pl) * a.dsz] = dd[t];
                const unsigned long long k = dkey(dd[t]); omin = k < omin ? k : omin; omax = k > omax ? k : omax;
            }
        }
    }
    block_minmax_commit(omin, omax, a.out_min, a.out_max);
}

struct InvSlabArgs {
    const double* lowx; const double* highx; long long bsy, bsz;   // band buffers, own planes start at kHaloInv
    double* dst; long long dsy, dsz;                               // local output (interleaved along z)
    int n0, n1, M;                                                 // box extents x, y; GLOBAL extent in z
    int pair_lo, pair_hi;
};

__global__ void __launch_bounds__(128) inv_zslab_kernel(InvSlabArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int i0 = a.pair_lo + blockIdx.z * kSR;
    if (x >= a.n0 || i0 >= a.pair_hi) return;
    const int Q = (a.M + 1) >> 1;
    const long long b = x + (long long)y * a.bsy;
    auto ld = [&](int j) -> double {
        return (j < Q) ? a.lowx[b + (long long)(j - a.pair_lo + kHaloInv) * a.bsz]
                       : a.highx[b + (long long)(j - Q - a.pair_lo + kHaloInv) * a.bsz];
    };
    double ev[kSR], od[kSR];
    inv_pairs<kSR>(ld, a.M, i0, ev, od);
#pragma unroll
    for (int t = 0; t < kSR; t++) {
        const int j = i0 + t;
        if (j >= a.pair_hi || j >= Q) continue;
        const long long o = x + (long long)y * a.dsy + (long long)(2 * (j - a.pair_lo)) * a.dsz;
        a.dst[o] = ev[t];
        if (2 * j + 1 < a.M) a.dst[o + a.dsz] = od[t];
    }
}

// band buffers of one inverse level: lowx[kHaloInv + p] = local box plane p, highx[kHaloInv + p] = plane nl + p;
// the low-low-low octant of the box is the previous (coarser) level's output held in `lll`
__global__ void __launch_bounds__(128) build_bands_kernel(const double* __restrict__ coef, long long ay, long long az,
                                                          const double* __restrict__ lll, long long lsy, long long lsz,
                                                          int q0, int q1, int n0, int n1, int nl,
                                                          double* __restrict__ lowx, double* __restrict__ highx,
                                                          long long bsy, long long bsz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, p = blockIdx.z;                   // p in [0, 2*nl)
    if (x >= n0) return;
    double v;
    if (lll != nullptr && x < q0 && y < q1 && p < nl) v = lll[x + (long long)y * lsy + (long long)p * lsz];
    else v = coef[x + (long long)y * ay + (long long)p * az];
    double* band = (p < nl) ? lowx : highx;
    const int pp = (p < nl) ? p : p - nl;
    band[x + (long long)y * bsy + (long long)(pp + kHaloInv) * bsz] = v;
}

// the same, with the coefficients rebuilt from the symbol planes (fld = (q0*deps0 + min0) + (q1*deps1 + min1) + ...,
// wrappers.cpp:480,513-514): no dequantise pass and no coefficient array in slab decoding
struct BandDequant { const uint8_t* sym; unsigned long long lstride; int nlay; double deps[kNLayMax], minval[kNLayMax]; };

__global__ void __launch_bounds__(128) build_bands_sym_kernel(BandDequant dq, long long ay, long long az,
                                                              const double* __restrict__ lll, long long lsy, long long lsz,
                                                              int q0, int q1, int n0, int n1, int nl,
                                                              double* __restrict__ lowx, double* __restrict__ highx,
                                                              long long bsy, long long bsz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, p = blockIdx.z;                   // p in [0, 2*nl)
    if (x >= n0) return;
    double v;
    if (lll != nullptr && x < q0 && y < q1 && p < nl) {
        v = lll[x + (long long)y * lsy + (long long)p * lsz];
    } else {
        const uint8_t* __restrict__ q = dq.sym + (x + (long long)y * ay + (long long)p * az);
        v = 0.0;
        for (int l = 0; l < dq.nlay; l++) {
            const double t = (double)q[(unsigned long long)l * dq.lstride] * dq.deps[l] + dq.minval[l];
            v = (l == 0) ? t : v + t;
        }
    }
    double* band = (p < nl) ? lowx : highx;
    const int pp = (p < nl) ? p : p - nl;
    band[x + (long long)y * bsy + (long long)(pp + kHaloInv) * bsz] = v;
}

template <class T>
__global__ void __launch_bounds__(256) copy_convert_kernel(const double* __restrict__ src, T* __restrict__ dst, unsigned long long n)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = (T)src[i];
}

static int slab_levels_ok(int nx, int ny, int nz, int nzl, int z0, int levels)
{
    if (levels == 0) return 1;
    if (nzl % (1 << (levels + 1)) != 0 || z0 % (1 << (levels + 1)) != 0) return 0;   // even local pair counts at every level
    if (nz % (1 << levels) != 0) return 0;
    (void)nx; (void)ny;
    return 1;
}

int wavelet_slab_supported(int nx, int ny, int nz, int z0, int nzl, int levels)
{
    return slab_levels_ok(nx, ny, nz, nzl, z0, levels);
}

// Forward, slab mode.  src: this rank's nzl planes (f32/f64, array strides nx, nx*ny).  tmp must hold
// (nzl + 7) * nx * ny doubles.  halo(user, buf, elem_bytes, plane_elems, nplanes_own, lo, hi) fills the lo planes
// before and hi planes after the own planes (which start at plane `lo` of buf) from the z-neighbours.
int wavelet_forward_slab(const void* src, int src_is_f32, double* coef, double* tmp, double* lllA, double* lllB, int nx,
                         int ny, int nz, int z0, int nzl, int levels, DevState* st, const SlabHooks& hk, cudaStream_t s,
                         void* halo1)
{
    const long long ay = nx, az = (long long)nx * ny;
    int n0 = nx, n1 = ny, n2l = nzl, n2g = nz, zg = z0;
    const void* cur = src;            // own planes of the level input
    void* cur_base = nullptr;         // same data inside a buffer with halo room (own planes at +kHaloLo), or null
    long long csy = ay, csz = az;
    bool cur_f32 = src_is_f32 != 0;
    for (int k = 1; k <= levels; k++) {
        const int m0 = (n0 + 1) / 2, m1 = (n1 + 1) / 2;
        const bool last = (k == levels);
        double* lll_base = last ? nullptr : ((k & 1) ? lllA : lllB);            // next level's input, with halo room
        const long long lsz = (long long)m0 * m1;
        double* lll = last ? nullptr : lll_base + kHaloLo * lsz;
        const int esz = cur_f32 ? 4 : 8;
        if (fused_forward_supported(n0, n1, n2g) && n2l >= 2) {
            // one pass: raw level input (+ halo planes from the neighbours) -> octants
            const size_t pb = (size_t)csz * esz;
            const void* hlo = nullptr; const void* hhi = nullptr;
            if (cur_base == nullptr && halo1 != nullptr) {
                // level 1: the caller's slab is read in place; the neighbours' planes arrive in a buffer of their own
                // (4 below, then 3 above) -- no copy of the slab just to gain halo room
                hlo = halo1; hhi = (char*)halo1 + (size_t)kHaloLo * pb;
                if (hk.nranks > 1) {
                    int rc = hk.halo(hk.user, cur, (const char*)cur + (size_t)(n2l - kHaloLo) * pb, halo1, (char*)halo1 + (size_t)kHaloLo * pb,
                                     (unsigned long long)kHaloHi * pb, (unsigned long long)kHaloLo * pb);
                    if (rc) return rc;
                }
                fused_forward_level(cur, cur_f32 ? 1 : 0, csy, csz, coef, ay, az, lll, n0, n1, n2g, &st->fmin_key, &st->fmax_key,
                                    &st->rmin_key[0], &st->rmax_key[0], s, zg - kHaloLo, zg / 2, n2l / 2, -1, hlo, hhi);
            } else {
                if (cur_base == nullptr) {                                     // level 1 without a halo buffer: copy into tmp
                    cur_base = tmp;
                    cudaMemcpyAsync((char*)tmp + (size_t)kHaloLo * pb, cur, (size_t)n2l * pb, cudaMemcpyDeviceToDevice, s);
                }
                if (hk.nranks > 1) {
                    int rc = halo_exchange_contiguous(hk, cur_base, pb, n2l, kHaloLo, kHaloHi);
                    if (rc) return rc;
                }
                fused_forward_level(cur_base, cur_f32 ? 1 : 0, csy, csz, coef, ay, az, lll, n0, n1, n2g,
                                    (k == 1) ? &st->fmin_key : nullptr, (k == 1) ? &st->fmax_key : nullptr, &st->rmin_key[0],
                                    &st->rmax_key[0], s, zg - kHaloLo, zg / 2, n2l / 2);
            }
        } else {
            // x: cur -> coef (local box used as scratch), y: coef -> tmp (compact, own planes after the lower halo)
            const long long tsy = n0, tsz = (long long)n0 * n1;
            wavelet_xy_passes(cur, cur_f32 ? 1 : 0, csy, csz, coef, ay, az, tmp + kHaloLo * tsz, tsy, tsz, n0, n1, n2l,
                              (k == 1) ? &st->fmin_key : nullptr, (k == 1) ? &st->fmax_key : nullptr, s);
            if (hk.nranks > 1) {
                int rc = halo_exchange_contiguous(hk, tmp, (size_t)tsz * 8, n2l, kHaloLo, kHaloHi);
                if (rc) return rc;
            }
            FwdSlabArgs a{};
            a.src = tmp; a.ssy = tsy; a.ssz = tsz; a.dst = coef; a.dsy = ay; a.dsz = az;
            a.lll = lll; a.lsy = m0; a.lsz = lsz;
            a.n0 = n0; a.n1 = n1; a.N = n2g; a.zoff = zg - kHaloLo;
            a.pair_lo = zg / 2; a.pair_hi = (zg + n2l) / 2; a.m0 = m0; a.m1 = m1;
            a.out_min = &st->rmin_key[0]; a.out_max = &st->rmax_key[0];
            dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, (n2l / 2 + kSR - 1) / kSR);
            fwd_zslab_kernel<<<grid, block, 0, s>>>(a);
            note_launch(1);
        }
        cur = lll; cur_base = lll_base; csy = m0; csz = lsz; cur_f32 = false;
        n0 = m0; n1 = m1; n2l /= 2; n2g /= 2; zg /= 2;
    }
    return 0;
}

// Inverse, slab mode.  ext must hold 2 * (nzl/2 + 4) * nx * ny doubles.
int wavelet_inverse_slab_two_pass_ok(int nx, int ny, int nz, int nzl, int levels)
{
    return inverse_two_pass_enabled() && wavelet_inverse_slab_fused_ok(nx, ny, nz, nzl, levels) && (nzl >> levels) >= 2;
}

int wavelet_inverse_slab_fused_ok(int nx, int ny, int nz, int nzl, int levels)
{
    if (levels < 1 || getenv("WRB_NO_FUSED_INVERSE") != nullptr || (long long)nx * ny >= (1ll << 31)) return 0;
    for (int k = 0; k < levels; k++) {
        const int n0 = (nx + (1 << k) - 1) >> k, n1 = (ny + (1 << k) - 1) >> k;
        if (!fused_inverse_supported(n0, n1, nz >> k) || ((nzl >> k) / 2) < 1) return 0;
    }
    return 1;
}

// the boundary coefficients of one level as doubles, for the neighbours: planes 0..3 go down to rank-1 (low[0], low[1],
// high[0], high[1] of my own pairs), planes 4..6 go up to rank+1 (low[last], high[last-1], high[last])
__global__ void __launch_bounds__(128) pack_boundary_kernel(BandDequant dq, long long ay, long long az,
                                                            const double* __restrict__ lll, long long lsy, long long lsz,
                                                            int q0, int q1, int n0, int nl, double* __restrict__ dst, long long plane)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, t = blockIdx.z;                   // t in [0, 7)
    if (x >= n0) return;
    const bool high = (t == 2 || t == 3 || t >= 5);
    const int pr = (t == 0 || t == 2) ? 0 : (t == 1 || t == 3) ? 1 : (t == 5 ? nl - 2 : nl - 1);
    double v;
    if (!high && lll != nullptr && x < q0 && y < q1) {
        v = lll[x + (long long)y * lsy + (long long)pr * lsz];
    } else {
        const uint8_t* __restrict__ q = dq.sym + (x + (long long)y * ay + (long long)(high ? nl + pr : pr) * az);
        v = 0.0;
        for (int l = 0; l < dq.nlay; l++) {
            const double tt = (double)q[(unsigned long long)l * dq.lstride] * dq.deps[l] + dq.minval[l];
            v = (l == 0) ? tt : v + tt;
        }
    }
    dst[x + (long long)y * n0 + (long long)t * plane] = v;
}

int wavelet_inverse_slab(double* coef, double* tmp, double* lllA, double* lllB, double* ext, void* out, int out_is_f32,
                         int nx, int ny, int nz, int z0, int nzl, int levels, const SlabHooks& hk, cudaStream_t s,
                         const uint8_t* sym, unsigned long long lstride, int nlay, const double* deps, const double* minval,
                         double* zb, size_t zb_bytes, double* hb)
{
    const long long ay = nx, az = (long long)nx * ny;
    const double* lll = nullptr; long long lsy = 0, lsz = 0;
    if (sym != nullptr && zb != nullptr && hb != nullptr && zb_bytes >= (size_t)nx * ny * 16 + 256 &&
        wavelet_inverse_slab_two_pass_ok(nx, ny, nz, nzl, levels)) {
        // per level: the boundary coefficients of the rank's own pairs go to the neighbours as doubles (one exchange of
        // 4 + 3 planes), then the two-pass level of wavelet_inv2.cu: streaming z kernel on the own symbol planes (halo
        // pairs from the received planes), TMA-staged y / x tile kernel.  No band buffers, no dequantise pass.
        BandDequant dq{};
        dq.sym = sym; dq.lstride = lstride; dq.nlay = nlay;
        for (int l = 0; l < nlay && l < kNLayMax; l++) { dq.deps[l] = deps[l]; dq.minval[l] = minval[l]; }
        for (int k = levels - 1; k >= 0; k--) {
            const int n0 = (nx + (1 << k) - 1) >> k, n1 = (ny + (1 << k) - 1) >> k;
            const int n2l = nzl >> k, n2g = nz >> k, zg = z0 >> k;
            const int q0 = n0 / 2, q1 = n1 / 2, nl = n2l / 2;
            const long long plane = (long long)n0 * n1;
            double* sendb = hb;                       // planes 0..6
            double* halo = hb + 7 * plane;            // planes 7..13: [3 from below | 4 from above]
            if (hk.nranks > 1) {
                dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, 7);
                pack_boundary_kernel<<<grid, block, 0, s>>>(dq, ay, az, lll, lsy, lsz, q0, q1, n0, nl, sendb, plane);
                note_launch(1);
                int rc = hk.halo(hk.user, sendb, sendb + 4 * plane, halo, halo + 3 * plane, 4ull * plane * 8, 3ull * plane * 8);
                if (rc) return rc;
            }
            double* nxt = (k & 1) ? lllA : lllB;
            inverse_level_two_pass(nullptr, ay, az, sym, lstride, nlay, deps, minval, lll, (k == 0) ? out : (void*)nxt,
                                   (k == 0) ? out_is_f32 : 0, (k == 0) ? ay : (long long)n0, (k == 0) ? az : (long long)n0 * n1,
                                   n0, n1, n2l, zb, zb_bytes, s, -1, -1, halo, n2g, zg / 2);
            lll = nxt; lsy = n0; lsz = (long long)n0 * n1;
        }
        return 0;
    }
    if (wavelet_inverse_slab_fused_ok(nx, ny, nz, nzl, levels)) {
        // per level: band buffers (own pairs + halo room) from the symbols or the coefficient array, two halo
        // exchanges, then ONE kernel for the z, y and x inverse lifting (wavelet_inv_fused.cu, band mode)
        BandDequant dq{};
        dq.sym = sym; dq.lstride = lstride; dq.nlay = nlay;
        for (int l = 0; l < nlay && l < kNLayMax && sym != nullptr; l++) { dq.deps[l] = deps[l]; dq.minval[l] = minval[l]; }
        for (int k = levels - 1; k >= 0; k--) {
            const int n0 = (nx + (1 << k) - 1) >> k, n1 = (ny + (1 << k) - 1) >> k;
            const int n2l = nzl >> k, n2g = nz >> k, zg = z0 >> k;
            const int q0 = n0 / 2, q1 = n1 / 2, nl = n2l / 2;
            const long long bsy = n0, bsz = (long long)n0 * n1;
            double* lowx = ext;
            double* highx = ext + (long long)(nl + 2 * kHaloInv) * bsz;
            dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, 2 * nl);
            if (sym != nullptr) build_bands_sym_kernel<<<grid, block, 0, s>>>(dq, ay, az, lll, lsy, lsz, q0, q1, n0, n1, nl, lowx, highx, bsy, bsz);
            else build_bands_kernel<<<grid, block, 0, s>>>(coef, ay, az, lll, lsy, lsz, q0, q1, n0, n1, nl, lowx, highx, bsy, bsz);
            note_launch(1);
            if (hk.nranks > 1) {
                int rc = halo_exchange_contiguous(hk, lowx, (size_t)bsz * 8, nl, kHaloInv, kHaloInv);
                if (rc) return rc;
                rc = halo_exchange_contiguous(hk, highx, (size_t)bsz * 8, nl, kHaloInv, kHaloInv);
                if (rc) return rc;
            }
            double* nxt = (k & 1) ? lllA : lllB;
            fused_inverse_level_bands(lowx, highx, bsy, bsz, kHaloInv, zg / 2, nl, (k == 0) ? out : (void*)nxt,
                                      (k == 0) ? out_is_f32 : 0, (k == 0) ? ay : (long long)n0,
                                      (k == 0) ? az : (long long)n0 * n1, n0, n1, n2g, s);
            lll = nxt; lsy = n0; lsz = (long long)n0 * n1;
        }
        return 0;
    }
    for (int k = levels - 1; k >= 0; k--) {
        const int n0 = (nx + (1 << k) - 1) >> k, n1 = (ny + (1 << k) - 1) >> k;
        const int n2l = nzl >> k, n2g = nz >> k, zg = z0 >> k;
        const int q0 = (n0 + 1) / 2, q1 = (n1 + 1) / 2, nl = n2l / 2;
        const long long bsy = n0, bsz = (long long)n0 * n1;
        double* lowx = ext;
        double* highx = ext + (long long)(nl + 2 * kHaloInv) * bsz;
        {
            dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, 2 * nl);
            build_bands_kernel<<<grid, block, 0, s>>>(coef, ay, az, lll, lsy, lsz, q0, q1, n0, n1, nl, lowx, highx, bsy, bsz);
            note_launch(1);
        }
        if (hk.nranks > 1) {
            int rc = halo_exchange_contiguous(hk, lowx, (size_t)bsz * 8, nl, kHaloInv, kHaloInv);
            if (rc) return rc;
            rc = halo_exchange_contiguous(hk, highx, (size_t)bsz * 8, nl, kHaloInv, kHaloInv);
            if (rc) return rc;
        }
        InvSlabArgs a{};
        a.lowx = lowx; a.highx = highx; a.bsy = bsy; a.bsz = bsz;
        a.dst = tmp; a.dsy = ay; a.dsz = az; a.n0 = n0; a.n1 = n1; a.M = n2g;
        a.pair_lo = zg / 2; a.pair_hi = (zg + n2l) / 2;
        dim3 block(128, 1, 1), grid((n0 + 127) / 128, n1, (nl + kSR - 1) / kSR);
        inv_zslab_kernel<<<grid, block, 0, s>>>(a);
        note_launch(1);
        // y: tmp -> coef (local box), x: coef -> next lll (compact) or the output slab
        double* nxt = (k & 1) ? lllA : lllB;
        wavelet_yx_inverse_passes(tmp, coef, ay, az, n0, n1, n2l, (k == 0) ? out : (void*)nxt, (k == 0) ? out_is_f32 : 0,
                                  (k == 0) ? ay : (long long)n0, (k == 0) ? az : (long long)n0 * n1, s);
        lll = nxt; lsy = n0; lsz = (long long)n0 * n1;
    }
    if (levels == 0) {
        const unsigned long long n = (unsigned long long)nx * ny * nzl;
        int blocks = (int)((n + 2047) / 2048); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
        if (out_is_f32) copy_convert_kernel<float><<<blocks, 256, 0, s>>>(coef, (float*)out, n);
        else copy_convert_kernel<double><<<blocks, 256, 0, s>>>(coef, (double*)out, n);
        note_launch(1);
    }
    return 0;
}

}  // namespace wrb
