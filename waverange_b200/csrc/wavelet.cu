// wavelet.cu -- 3-D CDF 9/7 lifting transform, general-shape kernels.
//
// Replaces waveletcdf97_3d() (reference src/waveletcdf97_3d/waveletcdf97_3d.c:38-468).
// The reference runs each 1-D lifting over whole lines, in place, one line at a time.  Here
// every thread produces R consecutive low/high output pairs of one line by evaluating the
// four lifting stages on a private register window (2R+7 input samples); because each lifting
// stage only combines neighbours, the window reproduces the whole-line result bit for bit as
// long as every + and * is rounded separately (-fmad=false) and the line-end formulas
// (mirror at both ends, extrapolated phantom sample for odd lengths) are applied at the same
// indices.  Passes are out of place: src -> dst, so no line has to be owned by one CTA.
//
// Data flow per level (see codec.cu): x-pass src->A, y-pass A->B, z-pass B->{coefficients,lll};
// the z-pass routes the low-low-low octant to a compact scratch (input of the next level) and
// every final coefficient to the coefficient array, reducing their min/max on the way.
#include <cstdlib>
#include "wr_common.cuh"
#include "wr_kernels.h"
#include "wavelet_pairs.cuh"

namespace wrb {

// ------------------------------------------------------------------------------------------
// Pass kernels.  Thread (tx, ty, tz): for DIM 0 tx indexes groups of R pairs along x; for
// DIM 1/2 tx indexes x (coalesced) and the group index comes from ty / tz.
// ------------------------------------------------------------------------------------------
struct FwdPassArgs {
    const void* src; long long ssy, ssz;     // input box (x stride 1)
    double* dst;     long long dsy, dsz;     // output (array strides)
    double* lll;     long long lsy, lsz;     // DIM 2 only: compact low-low-low scratch (or null)
    int n0, n1, n2;                          // box extents of this level
    int m0, m1;                              // low extents in x, y (for lll routing)
    unsigned long long* in_min;  unsigned long long* in_max;    // reduce over loaded samples (or null)
    unsigned long long* out_min; unsigned long long* out_max;   // reduce over values stored to dst (or null)
};

template <int DIM, class TIN, int R>
__global__ void __launch_bounds__(128) fwd_pass_kernel(FwdPassArgs a)
{
    const TIN* __restrict__ src = (const TIN*)a.src;
    const int nd[3] = {a.n0, a.n1, a.n2};
    const int N = nd[DIM], M = (N + 1) >> 1;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = blockIdx.y * blockDim.y + threadIdx.y;
    const int tz = blockIdx.z;
    int x, y, z, g;
    if (DIM == 0) { g = tx; x = 0; y = ty; z = tz; }
    else if (DIM == 1) { x = tx; g = ty; y = 0; z = tz; }
    else { x = tx; y = ty; g = tz; z = 0; }
    unsigned long long kmin = kKeyMinInit, kmax = kKeyMaxInit, omin = kKeyMinInit, omax = kKeyMaxInit;
    const int i0 = g * R;
    const int ngroups = (N == 1) ? 1 : (M + R - 1) / R;
    bool live = (g < ngroups) && (DIM == 0 || x < a.n0) && (DIM == 1 || y < a.n1) && (DIM == 2 || z < a.n2);
    if (live) {
        const long long sstride = (DIM == 0) ? 1 : (DIM == 1 ? a.ssy : a.ssz);
        const long long dstride = (DIM == 0) ? 1 : (DIM == 1 ? a.dsy : a.dsz);
        const long long sbase = x + (long long)y * a.ssy + (long long)z * a.ssz;
        const long long dbase = x + (long long)y * a.dsy + (long long)z * a.dsz;
        double so[R], dd[R];
        const bool track_in = a.in_min != nullptr;
        auto ld = [&](int j) -> double {
            double v = (double)src[sbase + (long long)j * sstride];
            if (track_in) { unsigned long long k = dkey(v); kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax; }
            return v;
        };
        auto st = [&](int j, double v) {        // j = output index along DIM
            if (DIM == 2 && a.lll != nullptr && x < a.m0 && y < a.m1 && j < M) {
                a.lll[x + (long long)y * a.lsy + (long long)j * a.lsz] = v;
            } else {
                a.dst[dbase + (long long)j * dstride] = v;
                if (a.out_min != nullptr) { unsigned long long k = dkey(v); omin = k < omin ? k : omin; omax = k > omax ? k : omax; }
            }
        };
        if (N == 1) {
            st(0, ld(0));                        // direction skipped (waveletcdf97_3d.c:82,146,210)
        } else {
            fwd_pairs<R>(ld, N, i0, so, dd);
#pragma unroll
            for (int t = 0; t < R; t++) {
                int j = i0 + t;
                if (j < M) {
                    st(j, so[t]);
                    if (2 * j + 1 < N) st(M + j, dd[t]);
                }
            }
        }
    }
    if (a.in_min != nullptr) block_minmax_commit(kmin, kmax, a.in_min, a.in_max);
    if (a.out_min != nullptr) block_minmax_commit(omin, omax, a.out_min, a.out_max);
}

struct InvPassArgs {
    const double* src; long long ssy, ssz;   // input box
    const double* lll; long long lsy, lsz;   // DIM 2 only: low-low-low octant comes from here (or null)
    void* dst;         long long dsy, dsz;
    int n0, n1, n2;                          // box extents of this level
    int q0, q1, q2;                          // low extents (lll routing)
};

template <int DIM, class TOUT, int R>
__global__ void __launch_bounds__(128) inv_pass_kernel(InvPassArgs a)
{
    TOUT* __restrict__ dst = (TOUT*)a.dst;
    const int nd[3] = {a.n0, a.n1, a.n2};
    const int M = nd[DIM], Q = (M + 1) >> 1;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = blockIdx.y * blockDim.y + threadIdx.y;
    const int tz = blockIdx.z;
    int x, y, z, g;
    if (DIM == 0) { g = tx; x = 0; y = ty; z = tz; }
    else if (DIM == 1) { x = tx; g = ty; y = 0; z = tz; }
    else { x = tx; y = ty; g = tz; z = 0; }
    const int i0 = g * R;
    const int ngroups = (M == 1) ? 1 : (Q + R - 1) / R;
    bool live = (g < ngroups) && (DIM == 0 || x < a.n0) && (DIM == 1 || y < a.n1) && (DIM == 2 || z < a.n2);
    if (!live) return;
    const long long sstride = (DIM == 0) ? 1 : (DIM == 1 ? a.ssy : a.ssz);
    const long long dstride = (DIM == 0) ? 1 : (DIM == 1 ? a.dsy : a.dsz);
    const long long sbase = x + (long long)y * a.ssy + (long long)z * a.ssz;
    const long long dbase = x + (long long)y * a.dsy + (long long)z * a.dsz;
    auto ld = [&](int j) -> double {
        if (DIM == 2 && a.lll != nullptr && x < a.q0 && y < a.q1 && j < a.q2)
            return a.lll[x + (long long)y * a.lsy + (long long)j * a.lsz];
        return a.src[sbase + (long long)j * sstride];
    };
    if (M == 1) { dst[dbase] = (TOUT)ld(0); return; }
    double ev[R], od[R];
    inv_pairs<R>(ld, M, i0, ev, od);
#pragma unroll
    for (int t = 0; t < R; t++) {
        int j = i0 + t;
        if (j < Q) {
            dst[dbase + (long long)(2 * j) * dstride] = (TOUT)ev[t];
            if (2 * j + 1 < M) dst[dbase + (long long)(2 * j + 1) * dstride] = (TOUT)od[t];
        }
    }
}

// widen / copy with min-max (wtflag == 0 path: the "transform" is the identity)
template <class TIN>
__global__ void __launch_bounds__(256) widen_minmax_kernel(const TIN* __restrict__ src, double* __restrict__ dst,
                                                           unsigned long long n, unsigned long long* in_min,
                                                           unsigned long long* in_max, unsigned long long* out_min,
                                                           unsigned long long* out_max)
{
    unsigned long long kmin = kKeyMinInit, kmax = kKeyMaxInit;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        double v = (double)src[i];
        dst[i] = v;
        unsigned long long k = dkey(v);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
    block_minmax_commit(kmin, kmax, in_min, in_max);
    if (out_min != nullptr) block_minmax_commit(kmin, kmax, out_min, out_max);
}

template <class TOUT>
__global__ void __launch_bounds__(256) narrow_copy_kernel(const double* __restrict__ src, TOUT* __restrict__ dst,
                                                          unsigned long long n)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = (TOUT)src[i];
}

// ------------------------------------------------------------------------------------------
// Streaming inverse z-lifting.  One thread per (x, y) column of the level box; the pairs (low[m], high[m])
// of its z-line are read once, in order, and the four inverse lifting stages roll through five doubles
// of state (waveletcdf97_3d.c:311-337, same grouping as inv_pairs); every plane access is coalesced in x.
// With FUSE the detail coefficients are not read from the coefficient array at all: they are rebuilt from
// the decoded symbol planes, fld = (0 + (q0*deps0 + min0)) + (q1*deps1 + min1) ... (wrappers.cpp:480,513-514),
// which removes the dequantise pass and 8 B/point of reads.  The low-low-low octant always comes from the
// previous (coarser) level's output.
// ------------------------------------------------------------------------------------------
struct DequantSrc {
    const uint8_t* sym;               // chunk-major padded symbol planes (null: read coefficients)
    unsigned long long layer_stride, chunk_len, pitch;
    int nlay;
    double deps[kNLayMax], minval[kNLayMax];
};

struct InvZArgs {
    const double* coef; long long ay, az;     // coefficient array (array strides)
    const double* lll; long long lsy, lsz;    // previous level's output (compact) or null
    double* dst; long long dsy, dsz;
    int n0, n1, M;                            // level box
    int q0, q1, q2;                           // low extents
    int zpairs;                               // output pairs per z-segment
};

template <bool FUSE>
__global__ void __launch_bounds__(128) inv_z_stream_kernel(InvZArgs a, DequantSrc dq)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.n0) return;
    const int Q = a.q2, NH = a.M - Q;
    const int e0 = blockIdx.z * a.zpairs, e1 = (e0 + a.zpairs < Q) ? e0 + a.zpairs : Q;
    const int mstart = (e0 - 3 > 0) ? e0 - 3 : 0;            // the stencil reaches three pairs back
    const bool in_lll = a.lll != nullptr && x < a.q0 && y < a.q1;
    const long long col = x + (long long)y * a.ay;
    // running (chunk, offset) of the low and the high element in the symbol container
    unsigned long long cl = 0, ol = 0, ch = 0, oh = 0, dc = 0, dof = 0;
    if (FUSE) {
        const unsigned long long jl = (unsigned long long)col + (unsigned long long)mstart * a.az;
        const unsigned long long jh = jl + (unsigned long long)Q * a.az;
        cl = jl / dq.chunk_len; ol = jl % dq.chunk_len;
        ch = jh / dq.chunk_len; oh = jh % dq.chunk_len;
        dc = (unsigned long long)a.az / dq.chunk_len; dof = (unsigned long long)a.az % dq.chunk_len;
    }
    auto deq = [&](unsigned long long c, unsigned long long o) -> double {
        const uint8_t* p = dq.sym + c * dq.pitch + o;
        double f = 0;
        for (int l = 0; l < dq.nlay; l++) f = f + ((double)p[(unsigned long long)l * dq.layer_stride] * dq.deps[l] + dq.minval[l]);
        return f;
    };
    double hp = 0, s1p = 0, d1p = 0, d1pp = 0, s2p = 0;      // h[m-1], s1[m-1], d1[m-1], d1[m-2], s2[m-2]
    double* __restrict__ out = a.dst + x + (long long)y * a.dsy;
    for (int m = mstart; m < e1 + 2 && m <= Q; m++) {
        if (m < Q) {
            double lv, hv = 0.0;
            if (in_lll) lv = a.lll[x + (long long)y * a.lsy + (long long)m * a.lsz];
            else lv = FUSE ? deq(cl, ol) : a.coef[col + (long long)m * a.az];
            if (m < NH) hv = FUSE ? deq(ch, oh) : a.coef[col + (long long)(Q + m) * a.az];      // phantom detail = 0 (:314)
            if (FUSE) {
                cl += dc; ol += dof; if (ol >= dq.chunk_len) { ol -= dq.chunk_len; cl++; }
                ch += dc; oh += dof; if (oh >= dq.chunk_len) { oh -= dq.chunk_len; ch++; }
            }
            const double l = lv * WRB_PSCL, h = hv * WRB_SCL;
            const double s1 = (m == 0) ? l - (WRB_LD * 2) * h : l - WRB_LD * (h + hp);                     // s1[m]
            if (m >= 1) {
                const double d1 = hp - WRB_LC * (s1 + s1p);                                                 // d1[m-1]
                const double s2 = (m == 1) ? s1p - (WRB_LB * 2) * d1 : s1p - WRB_LB * (d1 + d1p);           // s2[m-1]
                if (m >= 2) {
                    const double d2 = d1p - WRB_LA * (s2 + s2p);                                            // d2[m-2]
                    if (m - 2 >= e0 && m - 2 < e1) {
                        out[(long long)(2 * (m - 2)) * a.dsz] = s2p;
                        out[(long long)(2 * (m - 2) + 1) * a.dsz] = d2;
                    }
                }
                d1pp = d1p; d1p = d1; s2p = s2;
                (void)d1pp;
            }
            hp = h; s1p = s1;
        } else {                                             // m == Q: end of the line, flush pairs Q-2 and Q-1
            const double d1 = hp - (WRB_LC * 2) * s1p;                                                      // d1[Q-1]
            const double s2 = (Q == 1) ? s1p - (WRB_LB * 2) * d1 : s1p - WRB_LB * (d1 + d1p);               // s2[Q-1]
            if (Q >= 2) {
                const double d2 = d1p - WRB_LA * (s2 + s2p);                                                // d2[Q-2]
                if (Q - 2 >= e0 && Q - 2 < e1) {
                    out[(long long)(2 * (Q - 2)) * a.dsz] = s2p;
                    out[(long long)(2 * (Q - 2) + 1) * a.dsz] = d2;
                }
            }
            if (Q - 1 >= e0 && Q - 1 < e1) {
                out[(long long)(2 * (Q - 1)) * a.dsz] = s2;
                if (2 * (Q - 1) + 1 < a.M) out[(long long)(2 * (Q - 1) + 1) * a.dsz] = d1 - (WRB_LA * 2) * s2;   // d2[Q-1]
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
constexpr int kR = 4;

static void pass_geometry(int dim, int nx_threads, int n_second, int n_third, dim3& grid, dim3& block)
{
    int bx = 128;
    while (bx > 32 && bx / 2 >= nx_threads) bx >>= 1;
    int by = 128 / bx;
    block = dim3(bx, by, 1);
    grid = dim3((nx_threads + bx - 1) / bx, (n_second + by - 1) / by, n_third);
    (void)dim;
}

template <int DIM, class TIN>
static void launch_fwd_pass(const FwdPassArgs& a, cudaStream_t s)
{
    const int nd[3] = {a.n0, a.n1, a.n2};
    int N = nd[DIM], M = (N + 1) / 2;
    int groups = (N == 1) ? 1 : (M + kR - 1) / kR;
    dim3 grid, block;
    if (DIM == 0) pass_geometry(0, groups, a.n1, a.n2, grid, block);
    else if (DIM == 1) pass_geometry(1, a.n0, groups, a.n2, grid, block);
    else pass_geometry(2, a.n0, a.n1, groups, grid, block);
    fwd_pass_kernel<DIM, TIN, kR><<<grid, block, 0, s>>>(a);
    note_launch(1);
}

template <int DIM, class TOUT>
static void launch_inv_pass(const InvPassArgs& a, cudaStream_t s)
{
    const int nd[3] = {a.n0, a.n1, a.n2};
    int M = nd[DIM], Q = (M + 1) / 2;
    int groups = (M == 1) ? 1 : (Q + kR - 1) / kR;
    dim3 grid, block;
    if (DIM == 0) pass_geometry(0, groups, a.n1, a.n2, grid, block);
    else if (DIM == 1) pass_geometry(1, a.n0, groups, a.n2, grid, block);
    else pass_geometry(2, a.n0, a.n1, groups, grid, block);
    inv_pass_kernel<DIM, TOUT, kR><<<grid, block, 0, s>>>(a);
    note_launch(1);
}

static inline int half_up(int n) { return (n + 1) / 2; }

// Forward transform, `levels` levels.  src: raw field (f32 or f64, array strides nx, nx*ny).
// coef: coefficient array (array strides), tmp: scratch of the same size, lllA/lllB: compact
// scratches for the low-low-low octants (>= ceil(n/2)^3 and ceil(n/4)^3 doubles).
// Field extrema -> st->fmin/fmax keys, coefficient extrema -> st->rmin_key[0]/rmax_key[0].
void wavelet_forward(const void* src, int src_is_f32, double* coef, double* tmp, double* lllA, double* lllB,
                     int nx, int ny, int nz, int levels, DevState* st, cudaStream_t s, HostSource* pipe)
{
    const long long ay = nx, az = (long long)nx * ny;
    unsigned long long* fmin = &st->fmin_key; unsigned long long* fmax = &st->fmax_key;
    unsigned long long* cmin = &st->rmin_key[0]; unsigned long long* cmax = &st->rmax_key[0];
    const bool piecewise = pipe != nullptr && levels >= 1 && fused_forward_supported(nx, ny, nz) && nz / 2 >= 64;
    if (pipe != nullptr && !piecewise) cudaStreamWaitEvent(s, pipe->ev[3], 0);      // cannot follow the copy: wait for all of it
    if (levels == 0) {
        unsigned long long n = (unsigned long long)nx * ny * nz;
        int blocks = (int)((n + 256ull * 8 - 1) / (256ull * 8));
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (src_is_f32) widen_minmax_kernel<float><<<blocks, 256, 0, s>>>((const float*)src, coef, n, fmin, fmax, cmin, cmax);
        else widen_minmax_kernel<double><<<blocks, 256, 0, s>>>((const double*)src, coef, n, fmin, fmax, cmin, cmax);
        note_launch(1);
        return;
    }
    int n0 = nx, n1 = ny, n2 = nz;
    const void* cur = src; long long csy = ay, csz = az;   // level input
    bool cur_f32 = src_is_f32 != 0;
    for (int k = 1; k <= levels; k++) {
        const int m0 = half_up(n0), m1 = half_up(n1), m2 = half_up(n2);
        const bool last = (k == levels);
        double* lll = last ? nullptr : ((k & 1) ? lllA : lllB);
        if (fused_forward_supported(n0, n1, n2)) {          // one HBM round trip for this level
            if (k == 1 && piecewise) {
                // the output pairs [p0, p1) of piece i need input planes 2 p0 - 4 .. 2 p1 + 2: copy pieces 0 .. i+1
                for (int i = 0; i < 4; i++) {
                    const int p0 = (int)((long long)m2 * i / 4), p1 = (int)((long long)m2 * (i + 1) / 4);
                    cudaStreamWaitEvent(s, pipe->ev[i < 3 ? i + 1 : 3], 0);
                    fused_forward_level(cur, cur_f32 ? 1 : 0, csy, csz, coef + (long long)p0 * az, ay, az,
                                        lll ? lll + (long long)p0 * m0 * m1 : nullptr, n0, n1, n2, fmin, fmax, cmin, cmax, s,
                                        0, p0, p1 - p0, m2);
                }
                pipe->used = 1;
            } else
            fused_forward_level(cur, cur_f32 ? 1 : 0, csy, csz, coef, ay, az, lll, n0, n1, n2, (k == 1) ? fmin : nullptr,
                                (k == 1) ? fmax : nullptr, cmin, cmax, s);
            cur = lll; csy = m0; csz = (long long)m0 * m1; cur_f32 = false;
            n0 = m0; n1 = m1; n2 = m2;
            continue;
        }
        FwdPassArgs a{};
        a.n0 = n0; a.n1 = n1; a.n2 = n2; a.m0 = m0; a.m1 = m1;
        // x: cur -> coef (box region used as scratch)
        a.src = cur; a.ssy = csy; a.ssz = csz; a.dst = coef; a.dsy = ay; a.dsz = az;
        a.in_min = (k == 1) ? fmin : nullptr; a.in_max = (k == 1) ? fmax : nullptr;
        if (cur_f32) launch_fwd_pass<0, float>(a, s); else launch_fwd_pass<0, double>(a, s);
        // y: coef -> tmp
        a.in_min = a.in_max = nullptr;
        a.src = coef; a.ssy = ay; a.ssz = az; a.dst = tmp; a.dsy = ay; a.dsz = az;
        launch_fwd_pass<1, double>(a, s);
        // z: tmp -> coef (+ lll)
        a.src = tmp; a.dst = coef; a.lll = lll; a.lsy = m0; a.lsz = (long long)m0 * m1;
        a.out_min = cmin; a.out_max = cmax;
        launch_fwd_pass<2, double>(a, s);
        cur = lll; csy = m0; csz = (long long)m0 * m1; cur_f32 = false;
        n0 = m0; n1 = m1; n2 = m2;
    }
}

// x- then y-lifting of a local box: cur -> scratch (x), scratch -> dst (y).  Used by the slab path, which
// does its own z-lifting across slabs (wavelet_slab.cu).
void wavelet_xy_passes(const void* cur, int cur_is_f32, long long csy, long long csz, double* scratch, long long ay,
                       long long az, double* dst, long long dsy, long long dsz, int n0, int n1, int n2,
                       unsigned long long* in_min, unsigned long long* in_max, cudaStream_t s)
{
    FwdPassArgs a{};
    a.n0 = n0; a.n1 = n1; a.n2 = n2; a.m0 = (n0 + 1) / 2; a.m1 = (n1 + 1) / 2;
    a.src = cur; a.ssy = csy; a.ssz = csz; a.dst = scratch; a.dsy = ay; a.dsz = az;
    a.in_min = in_min; a.in_max = in_max;
    if (cur_is_f32) launch_fwd_pass<0, float>(a, s); else launch_fwd_pass<0, double>(a, s);
    a.in_min = a.in_max = nullptr;
    a.src = scratch; a.ssy = ay; a.ssz = az; a.dst = dst; a.dsy = dsy; a.dsz = dsz;
    launch_fwd_pass<1, double>(a, s);
}

// inverse y- then x-lifting of a local box: src -> scratch (y), scratch -> out (x)
void wavelet_yx_inverse_passes(const double* src, double* scratch, long long ay, long long az, int n0, int n1, int n2,
                               void* out, int out_is_f32, long long osy, long long osz, cudaStream_t s)
{
    InvPassArgs a{};
    a.n0 = n0; a.n1 = n1; a.n2 = n2; a.q0 = (n0 + 1) / 2; a.q1 = (n1 + 1) / 2; a.q2 = (n2 + 1) / 2;
    a.src = src; a.ssy = ay; a.ssz = az; a.dst = scratch; a.dsy = ay; a.dsz = az;
    launch_inv_pass<1, double>(a, s);
    a.src = scratch; a.dst = out; a.dsy = osy; a.dsz = osz;
    if (out_is_f32) launch_inv_pass<0, float>(a, s); else launch_inv_pass<0, double>(a, s);
}

static inline int ceil_shift(int n, int k) { return (int)(((long long)n + (1ll << k) - 1) >> k); }

// Inverse transform.  coef: coefficient array (destroyed), tmp: scratch, out: result (f32/f64,
// array strides).  reference: waveletcdf97_3d.c:281-466 (levels coarsest first, z then y then x)
void wavelet_inverse(double* coef, double* tmp, double* lllA, double* lllB, void* out, int out_is_f32,
                     int nx, int ny, int nz, int levels, cudaStream_t s, const uint8_t* sym,
                     unsigned long long layer_stride, unsigned long long chunk_len, unsigned long long pitch, int nlay,
                     const double* deps, const double* minval, HostSink* sink, double* zb, size_t zb_bytes)
{
    DequantSrc dq{};
    dq.sym = sym; dq.layer_stride = layer_stride; dq.chunk_len = chunk_len; dq.pitch = pitch; dq.nlay = nlay;
    for (int l = 0; l < nlay && l < kNLayMax && sym != nullptr; l++) { dq.deps[l] = deps[l]; dq.minval[l] = minval[l]; }
    const long long ay = nx, az = (long long)nx * ny;
    if (levels == 0) {
        unsigned long long n = (unsigned long long)nx * ny * nz;
        int blocks = (int)((n + 256ull * 8 - 1) / (256ull * 8));
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (out_is_f32) narrow_copy_kernel<float><<<blocks, 256, 0, s>>>(coef, (float*)out, n);
        else narrow_copy_kernel<double><<<blocks, 256, 0, s>>>(coef, (double*)out, n);
        note_launch(1);
        return;
    }
    // one fused kernel per level when every level's box is even and >= 8 (and the symbols, if given, are flat)
    bool fused = getenv("WRB_NO_FUSED_INVERSE") == nullptr && (long long)nx * ny < (1ll << 31) &&
                 (sym == nullptr || pitch == chunk_len);
    for (int k = 0; k < levels && fused; k++)
        fused = fused_inverse_supported(ceil_shift(nx, k), ceil_shift(ny, k), ceil_shift(nz, k));
    if (fused) {
        const double* prev = nullptr;
        // two-pass level (z stream, then TMA-staged y/x tiles) when its scratch is there; else the one-kernel level
        const bool two_pass = zb != nullptr && zb_bytes >= (size_t)nx * ny * 16 + 256 && inverse_two_pass_enabled();
        auto level = [&](const double* lllp, void* dst, int dst_f32, long long dsy, long long dsz, int n0, int n1, int n2,
                         int p0, int p1) {
            if (two_pass)
                inverse_level_two_pass(coef, ay, az, sym, layer_stride, nlay, deps, minval, lllp, dst, dst_f32, dsy, dsz, n0, n1,
                                       n2, zb, zb_bytes, s, p0, p1);
            else
                fused_inverse_level(coef, ay, az, sym, layer_stride, nlay, deps, minval, lllp, dst, dst_f32, dsy, dsz, n0, n1, n2,
                                    s, p0, p1);
        };
        for (int k = levels - 1; k >= 0; k--) {
            const int n0 = ceil_shift(nx, k), n1 = ceil_shift(ny, k), n2 = ceil_shift(nz, k);
            if (k == 0 && sink != nullptr && n2 / 2 >= 64) {
                // Four z-pieces; piece i leaves for the host on the copy stream while piece i+1 is computed (the copy of a
                // 512^3 float field takes ~10 ms, the level ~1 ms: only the first piece's compute stays exposed).
                const int q2 = n2 / 2;
                const size_t esz = out_is_f32 ? 4 : 8;
                for (int i = 0; i < 4; i++) {
                    const int p0 = (int)((long long)q2 * i / 4), p1 = (int)((long long)q2 * (i + 1) / 4);
                    level(prev, out, out_is_f32, ay, az, n0, n1, n2, p0, p1);
                    cudaEventRecord(sink->ev[i], s);
                    cudaStreamWaitEvent(sink->copy, sink->ev[i], 0);
                    const size_t off = (size_t)(2 * p0) * (size_t)az * esz, len = (size_t)(2 * (p1 - p0)) * (size_t)az * esz;
                    cudaMemcpyAsync((char*)sink->host + off, (const char*)out + off, len, cudaMemcpyDeviceToHost, sink->copy);
                }
                sink->used = 1;
            } else if (k == 0) {
                level(prev, out, out_is_f32, ay, az, n0, n1, n2, -1, -1);
            } else {
                double* nxt = (k & 1) ? lllA : lllB;
                level(prev, nxt, 0, n0, (long long)n0 * n1, n0, n1, n2, -1, -1);
                prev = nxt;
            }
        }
        return;
    }
    const double* lll = nullptr; long long lsy = 0, lsz = 0;
    for (int k = levels - 1; k >= 0; k--) {
        const int n0 = ceil_shift(nx, k), n1 = ceil_shift(ny, k), n2 = ceil_shift(nz, k);
        const int q0 = half_up(n0), q1 = half_up(n1), q2 = half_up(n2);
        InvPassArgs a{};
        a.n0 = n0; a.n1 = n1; a.n2 = n2; a.q0 = q0; a.q1 = q1; a.q2 = q2;
        // z: coef or symbols (+lll) -> tmp
        a.src = coef; a.ssy = ay; a.ssz = az; a.lll = lll; a.lsy = lsy; a.lsz = lsz;
        a.dst = tmp; a.dsy = ay; a.dsz = az;
        if (n2 > 1) {
            InvZArgs za{};
            za.coef = coef; za.ay = ay; za.az = az; za.lll = lll; za.lsy = lsy; za.lsz = lsz;
            za.dst = tmp; za.dsy = ay; za.dsz = az; za.n0 = n0; za.n1 = n1; za.M = n2; za.q0 = q0; za.q1 = q1; za.q2 = q2;
            const int bx = (n0 >= 128) ? 128 : ((n0 + 31) / 32) * 32;
            const int gx = (n0 + bx - 1) / bx;
            int zp = q2;                              // z-segments only when the columns alone cannot fill the GPU
            while (zp > 8 && (long long)gx * n1 * ((q2 + zp - 1) / zp) * bx < 148ll * 2048) zp = (zp + 1) / 2;
            za.zpairs = zp;
            dim3 grid(gx, n1, (q2 + zp - 1) / zp), block(bx, 1, 1);
            if (sym != nullptr) inv_z_stream_kernel<true><<<grid, block, 0, s>>>(za, dq);
            else inv_z_stream_kernel<false><<<grid, block, 0, s>>>(za, dq);
            note_launch(1);
        } else {
            launch_inv_pass<2, double>(a, s);       // extent-1 direction: plain copy (needs coefficients)
        }
        // y: tmp -> coef
        a.lll = nullptr;
        a.src = tmp; a.dst = coef;
        launch_inv_pass<1, double>(a, s);
        // x: coef -> next lll (compact) or the output array
        a.src = coef;
        if (k == 0) {
            a.dst = out; a.dsy = ay; a.dsz = az;
            if (out_is_f32) launch_inv_pass<0, float>(a, s); else launch_inv_pass<0, double>(a, s);
        } else {
            double* nxt = (k & 1) ? lllA : lllB;
            a.dst = nxt; a.dsy = n0; a.dsz = (long long)n0 * n1;
            launch_inv_pass<0, double>(a, s);
            lll = nxt; lsy = n0; lsz = (long long)n0 * n1;
        }
    }
}

}  // namespace wrb
