// wavelet_inv2.cu -- two-pass inverse CDF 9/7 level: a streaming z pass, then a TMA-staged tile pass for y and x.
//
// Same arithmetic, operation for operation, as the inverse line passes of wavelet.cu and the one-kernel version in
// wavelet_inv_fused.cu (reference waveletcdf97_3d.c:281-466: per level z, then y, then x; un-scale :312-313, four
// inverse lifting stages :317-330, interleave :333-337; accumulate loop of decoding_wrap(), wrappers.cpp:480,513-514).
//
// Why two kernels.  The one-kernel level keeps the z pipeline of every (x, y) position of its tile PLUS HALO in
// registers and rebuilds the coefficients from the symbol planes there, so the dequantiser and the z stage run on 1.6x
// the positions, and its y / x phases leave half the threads idle between barriers: 0.09 of the HBM roofline.  Here
//   * inv_z_kernel   streams along z with NO halo: a thread owns four consecutive x positions of one row, reads the
//                    symbol planes as aligned 32-bit words (or the coefficients / the previous level's output as
//                    doubles), rebuilds the coefficients, runs the rolling z pipeline in registers and writes the
//                    z-inverted values, still in coefficient layout in x and y, to a slab buffer of a few planes;
//   * inv_yx_kernel  takes (x, y) tiles of that slab: the four boxes a tile needs (low / high columns x low / high rows,
//                    halo included) arrive in shared memory through the TENSOR MEMORY ACCELERATOR (cp.async.bulk.tensor
//                    + mbarrier, double-buffered: the next two planes fly while these are lifted), the y inverse runs
//                    on columns of the tile, the x inverse on rows, and 32-byte runs of samples leave for the output.
// The slab buffer holds 2 * sp planes of the level's box (the whole level where that is <= ~1.1 GB: slabs small enough
// to stay in the 126 MB L2 between the two passes were measured slower -- 13 launch pairs instead of one at 512^3 --
// see inverse_two_pass_scratch_bytes).  Border tiles: the TMA fills out-of-range elements with zeros (and in-range ones with the
// neighbouring sub-band); the symmetric extension is then written over them in shared memory (fix_halo), which is
// what the index mirroring of the other kernels does (low: s[-k] = s[k], s[Q-1+k] = s[Q-k]; high: d[-k] = d[k-1],
// d[Q-1+k] = d[Q-1-k]).  Requires even box extents >= 8 (fused_inverse_supported); segments restart the z pipeline two
// pairs early, so any slab size gives bit-identical results.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <cuda.h>
#include "wr_common.cuh"
#include "wr_kernels.h"

namespace wrb {

// ------------------------------------------------------------------------------------------------------------------
// z pass
// ------------------------------------------------------------------------------------------------------------------
struct InvZArgs2 {
    const double* coef; long long ay, az;       // coefficient array (array strides), used when sym == null
    const uint8_t* sym;                         // flat symbol planes, layer l at sym + l*lstride (or null)
    unsigned long long lstride;
    double deps[kNLayMax], minval[kNLayMax];
    const double* lll; long long lsy, lsz;      // previous level's output (compact q0 x q1 x q2) or null
    double* zb;                                 // slab buffer: local plane p at zb + p*n0*n1, row pitch n0
    int n0, n1, n2;                             // box extents (even)
    int pair0, npairs;                          // the slab: output pairs [pair0, pair0 + npairs)
    int zpairs;                                 // output pairs per z-segment (blockIdx.z)
    // z-slab mode (SLAB): n2 is the GLOBAL extent of the level's box; this rank owns the pairs [own0, own0 + nown) and
    // its symbol / coefficient / lll planes are rank-local (z-low pair p at local plane p - own0, z-high pair p at
    // nown + p - own0).  The neighbours' boundary coefficients arrive as doubles in `halo`, seven planes of n0 x n1:
    // low[own0-1] | high[own0-2], high[own0-1] | low[own0+nown], low[own0+nown+1] | high[own0+nown], high[own0+nown+1]
    const double* halo; int own0, nown;
};

__device__ __forceinline__ int mirror_s2(int i, int Q)
{
    i = (i < 0) ? -i : i;
    i = (i >= Q) ? 2 * Q - 1 - i : i;
    return min(max(i, 0), Q - 1);
}
__device__ __forceinline__ int mirror_d2(int i, int Q)
{
    i = (i < 0) ? -i - 1 : i;
    i = (i >= Q) ? 2 * Q - 2 - i : i;
    return min(max(i, 0), Q - 1);
}

// VX consecutive x positions per thread (4: aligned word loads of the symbols, 16-byte loads / stores of doubles;
// 1: any shape and alignment).  NLAY > 0: coefficients rebuilt from NLAY symbol planes; 0: read from coef.
template <int NLAY, int VX, bool SLAB>
__global__ void __launch_bounds__(128) inv_z_kernel(InvZArgs2 a)
{
    constexpr bool FROM_SYM = NLAY > 0;
    constexpr int NRAW = FROM_SYM ? NLAY : 1;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * VX;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= a.n0 || y >= a.n1) return;
    const int q0 = a.n0 >> 1, q1 = a.n1 >> 1, q2 = a.n2 >> 1;
    const int e0 = a.pair0 + blockIdx.z * a.zpairs;
    const int e1 = min(e0 + a.zpairs, a.pair0 + a.npairs);
    if (e0 >= e1) return;
    const long long pos = x0 + (long long)y * a.ay;             // offset inside a plane of the coefficient / symbol array
    const bool inl = a.lll != nullptr && x0 < q0 && y < q1;      // VX == 4: q0 % 4 == 0, the quad does not straddle q0
    const long long lpos = x0 + (long long)y * a.lsy;
    const long long zplane = (long long)a.n0 * a.n1;
    double* __restrict__ zrow = a.zb + x0 + (long long)y * a.n0;

    double hp[VX], s1p[VX], d1p[VX], s2p[VX];
#pragma unroll
    for (int v = 0; v < VX; v++) { hp[v] = 0; s1p[v] = 0; d1p[v] = 0; s2p[v] = 0; }

    // raw inputs of one z step, loaded one step ahead of their use
    unsigned int qlo[NRAW], qhi[NRAW];            // VX symbols per word (VX == 1: one symbol)
    double clo[VX], chi[VX];
    bool lo_halo = false, hi_halo = false;        // SLAB: this step's low / high coefficients came from the halo planes (doubles)
    auto load4d = [&](const double* __restrict__ p, double (&dst)[VX]) {
        if (VX == 4) {
            const double2 a0 = *reinterpret_cast<const double2*>(p), a1 = *(reinterpret_cast<const double2*>(p) + 1);
            dst[0] = a0.x; dst[1] = a0.y; dst[VX > 2 ? 2 : 0] = a1.x; dst[VX > 3 ? 3 : 0] = a1.y;
        } else dst[0] = p[0];
    };
    auto load_step = [&](int m) {
        long long pl = mirror_s2(m, q2), ph = (long long)q2 + mirror_d2(m, q2);
        bool lh = false, hh = false;
        if (SLAB) {
            const int gl = (int)pl, gh = (int)(ph - q2);
            const long long hpos = x0 + (long long)y * a.n0;
            if (gl < a.own0) { lh = true; load4d(a.halo + hpos, clo); }                                     // low[own0-1] (also stands in for own0-2: feeds no output)
            else if (gl >= a.own0 + a.nown) { lh = true; load4d(a.halo + hpos + (long long)(3 + min(gl - a.own0 - a.nown, 1)) * zplane, clo); }
            else pl = gl - a.own0;
            if (gh < a.own0) { hh = true; load4d(a.halo + hpos + (long long)(1 + max(gh - (a.own0 - 2), 0)) * zplane, chi); }
            else if (gh >= a.own0 + a.nown) { hh = true; load4d(a.halo + hpos + (long long)(5 + min(gh - a.own0 - a.nown, 1)) * zplane, chi); }
            else ph = (long long)a.nown + gh - a.own0;
        }
        lo_halo = lh; hi_halo = hh;
        if (FROM_SYM) {
#pragma unroll
            for (int l = 0; l < NRAW; l++) {
                const uint8_t* __restrict__ sl = a.sym + (unsigned long long)l * a.lstride + pos;
                if (VX == 4) {
                    if (!hh) qhi[l] = __ldcs(reinterpret_cast<const unsigned int*>(sl + ph * a.az));
                    if (!inl && !lh) qlo[l] = __ldcs(reinterpret_cast<const unsigned int*>(sl + pl * a.az));
                } else {
                    if (!hh) qhi[l] = sl[ph * a.az];
                    if (!inl && !lh) qlo[l] = sl[pl * a.az];
                }
            }
        } else if (!hh) {
            const double* __restrict__ ch = a.coef + pos + ph * a.az;
            if (VX == 4) {
                const double2 h0 = __ldcs(reinterpret_cast<const double2*>(ch)), h1 = __ldcs(reinterpret_cast<const double2*>(ch) + 1);
                chi[0] = h0.x; chi[1] = h0.y; chi[VX > 2 ? 2 : 0] = h1.x; chi[VX > 3 ? 3 : 0] = h1.y;
            } else chi[0] = ch[0];
        }
        if (lh) return;
        if (inl) {
            load4d(a.lll + lpos + pl * a.lsz, clo);
        } else if (!FROM_SYM) {
            const double* __restrict__ cl = a.coef + pos + pl * a.az;
            if (VX == 4) {
                const double2 l0 = __ldcs(reinterpret_cast<const double2*>(cl)), l1 = __ldcs(reinterpret_cast<const double2*>(cl) + 1);
                clo[0] = l0.x; clo[1] = l0.y; clo[VX > 2 ? 2 : 0] = l1.x; clo[VX > 3 ? 3 : 0] = l1.y;
            } else clo[0] = cl[0];
        }
    };
    // fld = (q0*deps0 + min0) + (q1*deps1 + min1) + ...   (wrappers.cpp:480,513-514; 0 + t == t: t is never -0)
    auto deq = [&](const unsigned int (&q)[NRAW], int v) -> double {
        double f = 0.0;
#pragma unroll
        for (int l = 0; l < NRAW; l++) {
            const unsigned int b = (VX == 4) ? ((q[l] >> (8 * v)) & 0xFFu) : q[l];
            const double qd = __hiloint2double(0x43300000, (int)b) - 4503599627370496.0;       // (double)b, no conversion
            const double t = qd * a.deps[l] + a.minval[l];
            f = (l == 0) ? t : f + t;
        }
        return f;
    };

    // Pairs e0-2 .. e1+1 are fed: output pair i needs low[i-1..i+2] and high[i-2..i+2]; feeding pair m completes pair m-2.
    load_step(e0 - 2);
    for (int m = e0 - 2; m <= e1 + 1; m++) {
        double lv[VX], hv[VX];
#pragma unroll
        for (int v = 0; v < VX; v++) {
            if (FROM_SYM) {
                hv[v] = (SLAB && hi_halo) ? chi[v] : deq(qhi, v);
                lv[v] = (inl || (SLAB && lo_halo)) ? clo[v] : deq(qlo, v);
            } else { hv[v] = chi[v]; lv[v] = clo[v]; }
        }
        if (m <= e1) load_step(m + 1);
        const bool out = (m - 2 >= e0);
        double ev[VX], od[VX];
#pragma unroll
        for (int v = 0; v < VX; v++) {
            const double l = lv[v] * WRB_PSCL, h = hv[v] * WRB_SCL;
            const double s1 = l - WRB_LD * (h + hp[v]);                   // s1[m]
            const double d1 = hp[v] - WRB_LC * (s1 + s1p[v]);             // d1[m-1]
            const double s2 = s1p[v] - WRB_LB * (d1 + d1p[v]);            // s2[m-1]
            const double d2 = d1p[v] - WRB_LA * (s2 + s2p[v]);            // d2[m-2]
            ev[v] = s2p[v]; od[v] = d2;                                   // samples 2(m-2), 2(m-2)+1
            hp[v] = h; s1p[v] = s1; d1p[v] = d1; s2p[v] = s2;
        }
        if (out) {
            double* __restrict__ o = zrow + (long long)(2 * (m - 2 - a.pair0)) * zplane;
            if (VX == 4) {
                reinterpret_cast<double2*>(o)[0] = make_double2(ev[0], ev[1]);
                reinterpret_cast<double2*>(o)[1] = make_double2(ev[VX > 2 ? 2 : 0], ev[VX > 3 ? 3 : 0]);
                reinterpret_cast<double2*>(o + zplane)[0] = make_double2(od[0], od[1]);
                reinterpret_cast<double2*>(o + zplane)[1] = make_double2(od[VX > 2 ? 2 : 0], od[VX > 3 ? 3 : 0]);
            } else {
                o[0] = ev[0];
                o[zplane] = od[0];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// y / x pass
// ------------------------------------------------------------------------------------------------------------------
constexpr int YPX = 32;                       // output pairs per tile in x
constexpr int YPY = 8;                        // output pairs per tile in y
constexpr int YR = 4;                         // pairs per y / x lifting task
constexpr int YLX = YPX + 3, YHX = YPX + 4;   // low / high coefficient columns a tile needs (35, 36)
constexpr int YLY = YPY + 3, YHY = YPY + 4;   // low / high coefficient rows (11, 12)
constexpr int YBX = 36, YBY = 12;             // the TMA box: 36 x 12 doubles (288-byte rows), one box per sub-band quadrant
constexpr int YCX = YLX + YHX;                // 71 columns of z-inverted coefficients per tile row
constexpr int YBOX = YBX * YBY;               // doubles per box
constexpr int YPLANE = 4 * YBOX;              // doubles per plane of the tile: boxes LL, HL (high x), LH (high y), HH
constexpr int YTP = YCX + 2;                  // row pitch of the y-stage tile (73: odd)
constexpr int YROWS = 2 * YPY;                // 16 sample rows after the y inverse
constexpr int YTHREADS = 288;                 // 284 y tasks, 256 x tasks per round of two planes
constexpr int YSTAGES = 2;
constexpr int YSMEM = (YSTAGES * 2 * YPLANE + 2 * YROWS * YTP) * 8 + 64;

struct InvYXArgs {
    const double* zb;                         // slab buffer (n0 x n1 x nplanes doubles)
    int n0, n1, nplanes;
    void* dst; long long dsy, dsz;            // this level's output (x stride 1)
    long long out_plane0;                     // output plane of local plane 0
    int vec_ok;                               // output rows are 16-byte aligned: vector stores allowed
    int planes_per_cta;                       // even
};

template <class LDL, class LDH>
__device__ __forceinline__ void inv_window2(LDL ldl, LDH ldh, double (&ev)[YR], double (&od)[YR])
{
    double h[YR + 4], l[YR + 3];
#pragma unroll
    for (int t = 0; t < YR + 4; t++) h[t] = ldh(t) * WRB_SCL;
#pragma unroll
    for (int t = 0; t < YR + 3; t++) l[t] = ldl(t) * WRB_PSCL;
    double s1[YR + 3], d1[YR + 2], s2[YR + 1];
#pragma unroll
    for (int t = 0; t < YR + 3; t++) s1[t] = l[t] - WRB_LD * (h[t + 1] + h[t]);
#pragma unroll
    for (int t = 0; t < YR + 2; t++) d1[t] = h[t + 1] - WRB_LC * (s1[t + 1] + s1[t]);
#pragma unroll
    for (int t = 0; t < YR + 1; t++) s2[t] = s1[t + 1] - WRB_LB * (d1[t + 1] + d1[t]);
#pragma unroll
    for (int t = 0; t < YR; t++) {
        ev[t] = s2[t];
        od[t] = d1[t + 1] - WRB_LA * (s2[t + 1] + s2[t]);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Tile layout in shared memory, per plane: four dense boxes of YBY rows x YBX columns,
//   box 0: low  y rows (global low row  py0-1+r), low  x columns (global low column  px0-2+c: one column more than
//          needed on the left -- the TMA wants the first element of a box row 16-byte aligned in global memory, so
//          boxes of doubles start at even x; measured on this driver: an odd start raises "illegal instruction")
//   box 1: low  y rows,                           high x columns (global high column px0-2+c)
//   box 2: high y rows (global high row py0-2+r), low  x columns
//   box 3: high y rows,                           high x columns
// TMA == true: the boxes are fetched by cp.async.bulk.tensor from the 3-D tensor (x, y, plane) of the slab buffer;
// TMA == false: the same layout is filled by ordinary loads through mirrored indices (A/B and fallback).
template <class TOUT, bool TMA>
__global__ void __launch_bounds__(YTHREADS, 2) inv_yx_kernel(const __grid_constant__ CUtensorMap tmap, InvYXArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tz = reinterpret_cast<double*>(smem_raw);                       // [YSTAGES][2 planes][YPLANE]
    double* ty = tz + YSTAGES * 2 * YPLANE;                                 // [2 planes][YROWS][YTP]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(ty + 2 * YROWS * YTP);   // [YSTAGES]
    const int tid = threadIdx.x;
    const int q0 = a.n0 >> 1, q1 = a.n1 >> 1;
    const int px0 = blockIdx.x * YPX, py0 = blockIdx.y * YPY;
    const int p_begin = blockIdx.z * a.planes_per_cta;
    const int p_end = min(p_begin + a.planes_per_cta, a.nplanes);
    const int nrounds = (p_end - p_begin + 1) >> 1;
    // does the tile touch a line end in x or y?  (then the symmetric extension is written over what the TMA fetched)
    const bool border = (px0 - 2 < 0) || (px0 + YPX + 1 >= q0) || (py0 - 2 < 0) || (py0 + YPY + 1 >= q1);

    if (TMA) {
        if (tid == 0) {
            for (int s = 0; s < YSTAGES; s++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar[s])), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // TMA == true: one thread arms the stage's barrier with the byte count and issues the four box copies of each plane
    auto issue = [&](int round) {
        const int st = round % YSTAGES;
        const int p = p_begin + 2 * round;
        const int np = min(2, p_end - p);
        const uint32_t bar = smem_u32(&mbar[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(np * YPLANE * 8) : "memory");
        for (int i = 0; i < np; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int cx = (b & 1) ? q0 + px0 - 2 : px0 - 2;             // even: see the layout note above
                const int cy = (b & 2) ? q1 + py0 - 2 : py0 - 1;
                const uint32_t dst = smem_u32(tz + (st * 2 + i) * YPLANE + b * YBOX);
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                    :: "r"(dst), "l"(&tmap), "r"(cx), "r"(cy), "r"(p + i), "r"(bar) : "memory");
            }
        }
    };
    // TMA == false: every thread fetches its share of the boxes through mirrored indices
    auto fill = [&](int round) {
        const int st = round % YSTAGES;
        const int p = p_begin + 2 * round;
        const int np = min(2, p_end - p);
        for (int idx = tid; idx < np * YPLANE; idx += YTHREADS) {
            const int i = idx / YPLANE, rem = idx - i * YPLANE;
            const int b = rem / YBOX, e = rem - b * YBOX;
            const int r = e / YBX, c = e - r * YBX;
            const int gx = (b & 1) ? q0 + mirror_d2(px0 - 2 + c, q0) : mirror_s2(px0 - 2 + c, q0);
            const int gy = (b & 2) ? q1 + mirror_d2(py0 - 2 + r, q1) : mirror_s2(py0 - 1 + r, q1);
            tz[(st * 2 + i) * YPLANE + rem] = a.zb[gx + (long long)gy * a.n0 + (long long)(p + i) * a.n0 * a.n1];
        }
    };
    // Symmetric extension written over the out-of-range columns, then rows, of the boxes of one stage (border tiles, TMA
    // path).  Only the two positions next to either line end can reach an unmasked output (output pair i reads low
    // [i-1, i+2] and high [i-2, i+2]): global indices -2, -1, Q, Q+1 of a band -- at most 4 columns and 4 rows per box.
    auto fix_halo = [&](int st, int np) {
        for (int idx = tid; idx < np * 4 * YBY * 4; idx += YTHREADS) {        // columns: (plane, box, row, candidate)
            const int k = idx & 3, r = (idx >> 2) % YBY, pb = (idx >> 2) / YBY;      // pb = plane * 4 + box
            const int b = pb & 3;
            const int g = (k < 2) ? k - 2 : q0 + (k - 2);                        // -2, -1, q0, q0 + 1
            const int c = g - (px0 - 2);
            if (c >= 0 && c < YBX) {
                const int m = (b & 1) ? mirror_d2(g, q0) : mirror_s2(g, q0);
                const int cs = min(max(m - (px0 - 2), 0), YBX - 1);
                double* box = tz + (st * 2) * YPLANE + pb * YBOX;
                box[r * YBX + c] = box[r * YBX + cs];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < np * 4 * 4 * YBX; idx += YTHREADS) {        // rows: (plane, box, candidate, column)
            const int c = idx % YBX, k = (idx / YBX) & 3, pb = idx / (4 * YBX);
            const int b = pb & 3;
            const int g = (k < 2) ? k - 2 : q1 + (k - 2);
            const int base = (b & 2) ? py0 - 2 : py0 - 1;
            const int r = g - base;
            if (r >= 0 && r < YBY) {
                const int m = (b & 2) ? mirror_d2(g, q1) : mirror_s2(g, q1);
                const int rs = min(max(m - base, 0), YBY - 1);
                double* box = tz + (st * 2) * YPLANE + pb * YBOX;
                box[r * YBX + c] = box[rs * YBX + c];
            }
        }
        __syncthreads();
    };

    if (TMA) { if (tid == 0) issue(0); }
    else { fill(0); }
    for (int round = 0; round < nrounds; round++) {
        const int st = round % YSTAGES;
        const int p = p_begin + 2 * round;
        const int np = min(2, p_end - p);
        if (TMA) {
            // the other stage was last read by the y tasks of the previous round: every thread has passed the barrier
            // that followed them, so it may be overwritten now
            if (tid == 0 && round + 1 < nrounds) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(round + 1);
            }
            const uint32_t bar = smem_u32(&mbar[st]);
            const uint32_t parity = (uint32_t)((round / YSTAGES) & 1);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
            if (border) fix_halo(st, np);
        } else {
            __syncthreads();                                              // this round's boxes are in place
            if (round + 1 < nrounds) fill(round + 1);                     // (the other stage: free since the previous round's barrier)
        }
        const double* tzs = tz + st * 2 * YPLANE;
        // ---- y inverse: task = (plane, group of YR pairs, column); consecutive lanes on consecutive columns ----
        for (int t = tid; t < np * (YPY / YR) * YCX; t += YTHREADS) {
            const int pl = t / ((YPY / YR) * YCX), rem = t - pl * ((YPY / YR) * YCX);
            const int g = rem / YCX, c = rem - g * YCX;
            const bool hx = c >= YLX;
            const double* lo = tzs + pl * YPLANE + (hx ? YBOX : 0) + (hx ? c - YLX : c + 1);       // box 0 / 1 (low box: one spare column)
            const double* hi = lo + 2 * YBOX;                                                      // box 2 / 3
            auto ldl = [&](int j) -> double { return lo[(g * YR + j) * YBX]; };
            auto ldh = [&](int j) -> double { return hi[(g * YR + j) * YBX]; };
            double ev[YR], od[YR];
            inv_window2(ldl, ldh, ev, od);
            double* o = ty + pl * (YROWS * YTP) + (2 * g * YR) * YTP + c;
#pragma unroll
            for (int j = 0; j < YR; j++) { o[(2 * j) * YTP] = ev[j]; o[(2 * j + 1) * YTP] = od[j]; }
        }
        __syncthreads();
        // ---- x inverse: task = (plane, group of YR pairs, row); consecutive lanes on consecutive rows ----
        for (int t = tid; t < np * (YPX / YR) * YROWS; t += YTHREADS) {
            const int pl = t / ((YPX / YR) * YROWS), rem = t - pl * ((YPX / YR) * YROWS);
            const int g = rem / YROWS, r = rem - g * YROWS;
            const double* row = ty + pl * (YROWS * YTP) + r * YTP;
            auto ldl = [&](int j) -> double { return row[g * YR + j]; };
            auto ldh = [&](int j) -> double { return row[YLX + g * YR + j]; };
            double ev[YR], od[YR];
            inv_window2(ldl, ldh, ev, od);
            const int xp = px0 + g * YR;                                  // first output pair
            const int y = 2 * py0 + r;
            const long long z = a.out_plane0 + p + pl;
            if (y < a.n1 && xp < q0) {
                TOUT* __restrict__ o = (TOUT*)a.dst + 2 * xp + (long long)y * a.dsy + z * a.dsz;
                if (a.vec_ok && xp + YR <= q0) {
                    if (sizeof(TOUT) == 4) {
                        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
                        for (int j = 0; j < YR / 2; j++)
                            __stcs(o4 + j, make_float4((float)ev[2 * j], (float)od[2 * j], (float)ev[2 * j + 1], (float)od[2 * j + 1]));
                    } else {
                        double2* o2 = reinterpret_cast<double2*>(o);
#pragma unroll
                        for (int j = 0; j < YR; j++) __stcs(o2 + j, make_double2(ev[j], od[j]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < YR; j++)
                        if (xp + j < q0) { o[2 * j] = (TOUT)ev[j]; o[2 * j + 1] = (TOUT)od[j]; }
                }
            }
        }
        // ty is rewritten by the next round's y tasks, the other tz stage by the next issue / fill: both after this
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

// the slab buffer as a 3-D tensor (x, y, plane) of doubles with a 36 x 12 x 1 box
static bool make_slab_map(CUtensorMap* tm, const double* zb, int n0, int n1, int nplanes)
{
    EncodeTiledFn enc = encode_tiled_fn();
    // rows of the slab 16-byte aligned, and box starts at even x in both x-bands: n0 % 4 == 0
    if (!enc || (reinterpret_cast<size_t>(zb) & 15) != 0 || (n0 & 3)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)nplanes};
    const cuuint64_t strides[2] = {(cuuint64_t)n0 * 8, (cuuint64_t)n0 * n1 * 8};
    const cuuint32_t box[3] = {YBX, YBY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(zb), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t inverse_two_pass_scratch_bytes(int nx, int ny, int nz)
{
    // slab of the finest level: sp pairs (2 sp planes of nx x ny doubles), see pick_slab_pairs
    const char* e = getenv("WRB_INV_SLAB_MB");
    // measured (512^3 and 1024^3, r2f): fewer, larger slabs win -- a launch pair costs ~10 us, and the 126 MB L2 keeps too
    // little of a slab that is streamed through it once -- so: the whole level up to ~1.1 GB, 64+ planes beyond
    const double mb = (e && *e) ? atof(e) : 1100.0;
    const long long plane2 = (long long)nx * ny * 16;                     // bytes of one pair of planes
    long long sp = (long long)(mb * 1048576.0) / plane2;
    if (sp < 8) sp = 8;
    if (sp > nz / 2) sp = nz / 2;
    if (sp < 1) sp = 1;
    return (size_t)sp * (size_t)plane2 + 256;
}

static int pick_slab_pairs(int n0, int n1, int q2, size_t scratch_bytes)
{
    const long long plane2 = (long long)n0 * n1 * 16;
    long long sp = (long long)(scratch_bytes - 256) / plane2;
    if (sp > q2) sp = q2;
    if (sp < 1) sp = 1;
    return (int)sp;
}

bool inverse_two_pass_enabled()
{
    const char* e = getenv("WRB_INV_IMPL");                                // "fused": the one-kernel level (A/B timing, tests)
    return !(e && *e == 'f');
}

// One level: coefficients of box (n0, n1, n2) [symbols or coef, + lll for the low-low-low octant] -> dst, output pairs
// [seg_lo, seg_hi) along z.  zb: scratch of zb_bytes (>= one pair of planes of the box).
void inverse_level_two_pass(const double* coef, long long ay, long long az, const uint8_t* sym, unsigned long long lstride,
                            int nlay, const double* deps, const double* minval, const double* lll, void* dst,
                            int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2, double* zb, size_t zb_bytes,
                            cudaStream_t s, int seg_lo, int seg_hi, const double* halo, int n2_global, int own0)
{
    // z-slab mode (halo != null): n2 counts the rank's own planes, n2_global the level's box; output pairs are local
    const int q2 = n2 / 2;
    if (seg_lo < 0) seg_lo = 0;
    if (seg_hi < 0 || seg_hi > q2) seg_hi = q2;
    if (seg_hi <= seg_lo) return;
    InvZArgs2 za{};
    za.coef = coef; za.ay = ay; za.az = az; za.sym = sym; za.lstride = lstride;
    for (int l = 0; l < nlay && l < kNLayMax && sym != nullptr; l++) { za.deps[l] = deps[l]; za.minval[l] = minval[l]; }
    za.lll = lll; za.lsy = n0 / 2; za.lsz = (long long)(n0 / 2) * (n1 / 2);
    za.zb = zb; za.n0 = n0; za.n1 = n1; za.n2 = halo ? n2_global : n2;
    za.halo = halo; za.own0 = own0; za.nown = q2;
    // four positions per thread need 16-byte aligned rows of doubles, 4-byte aligned rows of symbols, and quads that
    // do not straddle the low / high boundary in x
    bool vec = (n0 % 8 == 0) && (ay % 4 == 0) && (az % 4 == 0) && (reinterpret_cast<size_t>(zb) % 16 == 0);
    if (sym != nullptr) vec = vec && (lstride % 4 == 0) && (reinterpret_cast<size_t>(sym) % 4 == 0);
    else vec = vec && (reinterpret_cast<size_t>(coef) % 16 == 0) && (ay % 2 == 0) && (az % 2 == 0);
    if (lll != nullptr) vec = vec && (reinterpret_cast<size_t>(lll) % 16 == 0) && ((n0 / 2) % 2 == 0);
    if (halo != nullptr) vec = vec && (reinterpret_cast<size_t>(halo) % 16 == 0);
    if (getenv("WRB_INV_NOVEC") != nullptr) vec = false;
    const int vx = vec ? 4 : 1;
    const int xthreads = (n0 + vx - 1) / vx;
    int bx = xthreads >= 128 ? 128 : (xthreads >= 64 ? 64 : 32);
    int by = 128 / bx;
    dim3 zblock(bx, by, 1);

    InvYXArgs ya{};
    ya.zb = zb; ya.n0 = n0; ya.n1 = n1; ya.dst = dst; ya.dsy = dsy; ya.dsz = dsz;
    const size_t esz = dst_is_f32 ? 4 : 8;
    ya.vec_ok = ((reinterpret_cast<size_t>(dst) % 16) == 0 && (dsy * esz) % 16 == 0 && (dsz * esz) % 16 == 0) ? 1 : 0;
    const int gx = (n0 / 2 + YPX - 1) / YPX, gy = (n1 / 2 + YPY - 1) / YPY;
    static DeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(inv_yx_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, YSMEM);
        cudaFuncSetAttribute(inv_yx_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, YSMEM);
        cudaFuncSetAttribute(inv_yx_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, YSMEM);
        cudaFuncSetAttribute(inv_yx_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, YSMEM);
    });
    const bool want_tma = getenv("WRB_INV_NOTMA") == nullptr;

    const int sp = pick_slab_pairs(n0, n1, q2, zb_bytes);
    for (int p0 = seg_lo; p0 < seg_hi; p0 += sp) {
        const int np = (p0 + sp <= seg_hi) ? sp : seg_hi - p0;
        // ---- z pass: pairs [p0, p0 + np) -> slab planes 0 .. 2 np ----
        za.pair0 = own0 + p0; za.npairs = np;
        int zp = np;              // z-segments inside the slab only when the (x, y) positions alone cannot fill the GPU
        while (zp > 8 && (long long)xthreads * n1 * ((np + zp - 1) / zp) < 148ll * 1536) zp = (zp + 1) / 2;
        za.zpairs = zp;
        dim3 zgrid((xthreads + bx - 1) / bx, (n1 + by - 1) / by, (np + zp - 1) / zp);
#define WRB_Z_LAUNCH(NL)                                                                            \
        do {                                                                                        \
            if (za.halo != nullptr) {                                                               \
                if (vec) inv_z_kernel<NL, 4, true><<<zgrid, zblock, 0, s>>>(za);                    \
                else inv_z_kernel<NL, 1, true><<<zgrid, zblock, 0, s>>>(za);                        \
            } else if (vec) inv_z_kernel<NL, 4, false><<<zgrid, zblock, 0, s>>>(za);                \
            else inv_z_kernel<NL, 1, false><<<zgrid, zblock, 0, s>>>(za);                           \
        } while (0)
        switch (sym != nullptr ? nlay : 0) {
        case 0: WRB_Z_LAUNCH(0); break;
        case 1: WRB_Z_LAUNCH(1); break;
        case 2: WRB_Z_LAUNCH(2); break;
        case 3: WRB_Z_LAUNCH(3); break;
        case 4: WRB_Z_LAUNCH(4); break;
        case 5: WRB_Z_LAUNCH(5); break;
        case 6: WRB_Z_LAUNCH(6); break;
        case 7: WRB_Z_LAUNCH(7); break;
        default: WRB_Z_LAUNCH(8); break;
        }
#undef WRB_Z_LAUNCH
        // ---- y / x pass over the slab ----
        ya.nplanes = 2 * np; ya.out_plane0 = 2ll * (p0 - 0);
        // planes per CTA: enough CTAs to fill the machine twice over, but at least four planes (two rounds) each so
        // that the TMA prefetch of the next round has something to hide behind
        int ppc = 2 * np;
        while (ppc > 4 && (long long)gx * gy * ((2 * np + ppc - 1) / ppc) < 148ll * 4) ppc = ((ppc / 2 + 1) / 2) * 2;
        // ... and whole waves: three CTAs are resident per SM (shared memory), all take about the same time, so a grid of
        // 2.3 waves (512^3, level 1: 1024 CTAs on 444 slots) runs for three.  Among the z-splits with at least as many CTAs,
        // take the one with the smallest waves x (planes per CTA + start-up) product.
        if (getenv("WRB_INV_WAVES") == nullptr || atoi(getenv("WRB_INV_WAVES")) != 0) {
            const long long tiles = (long long)gx * gy, slots = 148ll * 3;
            auto cost = [&](int pp) { const long long ctas = tiles * ((2 * np + pp - 1) / pp); return ((ctas + slots - 1) / slots) * (pp + 3); };
            int best = ppc;
            long long best_cost = cost(ppc);
            for (int pp = ppc - 2; pp >= 4 && pp >= ppc / 4; pp -= 2) {
                const long long c = cost(pp);
                if (c < best_cost) { best_cost = c; best = pp; }
            }
            ppc = best;
        }
        ya.planes_per_cta = ppc;
        dim3 ygrid(gx, gy, (2 * np + ppc - 1) / ppc);
        CUtensorMap tm;
        const bool tma = want_tma && make_slab_map(&tm, zb, n0, n1, 2 * np);
        if (!tma) memset(&tm, 0, sizeof(tm));
        if (dst_is_f32) {
            if (tma) inv_yx_kernel<float, true><<<ygrid, YTHREADS, YSMEM, s>>>(tm, ya);
            else inv_yx_kernel<float, false><<<ygrid, YTHREADS, YSMEM, s>>>(tm, ya);
        } else {
            if (tma) inv_yx_kernel<double, true><<<ygrid, YTHREADS, YSMEM, s>>>(tm, ya);
            else inv_yx_kernel<double, false><<<ygrid, YTHREADS, YSMEM, s>>>(tm, ya);
        }
        note_launch(2);
    }
}

}  // namespace wrb
