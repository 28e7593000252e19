// wavelet_inv_fused.cu -- one-pass-per-level inverse CDF 9/7 transform with the dequantiser folded in.
//
// Same arithmetic as the three inverse line passes of wavelet.cu (reference waveletcdf97_3d.c:281-466:
// per level z, then y, then x; un-scale :312-313, four inverse lifting stages :317-330, interleave
// :333-337) and as the accumulate loop of decoding_wrap() (wrappers.cpp:480,513-514), but one HBM round
// trip per level: the kernel reads the decoded SYMBOL planes (nlay bytes per coefficient) and the previous
// level's output, and writes this level's output -- the coefficient array is never materialised.
//
//   * a CTA (20 warps, one per SM) owns a tile of IPX x IPY output pairs in (x, y) and a z-segment of the box;
//   * z is a rolling pipeline in registers: every thread owns ISLOTS (x, y) coefficient positions of the
//     tile + halo and keeps four doubles of lifting state for each; feeding the pair (low[m], high[m])
//     completes the output planes 2(m-2) and 2(m-2)+1 at that position;
//   * the two finished planes go to shared memory, are inverse-lifted along y (tasks of IR pairs, register
//     windows, inv_window()) into a second tile, then along x, and leave as 32-byte runs;
//   * coefficients are rebuilt on the fly: fld = (q0*deps0 + min0) + (q1*deps1 + min1) + ...; the
//     low-low-low octant comes from the previous (coarser) level's output.
// For even extents the reference's line-end formulas equal the interior formula on the symmetric extension
// of the coefficient sequences (low: s[-k] = s[k], s[Q-1+k] = s[Q-k]; high: d[-k] = d[k-1],
// d[Q-1+k] = d[Q-1-k]; d*(h+h) == (2d)*h bit for bit), so halo positions are fetched through mirrored
// indices and no boundary code exists.  Segments restart the z pipeline two pairs early, so any
// segmentation gives bit-identical results.  Requires even box extents >= 8 at the level.
#include "wr_common.cuh"
#include "wr_kernels.h"

namespace wrb {

constexpr int IPX = 32;                       // output pairs per tile in x
constexpr int IPY = 8;                        // output pairs per tile in y
constexpr int IR = 4;                         // pairs per y/x lifting task
constexpr int ILX = IPX + 3, IHX = IPX + 4;   // low / high coefficient columns a tile needs
constexpr int ILY = IPY + 3, IHY = IPY + 4;
constexpr int ICX = ILX + IHX;                // 71
constexpr int ICY = ILY + IHY;                // 23
constexpr int ITHREADS = 640;                 // one CTA of 20 warps per SM (96 registers per thread)
constexpr int ISLOTS = (ICX * ICY + ITHREADS - 1) / ITHREADS;     // 3
constexpr int IZP = ICX + 2;                  // pitch of the z-stage tile (73: odd)
constexpr int IYP = ICX + 2;                  // pitch of the y-stage tile
constexpr int IYR = 2 * IPY;                  // rows of the y-stage tile (16)

struct FusedInvArgs {
    const double* coef; long long ay, az;     // coefficient array (array strides); ignored when sym != null
    const uint8_t* sym;                       // flat symbol planes, layer l at sym + l*lstride (or null)
    unsigned long long lstride;
    int nlay;
    double deps[kNLayMax], minval[kNLayMax];
    const double* lll; long long lsy, lsz;    // previous level's output (compact q0 x q1 x q2) or null
    void* dst; long long dsy, dsz;            // this level's output (x stride 1)
    int n0, n1, n2;                           // box extents (even)
    int zpairs;                               // output pairs per z-segment
    int seg_lo, seg_hi;                       // output pairs [seg_lo, seg_hi) this launch produces (whole box: 0, n2/2)
    int vec_ok;                               // output rows are 16-byte aligned: vector stores allowed
    // z-slab mode (NLAY == 0 only): the coefficients come from two band buffers that hold this rank's pairs
    // [pair_lo, pair_lo + nown) of the GLOBAL line plus `halo` planes of the neighbours on either side (own pair p at
    // plane p - pair_lo + halo); n2 is then the global extent and the output planes are rank-local.
    const double* lowb; const double* highb; long long bsz;
    int pair_lo, nown, halo;
};

// mirrored index of a low (s-type) / high (d-type) coefficient of a line with Q pairs, clamped
__device__ __forceinline__ int mirror_s(int i, int Q)
{
    i = (i < 0) ? -i : i;
    i = (i >= Q) ? 2 * Q - 1 - i : i;
    return min(max(i, 0), Q - 1);
}
__device__ __forceinline__ int mirror_d(int i, int Q)
{
    i = (i < 0) ? -i - 1 : i;
    i = (i >= Q) ? 2 * Q - 2 - i : i;
    return min(max(i, 0), Q - 1);
}

// Inverse lifting of IR output pairs from a window of IR+3 low and IR+4 high coefficients:
// l[t] = low[i0-1+t], h[t] = high[i0-2+t].  Same operations as inv_pairs() away from the line ends.
template <class LDL, class LDH>
__device__ __forceinline__ void inv_window(LDL ldl, LDH ldh, double (&ev)[IR], double (&od)[IR])
{
    double h[IR + 4], l[IR + 3];
#pragma unroll
    for (int t = 0; t < IR + 4; t++) h[t] = ldh(t) * WRB_SCL;
#pragma unroll
    for (int t = 0; t < IR + 3; t++) l[t] = ldl(t) * WRB_PSCL;
    double s1[IR + 3], d1[IR + 2], s2[IR + 1];
#pragma unroll
    for (int t = 0; t < IR + 3; t++) s1[t] = l[t] - WRB_LD * (h[t + 1] + h[t]);
#pragma unroll
    for (int t = 0; t < IR + 2; t++) d1[t] = h[t + 1] - WRB_LC * (s1[t + 1] + s1[t]);
#pragma unroll
    for (int t = 0; t < IR + 1; t++) s2[t] = s1[t + 1] - WRB_LB * (d1[t + 1] + d1[t]);
#pragma unroll
    for (int t = 0; t < IR; t++) {
        ev[t] = s2[t];
        od[t] = d1[t + 1] - WRB_LA * (s2[t + 1] + s2[t]);
    }
}

// NLAY > 0: coefficients are rebuilt from NLAY symbol planes; NLAY == 0: read from the coefficient array.
template <class TOUT, int NLAY>
__global__ void __launch_bounds__(ITHREADS, 1) inv_level_fused_kernel(FusedInvArgs a)
{
    constexpr bool FROM_SYM = NLAY > 0;
    constexpr int NRAW = FROM_SYM ? NLAY : 1;
    __shared__ double tz[2][ICY * IZP];       // z-inverted planes (even, odd) of the current pair, coefficient layout in x, y
    __shared__ double ty[2][IYR * IYP];       // after the y inverse: 16 sample rows x (low | high) columns
    const int tid = threadIdx.x;
    const int q0 = a.n0 >> 1, q1 = a.n1 >> 1, q2 = a.n2 >> 1;
    const int px0 = blockIdx.x * IPX, py0 = blockIdx.y * IPY;
    const bool band = (NLAY == 0) && a.lowb != nullptr;
    const int pair_lo = band ? a.pair_lo : 0, pair_hi = band ? a.pair_lo + a.nown : a.seg_hi;
    const int e0 = (band ? pair_lo : a.seg_lo) + blockIdx.z * a.zpairs;
    const int e1 = (e0 + a.zpairs < pair_hi) ? e0 + a.zpairs : pair_hi;
    // ---- coefficient positions of this thread: flat index tid + k*ITHREADS over ICY x ICX ----
    int coff[ISLOTS], loff[ISLOTS], soff[ISLOTS];
    bool inl[ISLOTS];                          // the z-low coefficient comes from the previous level's output
#pragma unroll
    for (int k = 0; k < ISLOTS; k++) {
        const int idx = tid + k * ITHREADS;
        const bool ok = idx < ICX * ICY;
        const int ry = ok ? idx / ICX : 0, rx = ok ? idx - ry * ICX : 0;
        const bool xl = rx < ILX, yl = ry < ILY;
        const int xc = xl ? mirror_s(px0 - 1 + rx, q0) : q0 + mirror_d(px0 - 2 + (rx - ILX), q0);
        const int yc = yl ? mirror_s(py0 - 1 + ry, q1) : q1 + mirror_d(py0 - 2 + (ry - ILY), q1);
        inl[k] = xl && yl && a.lll != nullptr;
        coff[k] = (int)(xc + (long long)yc * a.ay);
        loff[k] = inl[k] ? (int)(xc + (long long)yc * a.lsy) : 0;         // offset inside a plane of lll
        soff[k] = ok ? ry * IZP + rx : ICX;                               // unused slots park in a pad column
    }
    double hp[ISLOTS], s1p[ISLOTS], d1p[ISLOTS], s2p[ISLOTS];
#pragma unroll
    for (int k = 0; k < ISLOTS; k++) { hp[k] = 0; s1p[k] = 0; d1p[k] = 0; s2p[k] = 0; }

    // Raw inputs of one z step (all loads of a step are issued together, one step ahead of their use):
    // symbols of every layer for the z-low and z-high coefficient of every slot, or the coefficients themselves;
    // lraw holds the previous level's output where the z-low coefficient comes from there.
    unsigned int qlo[ISLOTS][NRAW], qhi[ISLOTS][NRAW];
    double clo[ISLOTS], chi[ISLOTS];
    auto load_step = [&](int m) {
        const long long pl = (long long)mirror_s(m, q2), ph = (long long)q2 + mirror_d(m, q2);
#pragma unroll
        for (int k = 0; k < ISLOTS; k++) {
            const long long jl = coff[k] + pl * a.az, jh = coff[k] + ph * a.az;
            if (FROM_SYM) {
#pragma unroll
                for (int l = 0; l < NRAW; l++) {
                    qhi[k][l] = a.sym[(unsigned long long)l * a.lstride + jh];
                    if (!inl[k]) qlo[k][l] = a.sym[(unsigned long long)l * a.lstride + jl];
                }
                if (inl[k]) clo[k] = a.lll[loff[k] + pl * a.lsz];
            } else if (band) {
                chi[k] = a.highb[coff[k] + (ph - q2 - pair_lo + a.halo) * a.bsz];
                clo[k] = a.lowb[coff[k] + (pl - pair_lo + a.halo) * a.bsz];
            } else {
                chi[k] = a.coef[jh];
                clo[k] = inl[k] ? a.lll[loff[k] + pl * a.lsz] : a.coef[jl];
            }
        }
    };
    // fld = (q0*deps0 + min0) + (q1*deps1 + min1) + ...   (wrappers.cpp:480,513-514; 0 + t == t: t is never -0)
    auto deq = [&](const unsigned int (&q)[NRAW]) -> double {
        double f = 0.0;
#pragma unroll
        for (int l = 0; l < NRAW; l++) {
            // (double)q through the 2^52 trick: no conversion instruction
            const double qd = __hiloint2double(0x43300000, (int)q[l]) - 4503599627370496.0;
            const double t = qd * a.deps[l] + a.minval[l];
            f = (l == 0) ? t : f + t;
        }
        return f;
    };

    // Pairs e0-2 .. e1+1 are fed: output pair i needs low[i-1..i+2] and high[i-2..i+2]; feeding pair m completes
    // output pair m-2.
    load_step(e0 - 2);
    for (int m = e0 - 2; m <= e1 + 1; m++) {
        const bool out = (m - 2 >= e0);
#pragma unroll
        for (int k = 0; k < ISLOTS; k++) {
            double lv, hv;
            if (FROM_SYM) { hv = deq(qhi[k]); lv = inl[k] ? clo[k] : deq(qlo[k]); }
            else { hv = chi[k]; lv = clo[k]; }
            const double l = lv * WRB_PSCL, h = hv * WRB_SCL;
            const double s1 = l - WRB_LD * (h + hp[k]);                   // s1[m]
            const double d1 = hp[k] - WRB_LC * (s1 + s1p[k]);             // d1[m-1]
            const double s2 = s1p[k] - WRB_LB * (d1 + d1p[k]);            // s2[m-1]
            const double d2 = d1p[k] - WRB_LA * (s2 + s2p[k]);            // d2[m-2]
            if (out) { tz[0][soff[k]] = s2p[k]; tz[1][soff[k]] = d2; }    // planes 2(m-2), 2(m-2)+1
            hp[k] = h; s1p[k] = s1; d1p[k] = d1; s2p[k] = s2;
        }
        if (m <= e1) load_step(m + 1);                                    // in flight during the y and x phases
        if (!out) continue;                                               // uniform: the whole CTA skips
        __syncthreads();
        // ---- y inverse: task = (plane, group of IR pairs, column); consecutive lanes on consecutive columns ----
        for (int t = tid; t < 2 * (IPY / IR) * ICX; t += ITHREADS) {
            const int pl = t / ((IPY / IR) * ICX), rem = t - pl * ((IPY / IR) * ICX);
            const int g = rem / ICX, c = rem - g * ICX;
            const double* col = &tz[pl][c];
            auto ldl = [&](int j) -> double { return col[(g * IR + j) * IZP]; };
            auto ldh = [&](int j) -> double { return col[(ILY + g * IR + j) * IZP]; };
            double ev[IR], od[IR];
            inv_window(ldl, ldh, ev, od);
            double* o = &ty[pl][(2 * g * IR) * IYP + c];
#pragma unroll
            for (int j = 0; j < IR; j++) { o[(2 * j) * IYP] = ev[j]; o[(2 * j + 1) * IYP] = od[j]; }
        }
        __syncthreads();
        // ---- x inverse: task = (plane, group of IR pairs, row); consecutive lanes on consecutive rows ----
        for (int t = tid; t < 2 * (IPX / IR) * IYR; t += ITHREADS) {
            const int pl = t / ((IPX / IR) * IYR), rem = t - pl * ((IPX / IR) * IYR);
            const int g = rem / IYR, r = rem - g * IYR;
            const double* row = &ty[pl][r * IYP];
            auto ldl = [&](int j) -> double { return row[g * IR + j]; };
            auto ldh = [&](int j) -> double { return row[ILX + g * IR + j]; };
            double ev[IR], od[IR];
            inv_window(ldl, ldh, ev, od);
            const int xp = px0 + g * IR;                                  // first output pair
            const int y = 2 * py0 + r;
            const long long z = 2 * (long long)(m - 2 - pair_lo) + pl;
            if (y < a.n1 && xp < q0) {
                TOUT* __restrict__ o = (TOUT*)a.dst + 2 * xp + (long long)y * a.dsy + z * a.dsz;
                if (a.vec_ok && xp + IR <= q0) {
                    if (sizeof(TOUT) == 4) {
                        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
                        for (int j = 0; j < IR / 2; j++)
                            o4[j] = make_float4((float)ev[2 * j], (float)od[2 * j], (float)ev[2 * j + 1], (float)od[2 * j + 1]);
                    } else {
                        double2* o2 = reinterpret_cast<double2*>(o);
#pragma unroll
                        for (int j = 0; j < IR; j++) o2[j] = make_double2(ev[j], od[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < IR; j++)
                        if (xp + j < q0) { o[2 * j] = (TOUT)ev[j]; o[2 * j + 1] = (TOUT)od[j]; }
                }
            }
        }
        // no trailing barrier: tz is next written after this iteration's second barrier by threads that have
        // finished their y tasks, and ty is next written after the first barrier of the next iteration
    }
}

bool fused_inverse_supported(int n0, int n1, int n2)
{
    return (n0 % 2 == 0) && (n1 % 2 == 0) && (n2 % 2 == 0) && n0 >= 8 && n1 >= 8 && n2 >= 8;
}

// One level: coefficients of box (n0,n1,n2) [symbols or coef, + lll for the low-low-low octant] -> dst.
void fused_inverse_level(const double* coef, long long ay, long long az, const uint8_t* sym, unsigned long long lstride,
                         int nlay, const double* deps, const double* minval, const double* lll, void* dst,
                         int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2, cudaStream_t s,
                         int seg_lo, int seg_hi)
{
    FusedInvArgs a{};
    a.coef = coef; a.ay = ay; a.az = az; a.sym = sym; a.lstride = lstride; a.nlay = nlay;
    for (int l = 0; l < nlay && l < kNLayMax && sym != nullptr; l++) { a.deps[l] = deps[l]; a.minval[l] = minval[l]; }
    a.lll = lll; a.lsy = n0 / 2; a.lsz = (long long)(n0 / 2) * (n1 / 2);
    a.dst = dst; a.dsy = dsy; a.dsz = dsz; a.n0 = n0; a.n1 = n1; a.n2 = n2;
    const size_t esz = dst_is_f32 ? 4 : 8;
    a.vec_ok = ((reinterpret_cast<size_t>(dst) % 16) == 0 && (dsy * esz) % 16 == 0 && (dsz * esz) % 16 == 0) ? 1 : 0;
    const int q0 = n0 / 2, q1 = n1 / 2, q2 = n2 / 2;
    const int gx = (q0 + IPX - 1) / IPX, gy = (q1 + IPY - 1) / IPY;
    a.seg_lo = (seg_lo < 0) ? 0 : seg_lo;
    a.seg_hi = (seg_hi < 0 || seg_hi > q2) ? q2 : seg_hi;
    const int npairs = a.seg_hi - a.seg_lo;
    if (npairs <= 0) return;
    // z-segments: shortest critical path for one resident CTA per SM (512^3 level 1: 4 segments, 7 waves)
    const int zp = pick_zpairs((long long)gx * gy, npairs, 148, 4, 4);
    a.zpairs = zp;
    dim3 grid(gx, gy, (npairs + zp - 1) / zp);
#define WRB_INV_LAUNCH(NL)                                                                          \
    do {                                                                                            \
        if (dst_is_f32) inv_level_fused_kernel<float, NL><<<grid, ITHREADS, 0, s>>>(a);             \
        else inv_level_fused_kernel<double, NL><<<grid, ITHREADS, 0, s>>>(a);                       \
    } while (0)
    switch (sym != nullptr ? nlay : 0) {
    case 0: WRB_INV_LAUNCH(0); break;
    case 1: WRB_INV_LAUNCH(1); break;
    case 2: WRB_INV_LAUNCH(2); break;
    case 3: WRB_INV_LAUNCH(3); break;
    case 4: WRB_INV_LAUNCH(4); break;
    case 5: WRB_INV_LAUNCH(5); break;
    case 6: WRB_INV_LAUNCH(6); break;
    case 7: WRB_INV_LAUNCH(7); break;
    default: WRB_INV_LAUNCH(8); break;
    }
#undef WRB_INV_LAUNCH
    note_launch(1);
}

// One level in z-slab mode: bands (coefficient layout in x, y; strides bsy, bsz; own pair p at plane p - pair_lo + halo)
// -> this rank's 2*nown output planes.  n2g: GLOBAL z extent of the level's box.
void fused_inverse_level_bands(const double* lowb, const double* highb, long long bsy, long long bsz, int halo, int pair_lo,
                               int nown, void* dst, int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2g,
                               cudaStream_t s)
{
    FusedInvArgs a{};
    a.coef = nullptr; a.ay = bsy; a.az = 0; a.sym = nullptr; a.nlay = 0; a.lll = nullptr;
    a.dst = dst; a.dsy = dsy; a.dsz = dsz; a.n0 = n0; a.n1 = n1; a.n2 = n2g;
    a.lowb = lowb; a.highb = highb; a.bsz = bsz; a.pair_lo = pair_lo; a.nown = nown; a.halo = halo;
    const size_t esz = dst_is_f32 ? 4 : 8;
    a.vec_ok = ((reinterpret_cast<size_t>(dst) % 16) == 0 && (dsy * esz) % 16 == 0 && (dsz * esz) % 16 == 0) ? 1 : 0;
    const int q0 = n0 / 2, q1 = n1 / 2;
    const int gx = (q0 + IPX - 1) / IPX, gy = (q1 + IPY - 1) / IPY;
    const int zp = pick_zpairs((long long)gx * gy, nown, 148, 4, 4);
    a.zpairs = zp;
    dim3 grid(gx, gy, (nown + zp - 1) / zp);
    if (dst_is_f32) inv_level_fused_kernel<float, 0><<<grid, ITHREADS, 0, s>>>(a);
    else inv_level_fused_kernel<double, 0><<<grid, ITHREADS, 0, s>>>(a);
    note_launch(1);
}

}  // namespace wrb