// wavelet_fused.cu -- one-pass-per-level forward CDF 9/7 transform.
//
// Same arithmetic as the three line passes of wavelet.cu (and therefore of the reference,
// waveletcdf97_3d.c:94-264), but one HBM round trip per level instead of three:
//   * a CTA owns an (x, y) tile of PX x PY output pairs and a z-segment of the level's box;
//   * for every input plane of its segment it loads the tile + 7-sample halo into shared memory,
//     lifts the rows (x) and then the columns (y) there, each thread using the same register-window
//     evaluation (fwd_pairs) as the line passes;
//   * the z direction is a rolling pipeline: every thread owns four (x, y, sub-band) positions and
//     keeps five doubles of lifting state for each; when the pair of planes (2m, 2m+1) has
//     arrived, outputs s[m-2], d[m-2] are complete and go straight to their octant (the
//     low-low-low octant to the compact scratch that feeds the next level);
//   * field extrema (level 1) and coefficient extrema are reduced on the way.
// Segments restart the z pipeline two pairs early (the lifting stencil reaches 4 samples back), so
// any segmentation gives bit-identical results.  Requires even box extents >= 8; other shapes take
// the general three-pass path.
#include "wr_common.cuh"
#include "wr_kernels.h"

#include "wavelet_pairs.cuh"

namespace wrb {

constexpr int FPX = 32;                 // output pairs per tile in x  -> 64 input columns + 7 halo
constexpr int FPY = 8;                  // output pairs per tile in y  -> 16 input rows + 7 halo
constexpr int FTX = 2 * FPX + 7;        // 71
constexpr int FTY = 2 * FPY + 7;        // 23
constexpr int FXR = 4;                  // pairs per x-lifting task
constexpr int FTHREADS = 256;

struct FusedFwdArgs {
    const void* src; long long ssy, ssz;       // level input (x stride 1)
    double* coef;    long long ay, az;          // coefficient array (array strides)
    double* lll;     long long lsy, lsz;        // compact low-low-low scratch, or null on the last level
    int n0, n1, n2;                             // box extents (even); n2 is the GLOBAL z extent
    int zoff;                                   // slab mode: global z of plane 0 of src (0 otherwise)
    int pair_lo, nl;                            // owned global pairs [pair_lo, pair_lo + nl) (0, n2/2 otherwise)
    int zpairs;                                 // output pairs per z-segment
    unsigned long long* in_min;  unsigned long long* in_max;     // field extrema keys (or null)
    unsigned long long* out_min; unsigned long long* out_max;    // coefficient extrema keys
};

// index reflection of the whole-sample symmetric extension (then clamped: far-out positions only
// feed outputs that are masked anyway)
__device__ __forceinline__ int mirror_idx(int i, int n)
{
    i = (i < 0) ? -i : i;
    i = (i >= n) ? 2 * (n - 1) - i : i;
    return min(max(i, 0), n - 1);
}

// For even line lengths the reference's line-end formulas (waveletcdf97_3d.c:113,116,121,124) are what
// the interior formula gives on the whole-sample symmetric extension of the line: V1[M-1] += a*2*V0[M-1]
// is a*(V0[M]+V0[M-1]) with V0[M] := V0[M-1], bit for bit (s+s and 2a are exact).  So the halo of the
// tile, and the planes fed to the z pipeline beyond the box, are fetched through mirrored indices and
// no boundary code exists below.
template <class TIN>
__global__ void __launch_bounds__(FTHREADS, 2) fwd_level_fused_kernel(FusedFwdArgs a)
{
    // odd row pitches: the x-lifting tasks run with consecutive lanes on consecutive ROWS, so an odd
    // pitch (in doubles) spreads them over the banks; the y-lifting reads run along a row
    __shared__ double tin[FTY][FTX + 2];        // input tile incl. halo (pitch 73)
    __shared__ double tx[FTY][2 * FPX + 1];     // after x-lifting: [row][low 32 | high 32] (pitch 65)
    const TIN* __restrict__ src = (const TIN*)a.src;
    const int tid = threadIdx.x, wrp = tid >> 5, lane = tid & 31;
    const int m0 = a.n0 >> 1, m1 = a.n1 >> 1;
    const int px0 = blockIdx.x * FPX, py0 = blockIdx.y * FPY;
    const int e0 = a.pair_lo + blockIdx.z * a.zpairs;
    const int e1 = (e0 + a.zpairs < a.pair_lo + a.nl) ? e0 + a.zpairs : a.pair_lo + a.nl;
    const int x0 = 2 * px0 - 4, y0 = 2 * py0 - 4;           // tile origin in input coordinates
    // ---- tile slots of this thread: rows wrp, wrp+8, wrp+16; columns lane, lane+32, lane+64 ----
    int toff[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int ry = wrp + 8 * r, rx = lane + 32 * q;
            toff[r * 3 + q] = (ry < FTY && rx < FTX)
                                  ? (int)(mirror_idx(x0 + rx, a.n0) + (long long)mirror_idx(y0 + ry, a.n1) * a.ssy) : -1;
        }
    // ---- z-pipeline ownership: column c of tx, y pairs j0, j0+1 ----
    const int c = tid & 63;
    const int j0 = py0 + 2 * (tid >> 6);
    const int xo = (c < FPX) ? px0 + c : m0 + px0 + (c - FPX);      // output x
    const bool xok = ((c < FPX) ? px0 + c : px0 + c - FPX) < m0;
    int ooff[4], loff[2];                          // output offsets inside a z-plane (< 2^31), -1: masked
#pragma unroll
    for (int v = 0; v < 4; v++) {
        const int j = j0 + (v & 1);
        const int yo = (v < 2) ? j : m1 + j;
        ooff[v] = (xok && j < m1) ? (int)(xo + (long long)yo * a.ay) : -1;
        if (v < 2) loff[v] = (xok && j < m1 && c < FPX && a.lll != nullptr) ? (int)(xo + (long long)yo * a.lsy) : -1;
    }
    double s0p[4] = {0, 0, 0, 0}, d0p[4] = {0, 0, 0, 0}, d1p[4] = {0, 0, 0, 0}, s1p[4] = {0, 0, 0, 0}, d2p[4] = {0, 0, 0, 0};
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    double fmn = kInf, fmx = -kInf, omn = kInf, omx = -kInf;      // fmin/fmax identities
    const bool track_in = a.in_min != nullptr;

    TIN nxt[9];
    auto load_plane = [&](int z) {
        const TIN* __restrict__ plane = src + (long long)(mirror_idx(z, a.n2) - a.zoff) * a.ssz;
#pragma unroll
        for (int k = 0; k < 9; k++) nxt[k] = (toff[k] >= 0) ? plane[toff[k]] : (TIN)0;
    };
    // x- and y-lifting of the plane held in nxt[]; prefetches plane znext meanwhile; returns the four
    // (x, y, sub-band) values this thread feeds into its z pipelines
    auto lift_plane = [&](int znext, bool more, double (&yv)[4]) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
                if (toff[r * 3 + q] >= 0) {
                    const double v = (double)nxt[r * 3 + q];
                    tin[wrp + 8 * r][lane + 32 * q] = v;
                    if (track_in) { fmn = fmin(fmn, v); fmx = fmax(fmx, v); }
                }
        __syncthreads();
        if (more) load_plane(znext);               // next plane in flight while this one is lifted
        // x-lifting: FXR pairs per task, consecutive lanes on consecutive rows
        for (int t = tid; t < FTY * (FPX / FXR); t += FTHREADS) {
            const int g = t / FTY, ry = t - g * FTY;
            double so[FXR], dd[FXR];
            const double* row = &tin[ry][4 + 2 * g * FXR];         // sample 2*i0 of the line
            auto ld = [&](int j) -> double { return row[j]; };
            fwd_pairs_interior<FXR>(ld, 0, so, dd);
#pragma unroll
            for (int k = 0; k < FXR; k++) { tx[ry][g * FXR + k] = so[k]; tx[ry][FPX + g * FXR + k] = dd[k]; }
        }
        __syncthreads();
        // y-lifting of column c for pairs j0, j0+1
        double sy[2], dy[2];
        const int rb = 4 + 4 * (tid >> 6);                           // tile row of sample 2*j0
        auto ldy = [&](int j) -> double { return tx[rb + j][c]; };
        fwd_pairs_interior<2>(ldy, 0, sy, dy);
        yv[0] = sy[0]; yv[1] = sy[1]; yv[2] = dy[0]; yv[3] = dy[1];
        // no trailing barrier: tin is next written only by threads that have passed the second barrier
        // (x-lifting done), and tx is next written after the first barrier of the next plane
    };
    // Pairs e0-2 .. e1+1 are fed: the pipeline restarts two pairs before the segment, and the last output
    // pair e1-1 completes when the even plane of pair e1+1 has been lifted (its odd plane is not needed).
    load_plane(2 * (e0 - 2));
    for (int m = e0 - 2; m <= e1 + 1; m++) {
        double yv[4];
        lift_plane(2 * m + 1, m <= e1, yv);                      // even plane: s0[m]; pair m-2 completes
        const int mo = m - 2;
        const bool out = (mo >= e0);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double s0n = yv[v];
            const double d1 = d0p[v] + WRB_LA * (s0n + s0p[v]);          // d1[m-1]
            const double s1 = s0p[v] + WRB_LB * (d1 + d1p[v]);           // s1[m-1]
            const double d2 = d1p[v] + WRB_LC * (s1 + s1p[v]);           // d2[m-2]
            const double s2 = s1p[v] + WRB_LD * (d2 + d2p[v]);           // s2[m-2]
            d2p[v] = d2; d1p[v] = d1; s1p[v] = s1; s0p[v] = s0n;
            if (out && ooff[v] >= 0) {
                const double lo = s2 * WRB_SCL, hi = d2 * WRB_PSCL;
                a.coef[ooff[v] + (long long)(a.nl + mo - a.pair_lo) * a.az] = hi;   // z-high: always a final coefficient
                omn = fmin(omn, hi); omx = fmax(omx, hi);
                if (v < 2 && loff[v & 1] >= 0) {
                    a.lll[loff[v & 1] + (long long)(mo - a.pair_lo) * a.lsz] = lo;   // low-low-low: next level's input
                } else {
                    a.coef[ooff[v] + (long long)(mo - a.pair_lo) * a.az] = lo;
                    omn = fmin(omn, lo); omx = fmax(omx, lo);
                }
            }
        }
        if (m <= e1) {                                               // odd plane: d0[m]
            lift_plane(2 * m + 2, true, yv);
#pragma unroll
            for (int v = 0; v < 4; v++) d0p[v] = yv[v];
        }
    }
    if (track_in) block_minmax_commit(fmn <= fmx ? dkey(fmn) : kKeyMinInit, fmn <= fmx ? dkey(fmx) : kKeyMaxInit, a.in_min, a.in_max);
    block_minmax_commit(omn <= omx ? dkey(omn) : kKeyMinInit, omn <= omx ? dkey(omx) : kKeyMaxInit, a.out_min, a.out_max);
}

bool fused_forward_supported(int n0, int n1, int n2)
{
    return (n0 % 2 == 0) && (n1 % 2 == 0) && (n2 % 2 == 0) && n0 >= 8 && n1 >= 8 && n2 >= 8;
}

// One level: src box (n0,n1,n2) -> coef (detail octants, final) + lll (or coef when lll == null).
// Slab mode (zoff/pair_lo/nl): n2 is the global z extent, src holds the owned planes plus the halo
// (local plane = global z - zoff) and outputs use rank-local plane indices.
void fused_forward_level(const void* src, int src_is_f32, long long ssy, long long ssz, double* coef, long long ay,
                         long long az, double* lll, int n0, int n1, int n2, unsigned long long* in_min,
                         unsigned long long* in_max, unsigned long long* out_min, unsigned long long* out_max,
                         cudaStream_t s, int zoff, int pair_lo, int nl)
{
    FusedFwdArgs a{};
    a.src = src; a.ssy = ssy; a.ssz = ssz; a.coef = coef; a.ay = ay; a.az = az;
    a.lll = lll; a.lsy = n0 / 2; a.lsz = (long long)(n0 / 2) * (n1 / 2);
    a.n0 = n0; a.n1 = n1; a.n2 = n2;
    a.zoff = zoff; a.pair_lo = pair_lo; a.nl = (nl < 0) ? n2 / 2 : nl;
    a.in_min = in_min; a.in_max = in_max; a.out_min = out_min; a.out_max = out_max;
    const int m0 = n0 / 2, m1 = n1 / 2, m2 = a.nl;
    const int gx = (m0 + FPX - 1) / FPX, gy = (m1 + FPY - 1) / FPY;
    // z-segments: enough CTAs to fill the machine (148 SMs x 2 resident), but segments of >= 16 pairs
    int zp = m2;
    while (zp > 16 && (long long)gx * gy * ((m2 + zp - 1) / zp) < 148 * 4) zp = (zp + 1) / 2;
    a.zpairs = zp;
    dim3 grid(gx, gy, (m2 + zp - 1) / zp);
    if (src_is_f32) fwd_level_fused_kernel<float><<<grid, FTHREADS, 0, s>>>(a);
    else fwd_level_fused_kernel<double><<<grid, FTHREADS, 0, s>>>(a);
    note_launch(1);
}

}  // namespace wrb
