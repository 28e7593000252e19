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
#include <cstdlib>
#include <type_traits>
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
constexpr int kFwdImplDefault = 2;          // see the kernel: 0 two barriers per plane, 1 software pipeline, 2 pipeline + native-type tile

struct FusedFwdArgs {
    const void* src; long long ssy, ssz;       // level input (x stride 1)
    double* coef;    long long ay, az;          // coefficient array (array strides)
    double* lll;     long long lsy, lsz;        // compact low-low-low scratch, or null on the last level
    int n0, n1, n2;                             // box extents (even); n2 is the GLOBAL z extent
    int zoff;                                   // slab mode: global z of plane 0 of src incl. halo (0 otherwise)
    const void* halo_lo; const void* halo_hi;   // slab mode, level 1: the 4 planes below / 3 above live here and src
    int nown;                                   // holds only the nown own planes (null: src holds halo + own + halo)
    int pair_lo, nl;                            // owned global pairs [pair_lo, pair_lo + nl) (0, n2/2 otherwise)
    int zpairs;                                 // output pairs per z-segment
    int hi_off;                                 // plane offset of the z-high band relative to the z-low band (nl, or the
                                                // global n2/2 when a piece of a whole box is produced)
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

// running extrema in the input's own type: for f32 input fminf/fmaxf are single ALU instructions, and
// widening to double afterwards is exact and monotonic
__device__ __forceinline__ void track_ext(float v, float& mn, float& mx) { mn = fminf(mn, v); mx = fmaxf(mx, v); }
__device__ __forceinline__ void track_ext(double v, double& mn, double& mx) { mn = dmin2(mn, v); mx = dmax2(mx, v); }

// (row pitch of the input tile: template parameter PITCH of the kernel, 73 = FTX + 2: odd)
constexpr int FSLOTS = (FTY * FTX + FTHREADS - 1) / FTHREADS;       // tile elements per thread (7)

// For even line lengths the reference's line-end formulas (waveletcdf97_3d.c:113,116,121,124) are what
// the interior formula gives on the whole-sample symmetric extension of the line: V1[M-1] += a*2*V0[M-1]
// is a*(V0[M]+V0[M-1]) with V0[M] := V0[M-1], bit for bit (s+s and 2a are exact).  So the halo of the
// tile, and the planes fed to the z pipeline beyond the box, are fetched through mirrored indices and
// no boundary code exists below.
//
// Everything that does not change from plane to plane is computed once per thread: the (mirrored) global
// offset and the shared-memory offset of its FSLOTS tile elements, the in-plane offsets and masks of its
// eight outputs.  Per plane a load is then one IMAD.WIDE + LDG and a store one IMAD.WIDE + STG.
// PITCH: row pitch of the input tile in doubles.  73 (odd) is what runs; 72 is kept as a measurement variant
// (WRB_FWD_PITCH=72): it is the densest pitch a TMA box could deliver (box rows are multiples of 16 bytes, so the pitch of a
// TMA-written tile of doubles is even) and shows what the x-lifting -- consecutive lanes on consecutive ROWS -- pays for it.
//
// PIPE: the three stages of a plane -- tile store, x-lifting, y-lifting + z pipeline -- run as a software pipeline over
// double-buffered tiles: round q stores plane q, x-lifts plane q-1 and y-lifts plane q-2, so ONE barrier per plane
// separates the rounds instead of two, and a warp without an x-lifting task (184 tasks on 256 threads) goes straight
// on to its y-lifting.  Same arithmetic on the same values: bit-identical.
constexpr int FTXP = 2 * FPX + 1;               // row pitch of the x-lifted tile (65)
//
// NATIVE: the input tile stays in the input's own type (a float field: half the shared-memory traffic of the tile store and
// of the x-lifting loads; the widening moves into the x-lifting), and the x-lifting tasks are numbered 24 per group of
// pairs instead of 23 (one idle lane per group): with row pitches 75 (float) / 73 (double) no two lanes of a warp / half
// warp then meet in a bank, where the dense numbering costs every straddling half warp a replay.  The r3 profile of the
// level-1 kernel showed why this matters: the shared-memory data pipe was the busiest unit (70 % of its wavefront peak, a
// quarter of the wavefronts bank-conflict replays) ahead of the FP64 pipe (42 %) and DRAM (30 %).
template <class TIN, bool NATIVE> struct FwdTile { typedef double type; };
template <class TIN> struct FwdTile<TIN, true> { typedef TIN type; };
template <class TIN, int PITCH, bool NATIVE> __host__ __device__ constexpr int fused_fwd_pitch() { return NATIVE ? (sizeof(TIN) == 4 ? 75 : 73) : PITCH; }
template <class TIN, bool PIPE, int PITCH, bool NATIVE> constexpr int fused_fwd_smem()
{
    return (PIPE ? 2 : 1) * FTY * (fused_fwd_pitch<TIN, PITCH, NATIVE>() * (int)sizeof(typename FwdTile<TIN, NATIVE>::type) + FTXP * (int)sizeof(double)) + 16;
}

template <class TIN, bool TRACK_IN, int PITCH = FTX + 2, bool PIPE = false, bool NATIVE = false>
__global__ void __launch_bounds__(FTHREADS, 2) fwd_level_fused_kernel(FusedFwdArgs a)
{
    typedef typename FwdTile<TIN, NATIVE>::type TT;
    constexpr int FTP = fused_fwd_pitch<TIN, PITCH, NATIVE>();
    constexpr int XTG = NATIVE ? FTY + 1 : FTY;  // x-lifting task numbers per group of FXR pairs
    constexpr int NB = PIPE ? 2 : 1;
    // odd row pitches: the x-lifting tasks run with consecutive lanes on consecutive ROWS, so an odd
    // pitch spreads them over the banks; the y-lifting reads run along a row
    extern __shared__ __align__(16) double fsm[];
    double* const tx = fsm;                     // [NB] after x-lifting: [row][low 32 | high 32], FTY x FTXP
    TT* const tin = reinterpret_cast<TT*>(fsm + NB * FTY * FTXP);   // [NB] input tile incl. halo, FTY x FTP
    const TIN* __restrict__ src = (const TIN*)a.src;
    const int tid = threadIdx.x;
    const int m0 = a.n0 >> 1, m1 = a.n1 >> 1;
    const int px0 = blockIdx.x * FPX, py0 = blockIdx.y * FPY;
    const int e0 = a.pair_lo + blockIdx.z * a.zpairs;
    const int e1 = (e0 + a.zpairs < a.pair_lo + a.nl) ? e0 + a.zpairs : a.pair_lo + a.nl;
    const int x0 = 2 * px0 - 4, y0 = 2 * py0 - 4;           // tile origin in input coordinates
    // ---- tile elements of this thread: flat index tid + k*FTHREADS over FTY x FTX ----
    int goff[FSLOTS], soff[FSLOTS];
#pragma unroll
    for (int k = 0; k < FSLOTS; k++) {
        const int idx = tid + k * FTHREADS;
        const bool ok = idx < FTY * FTX;
        const int ry = idx / FTX, rx = idx - ry * FTX;
        // unused slots re-read element (0,0) of the plane (a real field value: harmless for the extrema)
        // and park it in the pad column of row 0
        goff[k] = ok ? (int)(mirror_idx(x0 + rx, a.n0) + (long long)mirror_idx(y0 + ry, a.n1) * a.ssy) : 0;
        soff[k] = ok ? ry * FTP + rx : FTX;
    }
    // ---- z-pipeline ownership: column c of tx, y pairs j0, j0+1 ----
    const int c = tid & 63;
    const int j0 = py0 + 2 * (tid >> 6);
    const bool xlow = c < FPX;
    const int xo = xlow ? px0 + c : m0 + px0 + (c - FPX);           // output x
    const bool xok = (xlow ? px0 + c : px0 + c - FPX) < m0;
    const bool to_lll = xlow && a.lll != nullptr;                    // v < 2: low-low-low, next level's input
    int ooff[4], o01[2];                           // output offsets inside a z-plane (< 2^31)
    bool ook[4];
#pragma unroll
    for (int v = 0; v < 4; v++) {
        const int j = j0 + (v & 1);
        const int yo = (v < 2) ? j : m1 + j;
        ook[v] = xok && j < m1;
        ooff[v] = ook[v] ? (int)(xo + (long long)yo * a.ay) : 0;
        if (v < 2) o01[v] = ook[v] ? (to_lll ? (int)(xo + (long long)yo * a.lsy) : ooff[v]) : 0;
    }
    // plane pointers of output pair mo = e0 - 4 (advanced once per pair; dereferenced for mo >= e0 only)
    const long long s01 = to_lll ? a.lsz : a.az;
    double* plo = a.coef + (long long)(e0 - 4 - a.pair_lo) * a.az;
    double* phi = plo + (long long)a.hi_off * a.az;
    double* p01 = (to_lll ? a.lll : a.coef) + (long long)(e0 - 4 - a.pair_lo) * s01;
    double s0p[4] = {0, 0, 0, 0}, d0p[4] = {0, 0, 0, 0}, d1p[4] = {0, 0, 0, 0}, s1p[4] = {0, 0, 0, 0}, d2p[4] = {0, 0, 0, 0};
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    TIN fmn = (TIN)kInf, fmx = (TIN)(-kInf);                        // fmin/fmax identities
    double omn = kInf, omx = -kInf;

    TIN nxt[FSLOTS];
    auto load_plane = [&](int z) {
        const int lp = mirror_idx(z, a.n2) - a.zoff;                 // plane counted from the first lower-halo plane
        const TIN* __restrict__ plane = src + (long long)lp * a.ssz;
        if (a.halo_lo != nullptr) {                                   // uniform per CTA
            if (lp < 4) plane = (const TIN*)a.halo_lo + (long long)lp * a.ssz;
            else if (lp < 4 + a.nown) plane = src + (long long)(lp - 4) * a.ssz;
            else plane = (const TIN*)a.halo_hi + (long long)(lp - 4 - a.nown) * a.ssz;
        }
#pragma unroll
        for (int k = 0; k < FSLOTS; k++) nxt[k] = plane[goff[k]];
    };
    // the three stages of a plane (b: tile buffer)
    auto stage_store = [&](int b) {                // nxt[] -> input tile
        TT* const t = tin + b * (FTY * FTP);
#pragma unroll
        for (int k = 0; k < FSLOTS; k++) {
            const TIN v = nxt[k];
            if (TRACK_IN) track_ext(v, fmn, fmx);
            t[soff[k]] = (TT)v;
        }
    };
    auto stage_x = [&](int b) {                    // x-lifting: FXR pairs per task, consecutive lanes on consecutive rows
        const TT* const t = tin + b * (FTY * FTP);
        double* const o = tx + b * (FTY * FTXP);
        for (int tk = tid; tk < XTG * (FPX / FXR); tk += FTHREADS) {
            const int g = tk / XTG, ry = tk - g * XTG;
            if (NATIVE && ry >= FTY) continue;                       // the idle number of the group
            double so[FXR], dd[FXR];
            const TT* row = &t[ry * FTP + 4 + 2 * g * FXR];         // sample 2*i0 of the line
            auto ld = [&](int j) -> double { return (double)row[j]; };
            fwd_pairs_interior<FXR>(ld, 0, so, dd);
#pragma unroll
            for (int k = 0; k < FXR; k++) { o[ry * FTXP + g * FXR + k] = so[k]; o[ry * FTXP + FPX + g * FXR + k] = dd[k]; }
        }
    };
    auto stage_y = [&](int b, double (&yv)[4]) {   // y-lifting of column c for pairs j0, j0+1: the four values fed to the z pipelines
        const double* const o = tx + b * (FTY * FTXP);
        double sy[2], dy[2];
        const int rb = 4 + 4 * (tid >> 6);                           // tile row of sample 2*j0
        auto ldy = [&](int j) -> double { return o[(rb + j) * FTXP + c]; };
        fwd_pairs_interior<2>(ldy, 0, sy, dy);
        yv[0] = sy[0]; yv[1] = sy[1]; yv[2] = dy[0]; yv[3] = dy[1];
    };
    // z pipeline: the even plane of pair m has arrived (s0[m]); pair m-2 completes
    auto z_even = [&](int m, const double (&yv)[4]) {
        const bool out = (m - 2 >= e0);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double s0n = yv[v];
            const double d1 = d0p[v] + WRB_LA * (s0n + s0p[v]);          // d1[m-1]
            const double s1 = s0p[v] + WRB_LB * (d1 + d1p[v]);           // s1[m-1]
            const double d2 = d1p[v] + WRB_LC * (s1 + s1p[v]);           // d2[m-2]
            const double s2 = s1p[v] + WRB_LD * (d2 + d2p[v]);           // s2[m-2]
            d2p[v] = d2; d1p[v] = d1; s1p[v] = s1; s0p[v] = s0n;
            if (out && ook[v]) {
                const double lo = s2 * WRB_SCL, hi = d2 * WRB_PSCL;
                phi[ooff[v]] = hi;                                       // z-high: always a final coefficient
                omn = dmin2(omn, hi); omx = dmax2(omx, hi);
                if (v < 2) {
                    p01[o01[v]] = lo;
                    if (!to_lll) { omn = dmin2(omn, lo); omx = dmax2(omx, lo); }
                } else {
                    plo[ooff[v]] = lo;
                    omn = dmin2(omn, lo); omx = dmax2(omx, lo);
                }
            }
        }
        plo += a.az; phi += a.az; p01 += s01;
    };
    // Pairs e0-2 .. e1+1 are fed: the pipeline restarts two pairs before the segment, and the last output
    // pair e1-1 completes when the even plane of pair e1+1 has been lifted (its odd plane is not needed).
    if constexpr (!PIPE) {
        // one plane at a time: store, barrier, (next plane's loads issued,) x-lifting, barrier, y-lifting.  No trailing
        // barrier: tin is next written only by threads that have passed the second barrier (x-lifting done), and tx is
        // next written after the first barrier of the next plane
        auto lift_plane = [&](int znext, bool more, double (&yv)[4]) {
            stage_store(0);
            __syncthreads();
            if (more) load_plane(znext);               // next plane in flight while this one is lifted
            stage_x(0);
            __syncthreads();
            stage_y(0, yv);
        };
        load_plane(2 * (e0 - 2));
        for (int m = e0 - 2; m <= e1 + 1; m++) {
            double yv[4];
            lift_plane(2 * m + 1, m <= e1, yv);                      // even plane: s0[m]
            z_even(m, yv);
            if (m <= e1) {                                               // odd plane: d0[m]
                lift_plane(2 * m + 2, true, yv);
#pragma unroll
                for (int v = 0; v < 4; v++) d0p[v] = yv[v];
            }
        }
    } else {
        // planes i = 0 .. P-1 are z = 2*(e0-2) + i (P odd: the last one is the even plane of pair e1+1); round q stores
        // plane q (and issues the loads of plane q+1), x-lifts plane q-1 into tx[(q-1)&1] and y-lifts plane q-2 from
        // tx[q&1].  Every buffer a round writes was last read in the round before: one barrier per round.
        const int P = 2 * (e1 - e0) + 7, zb = 2 * (e0 - 2);
        load_plane(zb);
        for (int k = 0;; k++) {
            const int q0 = 2 * k;                                        // even round: planes q0 (even), q0-1 (odd), q0-2 (even)
            if (q0 < P) { stage_store(0); if (q0 + 1 < P) load_plane(zb + q0 + 1); }
            if (q0 >= 1 && q0 < P) stage_x(1);                           // plane q0-1 (odd); the last round q0 = P+1 has none
            if (q0 >= 2) {
                double yv[4];
                stage_y(0, yv);
                z_even(e0 - 3 + k, yv);                                  // plane q0-2 is the even plane of pair e0-2 + (q0-2)/2
            }
            if (q0 == P + 1) break;
            __syncthreads();
            if (q0 + 1 < P) { stage_store(1); load_plane(zb + q0 + 2); }   // odd round q0+1 <= P: plane q0+2 <= P-1 exists whenever q0+1 < P (P odd)
            stage_x(0);                                                  // plane q0 (even)
            if (q0 >= 1) {
                double yv[4];
                stage_y(1, yv);                                          // plane q0-1: the odd plane of pair e0-2 + (q0-2)/2
#pragma unroll
                for (int v = 0; v < 4; v++) d0p[v] = yv[v];
            }
            __syncthreads();
        }
    }
    if (TRACK_IN) {
        const double dmn = (double)fmn, dmx = (double)fmx;
        block_minmax_commit(dmn <= dmx ? dkey(dmn) : kKeyMinInit, dmn <= dmx ? dkey(dmx) : kKeyMaxInit, a.in_min, a.in_max);
    }
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
                         cudaStream_t s, int zoff, int pair_lo, int nl, int hi_off, const void* halo_lo, const void* halo_hi)
{
    FusedFwdArgs a{};
    a.halo_lo = halo_lo; a.halo_hi = halo_hi;
    a.src = src; a.ssy = ssy; a.ssz = ssz; a.coef = coef; a.ay = ay; a.az = az;
    a.lll = lll; a.lsy = n0 / 2; a.lsz = (long long)(n0 / 2) * (n1 / 2);
    a.n0 = n0; a.n1 = n1; a.n2 = n2;
    a.zoff = zoff; a.pair_lo = pair_lo; a.nl = (nl < 0) ? n2 / 2 : nl;
    a.hi_off = (hi_off < 0) ? a.nl : hi_off;
    a.nown = 2 * a.nl;
    a.in_min = in_min; a.in_max = in_max; a.out_min = out_min; a.out_max = out_max;
    const int m0 = n0 / 2, m1 = n1 / 2, m2 = a.nl;
    const int gx = (m0 + FPX - 1) / FPX, gy = (m1 + FPY - 1) / FPY;
    // z-segments: enough CTAs to fill the machine (148 SMs x 2 resident), but segments of >= 16 pairs (two CTAs share
    // an SM, so the restart overhead of short segments costs throughput: the wave model of pick_zpairs measured worse)
    int zp = m2;
    while (zp > 16 && (long long)gx * gy * ((m2 + zp - 1) / zp) < 148 * 4) zp = (zp + 1) / 2;
    // coarse levels have too few tiles to occupy the machine: there the serial depth per CTA is what counts
    while (zp > 4 && (long long)gx * gy * ((m2 + zp - 1) / zp) < 148) zp = (zp + 1) / 2;
    a.zpairs = zp;
    dim3 grid(gx, gy, (m2 + zp - 1) / zp);
    const bool track = in_min != nullptr;
    // measurement variants: WRB_FWD_PITCH=72 (see the kernel's comment), WRB_FWD_IMPL=plain|pipe (two barriers per plane / the
    // software pipeline with one)
    const int pitch = [] { const char* e = getenv("WRB_FWD_PITCH"); return (e && atoi(e) == 72) ? 72 : FTX + 2; }();
    // 0: two barriers per plane, 1: the software pipeline, 2: the pipeline over the native-type tile
    const int impl = [] { const char* e = getenv("WRB_FWD_IMPL"); return (e && *e) ? (e[0] == 'n' ? 2 : (e[1] == 'i' ? 1 : 0)) : kFwdImplDefault; }();
    auto launch = [&](auto pitch_tag, auto pipe_tag, auto native_tag) {
        constexpr int PT = decltype(pitch_tag)::value;
        constexpr bool PP = decltype(pipe_tag)::value, NT = decltype(native_tag)::value;
        constexpr int smem_f = fused_fwd_smem<float, PP, PT, NT>(), smem_d = fused_fwd_smem<double, PP, PT, NT>();
        if (smem_d > 48 * 1024) {                                    // opt-in size: the attribute is per device and per kernel
            static DeviceOnce once;
            once.run([] {
                cudaFuncSetAttribute(fwd_level_fused_kernel<float, true, PT, PP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
                cudaFuncSetAttribute(fwd_level_fused_kernel<float, false, PT, PP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
                cudaFuncSetAttribute(fwd_level_fused_kernel<double, true, PT, PP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
                cudaFuncSetAttribute(fwd_level_fused_kernel<double, false, PT, PP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
            });
        }
        if (src_is_f32) {
            if (track) fwd_level_fused_kernel<float, true, PT, PP, NT><<<grid, FTHREADS, smem_f, s>>>(a);
            else fwd_level_fused_kernel<float, false, PT, PP, NT><<<grid, FTHREADS, smem_f, s>>>(a);
        } else {
            if (track) fwd_level_fused_kernel<double, true, PT, PP, NT><<<grid, FTHREADS, smem_d, s>>>(a);
            else fwd_level_fused_kernel<double, false, PT, PP, NT><<<grid, FTHREADS, smem_d, s>>>(a);
        }
    };
    using std::integral_constant;
    typedef integral_constant<bool, true> Yes;
    typedef integral_constant<bool, false> No;
    if (pitch == 72) launch(integral_constant<int, 72>{}, No{}, No{});
    else if (impl == 2) launch(integral_constant<int, FTX + 2>{}, Yes{}, Yes{});
    else if (impl == 1) launch(integral_constant<int, FTX + 2>{}, Yes{}, No{});
    else launch(integral_constant<int, FTX + 2>{}, No{}, No{});
    note_launch(1);
}

}  // namespace wrb
