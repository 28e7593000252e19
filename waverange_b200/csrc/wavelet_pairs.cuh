// wavelet_pairs.cuh -- register-window evaluation of the CDF 9/7 lifting stages, shared by the
// line-pass kernels (wavelet.cu) and the fused per-level kernels (wavelet_fused.cu).
#pragma once
#include "wr_common.cuh"

namespace wrb {

// ------------------------------------------------------------------------------------------
// Forward lifting of output pairs [i0, i0+R) of a line of N > 1 samples.
// reference: waveletcdf97_3d.c:101-132 (split, phantom :109, stages :112-125, scale :128-132)
// ------------------------------------------------------------------------------------------
template <int R, class LD>
__device__ __forceinline__ void fwd_pairs(LD ld, int N, int i0, double (&so)[R], double (&dd)[R])
{
    const int M = (N + 1) >> 1;
    double s0[R + 4], d0[R + 3];
#pragma unroll
    for (int t = 0; t < R + 4; t++) {
        int j = i0 - 2 + t;
        s0[t] = (j >= 0 && j < M) ? ld(2 * j) : 0.0;
    }
#pragma unroll
    for (int t = 0; t < R + 3; t++) {
        int j = i0 - 2 + t;
        double v = 0.0;
        if (j >= 0 && 2 * j + 1 < N) v = ld(2 * j + 1);
        else if (t >= 1 && j == M - 1)            // N odd: phantom sample (:109)
            v = (s0[t - 1] * WRB_E0 + d0[t >= 1 ? t - 1 : 0] * WRB_E1) + s0[t] * WRB_E2;
        d0[t] = v;
    }
    double d1[R + 3], s1[R + 2], d2[R + 1];
#pragma unroll
    for (int t = 0; t < R + 3; t++) {
        int j = i0 - 2 + t;
        d1[t] = (j < M - 1) ? d0[t] + WRB_LA * (s0[t + 1] + s0[t]) : d0[t] + (WRB_LA * 2) * s0[t];
    }
#pragma unroll
    for (int t = 0; t < R + 2; t++) {
        int j = i0 - 1 + t;
        s1[t] = (j == 0) ? s0[t + 1] + (WRB_LB * 2) * d1[t + 1] : s0[t + 1] + WRB_LB * (d1[t + 1] + d1[t]);
    }
#pragma unroll
    for (int t = 0; t < R + 1; t++) {
        int j = i0 - 1 + t;
        d2[t] = (j < M - 1) ? d1[t + 1] + WRB_LC * (s1[t + 1] + s1[t]) : d1[t + 1] + (WRB_LC * 2) * s1[t];
    }
#pragma unroll
    for (int t = 0; t < R; t++) {
        int j = i0 + t;
        double s2 = (j == 0) ? s1[t + 1] + (WRB_LD * 2) * d2[t + 1] : s1[t + 1] + WRB_LD * (d2[t + 1] + d2[t]);
        so[t] = s2 * WRB_SCL;
        dd[t] = d2[t + 1] * WRB_PSCL;
    }
}

// Interior variant of fwd_pairs: valid when 2 <= i0 and i0 + R + 1 <= M - 1 (no line end within
// reach) -- no guards, one formula per stage; bit-identical to fwd_pairs there.
template <int R, class LD>
__device__ __forceinline__ void fwd_pairs_interior(LD ld, int i0, double (&so)[R], double (&dd)[R])
{
    double s0[R + 4], d0[R + 3];
#pragma unroll
    for (int t = 0; t < R + 4; t++) s0[t] = ld(2 * (i0 - 2 + t));
#pragma unroll
    for (int t = 0; t < R + 3; t++) d0[t] = ld(2 * (i0 - 2 + t) + 1);
    double d1[R + 3], s1[R + 2], d2[R + 1];
#pragma unroll
    for (int t = 0; t < R + 3; t++) d1[t] = d0[t] + WRB_LA * (s0[t + 1] + s0[t]);
#pragma unroll
    for (int t = 0; t < R + 2; t++) s1[t] = s0[t + 1] + WRB_LB * (d1[t + 1] + d1[t]);
#pragma unroll
    for (int t = 0; t < R + 1; t++) d2[t] = d1[t + 1] + WRB_LC * (s1[t + 1] + s1[t]);
#pragma unroll
    for (int t = 0; t < R; t++) {
        const double s2 = s1[t + 1] + WRB_LD * (d2[t + 1] + d2[t]);
        so[t] = s2 * WRB_SCL;
        dd[t] = d2[t + 1] * WRB_PSCL;
    }
}

// ------------------------------------------------------------------------------------------
// Inverse lifting: samples x[2j], x[2j+1] for j in [i0, i0+R) of a line of M > 1 coefficients
// stored low half [0,Q) then high half [Q,M).   reference: waveletcdf97_3d.c:311-337
// ------------------------------------------------------------------------------------------
template <int R, class LD>
__device__ __forceinline__ void inv_pairs(LD ld, int M, int i0, double (&ev)[R], double (&od)[R])
{
    const int Q = (M + 1) >> 1, NH = M - Q;
    double h[R + 4], l[R + 3];
#pragma unroll
    for (int t = 0; t < R + 4; t++) {
        int j = i0 - 2 + t;
        h[t] = (j >= 0 && j < NH) ? ld(Q + j) * WRB_SCL : 0.0;      // phantom detail = 0 (:314)
    }
#pragma unroll
    for (int t = 0; t < R + 3; t++) {
        int j = i0 - 1 + t;
        l[t] = (j >= 0 && j < Q) ? ld(j) * WRB_PSCL : 0.0;
    }
    double s1[R + 3], d1[R + 2], s2[R + 1];
#pragma unroll
    for (int t = 0; t < R + 3; t++) {
        int j = i0 - 1 + t;
        s1[t] = (j == 0) ? l[t] - (WRB_LD * 2) * h[t + 1] : l[t] - WRB_LD * (h[t + 1] + h[t]);
    }
#pragma unroll
    for (int t = 0; t < R + 2; t++) {
        int j = i0 - 1 + t;
        d1[t] = (j < Q - 1) ? h[t + 1] - WRB_LC * (s1[t + 1] + s1[t]) : h[t + 1] - (WRB_LC * 2) * s1[t];
    }
#pragma unroll
    for (int t = 0; t < R + 1; t++) {
        int j = i0 + t;
        s2[t] = (j == 0) ? s1[t + 1] - (WRB_LB * 2) * d1[t + 1] : s1[t + 1] - WRB_LB * (d1[t + 1] + d1[t]);
    }
#pragma unroll
    for (int t = 0; t < R; t++) {
        int j = i0 + t;
        ev[t] = s2[t];
        od[t] = (j < Q - 1) ? d1[t + 1] - WRB_LA * (s2[t + 1] + s2[t]) : d1[t + 1] - (WRB_LA * 2) * s2[t];
    }
}

}  // namespace wrb
