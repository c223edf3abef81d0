// Correlation-pyramid lookup of ONE source pixel by ONE warp (core/corr.py:30-51) + the flow operands of the motion
// encoder.  Shared by lookup_kernel (kernels.cu) and by the persistent refinement kernel (conv_tc.cu), which runs the
// lookup as tiles of its dataflow program; both must produce the same bits, so the arithmetic below only uses
// operations the compiler cannot contract differently in the two translation units (explicit fmaf, no a*b+c).
//
// Per level the 81 sample points of a pixel are the 9x9 integer offsets of ONE position, so they share its fractional
// part and a 10x10 integer neighbourhood of the pixel's correlation row.  The warp evaluates the position once per
// level (the oracle's round trip on the first sample; the other samples' own round trips differ from "first sample
// + k" by ~1e-6 px, four orders of magnitude below the fp16 rounding of the output), stages the four neighbourhoods in
// shared memory (400 loads per pixel instead of 4 x 324) and blends.
#pragma once
#include "kernels.h"

namespace mftb {

constexpr int kLkWin = 10;
constexpr int kLkLevelFloats = kLkWin * kLkWin + 4;
constexpr int kLkWinFloats = 4 * kLkLevelFloats;      // shared-memory floats per pixel in flight

struct LookupPixel {
    float wE[4], wS[4];
    float cx, cy;
    bool finite;
};

// bilinear_sampler's normalise -> grid_sample(align_corners=True) round trip (core/utils/utils.py:98-106)
__device__ __forceinline__ float lk_roundtrip_div(float c, float size_m1) {
    const float g = (2.0f * c) / size_m1 - 1.0f;
    return ((g + 1.0f) * 0.5f) * size_m1;
}

// Phase 1: request the four 10x10 windows of pixel `pp` (= pair * h*w + n) into `win` (this warp's kLkWinFloats floats).
// coords1 may have been written earlier in the same launch by another CTA: read through L2.
__device__ __forceinline__ void lookup_gather(const LookupArgs& a, long pp, int lane, float* win, LookupPixel& px) {
    const float2 c = __ldcg(reinterpret_cast<const float2*>(a.coords1 + pp * 2));
    px.cx = c.x;
    px.cy = c.y;
    px.finite = isfinite(c.x) && isfinite(c.y);
    int hl = a.h, wl = a.w;
    float div = 1.0f;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const float fxp = lk_roundtrip_div(c.x / div - 4.0f, static_cast<float>(wl - 1));
        const float fyp = lk_roundtrip_div(c.y / div - 4.0f, static_cast<float>(hl - 1));
        const float fx = floorf(fxp), fy = floorf(fyp);
        px.wE[l] = fxp - fx;
        px.wS[l] = fyp - fy;
        const int X0 = px.finite ? static_cast<int>(fminf(fmaxf(fx, -32.0f), static_cast<float>(wl + 16))) : 0;
        const int Y0 = px.finite ? static_cast<int>(fminf(fmaxf(fy, -32.0f), static_cast<float>(hl + 16))) : 0;
        const float* base = a.lvl[l] + pp * static_cast<long>(hl) * wl;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = lane + 32 * k;
            if (e < kLkWin * kLkWin) {
                const int ey = e / kLkWin, ex = e - ey * kLkWin;
                const int gx = X0 + ex, gy = Y0 + ey;
                win[l * kLkLevelFloats + e] = (px.finite && gx >= 0 && gx < wl && gy >= 0 && gy < hl)
                                                  ? __ldg(base + static_cast<long>(gy) * wl + gx) : 0.0f;
            }
        }
        hl >>= 1; wl >>= 1; div *= 2.0f;
    }
}

// Phase 2 (after a __syncwarp): blend, write corr16 [324 + 4 pad], the 7x7x2 flow neighbourhood for convf1 and the flow
// channels of the GRU record.
__device__ __forceinline__ void lookup_emit(const LookupArgs& a, long pp, int lane, const float* win, const LookupPixel& px) {
    const int npx = a.h * a.w;
    const int n = static_cast<int>(pp % npx);
    const int y = n / a.w, x = n - y * a.w;
    __half* out = a.corr16 + pp * 328;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const float* W = win + l * kLkLevelFloats;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int o = lane + 32 * k;
            if (o < 81) {
                const int i = o / 9, j = o - i * 9;          // x offset index i (columns), y offset index j (rows)
                const float* q = W + j * kLkWin + i;
                const float vnw = q[0], vne = q[1], vsw = q[kLkWin], vse = q[kLkWin + 1];
                const float top = fmaf(px.wE[l], vne - vnw, vnw), bot = fmaf(px.wE[l], vse - vsw, vsw);
                const float r = px.finite ? fmaf(px.wS[l], bot - top, top) : NAN;
                out[l * 81 + o] = __float2half_rn(r);
            }
        }
    }
    if (lane < 4) out[324 + lane] = __float2half_rn(0.0f);
    // flow = coords1 - coords0 (core/raft.py:179); 7x7x2 zero-padded neighbourhood for convf1
    const long pbase = pp - n;
    __half* fp = a.flowpatch16 + pp * 104;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
        const int k = lane + 32 * k4;
        if (k < 104) {
            float v = 0.0f;
            if (k < 98) {
                const int c = k & 1, t = k >> 1, ky = t / 7, kx = t - ky * 7;
                const int yy = y + ky - 3, xx = x + kx - 3;
                if (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w)
                    v = __ldcg(a.coords1 + (pbase + static_cast<long>(yy) * a.w + xx) * 2 + c) - static_cast<float>(c == 0 ? xx : yy);
            }
            fp[k] = __float2half_rn(v);
        }
    }
    if (lane < 2) a.X[pp * 512 + 382 + lane] = __float2half_rn((lane == 0 ? px.cx : px.cy) - static_cast<float>(lane == 0 ? x : y));
}

}  // namespace mftb
