// Correlation-pyramid lookup of ONE source pixel by ONE warp (core/corr.py:30-51) + the flow operands of the motion
// encoder.  Shared by lookup_kernel (kernels.cu) and by the persistent refinement kernel (conv_tc.cu), which can run the
// lookup as tiles of its dataflow program; both must produce the same bits, so the arithmetic below only uses
// operations the compiler cannot contract differently in the two translation units (explicit fmaf, no a*b+c).
//
// Per level the 81 sample points of a pixel are the 9x9 integer offsets of ONE position, so they share its fractional
// part and a 10x10 integer neighbourhood of the pixel's correlation row.  The warp evaluates the position once per
// level (the oracle's round trip on the first sample; the other samples' own round trips differ from "first sample
// + k" by ~1e-6 px, four orders of magnitude below the fp16 rounding of the output), stages the four neighbourhoods in
// shared memory (400 loads per pixel instead of 4 x 324) and blends.
//
// The kernel is instruction-issue bound (about 1300 instructions per lane and pixel in its first form), so everything
// that only depends on the lane -- which window elements it fetches, which outputs it blends, which entries of the
// 7x7 flow patch it writes -- is tabulated once per warp (LookupLane) and reused for every pixel the warp handles.
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

// Per-lane constants.  Window element e = lane + 32k (k < 4, e < 100) sits at (ey, ex); output o = lane + 32k (k < 3,
// o < 81) blends the 2x2 window cell at woff = (o % 9) * 10 + o / 9 (x offset index o / 9 = columns, core/corr.py:37-40);
// flow-patch entry f = lane + 32k (k < 4, f < 98): channel f & 1 of tap (f >> 1) = ky * 7 + kx.
struct LookupLane {
    int ex[4], ey[4];
    int woff[3];
    int fdx[4], fdy[4];
};

__device__ __forceinline__ LookupLane lookup_lane_init(int lane) {
    // e -> e + 32 is (ey + 3, ex + 2), o -> o + 32 is (i + 3, j + 5), tap -> tap + 16 is (ky + 2, kx + 2), each with one carry:
    // cheaper than a division per entry (the tables are rebuilt for every pixel when a warp handles only one)
    LookupLane t;
    int ey = (lane * 26) >> 8, ex = lane - ey * kLkWin;              // lane / 10, lane % 10 (exact for lane < 69)
    int i = (lane * 57) >> 9, j = lane - i * 9;                        // lane / 9, lane % 9
    const int tap = lane >> 1;
    int ky = (tap * 37) >> 8, kx = tap - ky * 7;                       // tap / 7, tap % 7 (tap < 16)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t.ex[k] = ex; t.ey[k] = ey;
        t.fdx[k] = kx - 3; t.fdy[k] = ky - 3;
        if (k < 3) t.woff[k] = j * kLkWin + i;
        ex += 2; ey += 3;
        if (ex >= kLkWin) { ex -= kLkWin; ++ey; }
        j += 5; i += 3;
        if (j >= 9) { j -= 9; ++i; }
        kx += 2; ky += 2;
        if (kx >= 7) { kx -= 7; ++ky; }
    }
    return t;
}

// bilinear_sampler's normalise -> grid_sample(align_corners=True) round trip (core/utils/utils.py:98-106).  The quotient
// uses the 2-ulp fast division: the round trip itself only perturbs the position by ~1e-6 px, far below what the fp16
// output resolves.
__device__ __forceinline__ float lk_roundtrip_div(float c, float size_m1) {
    const float g = __fadd_rn(__fdividef(2.0f * c, size_m1), -1.0f);      // (no contraction: same bits in every translation unit)
    return ((g + 1.0f) * 0.5f) * size_m1;
}

// Phase 1: request the four 10x10 windows of pixel `pp` (= pair * h*w + n) into `win` (this warp's kLkWinFloats floats).
// coords1 may have been written earlier in the same launch by another CTA: read through L2.
__device__ __forceinline__ void lookup_gather(const LookupArgs& a, const LookupLane& t, long pp, int lane, float* win,
                                              LookupPixel& px) {
    const float2 c = __ldcg(reinterpret_cast<const float2*>(a.coords1 + pp * 2));
    px.cx = c.x;
    px.cy = c.y;
    px.finite = isfinite(c.x) && isfinite(c.y);
    // lane l & 3 evaluates level l & 3 (position round trip, fractions, window origin); the level loop then only
    // broadcasts the four numbers instead of every lane redoing all four levels
    const int myl = lane & 3;
    const int mh = a.h >> myl, mw = a.w >> myl;
    const float inv = 1.0f / static_cast<float>(1 << myl);          // 1 / 2^level: exact
    const float fxp = lk_roundtrip_div(c.x * inv - 4.0f, static_cast<float>(mw - 1));
    const float fyp = lk_roundtrip_div(c.y * inv - 4.0f, static_cast<float>(mh - 1));
    const float fx = floorf(fxp), fy = floorf(fyp);
    const float my_wE = fxp - fx, my_wS = fyp - fy;
    // non-finite coordinates: park the window outside the map, every tap then reads as zero (the output is NaN anyway)
    const int my_X0 = px.finite ? static_cast<int>(fminf(fmaxf(fx, -32.0f), static_cast<float>(mw + 16))) : -64;
    const int my_Y0 = px.finite ? static_cast<int>(fminf(fmaxf(fy, -32.0f), static_cast<float>(mh + 16))) : -64;
    int hl = a.h, wl = a.w;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        px.wE[l] = __shfl_sync(0xffffffffu, my_wE, l);
        px.wS[l] = __shfl_sync(0xffffffffu, my_wS, l);
        const int X0 = __shfl_sync(0xffffffffu, my_X0, l), Y0 = __shfl_sync(0xffffffffu, my_Y0, l);
        const __half* base = a.lvl[l] + pp * static_cast<long>(hl * wl);
        asm volatile("" : "+l"(base));       // keep the row pointer in registers (else it is re-derived from pp for every tap)
        // unconditional loads (taps outside the map read element 0 and are zeroed afterwards): the address of a predicated
        // load is recomputed under its predicate, which tripled the integer work of this loop
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < 3 || lane < kLkWin * kLkWin - 96) {
                const unsigned gx = static_cast<unsigned>(X0 + t.ex[k]), gy = static_cast<unsigned>(Y0 + t.ey[k]);
                const bool inside = gx < static_cast<unsigned>(wl) && gy < static_cast<unsigned>(hl);
                const unsigned off = inside ? gy * static_cast<unsigned>(wl) + gx : 0u;
                const float v = __half2float(__ldg(base + off));
                win[l * kLkLevelFloats + lane + 32 * k] = inside ? v : 0.0f;
            }
        }
        hl >>= 1; wl >>= 1;
    }
}

// Phase 2 (after a __syncwarp): blend, write corr16 [324 + 4 pad], the 7x7x2 flow neighbourhood for convf1 and the flow
// channels of the GRU record.
// (x, y) = the pixel's position in its frame (pp = pair * h*w + y*w + x).
__device__ __forceinline__ void lookup_emit(const LookupArgs& a, const LookupLane& t, long pp, int x, int y, int lane,
                                            const float* win, const LookupPixel& px) {
    const int n = y * a.w + x;
    __half* out = a.corr16 + pp * 328 + lane;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const float* W = win + l * kLkLevelFloats;
        const float wE = px.wE[l], wS = px.wS[l];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < 2 || lane < 81 - 64) {
                const float* q = W + t.woff[k];
                const float vnw = q[0], vne = q[1], vsw = q[kLkWin], vse = q[kLkWin + 1];
                const float top = fmaf(wE, vne - vnw, vnw), bot = fmaf(wE, vse - vsw, vsw);
                const float r = px.finite ? fmaf(wS, bot - top, top) : NAN;
                out[l * 81 + 32 * k] = __float2half_rn(r);
            }
        }
    }
    if (lane < 4) out[324] = __float2half_rn(0.0f);
    // flow = coords1 - coords0 (core/raft.py:179); 7x7x2 zero-padded neighbourhood for convf1
    const float* cbase = a.coords1 + (pp - n) * 2;
    __half* fp = a.flowpatch16 + pp * 104 + lane;
    const int ch = lane & 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < 3 || lane < 104 - 96) {
            const unsigned xx = static_cast<unsigned>(x + t.fdx[k]), yy = static_cast<unsigned>(y + t.fdy[k]);
            const bool inside = (k < 3 || lane < 98 - 96) && xx < static_cast<unsigned>(a.w) && yy < static_cast<unsigned>(a.h);
            const unsigned off = inside ? (yy * static_cast<unsigned>(a.w) + xx) * 2u + static_cast<unsigned>(ch) : 0u;
            float v = __ldcg(cbase + off) - static_cast<float>(ch == 0 ? xx : yy);
            v = inside ? v : 0.0f;
            fp[32 * k] = __float2half_rn(v);
        }
    }
    if (lane < 2) a.X[pp * 512 + 382 + lane] = __float2half_rn((lane == 0 ? px.cx : px.cy) - static_cast<float>(lane == 0 ? x : y));
}

}  // namespace mftb
