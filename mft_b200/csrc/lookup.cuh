// Correlation-pyramid lookup (core/corr.py:30-51) + the flow operands of the motion encoder, for a GROUP of kLkGroup
// consecutive source pixels per warp pass.  Shared by lookup_kernel (kernels.cu) and by the persistent refinement kernel
// (conv_tc.cu), which can run the lookup as tiles of its dataflow program; both must produce the same bits, so the
// arithmetic below only uses operations the compiler cannot contract differently in the two translation units
// (explicit fmaf, no a*b+c).
//
// Per level the 81 sample points of a pixel are the 9x9 integer offsets of ONE position, so they share its fractional
// part and a 10x10 integer neighbourhood of the pixel's correlation row.  The warp evaluates the position once per
// (pixel, level) (the oracle's round trip on the first sample; the other samples' own round trips differ from "first
// sample + k" by ~1e-6 px, four orders of magnitude below the fp16 rounding of the output).
//
// Round 1 ran one pixel per warp, one fp16 element per load and one of the 81 outputs per lane and was instruction-issue
// bound (808 warp instructions per pixel, 35 us per iteration at 512^2).  Now, level by level for the whole group:
//   gather   lane task = (pixel, window row, aligned chunk of VEC = 4 | 2 | 1 fp16 elements): 8-byte loads where the
//            level's width allows it; a chunk lies either wholly inside the map or wholly outside (zero), so there is
//            no per-element bounds handling; the chunk lands in shared memory as fp32 [pixel][row][16 | 12 | 10];
//   blend    lane task = (pixel, output column i): separable -- 10 horizontal lerps H[row] = lerp(W[row][i], W[row][i+1])
//            down the column, 9 vertical lerps between consecutive rows (171 lerps per pixel and level instead of 243,
//            2 shared-memory reads per row instead of 4 per output), the column's 9 outputs are consecutive channels
//            (core/corr.py:37-40: x offset major) and leave as packed fp16 words.
// The loads of level l+1 are issued before level l is blended, so their latency overlaps the arithmetic.  The blend is
// bit-identical to the per-output form lerp(lerp(nw, ne), lerp(sw, se)) of round 1.
#pragma once
#include "kernels.h"

namespace mftb {

constexpr int kLkGroup = 4;                              // pixels per warp pass
constexpr int kLkWinFloats = kLkGroup * 10 * 16;         // one level's windows of a group: at most 2560 bytes per warp

// Per-lane constants of the 7x7x2 flow patch: entry f = lane + 32k (k < 4, f < 98): channel f & 1 of tap (f >> 1) = ky * 7 + kx.
struct LookupLane {
    int fdx[4], fdy[4];
};

__device__ __forceinline__ LookupLane lookup_lane_init(int lane) {
    // tap -> tap + 16 is (ky + 2, kx + 2) with one carry: cheaper than a division per entry
    LookupLane t;
    const int tap = lane >> 1;
    int ky = (tap * 37) >> 8, kx = tap - ky * 7;                       // tap / 7, tap % 7 (tap < 16)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t.fdx[k] = kx - 3; t.fdy[k] = ky - 3;
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

// Geometry of one level's gather for vector width VEC: a window row is covered by kPerRow aligned chunks of VEC elements;
// task T = (pixel p, row, chunk) = ((p * 10 + row) * kPerRow + chunk) lands at float T * VEC of the window buffer, i.e.
// the buffer is [pixel][row][kPitch = kPerRow * VEC] and needs no address arithmetic on the store side.
template <int VEC>
struct LkChunks {
    static constexpr int kPerRow = VEC == 4 ? 4 : (VEC == 2 ? 6 : 10);
    static constexpr int kPitch = kPerRow * VEC;                                // 16 | 12 | 10 floats per window row
    static constexpr int kTasks = kLkGroup * 10 * kPerRow;
    static constexpr int kRounds = (kTasks + 31) / 32;
    static constexpr int kWords = VEC == 4 ? 2 : 1;                             // 32-bit registers per chunk
};
// One register file for the three widths (a level uses exactly one of them): 13 words cover the worst case (VEC = 1).
struct LkRegs {
    unsigned w[13];
};

// Issues the loads of level `l`.  lane (p * 4 + l) holds X0 / Y0 of (pixel p, level l); base0 = the level's image of pixel pp0.
template <int VEC>
__device__ __forceinline__ void lookup_load_level(const __half* __restrict__ base0, int hl, int wl, unsigned valid_mask, int l,
                                                  int myX0, int myY0, int lane, LkRegs& c) {
    using C = LkChunks<VEC>;
    const int img = hl * wl;                                                    // (4 images: 32-bit offsets)
#pragma unroll
    for (int r = 0; r < C::kRounds; ++r) {
        const int T = r * 32 + lane;
        int p = T / (10 * C::kPerRow);
        const int t = T - p * (10 * C::kPerRow);
        const int row = t / C::kPerRow, ch = t - row * C::kPerRow;
        p = p < kLkGroup ? p : kLkGroup - 1;                                    // (tail lanes of the last round: clamped, masked below)
        const int X0 = __shfl_sync(0xffffffffu, myX0, p * 4 + l), Y0 = __shfl_sync(0xffffffffu, myY0, p * 4 + l);
        const int a = VEC == 4 ? (X0 & ~3) : (VEC == 2 ? (X0 & ~1) : X0);
        const int gx = a + ch * VEC, gy = Y0 + row;
        const bool inside = T < C::kTasks && ((valid_mask >> p) & 1u) && static_cast<unsigned>(gx) < static_cast<unsigned>(wl) &&
                            static_cast<unsigned>(gy) < static_cast<unsigned>(hl);
        uint2 v = make_uint2(0u, 0u);
        if (inside) {
            const __half* q = base0 + (p * img + gy * wl + gx);
            if constexpr (VEC == 4) v = __ldg(reinterpret_cast<const uint2*>(q));
            else if constexpr (VEC == 2) v.x = __ldg(reinterpret_cast<const unsigned*>(q));
            else v.x = __ldg(reinterpret_cast<const unsigned short*>(q));
        }
        c.w[r * C::kWords] = v.x;
        if constexpr (VEC == 4) c.w[r * C::kWords + 1] = v.y;
    }
}

// Converts the chunks to fp32 and stores them into the group's window buffer [pixel][row][kPitch].
template <int VEC, int PITCH = LkChunks<VEC>::kPitch>
__device__ __forceinline__ void lookup_store_level(const LkRegs& c, int lane, float* win) {
    using C = LkChunks<VEC>;
#pragma unroll
    for (int r = 0; r < C::kRounds; ++r) {
        const int T = r * 32 + lane;
        if (T < C::kTasks) {
            float* dst = win + (PITCH == C::kPitch ? T * VEC : (T / C::kPerRow) * PITCH + (T % C::kPerRow) * VEC);
            if constexpr (VEC == 4) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&c.w[2 * r]));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&c.w[2 * r + 1]));
                *reinterpret_cast<float4*>(dst) = make_float4(f0.x, f0.y, f1.x, f1.y);
            } else if constexpr (VEC == 2) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&c.w[r]));
                *reinterpret_cast<float2*>(dst) = f0;
            } else {
                const unsigned short h = static_cast<unsigned short>(c.w[r]);
                dst[0] = __half2float(*reinterpret_cast<const __half*>(&h));
            }
        }
    }
}

// channels 81 l + 9 i + (j0 ..): n consecutive halves at a 2-byte aligned offset -> 4-byte words plus a leading / trailing half
__device__ __forceinline__ void lk_store9(__half* dst, const float (&o)[9], bool word_aligned) {
    if (word_aligned) {
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<__half2*>(dst + 2 * k) = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
        dst[8] = __float2half_rn(o[8]);
    } else {
        dst[0] = __float2half_rn(o[0]);
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<__half2*>(dst + 1 + 2 * k) = __floats2half2_rn(o[1 + 2 * k], o[2 + 2 * k]);
    }
}

// Blends level `l` of the group out of `win` (row pitch `pitch` floats) and writes channels [81 l, 81 l + 81) of corr16.
// lane (p * 4 + l) holds wE / wS / the window's column offset within its aligned span / finiteness of (pixel p, level l).
// Round A: lane = (pixel p, output column i < 8): the whole column, 10 horizontal + 9 vertical lerps.  Round B: the ninth
// column of the four pixels, one output per lane (36 outputs: lanes 0..31, then lanes 0..3) -- a second column round would
// keep 4 of 32 lanes busy.
// PITCH: floats per window row (the gather's chunk geometry by default; the TMA variant uses 20: with 16 the four pixels of
// round A and the rows of round B fall into the same shared-memory banks, 4- to 5-way conflicts).
template <int VEC, int PITCH = LkChunks<VEC>::kPitch>
__device__ __forceinline__ void lookup_blend_level(__half* __restrict__ corr16, long pp0, unsigned valid_mask, int l, float my_wE, float my_wS,
                                                   int my_off, int my_finite, int lane, const float* win) {
    constexpr int pitch = PITCH;
    {
        const int p = lane >> 3, i = lane & 7;
        const int src = p * 4 + l;
        const float wE = __shfl_sync(0xffffffffu, my_wE, src), wS = __shfl_sync(0xffffffffu, my_wS, src);
        const int off = __shfl_sync(0xffffffffu, my_off, src), fin = __shfl_sync(0xffffffffu, my_finite, src);
        if ((valid_mask >> p) & 1u) {
            const float* q = win + p * 10 * pitch + off + i;
            float o[9];
            float prev = 0.0f;
#pragma unroll
            for (int row = 0; row < 10; ++row) {
                const float vw = q[row * pitch], ve = q[row * pitch + 1];
                const float h = fmaf(wE, ve - vw, vw);
                if (row > 0) o[row - 1] = fin ? fmaf(wS, h - prev, prev) : NAN;
                prev = h;
            }
            lk_store9(corr16 + (pp0 + p) * 328 + (l * 81 + i * 9), o, ((l + i) & 1) == 0);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int T = r * 32 + lane;                      // (pixel p, output row j) of column 8
        int p = (T * 57) >> 9;                            // T / 9 (T < 64)
        const int j = T - p * 9;
        const bool active = T < kLkGroup * 9;
        p = p < kLkGroup ? p : kLkGroup - 1;
        const int src = p * 4 + l;
        const float wE = __shfl_sync(0xffffffffu, my_wE, src), wS = __shfl_sync(0xffffffffu, my_wS, src);
        const int off = __shfl_sync(0xffffffffu, my_off, src), fin = __shfl_sync(0xffffffffu, my_finite, src);
        if (active && ((valid_mask >> p) & 1u)) {
            const float* q = win + (p * 10 + j) * pitch + off + 8;
            const float top = fmaf(wE, q[1] - q[0], q[0]), bot = fmaf(wE, q[pitch + 1] - q[pitch], q[pitch]);
            const float v = fin ? fmaf(wS, bot - top, top) : NAN;
            corr16[(pp0 + p) * 328 + (l * 81 + 72 + j)] = __float2half_rn(v);
        }
    }
}

__device__ __forceinline__ int lookup_vec(int wl) { return (wl & 3) == 0 ? 4 : ((wl & 1) == 0 ? 2 : 1); }

// The whole lookup of pixels pp0 .. pp0 + nvalid - 1 (global pixel indices pair * h*w + n; nvalid <= kLkGroup) by one warp.
// `win`: this warp's kLkGroup * 10 * PITCH4 floats of shared memory; PITCH4 = floats per window row for the 8-byte-chunk
// geometry (widths that are multiples of 4): 16 = dense, 20 = free of bank conflicts in the blend (25 % more shared memory).  coords1 may have been written earlier in the same launch by
// another CTA: read through L2.
template <int PITCH4 = 16>
__device__ __forceinline__ void lookup_group(const LookupArgs& a, const LookupLane& t, long pp0, int nvalid, int lane, float* win) {
    const unsigned valid_mask = (1u << nvalid) - 1u;
    const int npx = a.h * a.w;
    // ---- set-up: lane p * 4 + l evaluates (pixel p, level l): position round trip, fractions, window origin -------------------
    const int sp = (lane >> 2) & (kLkGroup - 1), sl = lane & 3;
    const bool sv = sp < nvalid;
    const long spp = pp0 + (sv ? sp : 0);
    const float2 c = __ldcg(reinterpret_cast<const float2*>(a.coords1 + spp * 2));
    const int my_finite = (isfinite(c.x) && isfinite(c.y)) ? 1 : 0;
    const int mh = a.h >> sl, mw = a.w >> sl;
    const float inv = 1.0f / static_cast<float>(1 << sl);            // 1 / 2^level: exact
    const float fxp = lk_roundtrip_div(c.x * inv - 4.0f, static_cast<float>(mw - 1));
    const float fyp = lk_roundtrip_div(c.y * inv - 4.0f, static_cast<float>(mh - 1));
    const float fx = floorf(fxp), fy = floorf(fyp);
    const float my_wE = fxp - fx, my_wS = fyp - fy;
    // non-finite coordinates: park the window outside the map, every tap then reads as zero (the output is NaN anyway)
    const int my_X0 = my_finite ? static_cast<int>(fminf(fmaxf(fx, -32.0f), static_cast<float>(mw + 16))) : -64;
    const int my_Y0 = my_finite ? static_cast<int>(fminf(fmaxf(fy, -32.0f), static_cast<float>(mh + 16))) : -64;
    // (a level is addressed by its row pitch: the pad columns hold zeros, exactly what an out-of-map tap reads)
    const int vec_mine = lookup_vec(sl == 0 ? a.pitch[0] : (sl == 1 ? a.pitch[1] : (sl == 2 ? a.pitch[2] : a.pitch[3])));
    const int my_off = my_X0 - (vec_mine == 4 ? (my_X0 & ~3) : (vec_mine == 2 ? (my_X0 & ~1) : my_X0));

    // ---- levels: the loads of level l + 1 are in flight while level l is blended -------------------------------------------------
    // (the three vector widths are separate instantiations; a level's width is warp-uniform)
    LkRegs regs;
    auto load = [&](int l) {
        const int hl = a.h >> l;
        // (no dynamic indexing of the kernel parameter: that would force a local-memory copy of the whole struct)
        const __half* lv = l == 0 ? a.lvl[0] : (l == 1 ? a.lvl[1] : (l == 2 ? a.lvl[2] : a.lvl[3]));
        const int wl = l == 0 ? a.pitch[0] : (l == 1 ? a.pitch[1] : (l == 2 ? a.pitch[2] : a.pitch[3]));       // row pitch
        const __half* base0 = lv + pp0 * (static_cast<long>(hl) * wl);
        const int v = lookup_vec(wl);
        if (v == 4) lookup_load_level<4>(base0, hl, wl, valid_mask, l, my_X0, my_Y0, lane, regs);
        else if (v == 2) lookup_load_level<2>(base0, hl, wl, valid_mask, l, my_X0, my_Y0, lane, regs);
        else lookup_load_level<1>(base0, hl, wl, valid_mask, l, my_X0, my_Y0, lane, regs);
    };
    load(0);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const int v = lookup_vec(l == 0 ? a.pitch[0] : (l == 1 ? a.pitch[1] : (l == 2 ? a.pitch[2] : a.pitch[3])));
        __syncwarp();                      // the previous level's blend is done with the buffer
        if (v == 4) lookup_store_level<4, PITCH4>(regs, lane, win);
        else if (v == 2) lookup_store_level<2>(regs, lane, win);
        else lookup_store_level<1>(regs, lane, win);
        if (l < 3) load(l + 1);
        __syncwarp();
        if (v == 4) lookup_blend_level<4, PITCH4>(a.corr16, pp0, valid_mask, l, my_wE, my_wS, my_off, my_finite, lane, win);
        else if (v == 2) lookup_blend_level<2>(a.corr16, pp0, valid_mask, l, my_wE, my_wS, my_off, my_finite, lane, win);
        else lookup_blend_level<1>(a.corr16, pp0, valid_mask, l, my_wE, my_wS, my_off, my_finite, lane, win);
    }
    // ---- per pixel: zero pad of corr16, the 7x7x2 zero-padded flow neighbourhood for convf1 (flow = coords1 - coords0,
    //      core/raft.py:179) and the flow channels of the GRU record ------------------------------------------------------------------
    for (int p = 0; p < nvalid; ++p) {
        const long pp = pp0 + p;
        const int n = static_cast<int>(static_cast<unsigned long>(pp) % static_cast<unsigned>(npx));
        const int y = n / a.w, x = n - y * a.w;
        if (lane < 4) a.corr16[pp * 328 + 324 + lane] = __float2half_rn(0.0f);
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
        if (lane < 2) {
            const float cc = __ldcg(a.coords1 + pp * 2 + lane);
            a.X[pp * 512 + 382 + lane] = __float2half_rn(cc - static_cast<float>(lane == 0 ? x : y));
        }
    }
}

}  // namespace mftb
