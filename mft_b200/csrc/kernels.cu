// Bandwidth-bound kernels of the MFT hot path.  Compiled with --fmad=false: the chaining /
// selection / lookup arithmetic follows a defined fp32 operation order (one rounding per
// operation, no contraction) so that results are bit-identical to oracle/mft_oracle.py.
#include "kernels.h"
#include "lookup.cuh"
#include "ptx.cuh"

#include <cmath>

namespace mftb {

// Programmatic dependent launch: every kernel of the path is launched with the stream-serialisation attribute,
// signals its dependents at entry and waits for its prerequisites before touching memory, so launch latency and
// CTA start-up overlap the tail of the previous kernel.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ==========================================================================================
// shared bilinear helper: the reference feeds pixel coordinates through a normalise ->
// grid_sample(align_corners=True) round trip (MFT/utils/interpolation.py:63-73 and
// MFT/RAFT/core/utils/utils.py:98-106); ATen undoes it as ((g+1)/2)*(size-1).
// ==========================================================================================
// (x / 2 is written x * 0.5f: bit-identical in IEEE fp32 and avoids the division sequence.)
__device__ __forceinline__ float roundtrip_mul(float c, float scale, float size_m1) {
    const float g = c * scale - 1.0f;                    // normalize_coords: x*(2/(W-1)) - 1
    return ((g + 1.0f) * 0.5f) * size_m1;
}
__device__ __forceinline__ float roundtrip_div(float c, float size_m1) {
    const float g = (2.0f * c) / size_m1 - 1.0f;         // bilinear_sampler: 2*x/(W-1) - 1
    return ((g + 1.0f) * 0.5f) * size_m1;
}

// ==========================================================================================
// chain + select  (MFT/MFT.py:114-142,233-239 ; MFT/results.py:87-136,250-265)
// ==========================================================================================
__global__ void __launch_bounds__(256)
chain_select_kernel(const ChainSelectArgs a, const float sx, const float sy) {
    pdl_enter();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= a.W) return;
    const long hw = static_cast<long>(a.H) * a.W;
    const long p = static_cast<long>(y) * a.W + x;
    const float gx = static_cast<float>(x), gy = static_cast<float>(y);
    const float wm1 = static_cast<float>(a.W - 1), hm1 = static_cast<float>(a.H - 1);
    const float ninf = -INFINITY;

    float bfx = 0.f, bfy = 0.f, bocc = 0.f, bsig = 0.f, bscore = 0.f;
    int bidx = 0;
    // fully unrolled with a uniform guard: the loads of the next chains are independent of the running selection,
    // so the compiler can keep several chains' gathers in flight (the kernel is latency bound, not bandwidth bound)
#pragma unroll
    for (int k = 0; k < kMaxChains; ++k) {
        if (k >= a.K) break;
        const float* L = a.left[k];
        const float lfx = __ldg(L + p), lfy = __ldg(L + hw + p), locc = __ldg(L + 2 * hw + p), lsig = __ldg(L + 3 * hw + p);
        const float px = gx + lfx, py = gy + lfy;
        const float ix = roundtrip_mul(px, sx, wm1), iy = roundtrip_mul(py, sy, hm1);
        float s0, s1, s2, s3;
        if (!(isfinite(ix) && isfinite(iy))) {
            s0 = s1 = s2 = s3 = NAN;
        } else {
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float wE = ix - x0f, wW = 1.0f - wE, wS = iy - y0f, wN = 1.0f - wS;
            const int x0 = static_cast<int>(fminf(fmaxf(x0f, -2.0f), static_cast<float>(a.W + 1)));
            const int y0 = static_cast<int>(fminf(fmaxf(y0f, -2.0f), static_cast<float>(a.H + 1)));
            const bool xw = x0 >= 0 && x0 < a.W, xe = x0 + 1 >= 0 && x0 + 1 < a.W;
            const bool yn = y0 >= 0 && y0 < a.H, ys = y0 + 1 >= 0 && y0 + 1 < a.H;
            const float* R = a.right + static_cast<long>(k) * 4 * hw;
            const long onw = static_cast<long>(y0) * a.W + x0;
            const float cnw = wW * wN, cne = wE * wN, csw = wW * wS, cse = wE * wS;
            float s[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* Rc = R + c * hw;
                const float vnw = (xw && yn) ? __ldg(Rc + onw) : 0.0f;
                const float vne = (xe && yn) ? __ldg(Rc + onw + 1) : 0.0f;
                const float vsw = (xw && ys) ? __ldg(Rc + onw + a.W) : 0.0f;
                const float vse = (xe && ys) ? __ldg(Rc + onw + a.W + 1) : 0.0f;
                float o = cnw * vnw;
                o = o + cne * vne;
                o = o + csw * vsw;
                o = o + cse * vse;
                s[c] = o;
            }
            s0 = s[0]; s1 = s[1]; s2 = s[2]; s3 = s[3];
        }
        const float fx = (px + s0) - gx;
        const float fy = (py + s1) - gy;
        // torch.maximum / np.maximum propagate NaN
        const float occ = (isnan(locc) || isnan(s2)) ? NAN : fmaxf(locc, s2);
        const float sig = sqrtf(lsig * lsig + s3 * s3);
        const float score = (occ > a.occlusion_threshold) ? ninf : -sig;
        const bool take = (k == 0) || (score > bscore) || (isnan(score) && !isnan(bscore));
        if (take) {
            bfx = fx; bfy = fy; bocc = occ; bsig = sig; bscore = score; bidx = k;
        }
    }
    const float ex = gx + bfx, ey = gy + bfy;
    if (ex < 0.0f || ey < 0.0f || ex >= static_cast<float>(a.W) || ey >= static_cast<float>(a.H)) bocc = 1.0f;
    a.out[p] = bfx;
    a.out[hw + p] = bfy;
    a.out[2 * hw + p] = bocc;
    a.out[3 * hw + p] = bsig;
    if (a.index != nullptr) a.index[p] = static_cast<uint8_t>(bidx);
}

cudaError_t launch_chain_select(const ChainSelectArgs& a, cudaStream_t stream) {
    const float sx = static_cast<float>(2.0 / (a.W - 1));
    const float sy = static_cast<float>(2.0 / (a.H - 1));
    dim3 grid((a.W + 255) / 256, a.H);
    return launch_pdl(chain_select_kernel, grid, dim3(256), 0, stream, a, sx, sy);
}

// ==========================================================================================
// warp_backward / chain / point sampling (MFT/results.py:87-188)
// ==========================================================================================
__device__ __forceinline__ float bilinear_zero_dev(const float* __restrict__ img, int H, int W, float ix, float iy) {
    if (!(isfinite(ix) && isfinite(iy))) return NAN;
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wE = ix - x0f, wW = 1.0f - wE, wS = iy - y0f, wN = 1.0f - wS;
    const int x0 = static_cast<int>(fminf(fmaxf(x0f, -2.0f), static_cast<float>(W + 1)));
    const int y0 = static_cast<int>(fminf(fmaxf(y0f, -2.0f), static_cast<float>(H + 1)));
    const bool xw = x0 >= 0 && x0 < W, xe = x0 + 1 >= 0 && x0 + 1 < W;
    const bool yn = y0 >= 0 && y0 < H, ys = y0 + 1 >= 0 && y0 + 1 < H;
    const long onw = static_cast<long>(y0) * W + x0;
    const float vnw = (xw && yn) ? __ldg(img + onw) : 0.0f;
    const float vne = (xe && yn) ? __ldg(img + onw + 1) : 0.0f;
    const float vsw = (xw && ys) ? __ldg(img + onw + W) : 0.0f;
    const float vse = (xe && ys) ? __ldg(img + onw + W + 1) : 0.0f;
    float o = (wW * wN) * vnw;
    o = o + (wE * wN) * vne;
    o = o + (wW * wS) * vsw;
    o = o + (wE * wS) * vse;
    return o;
}

__global__ void __launch_bounds__(256)
warp_backward_kernel(const float* __restrict__ flow, const float* __restrict__ img, int C, int H, int W, int add_flow,
                     float* __restrict__ out, float sx, float sy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const long hw = static_cast<long>(H) * W, p = static_cast<long>(y) * W + x;
    const float gx = static_cast<float>(x), gy = static_cast<float>(y);
    const float px = gx + flow[p], py = gy + flow[hw + p];
    const float ix = roundtrip_mul(px, sx, static_cast<float>(W - 1)), iy = roundtrip_mul(py, sy, static_cast<float>(H - 1));
    for (int c = 0; c < C; ++c) {
        float v = bilinear_zero_dev(img + c * hw, H, W, ix, iy);
        if (add_flow) v = ((c == 0 ? px : py) + v) - (c == 0 ? gx : gy);
        out[c * hw + p] = v;
    }
}

cudaError_t launch_warp_backward(const float* flow, const float* img, int C, int H, int W, int add_flow, float* out,
                          cudaStream_t stream) {
    dim3 grid((W + 255) / 256, H);
    warp_backward_kernel<<<grid, 256, 0, stream>>>(flow, img, C, H, W, add_flow, out, static_cast<float>(2.0 / (W - 1)),
                                                   static_cast<float>(2.0 / (H - 1)));
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
sample_points_kernel(const float* __restrict__ field, int C, int H, int W, const float* __restrict__ pts, int N,
                     int add_points, float* __restrict__ out, float sx, float sy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float px = pts[2 * i], py = pts[2 * i + 1];
    const float ix = roundtrip_mul(px, sx, static_cast<float>(W - 1)), iy = roundtrip_mul(py, sy, static_cast<float>(H - 1));
    const long hw = static_cast<long>(H) * W;
    for (int c = 0; c < C; ++c) {
        float v = bilinear_zero_dev(field + c * hw, H, W, ix, iy);
        if (add_points && c < 2) v = (c == 0 ? px : py) + v;
        out[static_cast<long>(c) * N + i] = v;
    }
}

cudaError_t launch_sample_points(const float* field, int C, int H, int W, const float* points_xy, int N, int add_points,
                          float* out, cudaStream_t stream) {
    if (N <= 0) return cudaSuccess;
    sample_points_kernel<<<(N + 255) / 256, 256, 0, stream>>>(field, C, H, W, points_xy, N, add_points, out,
                                                              static_cast<float>(2.0 / (W - 1)),
                                                              static_cast<float>(2.0 / (H - 1)));
    return cudaGetLastError();
}

// ==========================================================================================
// forward splat (MFT/results.py:190-248, MFT/utils/interpolation.py:234-309)
// ==========================================================================================
__global__ void __launch_bounds__(256)
splat_scatter_kernel(const float* __restrict__ flow, const float* __restrict__ img, const uint8_t* __restrict__ mask, int C,
                     int H, int W, float* __restrict__ accum, float* __restrict__ counts) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const long hw = static_cast<long>(H) * W, p = static_cast<long>(y) * W + x;
    if (mask != nullptr && mask[p] == 0) return;
    float px = static_cast<float>(x) + flow[p], py = static_cast<float>(y) + flow[hw + p];
    if (!(isfinite(px) && isfinite(py))) return;
    // corner indices from the UNCLAMPED floor, then position and corners clamped into the grid (interpolation.py:256-268)
    const float fx0 = fminf(fmaxf(floorf(px), -2.0f), static_cast<float>(W + 1));
    const float fy0 = fminf(fmaxf(floorf(py), -2.0f), static_cast<float>(H + 1));
    int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0), x1 = x0 + 1, y1 = y0 + 1;
    px = fminf(fmaxf(px, 0.0f), static_cast<float>(W - 1));
    py = fminf(fmaxf(py, 0.0f), static_cast<float>(H - 1));
    x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
    y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
    const float ax = static_cast<float>(x1) - px, bx = px - static_cast<float>(x0);
    const float ay = static_cast<float>(y1) - py, by = py - static_cast<float>(y0);
    const float w[4] = {ax * ay, ax * by, bx * ay, bx * by};
    const long cell[4] = {static_cast<long>(y0) * W + x0, static_cast<long>(y1) * W + x0, static_cast<long>(y0) * W + x1,
                          static_cast<long>(y1) * W + x1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (w[k] == 0.0f) continue;
        atomicAdd(counts + cell[k], w[k]);
        for (int c = 0; c < C; ++c) atomicAdd(accum + cell[k] * C + c, img[p * C + c] * w[k]);
    }
}

__global__ void __launch_bounds__(256)
splat_normalize_kernel(float* __restrict__ out, const float* __restrict__ counts, int C, long cells, int use_border, float border) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= cells * C) return;
    const float n = counts[i / C];
    out[i] = n > 0.0f ? out[i] / n : (use_border ? border : out[i]);
}

cudaError_t launch_warp_forward(const float* flow, const float* img, const uint8_t* mask, int C, int H, int W, int use_border,
                         float border, float* out, float* counts, cudaStream_t stream) {
    const long cells = static_cast<long>(H) * W;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * cells * C, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, sizeof(float) * cells, stream);
    if (e != cudaSuccess) return e;
    splat_scatter_kernel<<<dim3((W + 255) / 256, H), 256, 0, stream>>>(flow, img, mask, C, H, W, out, counts);
    splat_normalize_kernel<<<static_cast<unsigned>((cells * C + 255) / 256), 256, 0, stream>>>(out, counts, C, cells, use_border, border);
    return cudaGetLastError();
}

// ==========================================================================================
// frame -> im2col patches for the 7x7/2 first conv (MFT/raft.py:41-48, core/raft.py:122-124)
// ==========================================================================================
// one thread = 8 consecutive patch entries (one 16-byte store); 152 = 19 x 8 entries per output pixel
__global__ void __launch_bounds__(256)
frame_patches_kernel(const uint8_t* __restrict__ bgr, int H, int W, int Hp, int Wp, int pl, int pt,
                     __half* __restrict__ patches, long total8) {
    // 2 * (u / 255) - 1 for the 256 byte values, once per block (the same fp32 expression, so the same bits as per entry);
    // independent of the previous kernel: built before the dependency wait
    __shared__ __half lut[256];
    lut[threadIdx.x] = __float2half_rn(2.0f * (static_cast<float>(threadIdx.x) / 255.0f) - 1.0f);
    pdl_enter();
    __syncthreads();
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total8) return;
    const unsigned iu = static_cast<unsigned>(i);                    // total8 < 2^31: 32-bit divisions
    const unsigned op = iu / 19u;
    const int seg = static_cast<int>(iu - op * 19u);
    const unsigned Wo = static_cast<unsigned>(Wp / 2);
    const int oy = static_cast<int>(op / Wo), ox = static_cast<int>(op - static_cast<unsigned>(oy) * Wo);
    const __half zero = __float2half_rn(0.0f);
    __half v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = seg * 8 + j;
        __half v = zero;
        if (k < 147) {
            const int tap = (k * 171) >> 9, c = k - tap * 3, ky = (tap * 37) >> 8, kx = tap - ky * 7;      // k / 3 (k < 256), tap / 7 (tap < 49)
            const int yp = 2 * oy + ky - 3, xp = 2 * ox + kx - 3;
            if (yp >= 0 && yp < Hp && xp >= 0 && xp < Wp) {
                const int ys = min(max(yp - pt, 0), H - 1), xs = min(max(xp - pl, 0), W - 1);
                v = lut[__ldg(bgr + (ys * W + xs) * 3 + (2 - c))];
            }
        }
        v8[j] = v;
    }
    const __half2 p0 = __halves2half2(v8[0], v8[1]), p1 = __halves2half2(v8[2], v8[3]), p2 = __halves2half2(v8[4], v8[5]),
                  p3 = __halves2half2(v8[6], v8[7]);
    *reinterpret_cast<uint4*>(patches + i * 8) =
        make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                   *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
}

cudaError_t launch_frame_patches(const uint8_t* bgr, int H, int W, int Hp, int Wp, int pad_left, int pad_top, __half* patches,
                          cudaStream_t stream) {
    const long total8 = static_cast<long>(Hp / 2) * (Wp / 2) * 19;
    return launch_pdl(frame_patches_kernel, dim3(static_cast<unsigned>((total8 + 255) / 256)), dim3(256), 0, stream, bgr, H, W, Hp,
               Wp, pad_left, pad_top, patches, total8);
}

// ==========================================================================================
// instance norm (MFT/RAFT/core/extractor.py:28-32,129-130: biased variance, eps 1e-5, no affine)
// ==========================================================================================
constexpr int kStatThreads = 192;   // divisible by C/2 for C in {64, 96, 128}

__global__ void __launch_bounds__(kStatThreads)
instnorm_stats_kernel(const __half* __restrict__ raw, int P, int C, int pix_per_block, double* __restrict__ sums) {
    pdl_enter();
    __shared__ float red[2][kStatThreads * 2];
    const int b = blockIdx.y;
    const int c2 = C / 2;
    const int cp = threadIdx.x % c2, pl = threadIdx.x / c2, nrow = kStatThreads / c2;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const __half2* src = reinterpret_cast<const __half2*>(raw + static_cast<long>(b) * P * C);
    for (int p = p0 + pl; p < p1; p += nrow) {
        const float2 v = __half22float2(src[static_cast<long>(p) * c2 + cp]);
        s0 += v.x; s1 += v.y;
        q0 += v.x * v.x; q1 += v.y * v.y;
    }
    red[0][threadIdx.x * 2] = s0; red[0][threadIdx.x * 2 + 1] = s1;
    red[1][threadIdx.x * 2] = q0; red[1][threadIdx.x * 2 + 1] = q1;
    __syncthreads();
    if (pl == 0) {
        for (int r = 1; r < nrow; ++r) {
            const int t = r * c2 + cp;
            s0 += red[0][t * 2]; s1 += red[0][t * 2 + 1];
            q0 += red[1][t * 2]; q1 += red[1][t * 2 + 1];
        }
        double* o = sums + static_cast<long>(b) * 2 * C;
        atomicAdd(o + 2 * cp, static_cast<double>(s0));
        atomicAdd(o + 2 * cp + 1, static_cast<double>(s1));
        atomicAdd(o + C + 2 * cp, static_cast<double>(q0));
        atomicAdd(o + C + 2 * cp + 1, static_cast<double>(q1));
    }
}

cudaError_t launch_instnorm_stats(const __half* raw, int B, int P, int C, double* sums, cudaStream_t stream) {
    if (cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * B * 2 * C, stream)) return e;
    const int pix_per_block = 256;
    dim3 grid((P + pix_per_block - 1) / pix_per_block, B);
    return launch_pdl(instnorm_stats_kernel, grid, dim3(kStatThreads), 0, stream, raw, P, C, pix_per_block, sums);
}

// 16-byte accesses: one thread = 8 consecutive channels of one pixel (C % 8 == 0: 64 / 96 / 128), two per thread
__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const uint4* __restrict__ raw, const double* __restrict__ sums, int P, int C, int relu,
                      const uint4* __restrict__ res, uint4* __restrict__ out, long total8) {
    pdl_enter();
    // per-channel mean / rstd once per block (double: sum-of-squares minus square-of-sum cancels in fp32)
    __shared__ __align__(16) float s_mean[128];
    __shared__ __align__(16) float s_rstd[128];
    const int c8 = C / 8;
    const long per_img8 = static_cast<long>(P) * c8;
    const long i0 = static_cast<long>(blockIdx.x) * (blockDim.x * 2);
    // element counts < 2^31: 32-bit divisions.  A block may straddle two images: stragglers recompute below
    const int b = static_cast<int>(static_cast<unsigned>(i0) / static_cast<unsigned>(per_img8));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double* s = sums + static_cast<long>(b) * 2 * C;
        const double inv = 1.0 / P;
        const double m = s[c] * inv;
        const double v = fmax(s[C + c] * inv - m * m, 0.0);
        s_mean[c] = static_cast<float>(m);
        s_rstd[c] = static_cast<float>(1.0 / sqrt(v + 1e-5));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long i = i0 + static_cast<long>(k) * blockDim.x + threadIdx.x;
        if (i >= total8) break;
        const int c0 = static_cast<int>(static_cast<unsigned>(i) % static_cast<unsigned>(c8)) * 8;
        float m[8], r[8];
        *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(s_mean + c0);
        *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(s_mean + c0 + 4);
        *reinterpret_cast<float4*>(r) = *reinterpret_cast<const float4*>(s_rstd + c0);
        *reinterpret_cast<float4*>(r + 4) = *reinterpret_cast<const float4*>(s_rstd + c0 + 4);
        if (i >= static_cast<long>(b + 1) * per_img8) {            // rare: a pixel of the next image inside this block
            const double* s = sums + (i / per_img8) * 2 * C;
            const double inv = 1.0 / P;
            for (int j = 0; j < 8; ++j) {
                const double a0 = s[c0 + j] * inv;
                m[j] = static_cast<float>(a0);
                r[j] = static_cast<float>(1.0 / sqrt(fmax(s[C + c0 + j] * inv - a0 * a0, 0.0) + 1e-5));
            }
        }
        const uint4 xv = raw[i];
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t rw[4] = {0u, 0u, 0u, 0u};
        if (res != nullptr) {
            const uint4 rv = res[i];
            rw[0] = rv.x; rw[1] = rv.y; rw[2] = rv.z; rw[3] = rv.w;
        }
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&xw[j]));
            float y0 = (x.x - m[2 * j]) * r[2 * j], y1 = (x.y - m[2 * j + 1]) * r[2 * j + 1];
            if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            if (res != nullptr) {
                const float2 q = __half22float2(*reinterpret_cast<const __half2*>(&rw[j]));
                y0 = fmaxf(y0 + q.x, 0.f);
                y1 = fmaxf(y1 + q.y, 0.f);
            }
            const __half2 o = __floats2half2_rn(y0, y1);
            ow[j] = *reinterpret_cast<const uint32_t*>(&o);
        }
        out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

cudaError_t launch_instnorm_apply(const __half* raw, const double* sums, int B, int P, int C, int relu, const __half* res,
                           __half* out, cudaStream_t stream) {
    const long total8 = static_cast<long>(B) * P * C / 8;
    return launch_pdl(instnorm_apply_kernel, dim3(static_cast<unsigned>((total8 + 511) / 512)), dim3(256), 0, stream,
               reinterpret_cast<const uint4*>(raw), sums, P, C, relu, reinterpret_cast<const uint4*>(res),
               reinterpret_cast<uint4*>(out), total8);
}

// ==========================================================================================
// per-pair state set-up (core/raft.py:141-154): gather cached per-frame features, net/inp split,
// coords1 = coords0 = grid
// ==========================================================================================
// one warp per (pair, pixel): 16-byte vector copies of the cached per-frame features
__global__ void __launch_bounds__(256)
pair_setup_kernel(const PairSetup a) {
    pdl_enter();
    const int npx = a.h * a.w;
    const long pp = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);      // pair * npx + n
    if (pp >= static_cast<long>(a.n_pairs) * npx) return;
    const int lane = threadIdx.x & 31;
    const int pair = static_cast<int>(static_cast<unsigned>(pp) / static_cast<unsigned>(npx)), n = static_cast<int>(pp) - pair * npx;
    const int ls = a.slots[2 * pair], rs = a.slots[2 * pair + 1];
    const long lsrc = static_cast<long>(ls) * npx + n, rsrc = static_cast<long>(rs) * npx + n;
    // fmaps: 256 halves = 32 x uint4 per pixel
    reinterpret_cast<uint4*>(a.F1)[pp * 32 + lane] = reinterpret_cast<const uint4*>(a.fmap_slots)[lsrc * 32 + lane];
    reinterpret_cast<uint4*>(a.F2)[pp * 32 + lane] = reinterpret_cast<const uint4*>(a.fmap_slots)[rsrc * 32 + lane];
    // net: 128 floats = 32 x float4 -> fp32 master + fp16 operand copy (X[:, 0:128])
    const float4 hv = reinterpret_cast<const float4*>(a.net_slots)[lsrc * 32 + lane];
    reinterpret_cast<float4*>(a.h32)[pp * 32 + lane] = hv;
    const __half2 h0 = __floats2half2_rn(hv.x, hv.y), h1 = __floats2half2_rn(hv.z, hv.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
    reinterpret_cast<uint2*>(a.X + pp * 512)[lane] = pk;
    // inp: 128 halves = 16 x uint4 -> X[:, 128:256]
    if (lane < 16) reinterpret_cast<uint4*>(a.X + pp * 512 + 128)[lane] = reinterpret_cast<const uint4*>(a.inp_slots)[lsrc * 16 + lane];
    if (lane < 2) {
        float c = static_cast<float>(lane == 0 ? n % a.w : n / a.w);
        if (a.init_flow != nullptr) c = c + a.init_flow[(static_cast<long>(pair) * 2 + lane) * npx + n];
        a.coords1[pp * 2 + lane] = c;
    }
}

cudaError_t launch_pair_setup(const PairSetup& a, cudaStream_t stream) {
    const long total = static_cast<long>(a.n_pairs) * a.h * a.w;
    return launch_pdl(pair_setup_kernel, dim3(static_cast<unsigned>((total + 7) / 8)), dim3(256), 0, stream, a);
}

// ==========================================================================================
// correlation pyramid pooling (core/corr.py:26-28): avg_pool2d(2, stride 2), floor sizes
// ==========================================================================================
__global__ void __launch_bounds__(256)
corr_pool_kernel(const __half* __restrict__ L0, __half* __restrict__ L1, __half* __restrict__ L2, __half* __restrict__ L3,
                 int h, int w, int p1, int p2, int p3) {
    pdl_enter();
    extern __shared__ float sm[];
    const int h1 = h / 2, w1 = w / 2, h2 = h1 / 2, w2 = w1 / 2, h3 = h2 / 2, w3 = w2 / 2;
    float* s1 = sm;
    float* s2 = sm + h1 * w1;
    const long row = blockIdx.x;
    const __half* src = L0 + row * h * w;
    // (per-row base pointers once, 32-bit unsigned index arithmetic inside the loops: the kernel is instruction-issue bound)
    __half* const o1 = L1 + row * h1 * p1;
    __half* const o2 = L2 + row * h2 * p2;
    __half* const o3 = L3 + row * h3 * p3;
    const unsigned uw = static_cast<unsigned>(w), uw1 = static_cast<unsigned>(w1), uw2 = static_cast<unsigned>(w2), uw3 = static_cast<unsigned>(w3);
    // every level is the fp32 average of the level above as it is STORED (fp16), rounded once: what avg_pool2d of the
    // stored volume gives (core/corr.py:26-28)
    if ((w & 7) == 0) {
        // 16-byte loads: one thread = 8 columns of the row pair (2y, 2y+1) -> 4 outputs, one 8-byte store
        const unsigned w8 = uw >> 3, n1 = static_cast<unsigned>(h1) * w8;
        for (unsigned i = threadIdx.x; i < n1; i += blockDim.x) {
            const unsigned y = i / w8, g = i - y * w8;
            const uint4* p = reinterpret_cast<const uint4*>(src + 2u * y * uw) + g;
            const uint4 a = __ldg(p);
            const uint4 b = __ldg(p + w8);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
            __half r[4];
            float rf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&av[j]));
                const float2 u = __half22float2(*reinterpret_cast<const __half2*>(&bv[j]));
                r[j] = __float2half_rn((((t.x + t.y) + u.x) + u.y) * 0.25f);
                rf[j] = __half2float(r[j]);
            }
            // (one 16-byte store: four scalar stores at a 4-word stride were 4-way bank conflicts; w1 is a multiple of 4 here)
            *reinterpret_cast<float4*>(s1 + y * uw1 + 4u * g) = make_float4(rf[0], rf[1], rf[2], rf[3]);
            const __half2 p01 = __halves2half2(r[0], r[1]), p23 = __halves2half2(r[2], r[3]);
            *reinterpret_cast<uint2*>(o1 + y * static_cast<unsigned>(p1) + 4u * g) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&p01), *reinterpret_cast<const uint32_t*>(&p23));
        }
    } else {
        const unsigned n1 = static_cast<unsigned>(h1) * uw1;
        for (unsigned i = threadIdx.x; i < n1; i += blockDim.x) {
            const unsigned y = i / uw1, x = i - y * uw1;
            const __half* q = src + 2u * y * uw + 2u * x;
            const __half r = __float2half_rn((((__half2float(q[0]) + __half2float(q[1])) + __half2float(q[w])) + __half2float(q[w + 1])) * 0.25f);
            s1[i] = __half2float(r);
            o1[y * static_cast<unsigned>(p1) + x] = r;
        }
    }
    __syncthreads();
    {
        const unsigned n2 = static_cast<unsigned>(h2) * uw2;
        for (unsigned i = threadIdx.x; i < n2; i += blockDim.x) {
            const unsigned y = i / uw2, x = i - y * uw2;
            const float* q = s1 + 2u * y * uw1 + 2u * x;
            float2 t, u;
            if ((w1 & 1) == 0) {                                           // even row pitch: the pairs are 8-byte aligned
                t = *reinterpret_cast<const float2*>(q);
                u = *reinterpret_cast<const float2*>(q + w1);
            } else {
                t = make_float2(q[0], q[1]);
                u = make_float2(q[w1], q[w1 + 1]);
            }
            const __half r = __float2half_rn((((t.x + t.y) + u.x) + u.y) * 0.25f);
            s2[i] = __half2float(r);
            o2[y * static_cast<unsigned>(p2) + x] = r;
        }
    }
    __syncthreads();
    {
        const unsigned n3 = static_cast<unsigned>(h3) * uw3;
        for (unsigned i = threadIdx.x; i < n3; i += blockDim.x) {
            const unsigned y = i / uw3, x = i - y * uw3;
            const float* q = s2 + 2u * y * uw2 + 2u * x;
            o3[y * static_cast<unsigned>(p3) + x] = __float2half_rn((((q[0] + q[1]) + q[w2]) + q[w2 + 1]) * 0.25f);
        }
    }
}

cudaError_t launch_corr_pool(const __half* L0, __half* L1, __half* L2, __half* L3, long rows, int h, int w, int p1, int p2, int p3,
                             cudaStream_t stream) {
    const size_t smem = sizeof(float) * (static_cast<size_t>(h / 2) * (w / 2) + static_cast<size_t>(h / 4) * (w / 4));
    static bool attr = false;
    if (!attr) {
        if (cudaError_t e = cudaFuncSetAttribute(corr_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) return e;
        attr = true;
    }
    return launch_pdl(corr_pool_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), smem, stream, L0, L1, L2, L3, h, w, p1, p2, p3);
}

// ==========================================================================================
// pyramid lookup (core/corr.py:30-51) + flow operands of the motion encoder
// one warp per (pair, source pixel)
// ==========================================================================================
// The routine lives in lookup.cuh (shared with the persistent refinement kernel, which can run the same lookup as tiles
// of its dataflow program).  Here: one warp per group of kLkGroup consecutive (pair, source pixel)s, 8 groups per block.
__global__ void __launch_bounds__(256, 5)
lookup_kernel(const LookupArgs a, const long n_groups) {
    pdl_enter();
    __shared__ __align__(16) float win[8][kLkGroup * 10 * 20];        // window row pitch 20 floats: no bank conflicts in the blend
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long total = static_cast<long>(a.n_pairs) * a.h * a.w;
    const LookupLane t = lookup_lane_init(lane);
    // (one group per warp: a grid-stride loop over groups makes the compiler hoist the per-group set-up out of the loop
    // and spill it -- 300+ bytes of local memory)
    {
        const long g = static_cast<long>(blockIdx.x) * 8 + wib;
        if (g >= n_groups) return;
        const long pp0 = g * kLkGroup;
        const long left = total - pp0;
        lookup_group<20>(a, t, pp0, left < kLkGroup ? static_cast<int>(left) : kLkGroup, lane, win[wib]);
        __syncwarp();
    }
}

cudaError_t launch_lookup(const LookupArgs& a, cudaStream_t stream) {
    const long total = static_cast<long>(a.n_pairs) * a.h * a.w;
    const long n_groups = (total + kLkGroup - 1) / kLkGroup;
    const long blocks = (n_groups + 7) / 8;
    return launch_pdl(lookup_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, a, n_groups);
}

// ------------------------------------------------------------------------------------------
// TMA variant.  One warp per group of kLkGroup consecutive (pair, source pixel)s, as above, but the 16 windows of a group
// (4 pixels x 4 levels) are 16 bulk tensor copies issued by lanes 0..15 (lane p * 4 + l: the lane that evaluated the
// position of (pixel p, level l)) onto the warp's own mbarrier.  While they are in flight the warp writes the flow
// operands (7x7x2 neighbourhood for convf1, flow channels of the GRU record); then each level's boxes are widened to the
// fp32 window layout of lookup_blend_level and blended exactly as in lookup_group: same bits.
// ------------------------------------------------------------------------------------------
constexpr int kLtWarps = 4;
// The innermost box coordinate of a bulk tensor copy must be a multiple of 16 bytes (anything else raises "illegal
// instruction" on the B200: tools/tma_box_probe.cu), so a window that starts at column X0 is fetched from column X0 & ~7:
// its 10 columns then sit at offset X0 & 7 <= 7 of a 24-column box.
constexpr int kLtBoxCols = 24;
constexpr int kLtBoxBytes = 512;                          // a 24 x 10 fp16 box is 480 bytes; destinations are 128-byte aligned
constexpr int kLtBoxTx = kLtBoxCols * 10 * 2;

constexpr int kLtWinPitch = 20;                           // floats per fp32 window row: 20 (not 16) keeps the blends free of bank conflicts
constexpr int kLtRegion = kLkGroup * 10 * kLtWinPitch * 4; // the fp32 window of one level (3200 bytes), laid over the level's four boxes (2048 bytes)

__global__ void __launch_bounds__(kLtWarps * 32, 8)
lookup_tma_kernel(const __grid_constant__ LookupTmaArgs P, const long n_groups) {
    // Per warp: two regions of four box slots (+ 512 bytes each) and two mbarriers.  The levels run 3, 2, 1, 0; level l uses
    // region l & 1.  Levels 3 and 2 are fetched first; level 1 is fetched into region 1 as soon as level 3 has been blended
    // and level 0 into region 0 after level 2, so their latency runs under the blends in between.  The fp32 window of a level
    // (3200 bytes, row pitch 20 floats) is written over the level's own boxes (read into registers first) and the spare bytes
    // behind them: 6.25 KiB of shared memory per warp.
    __shared__ __align__(128) unsigned char boxes[kLtWarps][2 * kLtRegion];
    // the group's four output rows (328 fp16 each) are collected here and leave as 16-byte stores of one contiguous 2624-byte run
    // (written per level from the blend they were 2- and 4-byte stores at an 18-byte stride: ~8 L1 wavefronts per instruction)
    __shared__ __align__(16) __half outrow[kLtWarps][kLkGroup * 328];
    __shared__ __align__(8) uint64_t bar[kLtWarps][2];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 2) mbar_init(&bar[wib][lane], 1);
    if (lane == 0) fence_mbar_init();
    __syncwarp();
    pdl_enter();
    const long g = static_cast<long>(blockIdx.x) * kLtWarps + wib;
    if (g >= n_groups) return;
    const LookupArgs& a = P.a;
    const int npx = a.h * a.w;
    const long total = static_cast<long>(a.n_pairs) * npx;
    const long pp0 = g * kLkGroup;
    const int nvalid = total - pp0 < kLkGroup ? static_cast<int>(total - pp0) : kLkGroup;
    const unsigned valid_mask = (1u << nvalid) - 1u;

    // ---- set-up: lane p * 4 + l evaluates (pixel p, level l), as in lookup_group -----------------------------------------
    const int sp = (lane >> 2) & (kLkGroup - 1), sl = lane & 3;
    const bool sv = sp < nvalid;
    const long spp = pp0 + (sv ? sp : 0);
    const float2 c = __ldcg(reinterpret_cast<const float2*>(a.coords1 + spp * 2));
    const int my_finite = (isfinite(c.x) && isfinite(c.y)) ? 1 : 0;
    const int mh = a.h >> sl, mw = a.w >> sl;
    const float inv = 1.0f / static_cast<float>(1 << sl);
    const float fxp = lk_roundtrip_div(c.x * inv - 4.0f, static_cast<float>(mw - 1));
    const float fyp = lk_roundtrip_div(c.y * inv - 4.0f, static_cast<float>(mh - 1));
    const float fx = floorf(fxp), fy = floorf(fyp);
    const float my_wE = fxp - fx, my_wS = fyp - fy;
    const int my_X0 = my_finite ? static_cast<int>(fminf(fmaxf(fx, -32.0f), static_cast<float>(mw + 16))) : -64;
    const int my_Y0 = my_finite ? static_cast<int>(fminf(fmaxf(fy, -32.0f), static_cast<float>(mh + 16))) : -64;
    const int my_a4 = my_X0 & 4;                                           // first box column of the widened window (0 | 4)

    // ---- window fetch: the lanes (p, level) of one level issue its four boxes onto the level's region / barrier ---------------
    unsigned char* mybox = boxes[wib];
    const uint32_t level_tx = static_cast<uint32_t>(nvalid) * kLtBoxTx;
    auto fetch = [&](int l) {
        uint64_t* b = &bar[wib][l & 1];
        if (lane == 0) mbar_arrive_expect_tx(b, level_tx);
        __syncwarp();
        if (lane < 16 && sl == l && sv)
            tma_load_3d(mybox + (l & 1) * kLtRegion + sp * kLtBoxBytes, &P.tm[l], b, my_X0 & ~7, my_Y0, P.pix0 + static_cast<int>(spp));
    };
    fetch(3);
    fetch(2);

    // ---- flow operands while the windows are in flight: lane = tap (ky * 7 + kx) of the 7x7 neighbourhood, both channels.
    //      All eight loads of the group are issued before the first store (the compiler cannot prove that the fp16
    //      outputs do not alias coords1, so loads behind a store would wait for it: eight L2 round trips in a row).
    //      (The four pixels of a group are consecutive in one image row: the coarse width is a multiple of 64 here.)
    {
        const int n0 = static_cast<int>(static_cast<unsigned>(pp0) % static_cast<unsigned>(npx));      // (total < 2^31)
        const int y = n0 / a.w, x0 = n0 - y * a.w;
        const int t0 = lane, t1 = lane + 32;                               // taps 0..31 and 32..51 (49..51: zero pad of the row)
        const int ky0 = (t0 * 37) >> 8, ky1 = (t1 * 37) >> 8;             // tap / 7 (tap < 64)
        const int dx0 = t0 - ky0 * 7 - 3, dx1 = t1 - ky1 * 7 - 3;
        const unsigned ya = static_cast<unsigned>(y + ky0 - 3), yb = static_cast<unsigned>(y + ky1 - 3);
        const bool rowa = ya < static_cast<unsigned>(a.h), rowb = lane < 17 && yb < static_cast<unsigned>(a.h);
        const float2* cb = reinterpret_cast<const float2*>(a.coords1) + (pp0 - n0);
        float2 f0[kLkGroup], f1[kLkGroup];
#pragma unroll
        for (int p = 0; p < kLkGroup; ++p) {
            const unsigned xa = static_cast<unsigned>(x0 + p + dx0), xb = static_cast<unsigned>(x0 + p + dx1);
            const bool ia = rowa && xa < static_cast<unsigned>(a.w), ib = rowb && xb < static_cast<unsigned>(a.w);
            f0[p] = __ldcg(cb + (ia ? ya * static_cast<unsigned>(a.w) + xa : 0u));
            f1[p] = __ldcg(cb + (ib ? yb * static_cast<unsigned>(a.w) + xb : 0u));
        }
#pragma unroll
        for (int p = 0; p < kLkGroup; ++p) {
            if (p < nvalid) {
                const long pp = pp0 + p;
                unsigned* fp = reinterpret_cast<unsigned*>(a.flowpatch16 + pp * 104);
                const unsigned xa = static_cast<unsigned>(x0 + p + dx0), xb = static_cast<unsigned>(x0 + p + dx1);
                const bool ia = rowa && xa < static_cast<unsigned>(a.w), ib = rowb && xb < static_cast<unsigned>(a.w);
                const __half2 ha = __floats2half2_rn(ia ? f0[p].x - static_cast<float>(xa) : 0.0f, ia ? f0[p].y - static_cast<float>(ya) : 0.0f);
                const __half2 hb = __floats2half2_rn(ib ? f1[p].x - static_cast<float>(xb) : 0.0f, ib ? f1[p].y - static_cast<float>(yb) : 0.0f);
                fp[t0] = *reinterpret_cast<const unsigned*>(&ha);
                if (lane == 24) *reinterpret_cast<unsigned*>(a.X + pp * 512 + 382) = *reinterpret_cast<const unsigned*>(&ha);
                if (lane < 20) fp[t1] = *reinterpret_cast<const unsigned*>(&hb);
            }
        }
    }
    __half* orow = outrow[wib];
    if (lane < kLkGroup) *reinterpret_cast<uint2*>(orow + lane * 328 + 324) = make_uint2(0u, 0u);      // channels 324..327: zero pad

    // ---- widen + blend, level by level --------------------------------------------------------------------------------------
    // widening task T = lane + 32 r (T < 160): (pixel p, window row, 4-element chunk c) -> floats 4c .. 4c+3 of row (p, row) of
    // the fp32 window [pixel][row][20] = box columns a4 + 4c ..; the window proper starts at float X0 & 3 of its rows
    int src_off[5], dst_off[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int T = lane + 32 * r;
        const int p = T / 40, rem = T - p * 40, row = rem >> 2, ch = rem & 3;
        src_off[r] = p * kLtBoxBytes + row * (kLtBoxCols * 2) + ch * 8;
        dst_off[r] = (p * 10 + row) * kLtWinPitch + ch * 4;
    }
#pragma unroll
    for (int l = 3; l >= 0; --l) {
        if (!__all_sync(0xffffffffu, mbar_wait(&bar[wib][l & 1], l < 2 ? 1u : 0u))) {      // (each barrier completes twice: levels 3 | 2, then 1 | 0; warp-uniform verdict)
            if (lane == 0 && P.err_flag != nullptr) atomicExch(P.err_flag, 90);
            return;
        }
        unsigned char* region = mybox + (l & 1) * kLtRegion;
        float* w = reinterpret_cast<float*>(region);
        uint2 v[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int T = lane + 32 * r;
            const int a4 = __shfl_sync(0xffffffffu, my_a4, (T / 40) * 4 + l);
            v[r] = *reinterpret_cast<const uint2*>(region + src_off[r] + a4 * 2);
        }
        __syncwarp();                                                      // every box of this level is in registers: the window may overwrite them
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v[r].x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v[r].y));
            *reinterpret_cast<float4*>(w + dst_off[r]) = make_float4(f0.x, f0.y, f1.x, f1.y);
        }
        __syncwarp();
        lookup_blend_level<4, kLtWinPitch>(orow, 0, valid_mask, l, my_wE, my_wS, my_X0 & 3, my_finite, lane, w);
        if (l >= 2) {
            __syncwarp();                                                  // the blend is done with the region
            fence_proxy_async_smem();                                      // ... before the TMA unit writes it again
            fetch(l - 2);
        }
    }
    __syncwarp();
    {
        uint4* dst = reinterpret_cast<uint4*>(a.corr16 + pp0 * 328);
        const uint4* src = reinterpret_cast<const uint4*>(orow);
        const int n16 = nvalid * 41;                                       // 656 bytes per row
#pragma unroll
        for (int r = 0; r < (kLkGroup * 41 + 31) / 32; ++r) {
            const int T = lane + 32 * r;
            if (T < n16) dst[T] = src[T];
        }
    }
}

cudaError_t launch_lookup_tma(const LookupTmaArgs& a, cudaStream_t stream) {
    const long total = static_cast<long>(a.a.n_pairs) * a.a.h * a.a.w;
    const long n_groups = (total + kLkGroup - 1) / kLkGroup;
    const long blocks = (n_groups + kLtWarps - 1) / kLtWarps;
    return launch_pdl(lookup_tma_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kLtWarps * 32), 0, stream, a, n_groups);
}

// ==========================================================================================
// OU head input: [net | inp | corr | flow | delta_flow | motion] = 712 channels (update.py:197)
// ==========================================================================================
// 720 halves per pixel = 90 x 16 bytes: [X 0:256 | corr 0:320 | corr 320:324 + flow + delta | X 256:384 | zero pad]
__global__ void __launch_bounds__(256)
ou_pack_kernel(const OuPackArgs a) {
    pdl_enter();
    const int npx = a.h * a.w;
    const long pp = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (pp >= static_cast<long>(a.n_pairs) * npx) return;
    const int lane = threadIdx.x & 31;
    const int n = static_cast<int>(static_cast<unsigned>(pp) % static_cast<unsigned>(npx));
    const uint4* X = reinterpret_cast<const uint4*>(a.X + pp * 512);          // 64 chunks
    const uint4* C = reinterpret_cast<const uint4*>(a.corr16 + pp * 328);     // 41 chunks (row pitch 656 B)
    uint4* o = reinterpret_cast<uint4*>(a.packed + pp * 720);                  // 90 chunks
    for (int k = lane; k < 90; k += 32) {
        uint4 v;
        if (k < 32) {
            v = X[k];
        } else if (k < 72) {
            v = C[k - 32];
        } else if (k == 72) {
            const uint2 c4 = *reinterpret_cast<const uint2*>(a.corr16 + pp * 328 + 320);     // corr 320:324
            const float fx = a.coords1[pp * 2] - static_cast<float>(n % a.w), fy = a.coords1[pp * 2 + 1] - static_cast<float>(n / a.w);
            const __half2 f = __floats2half2_rn(fx, fy), d = __floats2half2_rn(a.delta32[pp * 2], a.delta32[pp * 2 + 1]);
            v.x = c4.x; v.y = c4.y;
            v.z = *reinterpret_cast<const uint32_t*>(&f);
            v.w = *reinterpret_cast<const uint32_t*>(&d);
        } else if (k < 89) {
            v = X[32 + (k - 73)];
        } else {
            v = make_uint4(0u, 0u, 0u, 0u);
        }
        o[k] = v;
    }
}

cudaError_t launch_ou_pack(const OuPackArgs& a, cudaStream_t stream) {
    const long total = static_cast<long>(a.n_pairs) * a.h * a.w;
    return launch_pdl(ou_pack_kernel, dim3(static_cast<unsigned>((total + 7) / 8)), dim3(256), 0, stream, a);
}

// ==========================================================================================
// convex 8x upsampling of flow / occlusion logits / uncertainty with one shared mask
// (core/raft.py:83-94,190-218) + post-processing and unpadding (MFT/raft.py:56-62)
// 64 threads per coarse pixel (one per 8x8 sub-position), 4 consecutive coarse pixels per block
// ==========================================================================================
__global__ void __launch_bounds__(256)
upsample_kernel(const UpsampleArgs a) {
    pdl_enter();
    // The 64 threads of a coarse pixel blend the same 3x3 neighbours: 36 threads of the block fetch them once
    // ([8 * flow x, 8 * flow y, occlusion logits 0 / 1, log-variance]; zeros outside the image = F.unfold's padding).
    __shared__ float nb[4][9][5];
    const int npx = a.h * a.w;
    const long total = static_cast<long>(a.n_pairs) * npx;
    if (threadIdx.x < 36) {
        const int qd = threadIdx.x / 9, k = threadIdx.x - qd * 9;
        const long pq = static_cast<long>(blockIdx.x) * 4 + qd;
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (pq < total) {
            const int pr = static_cast<int>(static_cast<unsigned>(pq) / static_cast<unsigned>(npx)), nq = static_cast<int>(pq) - pr * npx;
            const int yq = nq / a.w, xq = nq - yq * a.w;
            const int yy = yq + k / 3 - 1, xx = xq + k % 3 - 1;
            if (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w) {
                const long q = pq - nq + static_cast<long>(yy) * a.w + xx;
                const float2 c = *reinterpret_cast<const float2*>(a.coords1 + q * 2);
                const float4 ou = *reinterpret_cast<const float4*>(a.ou32 + q * 4);
                v[0] = 8.0f * (c.x - static_cast<float>(xx));
                v[1] = 8.0f * (c.y - static_cast<float>(yy));
                v[2] = ou.x; v[3] = ou.y; v[4] = ou.z;
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) nb[qd][k][j] = v[j];
    }
    __syncthreads();
    // thread -> (sub-row sy = warp, coarse pixel qd, sub-column sx): a warp writes ONE full-resolution row segment of the
    // block's 4 consecutive coarse pixels = 32 consecutive floats per plane (128-byte stores; round 1 had a warp cover
    // 4 rows x 8 columns = 32-byte runs) and reads four 32-byte runs of the mask per neighbour
    const int qd = (threadIdx.x >> 3) & 3;
    const long pp = static_cast<long>(blockIdx.x) * 4 + qd;
    if (pp >= total) return;
    const int sy = threadIdx.x >> 5, sx = threadIdx.x & 7;
    const int sub = sy * 8 + sx;
    const int pair = static_cast<int>(static_cast<unsigned>(pp) / static_cast<unsigned>(npx)), n = static_cast<int>(pp) - pair * npx;
    const int y = n / a.w, x = n - y * a.w;
    const float* m = a.mask32 + pp * 576 + sub;
    float mk[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        mk[k] = m[k * 64];
        mx = fmaxf(mx, mk[k]);
    }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        mk[k] = __expf(mk[k] - mx);          // ex2.approx: 2 ulp on weights that multiply fp16-operand results
        den += mk[k];
    }
    const float inv_den = 1.0f / den;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float wk = mk[k] * inv_den;
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[j] = fmaf(wk, nb[qd][k][j], acc[j]);
    }
    const int Y = 8 * y + sy - a.pad_top, Xo = 8 * x + sx - a.pad_left;
    if (Y < 0 || Y >= a.H || Xo < 0 || Xo >= a.W) return;
    const long hw = static_cast<long>(a.H) * a.W;
    float* o = a.out + static_cast<long>(pair) * 4 * hw + static_cast<long>(Y) * a.W + Xo;
    o[0] = acc[0];
    o[hw] = acc[1];
    const float lm = fmaxf(acc[2], acc[3]);
    const float e0 = expf(acc[2] - lm), e1 = expf(acc[3] - lm);
    o[2 * hw] = e1 / (e0 + e1);
    o[3 * hw] = sqrtf(expf(acc[4]));
}

cudaError_t launch_upsample(const UpsampleArgs& a, cudaStream_t stream) {
    const long total = static_cast<long>(a.n_pairs) * a.h * a.w;
    return launch_pdl(upsample_kernel, dim3(static_cast<unsigned>((total + 3) / 4)), dim3(256), 0, stream, a);
}

}  // namespace mftb
