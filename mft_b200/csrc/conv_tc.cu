// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.  See conv.h for the contract.
//
// Two kernels share the operand pipeline and the fused epilogues:
//   conv_tc_kernel    one launch per layer; CTA = 256 threads = one 128-pixel x n_tile output tile, two CTAs per SM.
//                     warp 0 = A-operand TMA producer (shifted activation box per tap), warp 2 = B-operand TMA producer
//                     (weight slab), warp 1 = TMEM allocator + tcgen05.mma issuer (4 x K=16 MMAs per 64-channel stage);
//                     `stages` smem slots guarded by full/empty mbarriers, accumulator hand-off through a third mbarrier
//                     signalled by tcgen05.commit; then all 8 warps drain the accumulator (tile_epilogue).
//   conv_prog_kernel  persistent: one CTA per SM runs a whole program of dependent layers (a GRU iteration) with
//                     tile-level dataflow, warp-specialised roles and two TMEM accumulators (see below).
#include <cstddef>
#include <cstdio>
#include <type_traits>
#include <cstring>

#include "conv.h"
#include "lookup.cuh"
#include "ptx.cuh"

namespace mftb {

// ------------------------------------------------------------------------------------------
// fused epilogue on 32 consecutive accumulator columns of one pixel
// ------------------------------------------------------------------------------------------
// GRU gate non-linearities on the SFU fast path (ex2.approx + rcp.approx, ~2 ulp): absolute error ~1e-7, far
// below the fp16 rounding of the operands that feed them.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_gate(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// The epilogue handles 4 consecutive output columns of one pixel per call.  The tensor-core kernel stages the
// accumulator tile through shared memory so that the 8 lanes sharing a pixel row touch 8 consecutive 16-byte
// (fp32) / 8-byte (fp16) segments: every global load and store of the epilogue is coalesced.  Global INPUTS of
// the epilogue (residual, GRU state) are fetched by epi_prefetch for several pixels before any is consumed, so
// their L2 latencies overlap.
struct EpiAux {
    float4 a, b;
};

// Pixel indices are 32-bit everywhere except the correlation volume (pixel * N^2 overflows): `pix` stays 64-bit in
// the signature but every mode other than EPI_F32 does its address arithmetic on the low 32 bits (M * stride < 2^31).
// LEAN: the layer has neither a residual input nor fused statistics (checked by the host): compiled out.
template <int MODE, bool LEAN = false>
__device__ __forceinline__ EpiAux epi_prefetch(const ConvEpi& e, int col, long pix) {
    EpiAux x;
    x.a = make_float4(0.f, 0.f, 0.f, 0.f);
    x.b = x.a;
    const uint32_t p32 = static_cast<uint32_t>(pix);
    if constexpr (MODE == EPI_F16) {
        if (!LEAN && e.res16 != nullptr && col + 4 <= e.n_valid) {
            const uint2 rr = __ldcg(reinterpret_cast<const uint2*>(e.res16 + (p32 * e.res_stride + e.res_coff + col)));
            const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(&rr.x));
            const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(&rr.y));
            x.a = make_float4(r0.x, r0.y, r1.x, r1.y);
        }
    } else if constexpr (MODE == EPI_GRU_ZR) {
        if (col >= 128) x.a = __ldcg(reinterpret_cast<const float4*>(e.h32 + (p32 * 128u + (col - 128))));
    } else if constexpr (MODE == EPI_GRU_Q) {
        x.a = __ldcg(reinterpret_cast<const float4*>(e.h32 + (p32 * 128u + col)));
        x.b = __ldcg(reinterpret_cast<const float4*>(e.z32 + (p32 * 128u + col)));
    } else if constexpr (MODE == EPI_FLOW) {
        if (col == 0) {
            const float2 c = __ldcg(reinterpret_cast<const float2*>(e.coords1 + p32 * 2u));
            x.a.x = c.x;
            x.a.y = c.y;
        }
    }
    return x;
}

__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
    return pk;
}

// `col` < n_valid is guaranteed by the caller; relu_lo = 0 (ReLU) or -inf (none) makes the activation branch-free.
template <int MODE, bool LEAN = false>
__device__ __forceinline__ void epilogue4(const ConvEpi& e, float4 v, const float4 bb, const EpiAux& ax, int col, long pix,
                                          bool full, float relu_lo) {
    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    const uint32_t p32 = static_cast<uint32_t>(pix);
    if constexpr (MODE == EPI_F16) {
        v.x = fmaxf(v.x, relu_lo); v.y = fmaxf(v.y, relu_lo); v.z = fmaxf(v.z, relu_lo); v.w = fmaxf(v.w, relu_lo);
        __half* o = e.out16 + (p32 * e.out16_stride + e.out16_coff + col);
        if (full) {
            if (!LEAN && e.res16 != nullptr) {
                v.x = fmaxf(v.x + ax.a.x, 0.f); v.y = fmaxf(v.y + ax.a.y, 0.f);
                v.z = fmaxf(v.z + ax.a.z, 0.f); v.w = fmaxf(v.w + ax.a.w, 0.f);
            }
            *reinterpret_cast<uint2*>(o) = pack_half4(v.x, v.y, v.z, v.w);
        } else {
            const float a[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4 && col + j < e.n_valid; ++j) {
                float x = a[j];
                if (!LEAN && e.res16 != nullptr) x = fmaxf(x + __half2float(e.res16[p32 * e.res_stride + e.res_coff + col + j]), 0.f);
                o[j] = __float2half_rn(x);
            }
        }
    } else if constexpr (MODE == EPI_F32) {
        v.x = fmaxf(v.x * e.scale, relu_lo); v.y = fmaxf(v.y * e.scale, relu_lo);
        v.z = fmaxf(v.z * e.scale, relu_lo); v.w = fmaxf(v.w * e.scale, relu_lo);
        if (e.out16 != nullptr) {          // same epilogue (64-bit addressing, scale), result stored as fp16: the correlation volume
            __half* o = e.out16 + pix * e.out16_stride + e.out16_coff + col;
            if (full && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
                *reinterpret_cast<uint2*>(o) = pack_half4(v.x, v.y, v.z, v.w);
            } else {
                const float a[4] = {v.x, v.y, v.z, v.w};
                for (int j = 0; j < 4 && col + j < e.n_valid; ++j) o[j] = __float2half_rn(a[j]);
            }
            return;
        }
        float* o = e.out32 + pix * e.out32_stride + e.out32_coff + col;
        if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            const float a[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4 && col + j < e.n_valid; ++j) o[j] = a[j];
        }
    } else if constexpr (MODE == EPI_CNET) {
        if (col < 128) {
            *reinterpret_cast<float4*>(e.out32 + (p32 * 128u + col)) = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
        } else {
            *reinterpret_cast<uint2*>(e.out16 + (p32 * 128u + (col - 128))) =
                pack_half4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        }
    } else if constexpr (MODE == EPI_GRU_ZR) {
        if (col < 128) {
            *reinterpret_cast<float4*>(e.z32 + (p32 * 128u + col)) =
                make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
        } else {
            *reinterpret_cast<uint2*>(e.out16 + (p32 * e.out16_stride + e.out16_coff + (col - 128))) =
                pack_half4(sigmoidf_(v.x) * ax.a.x, sigmoidf_(v.y) * ax.a.y, sigmoidf_(v.z) * ax.a.z, sigmoidf_(v.w) * ax.a.w);
        }
    } else if constexpr (MODE == EPI_GRU_Q) {
        const float4 hv = ax.a, zv = ax.b;
        float4 n;
        n.x = (1.0f - zv.x) * hv.x + zv.x * tanh_gate(v.x);
        n.y = (1.0f - zv.y) * hv.y + zv.y * tanh_gate(v.y);
        n.z = (1.0f - zv.z) * hv.z + zv.z * tanh_gate(v.z);
        n.w = (1.0f - zv.w) * hv.w + zv.w * tanh_gate(v.w);
        *reinterpret_cast<float4*>(e.h32 + (p32 * 128u + col)) = n;
        *reinterpret_cast<uint2*>(e.out16 + (p32 * e.out16_stride + e.out16_coff + col)) = pack_half4(n.x, n.y, n.z, n.w);
    } else if constexpr (MODE == EPI_FLOW) {
        if (col == 0) {
            *reinterpret_cast<float2*>(e.delta32 + p32 * 2u) = make_float2(v.x, v.y);
            *reinterpret_cast<float2*>(e.coords1 + p32 * 2u) = make_float2(ax.a.x + v.x, ax.a.y + v.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// epilogue of one 128-pixel x n_tile accumulator tile, run by all 8 warps of the CTA
// ------------------------------------------------------------------------------------------
// Warp w reads TMEM lane quarter w%4 (hardware restriction) and the 32-column chunks of parity w/4, so each quarter is
// drained by two warps.  The pipeline slots are idle by now (every TMA landed and every MMA read it): they are reused
// as the per-warp staging area, 2 x [32 rows][32 fp32] with the 16-byte chunks XOR-swizzled by row.
// STG = staging floats per warp: 2048 = double buffered, 1024 = single.  PRE: the epilogue's global inputs (GRU state,
// residual) of a whole column chunk are requested BEFORE the accumulator chunk is pulled out of TMEM and staged, so their
// L2 latency overlaps that work (the persistent kernel has the registers for it); bias then comes straight from global.
template <int MODE, int STG = 2048, int PRE = 0, bool LEAN = false>      // PRE: 0 = off, 1 = bias from global only, 2 = bias + input prefetch
__device__ __forceinline__ void tile_epilogue(const ConvGeom& g, const ConvEpi& e, uint8_t* smem, const float* bw,
                                              float* stat_s, uint32_t tmem_base, int warp, int lane, int tx, int ty,
                                              int b, int ny, long long* tstamp, const CUtensorMap* tmO = nullptr) {
    const int q = warp & 3;
    const int cpar = warp >> 2;
    float* stg = reinterpret_cast<float*>(smem) + warp * STG;
    const int sub = lane >> 3, cq = lane & 7;
    const int tw_mask = g.tile_w - 1;
    const int nchunk = (g.n_tile + 31) / 32;
    // the 8 pixel rows this lane serves are the same for every column chunk (32-bit except for the correlation volume)
    using PixT = typename std::conditional<MODE == EPI_F32, long, uint32_t>::type;
    PixT pixr[8];
    unsigned valid_bits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int row = q * 32 + k * 4 + sub;
        const int y = ty * g.tile_h + (row >> g.tile_w_log2), x = tx * g.tile_w + (row & tw_mask);
        if (y < g.H && x < g.W && b < g.b0 + g.nbatch) valid_bits |= 1u << k;
        pixr[k] = static_cast<PixT>((static_cast<PixT>(b) * g.H + y) * g.W + x);
    }
    for (int c = cpar; c < nchunk; c += 2) {
        EpiAux axp[PRE == 2 ? 8 : 1];
        if constexpr (PRE == 2) {
            const int colp = ny * g.n_tile + c * 32 + cq * 4;
            if (colp < e.n_valid) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((valid_bits >> k) & 1u) axp[k] = epi_prefetch<MODE, LEAN>(e, colp, pixr[k]);
            }
        }
        uint32_t r[32];
        const bool tt = tstamp && warp == 2 && lane == 0 && c < 4;
        if (tt) tstamp[8 + (c >> 1) * 4] = clock64();
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        if (tt) tstamp[9 + (c >> 1) * 4] = clock64();
        if constexpr (STG < 2048) __syncwarp();      // single staging buffer: everyone is done reading the previous chunk
        float* buf = stg + (STG >= 2048 ? ((c >> 1) & 1) * 1024 : 0);
        if constexpr (MODE == EPI_F32 && STG >= 2048) {
            if (tmO != nullptr) {
                // Bulk-store path (one-row tiles: the warp's 32 accumulator rows are 32 consecutive rows of the output
                // matrix).  thread = row: scale / bias / relu in registers, the 32 x 32 block staged in TMA's 128-byte
                // swizzle (the same XOR pattern as the transposing path), one cp.async.bulk.tensor store per block.
                const int col0 = ny * g.n_tile + c * 32;
                if (lane == 0) tma_store_wait_read<1>();          // the store that read this buffer two blocks ago is done
                __syncwarp();
                const float lo = e.relu ? 0.0f : -INFINITY;
                if (e.out16 != nullptr) {
                    // fp16 output (the correlation volume): 32 x 32 halves = 64-byte rows in TMA's 64-byte swizzle
                    // (16-byte chunk index ^ address bits 7..8, i.e. ^ (row >> 1) & 3 in a 512-byte aligned buffer)
                    uint8_t* hb = reinterpret_cast<uint8_t*>(buf);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float bia = e.bias != nullptr ? __ldg(e.bias + col0 + 8 * j + i) : 0.0f;
                            w[i] = fmaxf((__uint_as_float(r[8 * j + i]) + bia) * e.scale, lo);
                        }
                        const uint2 lo4 = pack_half4(w[0], w[1], w[2], w[3]), hi4 = pack_half4(w[4], w[5], w[6], w[7]);
                        *reinterpret_cast<uint4*>(hb + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(lo4.x, lo4.y, hi4.x, hi4.y);
                    }
                } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e.bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + 4 * j));
                    float4 v;
                    v.x = fmaxf((__uint_as_float(r[4 * j]) + bb.x) * e.scale, lo);
                    v.y = fmaxf((__uint_as_float(r[4 * j + 1]) + bb.y) * e.scale, lo);
                    v.z = fmaxf((__uint_as_float(r[4 * j + 2]) + bb.z) * e.scale, lo);
                    v.w = fmaxf((__uint_as_float(r[4 * j + 3]) + bb.w) * e.scale, lo);
                    *reinterpret_cast<float4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) = v;
                }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(tmO, buf, col0, tx * g.tile_w + q * 32, b);      // rows past the batch entry's end are clipped
                    tma_store_commit();
                }
                continue;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        if (tt) tstamp[10 + (c >> 1) * 4] = clock64();
        const int col = ny * g.n_tile + c * 32 + cq * 4;
        float4 bb;
        if constexpr (PRE != 0) bb = e.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(e.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        else bb = *reinterpret_cast<const float4*>(bw + c * 32 + cq * 4);
        const bool col_ok = col < e.n_valid, full = col + 4 <= e.n_valid;
        const float relu_lo = e.relu ? 0.0f : -INFINITY;
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
        if (col_ok) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 v[4];
            PixT pix[4];
            bool val[4];
            EpiAux ax[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rr = (half * 4 + k) * 4 + sub;
                v[k] = *reinterpret_cast<const float4*>(buf + rr * 32 + ((cq ^ (rr & 7)) << 2));
                val[k] = (valid_bits >> (half * 4 + k)) & 1u;
                pix[k] = pixr[half * 4 + k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if constexpr (PRE == 2) ax[k] = axp[half * 4 + k];
                else if (val[k]) ax[k] = epi_prefetch<MODE, LEAN>(e, col, pix[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (val[k]) epilogue4<MODE, LEAN>(e, v[k], bb, ax[k], col, pix[k], full, relu_lo);
            if constexpr (MODE == EPI_F16 && !LEAN) {
                if (e.stats != nullptr) {            // instance-norm statistics of the conv output (bias included)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (val[k]) {
                            const float w[4] = {v[k].x + bb.x, v[k].y + bb.y, v[k].z + bb.z, v[k].w + bb.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) { s4[j] += w[j]; q4[j] += w[j] * w[j]; }
                        }
                }
            }
        }
        }
        if constexpr (MODE == EPI_F16 && !LEAN) {
            if (e.stats != nullptr) {
                // deterministic: fixed-order shuffle over the 4 lanes that share these columns, one smem slot per
                // (lane quarter, column) written exactly once, quarters summed in order at the end of the CTA
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s4[j] += __shfl_xor_sync(0xffffffffu, s4[j], 8);
                    s4[j] += __shfl_xor_sync(0xffffffffu, s4[j], 16);
                    q4[j] += __shfl_xor_sync(0xffffffffu, q4[j], 8);
                    q4[j] += __shfl_xor_sync(0xffffffffu, q4[j], 16);
                }
                if (sub == 0) {
                    float* dst = stat_s + q * 512 + c * 32 + cq * 4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { dst[j] = s4[j]; dst[256 + j] = q4[j]; }
                }
            }
        }
        if (tt) tstamp[11 + (c >> 1) * 4] = clock64();
    }
    if constexpr (MODE == EPI_F32 && STG >= 2048) {
        if (tmO != nullptr) {
            if (lane == 0) tma_store_wait_read<0>();              // the staging area may be reused / the CTA may exit
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// tensor-core kernel
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 256;   // warp 0: TMA producer, warp 1: MMA issuer; afterwards ALL 8 warps run the epilogue

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const ConvGeom g, const ConvEpi e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = kTileM * 128;
    const uint32_t b_bytes = static_cast<uint32_t>(g.n_tile) * 128;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    size_t pipe_bytes = static_cast<size_t>(g.stages) * stage_bytes;
    if (pipe_bytes < 72 * 1024) pipe_bytes = 72 * 1024;            // room for the epilogue staging area (8 warps x 8 KiB) + channel sums
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + pipe_bytes);
    uint64_t* empty = full + g.stages;
    uint64_t* accum_ready = empty + g.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_ready + 1);
    float* bias_s = reinterpret_cast<float*>(smem + pipe_bytes + 256);   // [8 warps][256]
    // [4 lane quarters][2][256] channel sums (stats mode): lives in the idle pipeline area right after the 64 KiB of
    // epilogue staging; every slot that is read is written exactly once, so it needs no zeroing
    float* stat_s = reinterpret_cast<float*>(smem + 64 * 1024);

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    long long* tstamp = e.timing ? e.timing + (static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
    if (tstamp && threadIdx.x == 0) {
        tstamp[0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        tstamp[7] = static_cast<long long>(gt);
    }

    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = g.b0 + t / g.tiles_y;
    const int ny = blockIdx.y;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full[s], 2);                                   // A producer + B producer
            mbar_init(&empty[s], static_cast<uint32_t>(g.cluster));   // one release per consumer CTA of the cluster
        }
        mbar_init(accum_ready, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, static_cast<uint32_t>(g.tmem_cols));
        tmem_relinquish();
    }
    tc_fence_before();
    if (g.cluster > 1) cluster_sync_all(); else __syncthreads();   // barriers visible cluster-wide before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail
    // of the previous kernel in the stream; from here on we touch data it produced.
    pdl_launch_dependents();
    pdl_wait();
    if (tstamp && threadIdx.x == 0) tstamp[1] = clock64();
    const int T = g.ntaps * g.kchunks;
    const uint32_t crank = g.cluster > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = static_cast<uint16_t>((1u << g.cluster) - 1u);
    bool ok = true;

    // Warp roles are warp-uniform (warp index via shuffle, loops run by all 32 lanes, the asynchronous instructions are
    // issued by one elected lane): the compiler then keeps descriptors, coordinates and barrier addresses in uniform
    // registers and emits UTMALDG / UTCHMMA / UTCBAR back to back.  Issuing them from a divergent `lane == 0` branch
    // instead costs an ELECT / R2UR / branch "waterfall" per instruction: ~67 cycles per MMA for a single thread,
    // which capped a CTA at one K=64 stage per ~600 cycles whatever N (tools/mma_probe.cu).
    if (warp == 0) {
        // A-operand producer: one shifted activation box per (tap, 64-channel chunk)
        const int x0 = g.stride * tx * g.tile_w;
        const int y0 = g.stride * ty * g.tile_h;
        int kc = 0, kx = 0, ky = 0, s = 0;
        uint32_t ph = 0;
        const int rx = g.kw / 2, ry = g.kh / 2;
        for (int it = 0; it < T; ++it) {
            if (!__all_sync(0xffffffffu, mbar_wait(&empty[s], ph ^ 1))) { ok = false; break; }
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], a_bytes);
                tma_load_4d(smem + static_cast<size_t>(s) * stage_bytes, &tmA, &full[s], kc * kChunkK, x0 + kx - rx,
                            y0 + ky - ry, b);
            }
            __syncwarp();
            if (++kc == g.kchunks) {
                kc = 0;
                if (++kx == g.kw) { kx = 0; ++ky; }
            }
            if (++s == g.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 2) {
        // B-operand producer (its own warp: the two box streams are issued independently)
        const int brow = b * g.b_rows_per_batch + ny * g.n_tile;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < T; ++it) {
            if (!__all_sync(0xffffffffu, mbar_wait(&empty[s], ph ^ 1))) { ok = false; break; }
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], b_bytes);
                uint8_t* sb = smem + static_cast<size_t>(s) * stage_bytes + a_bytes;
                if (g.cluster > 1) {
                    // each CTA fetches 1/cluster of the B slab and multicasts it to all CTAs of the cluster
                    const int slice = g.n_tile / g.cluster;
                    tma_load_2d_mc(sb + crank * slice * 128, &tmB, &full[s], it * kChunkK,
                                   brow + static_cast<int>(crank) * slice, cmask);
                } else {
                    tma_load_2d(sb, &tmB, &full[s], it * kChunkK, brow);
                }
            }
            __syncwarp();
            if (++s == g.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_f16(kTileM, g.n_tile);
        // descriptors: constant high words, low word = (address >> 4) | LBO field; +2 per K=16 step (32 bytes)
        const uint64_t d0 = umma_desc_k128(smem_u32(smem));
        const uint32_t dhi = static_cast<uint32_t>(d0 >> 32);
        uint32_t alo = static_cast<uint32_t>(d0);
        const uint32_t alo0 = alo, stage_lo = stage_bytes >> 4, b_off_lo = a_bytes >> 4;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < T; ++it) {
            if (!__all_sync(0xffffffffu, mbar_wait(&full[s], ph))) { ok = false; break; }
            tc_fence_after();
            if (elect_one()) {
                if (tstamp && it == 0) tstamp[2] = clock64();
                const uint32_t blo = alo + b_off_lo;
                umma_f16_lohi(tmem_base, alo, dhi, blo, dhi, idesc, it != 0 ? 1u : 0u);
                umma_f16_lohi(tmem_base, alo + 2, dhi, blo + 2, dhi, idesc, 1u);
                umma_f16_lohi(tmem_base, alo + 4, dhi, blo + 4, dhi, idesc, 1u);
                umma_f16_lohi(tmem_base, alo + 6, dhi, blo + 6, dhi, idesc, 1u);
                // slot reusable once these MMAs have read it (in every CTA that multicasts into it)
                if (g.cluster > 1) umma_commit_mc(&empty[s], cmask); else umma_commit(&empty[s]);
            }
            __syncwarp();
            alo += stage_lo;
            if (++s == g.stages) { s = 0; ph ^= 1; alo = alo0; }
        }
        if (elect_one()) {
            if (tstamp) tstamp[3] = clock64();
            umma_commit(accum_ready);     // accumulator complete
        }
        __syncwarp();
    }
    {
        // ---- epilogue: all 8 warps.  Warp w reads TMEM lane quarter w%4 (hardware restriction) and the
        //      32-column chunks of parity w/4, so each quarter is drained by two warps.
        // bias of this CTA's couts -> this warp's smem copy, while the main loop runs
        float* bw = bias_s + warp * 256;
        for (int i = lane; i < 256; i += 32)
            bw[i] = (e.bias != nullptr && i < ((g.n_tile + 31) & ~31)) ? __ldg(e.bias + ny * g.n_tile + i) : 0.0f;
        __syncwarp();
        const bool ok_acc = mbar_wait(accum_ready, 0);
        ok = ok && ok_acc;
        if (tstamp && warp == 2 && lane == 0) tstamp[4] = clock64();
        tc_fence_after();
        if (ok_acc) tile_epilogue<MODE>(g, e, smem, bw, stat_s, tmem_base, warp, lane, tx, ty, b, ny, tstamp, e.tma_store ? &tmO : nullptr);
    }
    if (tstamp && warp == 2 && lane == 0) tstamp[5] = clock64();
    if (!ok && e.err_flag != nullptr) atomicExch(e.err_flag, 1 + warp);
    tc_fence_before();
    // no CTA may exit while a peer can still multicast into its smem / arrive on its barriers
    if (g.cluster > 1) cluster_sync_all(); else __syncthreads();
    if (e.stats != nullptr) {
        const int cg = ny * g.n_tile + threadIdx.x;
        if (threadIdx.x < g.n_tile && cg < e.n_valid) {
            const float* p0 = stat_s + threadIdx.x;
            const float sum = ((p0[0] + p0[512]) + p0[1024]) + p0[1536];
            const float sq = ((p0[256] + p0[768]) + p0[1280]) + p0[1792];
            atomicAdd(e.stats + cg, static_cast<double>(sum));
            atomicAdd(e.stats + e.n_valid + cg, static_cast<double>(sq));
        }
    }
    if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(g.tmem_cols));
    if (tstamp && threadIdx.x == 0) tstamp[6] = clock64();
}

// ------------------------------------------------------------------------------------------
// persistent layer-program kernel: tile-level dataflow between dependent convolutions
// ------------------------------------------------------------------------------------------
// One launch runs a whole chain of convolutions (the 11 of a GRU iteration).  One resident CTA per SM pops work items
// (layer, batch entry, tile) from a ready queue.  A tile of layer L becomes ready -- is pushed -- as soon as the 3x3
// tile neighbourhood of its predecessor layer(s) is complete (per-tile arrival counters bumped with gpu-scope acq_rel
// atomics + generic->async proxy fences, because the producer writes with st.global and the consumer reads with
// TMA) -- not when the whole previous layer is.  Compared with one launch per layer this removes the per-launch drain /
// fill bubbles (~7 us of every ~20 us launch at 512^2) and the 224-tiles-on-148-SMs wave quantisation: an SM that
// is done with its share of layer L moves on to layer L+1.
// A CTA only ever waits for a queue slot to be filled, i.e. for some running tile to complete, never for work that
// has not been handed out, so the scheme cannot deadlock whatever the number of resident CTAs; every wait is bounded
// and aborts the launch through err_flag instead of hanging.
// The same 3x3 rule also covers the write-after-read hazards of the in-place GRU record (h and r*h slices of X are
// overwritten by a later layer of the chain only after every tile that reads their halo is complete).
// Shared-memory map (one CTA per SM, 384 threads):
//   [0, 192K)     operand ring: 4 slots x 48 KiB (A box 16 KiB + B slab <= 32 KiB), the same slots for every layer
//   [192K, 224K)  epilogue staging, 8 warps x 4 KiB
//   [224K, 225K)  control block: mbarriers, TMEM base, ticket ring
// Warp roles:  0 = A-operand producer   1 = MMA issuer   2 = B-operand producer   3 = scheduler (tickets + dependencies)
//              4..11 = epilogue (warp w drains TMEM lane quarter w % 4, column chunks of parity (w - 4) / 4)
// The last epilogue warp to finish a tile raises its completion flag.
// Everything is decoupled by mbarriers, so that while the epilogue warps drain tile n from one TMEM accumulator the
// MMA warp already accumulates tile n+1 into the other, the producers prefetch its operands and the scheduler waits
// for the dependencies of tile n+2.
constexpr int kProgThreads = 384;      // 12 warps = 3 per SM sub-partition: up to 168 registers per thread
constexpr int kProgStages = 4;
constexpr uint32_t kProgSlotBytes = 48 * 1024;
constexpr uint32_t kProgStagingOff = kProgStages * kProgSlotBytes;
constexpr uint32_t kProgCtlOff = kProgStagingOff + 8 * 4096;
constexpr uint32_t kProgSmemBytes = kProgCtlOff + 1024;
constexpr int kProgTickets = 4;
constexpr uint32_t kTicketEnd = 0xffffffffu;

struct ProgCtl {
    uint64_t full[kProgStages], empty[kProgStages];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t tk_full[kProgTickets], tk_empty[kProgTickets];
    uint32_t ticket[kProgTickets];
    uint32_t stored[kProgTickets];            // epilogue warps that have stored their part of the tile
    uint32_t tmem_base;
    uint32_t abort;
    uint32_t depth;                           // tickets in flight per CTA (<= kProgTickets)
};
static_assert(sizeof(ProgCtl) <= 1024, "control block");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Correlation-pyramid lookup of one tile's 128 pixels by the 8 epilogue warps: 16 pixels per warp in groups of kLkGroup
// consecutive pixels of one tile row (kLkWinFloats floats of the warp's staging area hold a level's windows).  Out of
// line: its registers are then allocated apart from the convolution epilogues' (a spill costs an L2 round trip here --
// the L1 is carved out for shared memory).
__device__ __noinline__ void prog_lookup_tile(const LookupArgs& lk, const ConvGeom& g, float* win, int ew, int lane, int tx,
                                              int ty, int b) {
    const LookupLane t = lookup_lane_init(lane);
    for (int r = 0; r < 16; r += kLkGroup) {
        const int row = ew * 16 + r;                                   // kLkGroup divides the tile width: one tile row per group
        const int y = ty * g.tile_h + (row >> g.tile_w_log2), x = tx * g.tile_w + (row & (g.tile_w - 1));
        if (y >= g.H || x >= g.W) continue;
        const int nvalid = g.W - x < kLkGroup ? g.W - x : kLkGroup;
        lookup_group(lk, t, (static_cast<long>(b) * g.H + y) * g.W + x, nvalid, lane, win);
        __syncwarp();
    }
}

// item = ((iteration * n_layers + layer) * nbatch + batch) * tiles + tile
__device__ __forceinline__ void prog_decode(const ConvProgram& P, uint32_t item, int& it, int& l, int& b, int& tile) {
    const int tpp = P.tiles_x * P.tiles_y;
    const int per_layer = P.nbatch * tpp;
    const int il = static_cast<int>(item / static_cast<uint32_t>(per_layer));
    const int rem = static_cast<int>(item) - il * per_layer;
    it = il / P.n_layers;
    l = il - it * P.n_layers;
    const int pair = rem / tpp;
    tile = rem - pair * tpp;
    b = P.b0 + pair;
}
__device__ __forceinline__ void prog_decode(const ConvProgram& P, uint32_t item, int& l, int& b, int& tile) {
    int it;
    prog_decode(P, item, it, l, b, tile);
}

// Waits for ticket slot n % kProgTickets and returns its ticket (kTicketEnd on time-out, which ends the role).
__device__ __forceinline__ uint32_t prog_take_ticket(volatile ProgCtl* ctl, uint32_t n) {
    const uint32_t slot = n % ctl->depth, par = (n / ctl->depth) & 1u;
    if (!__all_sync(0xffffffffu, mbar_wait(const_cast<uint64_t*>(&ctl->tk_full[slot]), par))) {
        ctl->abort = 50;
        return kTicketEnd;
    }
    return ctl->ticket[slot];
}
__device__ __forceinline__ void prog_release_ticket(volatile ProgCtl* ctl, uint32_t n) {
    __syncwarp();
    if (elect_one()) mbar_arrive(const_cast<uint64_t*>(&ctl->tk_empty[n % ctl->depth]));
}

// LOOKUP: the program may contain a lookup layer (compiled apart: its code would only cost the plain program registers)
template <bool LOOKUP>
__global__ void __launch_bounds__(kProgThreads, 1)
conv_prog_kernel(const __grid_constant__ ConvProgram P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    volatile ProgCtl* ctl = reinterpret_cast<volatile ProgCtl*>(smem + kProgCtlOff);
    uint64_t* full = const_cast<uint64_t*>(ctl->full);
    uint64_t* empty = const_cast<uint64_t*>(ctl->empty);
    uint64_t* acc_full = const_cast<uint64_t*>(ctl->acc_full);
    uint64_t* acc_empty = const_cast<uint64_t*>(ctl->acc_empty);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kProgStages; ++s) {
            mbar_init(&full[s], 2);              // A producer + B producer
            mbar_init(&empty[s], 1);             // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);          // tcgen05.commit
            mbar_init(&acc_empty[i], 8);         // one arrive per epilogue warp
        }
        for (int i = 0; i < kProgTickets; ++i) {
            mbar_init(const_cast<uint64_t*>(&ctl->tk_full[i]), 1);      // scheduler
            mbar_init(const_cast<uint64_t*>(&ctl->tk_empty[i]), 11);    // A, B, MMA + 8 epilogue warps
            ctl->stored[i] = 0;
        }
        fence_mbar_init();
        ctl->abort = 0;
        ctl->depth = static_cast<uint32_t>(P.tickets);
    }
    if (threadIdx.x < 2 * P.n_layers) {
        const ProgLayer& L = P.L[threadIdx.x >> 1];
        tma_prefetch_desc((threadIdx.x & 1) ? &L.tmB : &L.tmA);
    }
    if (warp == 1) {
        tmem_alloc(const_cast<uint32_t*>(&ctl->tmem_base), 512);     // two 256-column accumulators
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();
    pdl_wait();

    const uint32_t total = static_cast<uint32_t>(P.iters * P.n_layers * P.nbatch * P.tiles_x * P.tiles_y);   // < 2^31 (host-checked)
    const int tpp = P.tiles_x * P.tiles_y;

    if (warp == 3) {
        // ================= scheduler: ready queue in, tickets out, completions -> successor tiles ======================
        const long set_elems = static_cast<long>(kMaxProgLayers) * P.max_batch * tpp;
        int* arr_cur = P.arrivals + static_cast<long>(P.epoch & 1u) * set_elems;
        {
            int* arr_nxt = P.arrivals + static_cast<long>((P.epoch & 1u) ^ 1u) * set_elems;      // cleared for the next launch
            for (long i = static_cast<long>(blockIdx.x) * 32 + lane; i < set_elems; i += static_cast<long>(gridDim.x) * 32) arr_nxt[i] = 0;
        }
        const unsigned long long tag = static_cast<unsigned long long>(P.epoch) << 32;
        auto push = [&](uint32_t item) {
            if (P.static_order) return;             // the item's own CTA polls its arrival counter
            const unsigned long long slot = atomicAdd(P.tail, 1ull) - P.tail_base;
            st_release_gpu_u64(P.queue + slot, tag | item);
        };
        const int per_layer = P.nbatch * tpp;
        for (int l = 0; l < P.n_layers; ++l) {          // roots: iteration 0 of the layers without same-iteration predecessors
            if (P.L[l].n_dep != 0 && P.L[l].iter_shift == 0) continue;
            for (int r = static_cast<int>(blockIdx.x) * 32 + lane; r < per_layer; r += static_cast<int>(gridDim.x) * 32)
                push(static_cast<uint32_t>(l * per_layer + r));
        }
        const uint32_t depth = static_cast<uint32_t>(P.tickets);
        uint32_t n_issue = 0, n_done = 0, my_item = 0, idx = 0, spins = 0, give_up = 0;
        unsigned long long t_idle = 0;
        bool have_idx = false, end_posted = false;
        for (;;) {
            bool progress = false;
            // ---- completed tiles of this CTA, in ticket order: count in at every successor tile, push those that become ready
            while (n_done < n_issue) {
                const uint32_t slot = n_done % depth;
                uint32_t st;
                asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];"
                             : "=r"(st) : "r"(smem_u32(const_cast<uint32_t*>(&ctl->stored[slot]))) : "memory");
                if (__shfl_sync(0xffffffffu, st, 0) != 8u) break;
                const uint32_t item = __shfl_sync(0xffffffffu, my_item, static_cast<int>(slot));
                int it, l, b, tile;
                prog_decode(P, item, it, l, b, tile);
                const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
                const ProgLayer& L = P.L[l];
                if (P.timing != nullptr && lane == 0) {                     // tuning aid: when did this layer's last tile finish
                    unsigned long long gt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                    atomicMax(reinterpret_cast<unsigned long long*>(P.timing) + 4096 + 2 * l + 1, gt);
                }
#pragma unroll
                for (int si = 0; si < 4; ++si) {            // one pass per successor layer, one lane per tile of its 5x5 reach
                    const int sl = L.succ[si];
                    if (sl < 0 || lane >= 25) continue;
                    const ProgLayer& S = P.L[sl];
                    const int dy = lane / 5 - 2, dx = lane - (lane / 5) * 5 - 2;
                    const int nty = ty + dy, ntx = tx + dx;
                    if (abs(dy) > S.ry || abs(dx) > S.rx || nty < 0 || nty >= P.tiles_y || ntx < 0 || ntx >= P.tiles_x) continue;
                    // the successor tile of iteration `its` is ready when all (its - shift + 1) rounds of arrivals are in
                    // (a round = every predecessor layer x every tile within the dependency radius); rounds cannot mix because
                    // round r+1 only starts arriving after the tile itself has run in round r (it is their ancestor)
                    const int its = it + S.iter_shift;
                    if (its >= P.iters) continue;
                    const int nn = (1 + min(S.rx, ntx) + min(S.rx, P.tiles_x - 1 - ntx)) * (1 + min(S.ry, nty) + min(S.ry, P.tiles_y - 1 - nty));
                    const int nt = nty * P.tiles_x + ntx;
                    const int old = atom_add_acq_rel_gpu(arr_cur + (static_cast<long>(sl * P.max_batch + b) * tpp + nt), 1);
                    if (old + 1 == (it + 1) * S.n_dep * nn)
                        push(static_cast<uint32_t>(((its * P.n_layers + sl) * P.nbatch + (b - P.b0)) * tpp + nt));
                }
                __syncwarp();
                if (lane == 0) ctl->stored[slot] = 0;
                ++n_done;
                progress = true;
            }
            if (end_posted) {
                if (n_done == n_issue) break;
            } else if (n_issue - n_done < depth) {
                const uint32_t slot = n_issue % depth;
                uint32_t room = 0;
                if (lane == 0) room = mbar_try_wait(const_cast<uint64_t*>(&ctl->tk_empty[slot]), ((n_issue / depth) & 1u) ^ 1u) ? 1u : 0u;
                if (__shfl_sync(0xffffffffu, room, 0) != 0u) {
                    if (!have_idx) {
                        if (lane == 0) {
                            const unsigned long long d = P.static_order ? static_cast<unsigned long long>(blockIdx.x) + static_cast<unsigned long long>(n_issue) * gridDim.x
                                                                        : atomicAdd(P.head, 1ull) - P.head_base;
                            idx = d < total ? static_cast<uint32_t>(d) : kTicketEnd;
                        }
                        idx = __shfl_sync(0xffffffffu, idx, 0);
                        have_idx = true;
                        progress = true;
                    }
                    if (idx == kTicketEnd || ctl->abort != 0) {
                        if (lane == 0) {
                            ctl->ticket[slot] = kTicketEnd;
                            mbar_arrive(const_cast<uint64_t*>(&ctl->tk_full[slot]));
                        }
                        end_posted = true;
                        progress = true;
                    } else {
                        unsigned long long entry = 0;
                        if (P.static_order) {
                            // item idx itself: ready when every round of arrivals up to its iteration is in (roots: at once)
                            if (lane == 0) {
                                int its, sl, b, nt;
                                prog_decode(P, idx, its, sl, b, nt);
                                const ProgLayer& S = P.L[sl];
                                const int rounds = its - S.iter_shift + 1;
                                bool ready = S.n_dep == 0 || rounds <= 0;
                                if (!ready) {
                                    const int nty = nt / P.tiles_x, ntx = nt - nty * P.tiles_x;
                                    const int nn = (1 + min(S.rx, ntx) + min(S.rx, P.tiles_x - 1 - ntx)) * (1 + min(S.ry, nty) + min(S.ry, P.tiles_y - 1 - nty));
                                    ready = ld_acquire_gpu(arr_cur + (static_cast<long>(sl * P.max_batch + b) * tpp + nt)) >= rounds * S.n_dep * nn;
                                }
                                entry = ready ? ((static_cast<unsigned long long>(P.epoch) << 32) | idx) : 0ull;
                            }
                        } else if (lane == 0) {
                            entry = ld_acquire_gpu_u64(P.queue + idx);
                        }
                        entry = __shfl_sync(0xffffffffu, entry, 0);
                        if ((entry >> 32) == P.epoch) {
                            const uint32_t item = static_cast<uint32_t>(entry);
                            if (P.timing != nullptr && lane == 0) {         // tuning aid: when was this layer's first tile handed out
                                unsigned long long gt;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                                int it2, l2, b2, t2;
                                prog_decode(P, item, it2, l2, b2, t2);
                                atomicMin(reinterpret_cast<unsigned long long*>(P.timing) + 4096 + 2 * l2, gt);
                            }
                            fence_proxy_async_all();
                            if (lane == static_cast<int>(slot)) my_item = item;
                            if (lane == 0) {
                                ctl->ticket[slot] = item;
                                mbar_arrive(const_cast<uint64_t*>(&ctl->tk_full[slot]));   // release: ticket + everything acquired above
                            }
                            ++n_issue;
                            have_idx = false;
                            progress = true;
                        }
                    }
                }
            }
            if (progress) {
                spins = 0;
            } else {
                // time-based bound (see mbar_wait): after 2 s without progress raise the abort flag, give the role warps
                // a few more rounds to see it, then leave
                if (++spins == 1024u) t_idle = global_timer_ns();
                if (spins > 1024u && (spins & 255u) == 0u && global_timer_ns() - t_idle > kWaitTimeoutNs) {
                    ctl->abort = 70;
                    if (++give_up > 64u) break;
                }
                __nanosleep(20);
            }
        }
    } else if (warp == 0) {
        // ================= A-operand producer ==========================================================================
        uint32_t it_glob = 0;                    // stage counter over the whole launch: slot = it % 4, parity = (it / 4) & 1
        for (uint32_t n = 0;; ++n) {
            const uint32_t item = prog_take_ticket(ctl, n);
            if (item == kTicketEnd) break;
            int l, b, tile;
            prog_decode(P, item, l, b, tile);
            const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
            const ProgLayer& L = P.L[l];
            const ConvGeom& g = L.g;
            fence_proxy_async_all();
            const int x0 = g.stride * tx * g.tile_w, y0 = g.stride * ty * g.tile_h;
            const int rx = g.kw / 2, ry = g.kh / 2;
            const int T = L.kind == 0 ? g.ntaps * g.kchunks : 0;        // lookup tiles have no operands
            int kc = 0, kx = 0, ky = 0;
            bool ok = true;
            for (int it = 0; it < T; ++it, ++it_glob) {
                const uint32_t s = it_glob % kProgStages;
                if (!__all_sync(0xffffffffu, mbar_wait(&empty[s], ((it_glob / kProgStages) & 1u) ^ 1u))) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], kTileM * 128);
                    tma_load_4d(smem + s * kProgSlotBytes, &L.tmA, &full[s], kc * kChunkK, x0 + kx - rx, y0 + ky - ry, b);
                }
                __syncwarp();
                if (++kc == g.kchunks) {
                    kc = 0;
                    if (++kx == g.kw) { kx = 0; ++ky; }
                }
            }
            if (!ok) { ctl->abort = 1; break; }
            prog_release_ticket(ctl, n);
        }
    } else if (warp == 2) {
        // ================= B-operand producer ==========================================================================
        uint32_t it_glob = 0;
        for (uint32_t n = 0;; ++n) {
            const uint32_t item = prog_take_ticket(ctl, n);
            if (item == kTicketEnd) break;
            int l, b, tile;
            prog_decode(P, item, l, b, tile);
            const ProgLayer& L = P.L[l];
            const ConvGeom& g = L.g;
            const int T = L.kind == 0 ? g.ntaps * g.kchunks : 0;
            const int brow = b * g.b_rows_per_batch + L.ny * g.n_tile;
            const uint32_t b_bytes = static_cast<uint32_t>(g.n_tile) * 128;
            bool ok = true;
            for (int it = 0; it < T; ++it, ++it_glob) {
                const uint32_t s = it_glob % kProgStages;
                if (!__all_sync(0xffffffffu, mbar_wait(&empty[s], ((it_glob / kProgStages) & 1u) ^ 1u))) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[s], b_bytes);
                    tma_load_2d(smem + s * kProgSlotBytes + kTileM * 128, &L.tmB, &full[s], it * kChunkK, brow);
                }
                __syncwarp();
            }
            if (!ok) { ctl->abort = 3; break; }
            prog_release_ticket(ctl, n);
        }
    } else if (warp == 1) {
        // ================= MMA issuer ==================================================================================
        const uint32_t tmem_base = ctl->tmem_base;
        const uint64_t d0 = umma_desc_k128(smem_u32(smem));
        const uint32_t dhi = static_cast<uint32_t>(d0 >> 32), alo0 = static_cast<uint32_t>(d0);
        uint32_t it_glob = 0;
        long long t_ticket = 0, t_acc = 0, t_full = 0, t_all = clock64(), n_tiles = 0, n_stages = 0;
        for (uint32_t n = 0;; ++n) {
            long long t0 = clock64();
            const uint32_t item = prog_take_ticket(ctl, n);
            t_ticket += clock64() - t0;
            if (item == kTicketEnd) break;
            int l, b, tile;
            prog_decode(P, item, l, b, tile);
            const ConvGeom& g = P.L[l].g;
            const int T = P.L[l].kind == 0 ? g.ntaps * g.kchunks : 0;   // lookup tile: just hand the (unused) accumulator over
            const uint32_t idesc = umma_idesc_f16(kTileM, g.n_tile);
            const uint32_t buf = n & 1u;
            t0 = clock64();
            bool ok = __all_sync(0xffffffffu, mbar_wait(&acc_empty[buf], ((n >> 1) & 1u) ^ 1u));     // accumulator drained
            t_acc += clock64() - t0;
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * 256;
            ++n_tiles;
            n_stages += T;
            for (int it = 0; ok && it < T; ++it, ++it_glob) {
                const uint32_t s = it_glob % kProgStages;
                t0 = clock64();
                if (!__all_sync(0xffffffffu, mbar_wait(&full[s], (it_glob / kProgStages) & 1u))) { ok = false; break; }
                t_full += clock64() - t0;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t alo = alo0 + s * (kProgSlotBytes >> 4), blo = alo + ((kTileM * 128) >> 4);
                    umma_f16_lohi(acc, alo, dhi, blo, dhi, idesc, it != 0 ? 1u : 0u);
                    umma_f16_lohi(acc, alo + 2, dhi, blo + 2, dhi, idesc, 1u);
                    umma_f16_lohi(acc, alo + 4, dhi, blo + 4, dhi, idesc, 1u);
                    umma_f16_lohi(acc, alo + 6, dhi, blo + 6, dhi, idesc, 1u);
                    umma_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (!ok) { ctl->abort = 2; break; }
            if (elect_one()) umma_commit(&acc_full[buf]);
            prog_release_ticket(ctl, n);
        }
        if (lane == 0 && P.timing != nullptr) {
            long long* o = P.timing + blockIdx.x * 16;
            o[0] = clock64() - t_all; o[1] = t_ticket; o[2] = t_acc; o[3] = t_full; o[4] = n_tiles; o[5] = n_stages;
        }
    } else {
        // ================= epilogue warps 4..11 ========================================================================
        const int ew = warp - 4;
        for (uint32_t n = 0;; ++n) {
            const uint32_t item = prog_take_ticket(ctl, n);
            if (item == kTicketEnd) break;
            int l, b, tile;
            prog_decode(P, item, l, b, tile);
            const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
            const ProgLayer& L = P.L[l];
            const ConvGeom& g = L.g;
            const ConvEpi& e = L.e;
            const uint32_t buf = n & 1u;
            const bool ok_acc = __all_sync(0xffffffffu, mbar_wait(&acc_full[buf], (n >> 1) & 1u));
            if (!ok_acc) ctl->abort = 4 + ew;      // keep going: the scheduler ends the launch
            tc_fence_after();
            const uint32_t acc = ctl->tmem_base + buf * 256;
            uint8_t* stg = smem + kProgStagingOff;
            if (LOOKUP && ok_acc && L.kind == 1) {
                const long long t0 = clock64();
                prog_lookup_tile(P.lk, g, reinterpret_cast<float*>(stg) + ew * 1024, ew, lane, tx, ty, b);
                if (warp == 4 && lane == 0 && P.timing != nullptr) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(P.timing + blockIdx.x * 16 + 8), static_cast<unsigned long long>(clock64() - t0));
                    atomicAdd(reinterpret_cast<unsigned long long*>(P.timing + blockIdx.x * 16 + 9), 1ull);
                }
            } else if (ok_acc) {
                switch (L.mode) {
                    case EPI_F16: tile_epilogue<EPI_F16, 1024, 1, true>(g, e, stg, nullptr, nullptr, acc, ew, lane, tx, ty, b, L.ny, nullptr); break;
                    case EPI_F32: tile_epilogue<EPI_F32, 1024, 1, true>(g, e, stg, nullptr, nullptr, acc, ew, lane, tx, ty, b, L.ny, nullptr); break;
                    case EPI_GRU_ZR: tile_epilogue<EPI_GRU_ZR, 1024, LOOKUP ? 1 : 2, true>(g, e, stg, nullptr, nullptr, acc, ew, lane, tx, ty, b, L.ny, nullptr); break;
                    case EPI_GRU_Q: tile_epilogue<EPI_GRU_Q, 1024, 1, true>(g, e, stg, nullptr, nullptr, acc, ew, lane, tx, ty, b, L.ny, nullptr); break;
                    case EPI_FLOW: tile_epilogue<EPI_FLOW, 1024, 2, true>(g, e, stg, nullptr, nullptr, acc, ew, lane, tx, ty, b, L.ny, nullptr); break;
                    default: break;
                }
            }
            // accumulator free for tile n+2 as soon as this warp's TMEM loads are done (tcgen05.wait::ld inside)
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(&acc_empty[buf]);
            // every thread orders its generic-proxy stores before later async-proxy reads (TMA of a consumer tile), then the
            // warp counts in (release); the scheduler acquires the count and announces the tile with gpu-scope acq_rel
            // atomics, which are cumulative over everything the eight warps stored
            fence_proxy_async_all();
            __syncwarp();
            if (elect_one())
                asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;"
                             ::"r"(smem_u32(const_cast<uint32_t*>(&ctl->stored[n % ctl->depth]))) : "memory");
            prog_release_ticket(ctl, n);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && ctl->abort != 0 && P.err_flag != nullptr) atomicExch(P.err_flag, static_cast<int>(ctl->abort));
    if (warp == 1) tmem_dealloc(ctl->tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// variant 2: 256-pixel tile (16 x 16), haloed A operand, B slab shared by the two 128-row halves
// ------------------------------------------------------------------------------------------
// The main loop of variant 1 is bound by shared-memory ingress (A box + B slab per tap): 96 B/clk/SM wanted,
// ~50 delivered.  Here one CTA owns a 16x16 pixel tile = two 128-row MMA halves (columns 0-7 / 8-15):
//  * per 64-channel chunk ONE haloed activation box (pitch `pxp` pixels x `py` rows) is loaded; every tap of the
//    window addresses it through a row-shifted UMMA descriptor (8-row groups = 8 consecutive x of one tile row,
//    group stride = haloed row pitch, base-offset = the shift within the 1024-byte swizzle atom);
//  * each weight slab is loaded once and multiplied into both halves (two TMEM accumulators).
// 3x3 / N=256: 343 KiB instead of 864 KiB of smem ingress per 256 pixels and chunk.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGeom g,
                const ConvEpi e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = static_cast<uint32_t>(g.pxp * g.py) * 128;
    const uint32_t b_bytes = static_cast<uint32_t>(g.n_tile) * 128;
    uint8_t* smem_b = smem + static_cast<size_t>(g.na) * a_bytes;
    size_t pipe_bytes = static_cast<size_t>(g.na) * a_bytes + static_cast<size_t>(g.nb) * b_bytes;
    if (pipe_bytes < 64 * 1024) pipe_bytes = 64 * 1024;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + pipe_bytes);
    uint64_t* a_empty = a_full + g.na;
    uint64_t* b_full = a_empty + g.na;
    uint64_t* b_empty = b_full + g.nb;
    uint64_t* accum_ready = b_empty + g.nb;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_ready + 1);
    float* bias_s = reinterpret_cast<float*>(smem + pipe_bytes + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    long long* tstamp = e.timing ? e.timing + (static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
    if (tstamp && threadIdx.x == 0) {
        tstamp[0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        tstamp[7] = static_cast<long long>(gt);
    }
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = g.b0 + t / g.tiles_y;
    const int ny = blockIdx.y;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < g.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < g.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(accum_ready, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, static_cast<uint32_t>(g.tmem_cols));
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    if (tstamp && threadIdx.x == 0) tstamp[1] = clock64();
    const int rx = g.kw / 2, ry = g.kh / 2;
    bool ok = true;

    // warp-uniform role loops, asynchronous instructions issued by one elected lane (see conv_tc_kernel)
    if (warp == 0) {
        // A producer: one haloed activation box per 64-channel chunk
        const int x0 = tx * 16 - rx, y0 = ty * 16 - ry;
        int sa = 0;
        uint32_t pa = 0;
        for (int kc = 0; kc < g.kchunks; ++kc) {
            if (!__all_sync(0xffffffffu, mbar_wait(&a_empty[sa], pa ^ 1))) { ok = false; break; }
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                tma_load_4d(smem + static_cast<size_t>(sa) * a_bytes, &tmA, &a_full[sa], kc * kChunkK, x0, y0, b);
            }
            __syncwarp();
            if (++sa == g.na) { sa = 0; pa ^= 1; }
        }
    } else if (warp == 2) {
        // B producer: one weight slab per (chunk, tap), shared by both 128-row halves
        const int brow = ny * g.n_tile;
        int sb = 0;
        uint32_t pb = 0;
        for (int kc = 0; kc < g.kchunks && ok; ++kc) {
            for (int tap = 0; tap < g.ntaps; ++tap) {
                if (!__all_sync(0xffffffffu, mbar_wait(&b_empty[sb], pb ^ 1))) { ok = false; break; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(&b_full[sb], b_bytes);
                    tma_load_2d(smem_b + static_cast<size_t>(sb) * b_bytes, &tmB, &b_full[sb], (tap * g.kchunks + kc) * kChunkK, brow);
                }
                __syncwarp();
                if (++sb == g.nb) { sb = 0; pb ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_f16(kTileM, g.n_tile);
        const uint32_t sbo = static_cast<uint32_t>(g.pxp) * 128;
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        bool first = true;
        for (int kc = 0; kc < g.kchunks && ok; ++kc) {
            if (!__all_sync(0xffffffffu, mbar_wait(&a_full[sa], pa))) { ok = false; break; }
            const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(sa) * a_bytes);
            int kx = 0, ky = 0;
            for (int tap = 0; tap < g.ntaps; ++tap) {
                if (!__all_sync(0xffffffffu, mbar_wait(&b_full[sb], pb))) { ok = false; break; }
                tc_fence_after();
                if (elect_one()) {
                    if (tstamp && first) tstamp[2] = clock64();
                    const uint32_t b_addr = smem_u32(smem_b + static_cast<size_t>(sb) * b_bytes);
                    const uint32_t a_tap = a_addr + static_cast<uint32_t>(ky * g.pxp + kx) * 128;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(tmem_base + half * g.n_tile,
                                     umma_desc_k128_ex(a_tap + half * 8 * 128 + k * 32, sbo, g.base_off_mode),
                                     umma_desc_k128(b_addr + k * 32), idesc, (!first || k != 0) ? 1u : 0u);
                    }
                    umma_commit(&b_empty[sb]);
                }
                __syncwarp();
                first = false;
                if (++kx == g.kw) { kx = 0; ++ky; }
                if (++sb == g.nb) { sb = 0; pb ^= 1; }
            }
            if (!ok) break;
            if (elect_one()) umma_commit(&a_empty[sa]);
            __syncwarp();
            if (++sa == g.na) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) {
            if (tstamp) tstamp[3] = clock64();
            umma_commit(accum_ready);
        }
        __syncwarp();
    }
    {
        const int q = warp & 3;
        const int cpar = warp >> 2;
        float* bw = bias_s + warp * 256;
        for (int i = lane; i < 256; i += 32)
            bw[i] = (e.bias != nullptr && i < ((g.n_tile + 31) & ~31)) ? __ldg(e.bias + ny * g.n_tile + i) : 0.0f;
        __syncwarp();
        const bool ok_acc = mbar_wait(accum_ready, 0);
        ok = ok && ok_acc;
        if (tstamp && warp == 2 && lane == 0) tstamp[4] = clock64();
        tc_fence_after();
        if (ok_acc) {
            float* stg = reinterpret_cast<float*>(smem) + warp * 2048;
            const int sub = lane >> 3, cq = lane & 7;
            const int nchunk = (g.n_tile + 31) / 32;
            // rows served by this lane: m = q*32 + k*4 + sub  ->  tile pixel (m >> 3, (m & 7) + 8*half)
            long pixr[8];
            unsigned valid_bits = 0;      // bit k: half 0, bit 8+k: half 1
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int m = q * 32 + k * 4 + sub;
                const int y = ty * 16 + (m >> 3), x = tx * 16 + (m & 7);
                if (y < g.H && b < g.b0 + g.nbatch) {
                    if (x < g.W) valid_bits |= 1u << k;
                    if (x + 8 < g.W) valid_bits |= 1u << (8 + k);
                }
                pixr[k] = (static_cast<long>(b) * g.H + y) * g.W + x;
            }
            int n_units = 0;
            for (int u = cpar; u < 2 * nchunk; u += 2, ++n_units) {
                const int half = u >= nchunk ? 1 : 0;
                const int c = u - half * nchunk;
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * g.n_tile + c * 32), r);
                tmem_ld_wait();
                float* buf = stg + (n_units & 1) * 1024;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(buf + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                const int col = ny * g.n_tile + c * 32 + cq * 4;
                const float4 bb = *reinterpret_cast<const float4*>(bw + c * 32 + cq * 4);
                const unsigned vb = valid_bits >> (8 * half);
                const bool col_ok = col < e.n_valid, full = col + 4 <= e.n_valid;
                const float relu_lo = e.relu ? 0.0f : -INFINITY;
                if (col_ok) {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    float4 v[4];
                    long pix[4];
                    bool val[4];
                    EpiAux ax[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = (h2 * 4 + k) * 4 + sub;
                        v[k] = *reinterpret_cast<const float4*>(buf + rr * 32 + ((cq ^ (rr & 7)) << 2));
                        val[k] = (vb >> (h2 * 4 + k)) & 1u;
                        pix[k] = pixr[h2 * 4 + k] + 8 * half;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (val[k]) ax[k] = epi_prefetch<MODE>(e, col, pix[k]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (val[k]) epilogue4<MODE>(e, v[k], bb, ax[k], col, pix[k], full, relu_lo);
                }
                }
            }
        }
    }
    if (tstamp && warp == 2 && lane == 0) tstamp[5] = clock64();
    if (!ok && e.err_flag != nullptr) atomicExch(e.err_flag, 1 + warp);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(g.tmem_cols));
    if (tstamp && threadIdx.x == 0) tstamp[6] = clock64();
}

// ------------------------------------------------------------------------------------------
// SIMT cross-check kernel (tests only): same geometry, same epilogue, scalar fp32 FMAs.
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128)
conv_simt_kernel(const __half* __restrict__ A, int a_pitch, int a_cin, int in_H, int in_W,
                 const __half* __restrict__ Bw, int ktot, const ConvGeom g, const ConvEpi e) {
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = g.b0 + t / g.tiles_y;
    const int ny = blockIdx.y;
    const int row = threadIdx.x;
    const int yy = row / g.tile_w, xx = row - yy * g.tile_w;
    const int y = ty * g.tile_h + yy, x = tx * g.tile_w + xx;
    if (y >= g.H || x >= g.W) return;
    const long pix = (static_cast<long>(b) * g.H + y) * g.W + x;
    const int brow0 = b * g.b_rows_per_batch + ny * g.n_tile;
    for (int c0 = 0; c0 < g.n_tile; c0 += 32) {
        float acc[32];
        for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
        for (int tap = 0; tap < g.ntaps; ++tap) {
            const int iy = g.stride * y + tap / g.kw - g.kh / 2, ix = g.stride * x + tap % g.kw - g.kw / 2;
            if (iy < 0 || iy >= in_H || ix < 0 || ix >= in_W) continue;
            const __half* arow = A + ((static_cast<long>(b) * in_H + iy) * in_W + ix) * a_pitch;
            const int kmax = min(a_cin, g.kchunks * kChunkK);
            for (int k = 0; k < kmax; ++k) {
                const float a = __half2float(arow[k]);
                const long kk = static_cast<long>(tap) * g.kchunks * kChunkK + k;
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < g.n_tile) acc[j] = fmaf(a, __half2float(Bw[(brow0 + c0 + j) * static_cast<long>(ktot) + kk]), acc[j]);
            }
        }
        for (int j = 0; j < 32; j += 4) {
            const int col = ny * g.n_tile + c0 + j;
            if (col >= e.n_valid) continue;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias != nullptr) bb = *reinterpret_cast<const float4*>(e.bias + col);
            const EpiAux ax = epi_prefetch<MODE>(e, col, pix);
            epilogue4<MODE>(e, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]), bb, ax, col, pix,
                            col + 4 <= e.n_valid, e.relu ? 0.0f : -INFINITY);
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
TapList taps_rect(int kh, int kw) {
    TapList t{};
    t.n = kh * kw;
    t.kh = kh;
    t.kw = kw;
    return t;
}

void choose_tile(int H, int W, int* tile_h, int* tile_w) {
    static const int cand[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
    long best = -1;
    for (auto& c : cand) {
        const long tiles = static_cast<long>((H + c[0] - 1) / c[0]) * ((W + c[1] - 1) / c[1]);
        if (best < 0 || tiles < best) {
            best = tiles;
            *tile_h = c[0];
            *tile_w = c[1];
        }
    }
}

// Variant 2 is correct but was not faster when it was measured -- BEFORE the warp-uniform issue fix, when its TMA / MMA
// loops still issued from a divergent `lane == 0` branch (~67 cycles per MMA): off by default, kept under test.  Its
// role loops are warp-uniform now (parity tests green) but it has NOT been re-timed yet.  With uniform issue the
// variant-1 main loop is bound by the ~70-95 B/clk an SM ingests from L2, which is exactly what the haloed tile and the
// shared weight slab cut (3.8x less ingress for the 64-channel encoder layers); to serve fnet it still needs an epilogue
// that accumulates the instance-norm statistics.
// Measured: the UMMA swizzle is address based, so the base-offset field must stay 0.
static int g_use_v2 = 0, g_v2_base_off = 0;
void conv_set_v2(int on, int base_off_mode) { g_use_v2 = on; g_v2_base_off = base_off_mode; }
static int g_use_pdl = 1;
void conv_set_pdl(int on) { g_use_pdl = on; }
static int g_forced_cluster = 0;
static int g_smem_cap_kib = 0;
void conv_set_forced_cluster(int c) { g_forced_cluster = c; }
void conv_set_smem_cap_kib(int kib) { g_smem_cap_kib = kib; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static const char* encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                          const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return "cuTensorMapEncodeTiled entry point not available";
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                    strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static thread_local char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu %llu, box %u %u)",
                 static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return buf;
    }
    return nullptr;
}

// Un-swizzled fp16 tensor map (zero fill outside the tensor, no L2 promotion): small gather boxes such as the lookup's
// 16 x 10 correlation windows.
const char* encode_tensor_map_plain(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                                    const unsigned long long* strides_bytes, const unsigned* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return "cuTensorMapEncodeTiled entry point not available";
    cuuint64_t d[5], st[5];
    cuuint32_t b[5], es[5];
    if (rank < 1 || rank > 5) return "encode_tensor_map_plain: bad rank";
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; if (i + 1 < rank) st[i] = strides_bytes[i]; }
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, st, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? nullptr : "encode_tensor_map_plain: cuTensorMapEncodeTiled failed";
}

const char* conv_plan_init(ConvPlan* p, const __half* a_base, int a_pitch, int a_cin, int in_H, int in_W, int batch,
                           int stride, const TapList& taps, const __half* wt, int cout_pad, int n_tile,
                           int b_rows_per_batch, int force_tile_h, int force_tile_w) {
    memset(p, 0, sizeof *p);
    if (taps.n < 1 || taps.n > kMaxTaps) return "conv_plan_init: bad tap count";
    if (n_tile % 16 != 0 || n_tile < 16 || n_tile > 256) return "conv_plan_init: bad n_tile";
    if (b_rows_per_batch == 0 && cout_pad % n_tile != 0) return "conv_plan_init: cout_pad must be a multiple of n_tile";
    if (a_pitch % 8 != 0 || (reinterpret_cast<uintptr_t>(a_base) & 15) != 0) return "conv_plan_init: A view not 16B aligned";
    if (stride != 1 && stride != 2) return "conv_plan_init: stride must be 1 or 2";
    ConvGeom& g = p->g;
    g.H = (in_H + stride - 1) / stride;
    g.W = (in_W + stride - 1) / stride;
    g.nbatch = batch;
    if (force_tile_h > 0) {
        g.tile_h = force_tile_h;
        g.tile_w = force_tile_w;
    } else {
        choose_tile(g.H, g.W, &g.tile_h, &g.tile_w);
    }
    if (g.tile_h * g.tile_w != kTileM) return "conv_plan_init: tile must cover 128 pixels";
    g.tile_w_log2 = 0;
    while ((1 << g.tile_w_log2) < g.tile_w) ++g.tile_w_log2;
    if ((1 << g.tile_w_log2) != g.tile_w) return "conv_plan_init: tile width must be a power of two";
    g.tiles_x = (g.W + g.tile_w - 1) / g.tile_w;
    g.tiles_y = (g.H + g.tile_h - 1) / g.tile_h;
    g.stride = stride;
    g.ntaps = taps.n;
    g.kchunks = (a_cin + kChunkK - 1) / kChunkK;
    g.kh = taps.kh;
    g.kw = taps.kw;
    g.n_tile = n_tile;
    g.n_tiles = (cout_pad + n_tile - 1) / n_tile;
    g.b_rows_per_batch = b_rows_per_batch;
    // cluster size: CTAs of a cluster share the B slab (weights are common to every tile; the correlation's
    // B operand is common to the tiles of one pair).  Slices must be whole 8-row swizzle atoms.
    {
        int c = g_forced_cluster > 0 ? g_forced_cluster : 1;
        while (c > 1 && (n_tile % (8 * c) != 0 || (b_rows_per_batch > 0 && (g.tiles_x * g.tiles_y) % c != 0))) c /= 2;
        g.cluster = c;
    }
    const int stage_bytes = kTileM * 128 + n_tile * 128;
    // default cap ~half an SM: two CTAs stay co-resident, so one CTA's prologue / epilogue overlaps the
    // other's main loop (measured +10..25 % on the GRU / motion-encoder layers vs one 200 KiB CTA per SM)
    int stages = ((g_smem_cap_kib > 0 ? g_smem_cap_kib : 104) * 1024) / stage_bytes;
    if (stages < 1) stages = 1;
    if (stages > 6) stages = 6;
    const int T = g.ntaps * g.kchunks;
    if (stages > T) stages = T;
    g.stages = stages;
    int cols = 32;
    while (cols < n_tile) cols *= 2;
    g.tmem_cols = cols;

    p->a_base = a_base; p->a_pitch = a_pitch; p->a_cin = a_cin; p->in_H = in_H; p->in_W = in_W;
    p->b_base = wt; p->ktot = T * kChunkK;

    {   // activations: (C, W, H, B)
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(a_cin), static_cast<cuuint64_t>(in_W),
                              static_cast<cuuint64_t>(in_H), static_cast<cuuint64_t>(batch)};
        cuuint64_t str[3] = {static_cast<cuuint64_t>(a_pitch) * 2, static_cast<cuuint64_t>(in_W) * a_pitch * 2,
                             static_cast<cuuint64_t>(in_H) * in_W * a_pitch * 2};
        cuuint32_t box[4] = {kChunkK, static_cast<cuuint32_t>(g.tile_w * stride), static_cast<cuuint32_t>(g.tile_h * stride), 1};
        cuuint32_t es[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
        if (const char* err = encode(&p->tmA, a_base, 4, dims, str, box, es)) return err;
    }
    {   // weights / B matrix: (K, rows)
        const long rows = b_rows_per_batch > 0 ? static_cast<long>(b_rows_per_batch) * batch : cout_pad;
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(p->ktot), static_cast<cuuint64_t>(rows)};
        cuuint64_t str[1] = {static_cast<cuuint64_t>(p->ktot) * 2};
        cuuint32_t box[2] = {kChunkK, static_cast<cuuint32_t>(n_tile / g.cluster)};
        cuuint32_t es[2] = {1, 1};
        if (const char* err = encode(&p->tmB, wt, 2, dims, str, box, es)) return err;
    }
    // ---- variant 2: stride-1 convolutions with enough tiles to matter -------------------------------------------
    p->variant = 1;
    if (g_use_v2 && stride == 1 && b_rows_per_batch == 0 && g.cluster == 1 && force_tile_h == 0 && taps.kw <= 5 &&
        taps.kh <= 5 && in_H >= 16 && in_W >= 16) {
        ConvGeom& q = p->g2;
        q = g;
        q.tile_h = 16; q.tile_w = 16; q.tile_w_log2 = 4;
        q.tiles_x = (g.W + 15) / 16;
        q.tiles_y = (g.H + 15) / 16;
        q.pxp = taps.kw > 1 ? 24 : 16;
        q.py = 16 + 2 * (taps.kh / 2);
        q.base_off_mode = g_v2_base_off;
        const int a2 = q.pxp * q.py * 128, b2 = n_tile * 128;
        q.na = g.kchunks >= 2 ? 2 : 1;
        int nb = (196 * 1024 - q.na * a2) / b2;
        if (nb > 8) nb = 8;
        if (nb > T) nb = T;
        q.nb = nb;
        int cols2 = 32;
        while (cols2 < 2 * n_tile) cols2 *= 2;
        q.tmem_cols = cols2;
        const long tiles2 = static_cast<long>(q.tiles_x) * q.tiles_y * batch;
        if (nb >= 2 && cols2 <= 512 && tiles2 >= 48) {
            cuuint64_t dims[4] = {static_cast<cuuint64_t>(a_cin), static_cast<cuuint64_t>(in_W),
                                  static_cast<cuuint64_t>(in_H), static_cast<cuuint64_t>(batch)};
            cuuint64_t str[3] = {static_cast<cuuint64_t>(a_pitch) * 2, static_cast<cuuint64_t>(in_W) * a_pitch * 2,
                                 static_cast<cuuint64_t>(in_H) * in_W * a_pitch * 2};
            cuuint32_t box[4] = {kChunkK, static_cast<cuuint32_t>(q.pxp), static_cast<cuuint32_t>(q.py), 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            if (const char* err = encode(&p->tmA2, a_base, 4, dims, str, box, es)) return err;
            p->variant = 2;
        }
    }
    return nullptr;
}

template <int MODE>
static const char* launch_mode(const ConvPlan& p, int nbatch, cudaStream_t stream, int use_simt, int b0) {
    ConvGeom g = p.g;
    g.nbatch = nbatch;
    g.b0 = b0;
    dim3 grid(static_cast<unsigned>(g.tiles_x * g.tiles_y * nbatch), static_cast<unsigned>(g.n_tiles));
    if (use_simt) {
        conv_simt_kernel<MODE><<<grid, 128, 0, stream>>>(p.a_base, p.a_pitch, p.a_cin, p.in_H, p.in_W, p.b_base,
                                                         p.ktot, g, p.e);
    } else if (p.variant == 2) {
        ConvGeom g2 = p.g2;
        g2.nbatch = nbatch;
        g2.b0 = b0;
        size_t pipe = static_cast<size_t>(g2.na) * g2.pxp * g2.py * 128 + static_cast<size_t>(g2.nb) * g2.n_tile * 128;
        if (pipe < 64 * 1024) pipe = 64 * 1024;
        const size_t smem = pipe + 256 + 8 * 256 * sizeof(float) + 1024;
        static bool attr_set2 = false;
        if (!attr_set2) {
            cudaError_t err = cudaFuncSetAttribute(conv2_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (err != cudaSuccess) return cudaGetErrorString(err);
            attr_set2 = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(g2.tiles_x * g2.tiles_y * nbatch), static_cast<unsigned>(g2.n_tiles));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_use_pdl ? 1 : 0;
        cudaError_t lerr = cudaLaunchKernelEx(&cfg, conv2_tc_kernel<MODE>, p.tmA2, p.tmB, g2, p.e);
        if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    } else {
        size_t pipe = static_cast<size_t>(g.stages) * (kTileM * 128 + g.n_tile * 128);
        if (pipe < 72 * 1024) pipe = 72 * 1024;                    // the epilogue stages 8 warps x 2 x 4 KiB (+ 8 KiB channel sums) in the idle pipeline slots
        const size_t smem = pipe + 256 /* barriers + TMEM slot */ + 8 * 256 * sizeof(float) /* bias copies */ + 1024;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (err != cudaSuccess) return cudaGetErrorString(err);
            attr_set = true;
        }
        if (g.cluster > 1) grid.x = (grid.x + g.cluster - 1) / g.cluster * g.cluster;   // phantom CTAs keep the cluster whole
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (g_use_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (g.cluster > 1) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = static_cast<unsigned>(g.cluster);
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        cudaError_t lerr = cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE>, p.tmA, p.tmB, p.tmO, g, p.e);
        if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

// ------------------------------------------------------------------------------------------
// persistent all-pairs correlation GEMM (core/corr.py:53-69): D[n1, n2] = <F1[n1,:], F2[n2,:]> * scale, fp16 out
// ------------------------------------------------------------------------------------------
// conv_tc_kernel runs the volume as (source tile, target slice) CTAs of four K stages each: at 1080p 226 k CTAs whose
// set-up + operand fill + epilogue chain (~10 us, two co-resident) never overlaps (7.9 ms, 0.5 PFLOP/s, 2 TB/s written).
// Here one CTA per SM stays resident and walks work items (pair, 128-pixel source tile, group of kCorrGroup target slices):
//   * the source tile (128 x 256 channels, 64 KiB) is loaded ONCE per item and stays in shared memory for all its slices;
//   * the target slices (256 x 64-channel stages of 32 KiB) stream through a 3-stage ring;
//   * two 256-column TMEM accumulators: the 8 epilogue warps drain slice s (scale, fp16, 32 x 32 bulk tensor stores: the
//     bulk-store branch of tile_epilogue) while the MMA warp accumulates slice s + 1.
// Items are ordered (pair, slice group, tile) with the tile fastest, so the CTAs that run at the same time read the same
// target slices out of L2.  Same K order and the same epilogue code as conv_tc_kernel: bit-identical volume.
constexpr int kCorrThreads = 384;       // warp 0: A producer, warp 2: B producer, warp 1: MMA issuer, warps 4..11: epilogue
constexpr int kCorrBStages = 3;
constexpr int kCorrABytes = 4 * kTileM * 128;              // 4 K chunks of 128 rows x 128 bytes
constexpr int kCorrBBytes = 256 * 128;                     // one stage: 256 target rows x 64 channels
constexpr int kCorrStageOff = kCorrABytes + kCorrBStages * kCorrBBytes;    // epilogue staging: 8 warps x 8 KiB
constexpr int kCorrCtlOff = kCorrStageOff + 8 * 8192;

struct CorrCtl {
    uint64_t a_full, a_empty, b_full[kCorrBStages], b_empty[kCorrBStages], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kCorrThreads, 1)
corr_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const ConvGeom g, const ConvEpi e, const int group) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    CorrCtl* ctl = reinterpret_cast<CorrCtl*>(smem + kCorrCtlOff);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        mbar_init(&ctl->a_full, 1);
        mbar_init(&ctl->a_empty, 1);
        for (int s = 0; s < kCorrBStages; ++s) { mbar_init(&ctl->b_full[s], 1); mbar_init(&ctl->b_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&ctl->acc_full[i], 1); mbar_init(&ctl->acc_empty[i], 8); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(&ctl->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    pdl_launch_dependents();
    pdl_wait();

    const int n_groups = (g.n_tiles + group - 1) / group;
    const int per_pair = n_groups * g.tiles_x;
    const int n_items = g.nbatch * per_pair;
    bool ok = true;
    // item i -> (pair b, slice group sg, source tile tx), tx fastest
    auto decode = [&](int i, int& b, int& sg, int& tx) {
        b = g.b0 + i / per_pair;
        const int r = i - (i / per_pair) * per_pair;
        sg = r / g.tiles_x;
        tx = r - sg * g.tiles_x;
    };
    if (warp == 0) {
        // ---- A producer: the item's source tile, all four K chunks on one barrier --------------------------------------------
        uint32_t n = 0;
        for (int i = blockIdx.x; i < n_items && ok; i += gridDim.x, ++n) {
            int b, sg, tx;
            decode(i, b, sg, tx);
            if (!__all_sync(0xffffffffu, mbar_wait(&ctl->a_empty, (n & 1u) ^ 1u))) { ok = false; break; }
            if (elect_one()) {
                mbar_arrive_expect_tx(&ctl->a_full, kCorrABytes);
                for (int k = 0; k < 4; ++k) tma_load_4d(smem + k * (kTileM * 128), &tmA, &ctl->a_full, k * kChunkK, tx * g.tile_w, 0, b);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ---- B producer: target slices, 64-channel stages through the ring --------------------------------------------------
        uint32_t st = 0, ph = 0;
        for (int i = blockIdx.x; i < n_items && ok; i += gridDim.x) {
            int b, sg, tx;
            decode(i, b, sg, tx);
            const int ny_end = min(g.n_tiles, (sg + 1) * group);
            for (int ny = sg * group; ny < ny_end && ok; ++ny) {
                const int brow = b * g.b_rows_per_batch + ny * g.n_tile;
                for (int k = 0; k < 4; ++k) {
                    if (!__all_sync(0xffffffffu, mbar_wait(&ctl->b_empty[st], ph ^ 1u))) { ok = false; break; }
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&ctl->b_full[st], kCorrBBytes);
                        tma_load_2d(smem + kCorrABytes + st * kCorrBBytes, &tmB, &ctl->b_full[st], k * kChunkK, brow);
                    }
                    __syncwarp();
                    if (++st == kCorrBStages) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----------------------------------------------------------------------------------------------------------
        const uint32_t idesc = umma_idesc_f16(kTileM, 256);
        const uint64_t d0 = umma_desc_k128(smem_u32(smem));
        const uint32_t dhi = static_cast<uint32_t>(d0 >> 32), alo0 = static_cast<uint32_t>(d0);
        const uint32_t blo0 = alo0 + (kCorrABytes >> 4);
        uint32_t st = 0, ph = 0, n = 0, nacc = 0;
        for (int i = blockIdx.x; i < n_items && ok; i += gridDim.x, ++n) {
            int b, sg, tx;
            decode(i, b, sg, tx);
            if (!__all_sync(0xffffffffu, mbar_wait(&ctl->a_full, n & 1u))) { ok = false; break; }
            const int ny_end = min(g.n_tiles, (sg + 1) * group);
            for (int ny = sg * group; ny < ny_end && ok; ++ny, ++nacc) {
                const uint32_t buf = nacc & 1u;
                if (!__all_sync(0xffffffffu, mbar_wait(&ctl->acc_empty[buf], ((nacc >> 1) & 1u) ^ 1u))) { ok = false; break; }
                tc_fence_after();
                for (int k = 0; k < 4; ++k) {
                    if (!__all_sync(0xffffffffu, mbar_wait(&ctl->b_full[st], ph))) { ok = false; break; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t alo = alo0 + static_cast<uint32_t>(k) * ((kTileM * 128) >> 4);
                        const uint32_t blo = blo0 + st * (kCorrBBytes >> 4);
                        const uint32_t acc = tmem_base + buf * 256;
                        umma_f16_lohi(acc, alo, dhi, blo, dhi, idesc, k != 0 ? 1u : 0u);
                        umma_f16_lohi(acc, alo + 2, dhi, blo + 2, dhi, idesc, 1u);
                        umma_f16_lohi(acc, alo + 4, dhi, blo + 4, dhi, idesc, 1u);
                        umma_f16_lohi(acc, alo + 6, dhi, blo + 6, dhi, idesc, 1u);
                        umma_commit(&ctl->b_empty[st]);
                    }
                    __syncwarp();
                    if (++st == kCorrBStages) { st = 0; ph ^= 1u; }
                }
                if (ok && elect_one()) umma_commit(&ctl->acc_full[buf]);
                __syncwarp();
            }
            if (ok && elect_one()) umma_commit(&ctl->a_empty);          // every MMA that reads this source tile has completed
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---- epilogue warps: accumulator -> scale -> fp16 -> bulk tensor stores --------------------------------------------------
        const int ew = warp - 4;
        uint32_t nacc = 0;
        for (int i = blockIdx.x; i < n_items && ok; i += gridDim.x) {
            int b, sg, tx;
            decode(i, b, sg, tx);
            const int ny_end = min(g.n_tiles, (sg + 1) * group);
            for (int ny = sg * group; ny < ny_end; ++ny, ++nacc) {
                const uint32_t buf = nacc & 1u;
                if (!__all_sync(0xffffffffu, mbar_wait(&ctl->acc_full[buf], (nacc >> 1) & 1u))) { ok = false; break; }
                tc_fence_after();
                tile_epilogue<EPI_F32>(g, e, smem + kCorrStageOff, nullptr, nullptr, tmem_base + buf * 256, ew, lane, tx, 0, b, ny, nullptr, &tmO);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
            }
        }
    }
    if (!ok && e.err_flag != nullptr) atomicExch(e.err_flag, 40 + warp);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

const char* corr_gemm_launch(const ConvPlan& p, int nbatch, int b0, cudaStream_t stream) {
    if (p.mode != EPI_F32 || p.variant != 1 || !p.e.tma_store || p.e.out16 == nullptr || p.g.n_tile != 256 || p.g.kchunks != 4 ||
        p.g.ntaps != 1 || p.g.tile_h != 1 || p.g.b_rows_per_batch == 0)
        return "corr_gemm_launch: not a bulk-store correlation plan (256 channels, 256-column slices)";
    ConvGeom g = p.g;
    g.nbatch = nbatch;
    g.b0 = b0;
    const int group = g.n_tiles <= 32 ? 4 : 8;
    const int n_items = nbatch * ((g.n_tiles + group - 1) / group) * g.tiles_x;
    static int n_sm = 0;
    static bool attr_set = false;
    const size_t smem = kCorrCtlOff + sizeof(CorrCtl) + 1024;
    if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t err = cudaFuncSetAttribute(corr_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return cudaGetErrorString(err);
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(n_items < n_sm ? n_items : n_sm));
    cfg.blockDim = dim3(kCorrThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaError_t lerr = cudaLaunchKernelEx(&cfg, corr_gemm_kernel, p.tmA, p.tmB, p.tmO, g, p.e, group);
    if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* conv_plan_enable_tma_store(ConvPlan* p, long rows) {
    p->e.tma_store = 0;
    if (p->mode != EPI_F32 || p->variant != 1) return "tma store: EPI_F32 plans of the 128-pixel kernel only";
    if (p->g.tile_h != 1 || p->g.H != 1) return "tma store: needs one-row tiles (GEMM view)";
    const bool half = p->e.out16 != nullptr;           // fp16 output: 64-byte block rows, 64-byte swizzle
    void* base = half ? static_cast<void*>(p->e.out16) : static_cast<void*>(p->e.out32);
    const long stride = half ? p->e.out16_stride : p->e.out32_stride;
    const int coff = half ? p->e.out16_coff : p->e.out32_coff;
    const int esz = half ? 2 : 4;
    if (base == nullptr || coff != 0 || (stride * esz) % 16 != 0 || p->e.n_valid != stride ||
        (reinterpret_cast<uintptr_t>(base) & 15) != 0)
        return "tma store: output must be a dense, 16-byte aligned row-major matrix";
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return "cuTensorMapEncodeTiled entry point not available";
    // (columns, rows of one batch entry, batch entries): a block never spills into the next batch entry's rows
    const long batches = rows / p->g.W;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(stride), static_cast<cuuint64_t>(p->g.W), static_cast<cuuint64_t>(batches)};
    cuuint64_t str[2] = {static_cast<cuuint64_t>(stride) * esz, static_cast<cuuint64_t>(stride) * esz * p->g.W};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&p->tmO, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, str, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "tma store: cuTensorMapEncodeTiled failed";
    p->e.tma_store = 1;
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// layer programs (host side)
// ------------------------------------------------------------------------------------------
const char* conv_prog_add(ConvProgram* prog, const ConvPlan& p, int dep0, int dep1, bool exact_halo, int ny) {
    if (prog->n_layers >= kMaxProgLayers) return "conv_prog_add: too many layers";
    const ConvGeom& g = p.g;
    if (p.variant != 1 || g.cluster != 1 || ny < 0 || ny >= g.n_tiles) return "conv_prog_add: layer needs the plain 128-pixel kernel / bad cout slice";
    if (p.mode == EPI_CNET || p.e.stats != nullptr || p.e.res16 != nullptr) return "conv_prog_add: unsupported epilogue";
    if (kTileM * 128 + g.n_tile * 128 > static_cast<int>(kProgSlotBytes)) return "conv_prog_add: stage does not fit a ring slot";
    if (prog->n_layers == 0) {
        prog->tiles_x = g.tiles_x;
        prog->tiles_y = g.tiles_y;
    } else if (prog->tiles_x != g.tiles_x || prog->tiles_y != g.tiles_y) {
        return "conv_prog_add: layers of one program must share the tile grid";
    }
    if (dep0 >= prog->n_layers || dep1 >= prog->n_layers) return "conv_prog_add: dependency on a later layer";
    ProgLayer& L = prog->L[prog->n_layers++];
    L.tmA = p.tmA; L.tmB = p.tmB; L.g = p.g; L.e = p.e; L.mode = p.mode; L.dep0 = dep0; L.dep1 = dep1;
    for (int& x : L.succ) x = -1;
    L.ny = ny;
    L.pad_[0] = L.pad_[1] = L.pad_[2] = 0;
    L.n_dep = (dep0 >= 0) + (dep1 >= 0);
    L.kind = 0;
    L.iter_shift = 0;
    // dependency radius in tiles = the window's reach in pixels over the tile size (at least the 3x3 neighbourhood when
    // iterations overlap inside one launch); the scheduler enumerates a 5x5 reach at most
    L.ry = (g.kh / 2 + g.tile_h - 1) / g.tile_h;
    L.rx = (g.kw / 2 + g.tile_w - 1) / g.tile_w;
    if (!exact_halo) { L.ry = L.ry > 1 ? L.ry : 1; L.rx = L.rx > 1 ? L.rx : 1; }
    if (L.ry > 2 || L.rx > 2) { --prog->n_layers; return "conv_prog_add: tile too small for this layer's halo"; }
    L.g.b0 = 0;
    return nullptr;
}

const char* conv_prog_add_lookup(ConvProgram* prog, const ConvPlan& like, const LookupArgs& lk, int dep_prev_iter) {
    if (prog->n_layers >= kMaxProgLayers) return "conv_prog_add_lookup: too many layers";
    const ConvGeom& g = like.g;
    if (prog->n_layers == 0) {
        prog->tiles_x = g.tiles_x;
        prog->tiles_y = g.tiles_y;
    } else if (prog->tiles_x != g.tiles_x || prog->tiles_y != g.tiles_y) {
        return "conv_prog_add_lookup: layers of one program must share the tile grid";
    }
    ProgLayer& L = prog->L[prog->n_layers++];
    memset(&L, 0, sizeof L);
    L.g = like.g;
    L.g.b0 = 0;
    L.mode = EPI_F16;
    L.dep0 = dep_prev_iter; L.dep1 = -1;
    for (int& x : L.succ) x = -1;
    L.n_dep = 1;
    L.kind = 1;
    L.iter_shift = 1;
    L.ry = (3 + g.tile_h - 1) / g.tile_h;      // the 7x7 flow patch reaches 3 pixels into the neighbouring tiles
    L.rx = (3 + g.tile_w - 1) / g.tile_w;
    if (L.ry > 2 || L.rx > 2) return "conv_prog_add_lookup: tile too small for the lookup's halo";
    prog->lk = lk;
    return nullptr;
}

const char* conv_prog_finish(ConvProgram* prog) {
    for (int i = 0; i < prog->n_layers; ++i)
        for (int& x : prog->L[i].succ) x = -1;
    for (int i = 0; i < prog->n_layers; ++i) {
        const ProgLayer& L = prog->L[i];
        for (int d : {L.dep0, L.dep1}) {
            if (d < 0) continue;
            if (d >= prog->n_layers) return "conv_prog_finish: dependency on a missing layer";
            if (L.iter_shift == 0 && d >= i) return "conv_prog_finish: same-iteration dependency on a later layer";
            ProgLayer& D = prog->L[d];
            int k = 0;
            while (k < 4 && D.succ[k] >= 0) ++k;
            if (k == 4) return "conv_prog_finish: a layer can feed at most four others";
            D.succ[k] = i;
        }
    }
    return nullptr;
}

long conv_prog_items(const ConvProgram& prog, int nbatch) {
    return static_cast<long>(prog.n_layers) * nbatch * prog.tiles_x * prog.tiles_y;
}

const char* conv_prog_launch(ConvProgram* prog, int nbatch, int b0, int iters, cudaStream_t stream) {
    static int n_sm = 0;
    static bool attr_set = false;
    const size_t smem = kProgSmemBytes + 1024;        // + alignment slack: 227 KiB in all, the per-CTA maximum
    if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t err = cudaFuncSetAttribute(conv_prog_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(conv_prog_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err != cudaSuccess) return cudaGetErrorString(err);
        attr_set = true;
    }
    if (nbatch < 1 || b0 < 0 || b0 + nbatch > prog->max_batch) return "conv_prog_launch: bad batch range";

    if (iters < 1) return "conv_prog_launch: bad iteration count";
    for (int i = 0; i < prog->n_layers; ++i)
        if (iters > 1 && prog->L[i].n_dep == 0) return "conv_prog_launch: a root layer cannot be iterated";
    const long total = conv_prog_items(*prog, nbatch) * iters;
    long grid = n_sm;
    if (grid > total) grid = total;
    if (total > prog->queue_cap) return "conv_prog_launch: ready queue too small";
    prog->nbatch = nbatch;
    prog->b0 = b0;
    prog->iters = iters;
    if (prog->tickets < 1 || prog->tickets > kProgTickets) prog->tickets = kProgTickets;
    prog->epoch += 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kProgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    bool has_lookup = false;
    for (int i = 0; i < prog->n_layers; ++i) has_lookup = has_lookup || prog->L[i].kind == 1;
    cudaError_t lerr = has_lookup ? cudaLaunchKernelEx(&cfg, conv_prog_kernel<true>, *prog)
                                  : cudaLaunchKernelEx(&cfg, conv_prog_kernel<false>, *prog);
    // every CTA pops until it sees a ticket >= total: head advances by total + grid per launch, tail by total
    if (!prog->static_order) {                     // (the static order touches neither counter)
        prog->head_base += static_cast<unsigned long long>(total + grid);
        prog->tail_base += static_cast<unsigned long long>(total);
    }
    if (lerr != cudaSuccess) return cudaGetErrorString(lerr);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* conv_launch(const ConvPlan& p, int nbatch, cudaStream_t stream, int use_simt, int b0) {
    switch (p.mode) {
        case EPI_F16: return launch_mode<EPI_F16>(p, nbatch, stream, use_simt, b0);
        case EPI_F32: return launch_mode<EPI_F32>(p, nbatch, stream, use_simt, b0);
        case EPI_CNET: return launch_mode<EPI_CNET>(p, nbatch, stream, use_simt, b0);
        case EPI_GRU_ZR: return launch_mode<EPI_GRU_ZR>(p, nbatch, stream, use_simt, b0);
        case EPI_GRU_Q: return launch_mode<EPI_GRU_Q>(p, nbatch, stream, use_simt, b0);
        case EPI_FLOW: return launch_mode<EPI_FLOW>(p, nbatch, stream, use_simt, b0);
    }
    return "conv_launch: unknown epilogue mode";
}

}  // namespace mftb
