// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.  See conv.h for the contract.
//
// CTA = 192 threads, one 128-pixel x n_tile output tile:
//   warp 0   : TMA producer  (one elected lane; A box = shifted activation tile, B box = weight slab)
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one elected lane, 4 x K=16 MMAs per stage)
//   warps 2-5: epilogue (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> HBM)
// Pipeline: `stages` smem slots guarded by full/empty mbarriers; accumulator hand-off through a
// third mbarrier signalled by tcgen05.commit.
#include <cstdio>
#include <cstring>

#include "conv.h"
#include "ptx.cuh"

namespace mftb {

// ------------------------------------------------------------------------------------------
// fused epilogue on 32 consecutive accumulator columns of one pixel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int MODE>
__device__ __forceinline__ void epilogue32(const ConvEpi& e, float (&v)[32], int col0, long pix) {
    if (col0 >= e.n_valid) return;
    if (e.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(e.bias + col0 + j);
    }
    if constexpr (MODE == EPI_F16) {
        if (e.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (e.res16 != nullptr) {
            const __half* r = e.res16 + pix * e.res_stride + e.res_coff + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < e.n_valid) v[j] = fmaxf(v[j] + __half2float(r[j]), 0.0f);
        }
        __half* o = e.out16 + pix * e.out16_stride + e.out16_coff + col0;
        if (col0 + 32 <= e.n_valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                __align__(16) __half2 h[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                *reinterpret_cast<uint4*>(o + j) = *reinterpret_cast<const uint4*>(h);
            }
        } else {
            for (int j = 0; j < 32 && col0 + j < e.n_valid; ++j) o[j] = __float2half_rn(v[j]);
        }
    } else if constexpr (MODE == EPI_F32) {
        float* o = e.out32 + pix * e.out32_stride + e.out32_coff + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = v[j] * e.scale;
            if (e.relu) x = fmaxf(x, 0.0f);
            v[j] = x;
        }
        if (col0 + 32 <= e.n_valid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
            for (int j = 0; j < 32 && col0 + j < e.n_valid; ++j) o[j] = v[j];
        }
    } else if constexpr (MODE == EPI_CNET) {
        if (col0 < 128) {
            float* o = e.out32 + pix * 128 + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(tanhf(v[j]), tanhf(v[j + 1]), tanhf(v[j + 2]), tanhf(v[j + 3]));
        } else {
            __half* o = e.out16 + pix * 128 + (col0 - 128);
#pragma unroll
            for (int j = 0; j < 32; j += 2)
                *reinterpret_cast<__half2*>(o + j) = __floats2half2_rn(fmaxf(v[j], 0.0f), fmaxf(v[j + 1], 0.0f));
        }
    } else if constexpr (MODE == EPI_GRU_ZR) {
        if (col0 < 128) {
            float* o = e.z32 + pix * 128 + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) =
                    make_float4(sigmoidf_(v[j]), sigmoidf_(v[j + 1]), sigmoidf_(v[j + 2]), sigmoidf_(v[j + 3]));
        } else {
            const float* h = e.h32 + pix * 128 + (col0 - 128);
            __half* o = e.out16 + pix * e.out16_stride + e.out16_coff + (col0 - 128);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 hv = *reinterpret_cast<const float4*>(h + j);
                *reinterpret_cast<__half2*>(o + j) = __floats2half2_rn(sigmoidf_(v[j]) * hv.x, sigmoidf_(v[j + 1]) * hv.y);
                *reinterpret_cast<__half2*>(o + j + 2) =
                    __floats2half2_rn(sigmoidf_(v[j + 2]) * hv.z, sigmoidf_(v[j + 3]) * hv.w);
            }
        }
    } else if constexpr (MODE == EPI_GRU_Q) {
        float* h = e.h32 + pix * 128 + col0;
        const float* z = e.z32 + pix * 128 + col0;
        __half* o = e.out16 + pix * e.out16_stride + e.out16_coff + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h + j);
            const float4 zv = *reinterpret_cast<const float4*>(z + j);
            float4 n;
            n.x = (1.0f - zv.x) * hv.x + zv.x * tanhf(v[j]);
            n.y = (1.0f - zv.y) * hv.y + zv.y * tanhf(v[j + 1]);
            n.z = (1.0f - zv.z) * hv.z + zv.z * tanhf(v[j + 2]);
            n.w = (1.0f - zv.w) * hv.w + zv.w * tanhf(v[j + 3]);
            *reinterpret_cast<float4*>(h + j) = n;
            *reinterpret_cast<__half2*>(o + j) = __floats2half2_rn(n.x, n.y);
            *reinterpret_cast<__half2*>(o + j + 2) = __floats2half2_rn(n.z, n.w);
        }
    } else if constexpr (MODE == EPI_FLOW) {
        if (col0 == 0) {
            float2 c = *reinterpret_cast<float2*>(e.coords1 + pix * 2);
            *reinterpret_cast<float2*>(e.delta32 + pix * 2) = make_float2(v[0], v[1]);
            c.x += v[0];
            c.y += v[1];
            *reinterpret_cast<float2*>(e.coords1 + pix * 2) = c;
        }
    }
}

// ------------------------------------------------------------------------------------------
// tensor-core kernel
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 192;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGeom g,
               const ConvEpi e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = kTileM * 128;
    const uint32_t b_bytes = static_cast<uint32_t>(g.n_tile) * 128;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(g.stages) * stage_bytes);
    uint64_t* empty = full + g.stages;
    uint64_t* accum_ready = empty + g.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_ready + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = t / g.tiles_y;
    const int ny = blockIdx.y;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(&full[s], 1);
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
    const int T = g.ntaps * g.kchunks;
    const uint32_t crank = g.cluster > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = static_cast<uint16_t>((1u << g.cluster) - 1u);
    bool ok = true;

    if (warp == 0) {
        if (lane == 0) {
            const int x0 = g.stride * tx * g.tile_w;
            const int y0 = g.stride * ty * g.tile_h;
            const int brow = b * g.b_rows_per_batch + ny * g.n_tile;
            int tap = 0, kc = 0;
            for (int it = 0; it < T; ++it) {
                const int s = it % g.stages;
                const uint32_t ph = (it / g.stages) & 1;
                if (!mbar_wait(&empty[s], ph ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
                tma_load_4d(sa, &tmA, &full[s], kc * kChunkK, x0 + g.dx[tap], y0 + g.dy[tap], b);
                if (g.cluster > 1) {
                    // each CTA fetches 1/cluster of the B slab and multicasts it to all CTAs of the cluster
                    const int slice = g.n_tile / g.cluster;
                    tma_load_2d_mc(sa + a_bytes + crank * slice * 128, &tmB, &full[s], it * kChunkK,
                                   brow + static_cast<int>(crank) * slice, cmask);
                } else {
                    tma_load_2d(sa + a_bytes, &tmB, &full[s], it * kChunkK, brow);
                }
                if (++kc == g.kchunks) { kc = 0; ++tap; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(kTileM, g.n_tile);
            for (int it = 0; it < T; ++it) {
                const int s = it % g.stages;
                const uint32_t ph = (it / g.stages) & 1;
                if (!mbar_wait(&full[s], ph)) { ok = false; break; }
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32), idesc,
                             (it | k) != 0 ? 1u : 0u);
                // slot reusable once these MMAs have read it (in every CTA that multicasts into it)
                if (g.cluster > 1) umma_commit_mc(&empty[s], cmask); else umma_commit(&empty[s]);
            }
            umma_commit(accum_ready);     // accumulator complete
        }
    } else {
        const int q = warp & 3;           // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int yy = row / g.tile_w, xx = row - yy * g.tile_w;
        const int y = ty * g.tile_h + yy, x = tx * g.tile_w + xx;
        const bool valid = (y < g.H) && (x < g.W) && (b < g.nbatch);
        const long pix = (static_cast<long>(b) * g.H + y) * g.W + x;
        ok = mbar_wait(accum_ready, 0);
        tc_fence_after();
        if (ok) {
            const int nchunk = (g.n_tile + 31) / 32;
            for (int c = 0; c < nchunk; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
                tmem_ld_wait();
                if (valid) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    epilogue32<MODE>(e, v, ny * g.n_tile + c * 32, pix);
                }
            }
        }
    }
    if (!ok && e.err_flag != nullptr) atomicExch(e.err_flag, 1 + warp);
    tc_fence_before();
    // no CTA may exit while a peer can still multicast into its smem / arrive on its barriers
    if (g.cluster > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(g.tmem_cols));
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
    const int b = t / g.tiles_y;
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
            const int iy = g.stride * y + g.dy[tap], ix = g.stride * x + g.dx[tap];
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
        epilogue32<MODE>(e, acc, ny * g.n_tile + c0, pix);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
TapList taps_rect(int kh, int kw) {
    TapList t{};
    t.n = 0;
    for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
            t.dy[t.n] = static_cast<int8_t>(ky - kh / 2);
            t.dx[t.n] = static_cast<int8_t>(kx - kw / 2);
            ++t.n;
        }
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
    g.tiles_x = (g.W + g.tile_w - 1) / g.tile_w;
    g.tiles_y = (g.H + g.tile_h - 1) / g.tile_h;
    g.stride = stride;
    g.ntaps = taps.n;
    g.kchunks = (a_cin + kChunkK - 1) / kChunkK;
    for (int i = 0; i < taps.n; ++i) {
        g.dy[i] = taps.dy[i];
        g.dx[i] = taps.dx[i];
    }
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
    return nullptr;
}

template <int MODE>
static const char* launch_mode(const ConvPlan& p, int nbatch, cudaStream_t stream, int use_simt) {
    ConvGeom g = p.g;
    g.nbatch = nbatch;
    dim3 grid(static_cast<unsigned>(g.tiles_x * g.tiles_y * nbatch), static_cast<unsigned>(g.n_tiles));
    if (use_simt) {
        conv_simt_kernel<MODE><<<grid, 128, 0, stream>>>(p.a_base, p.a_pitch, p.a_cin, p.in_H, p.in_W, p.b_base,
                                                         p.ktot, g, p.e);
    } else {
        const size_t smem = static_cast<size_t>(g.stages) * (kTileM * 128 + g.n_tile * 128) + (2 * g.stages + 1) * 8 + 16 + 1024;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (err != cudaSuccess) return cudaGetErrorString(err);
            attr_set = true;
        }
        if (g.cluster > 1) {
            grid.x = (grid.x + g.cluster - 1) / g.cluster * g.cluster;   // phantom CTAs keep the cluster whole
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = static_cast<unsigned>(g.cluster);
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaError_t err = cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE>, p.tmA, p.tmB, g, p.e);
            if (err != cudaSuccess) return cudaGetErrorString(err);
        } else {
            conv_tc_kernel<MODE><<<grid, kThreads, smem, stream>>>(p.tmA, p.tmB, g, p.e);
        }
    }
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* conv_launch(const ConvPlan& p, int nbatch, cudaStream_t stream, int use_simt) {
    switch (p.mode) {
        case EPI_F16: return launch_mode<EPI_F16>(p, nbatch, stream, use_simt);
        case EPI_F32: return launch_mode<EPI_F32>(p, nbatch, stream, use_simt);
        case EPI_CNET: return launch_mode<EPI_CNET>(p, nbatch, stream, use_simt);
        case EPI_GRU_ZR: return launch_mode<EPI_GRU_ZR>(p, nbatch, stream, use_simt);
        case EPI_GRU_Q: return launch_mode<EPI_GRU_Q>(p, nbatch, stream, use_simt);
        case EPI_FLOW: return launch_mode<EPI_FLOW>(p, nbatch, stream, use_simt);
    }
    return "conv_launch: unknown epilogue mode";
}

}  // namespace mftb
