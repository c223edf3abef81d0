// Micro-benchmarks that bound the conv kernel's main loop on B200 (build: see tools/build_probe.sh).
//   mma   : one thread issues R tcgen05.mma (M=128, N, K=16, fp16) on resident smem, no TMA: cycles per MMA
//           for N in {16..256}, 1 or 2 CTAs per SM, commit every `cpe` MMAs, 1 or 2 issuing warps per CTA.
//   tma   : one thread streams [rows x 128 B] boxes (row pitch `pitch` bytes) through an S-stage ring: cycles per
//           stage and bytes/clk per CTA, 1 or 2 CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mft_b200/csrc/ptx.cuh"

using namespace mftb;

__global__ void __launch_bounds__(128) mma_probe(int N, int reps, int commit_every, int issuers, int distinct,
                                                 long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[4];
    __shared__ uint64_t ring[16];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < 16; ++i) mbar_init(&ring[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 256);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp < issuers && lane == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t da = umma_desc_k128(smem_u32(smem));
        const uint64_t db = umma_desc_k128(smem_u32(smem + 16 * 1024));
        const uint32_t dhi = static_cast<uint32_t>(da >> 32);
        const uint32_t alo = static_cast<uint32_t>(da), blo = static_cast<uint32_t>(db);
        const uint32_t acc = tmem_base + (issuers > 1 ? warp * 128 : 0);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t o = distinct ? ((r & 3) * 2) : 0;
            umma_f16_lohi(acc, alo + o, dhi, blo + o, dhi, idesc, r != 0 ? 1u : 0u);
            if (commit_every > 0 && (r % commit_every) == commit_every - 1) umma_commit(&ring[(warp * 8 + (r / commit_every)) & 15]);
        }
        const long long t1 = clock64();
        umma_commit(&bars[warp]);
        mbar_wait(&bars[warp], 0);
        const long long t2 = clock64();
        out[(blockIdx.x * 2 + warp) * 2] = t1 - t0;
        out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
    }
    if (warp == 3 && lane == 0 && commit_every > 0 && distinct == 2) {
        const int groups = reps / commit_every;
        uint32_t ph[16] = {0};
        for (int gidx = 0; gidx < groups; ++gidx) {
            const int b = gidx & 15;
            mbar_wait(&ring[b], ph[b]);
            ph[b] ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap tm, int rows, int stages, int reps,
                                                 int nrow_tiles, int producers, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_all[4][8];
    const int pw = threadIdx.x >> 5;
    uint64_t* full = full_all[pw];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) mbar_init(&full_all[0][0] + i, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && pw < producers) {
        const uint32_t bytes = rows * 128;
        smem += pw * stages * bytes;
        const long long t0 = clock64();
        int issued = 0, done = 0;
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        while (done < reps) {
            while (issued < reps && issued - done < stages) {
                const int s = issued % stages;
                mbar_arrive_expect_tx(&full[s], bytes);
                const int tile = (blockIdx.x * 37 + issued * 11) % nrow_tiles;
                tma_load_2d(smem + s * bytes, &tm, &full[s], (issued % 6) * 64, tile * rows);
                ++issued;
            }
            const int s = done % stages;
            mbar_wait(&full[s], ph[s]);
            ph[s] ^= 1;
            ++done;
        }
        if (pw == 0) out[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* out;
    cudaMalloc(&out, 1 << 20);
    std::vector<long long> h(4096);
    cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 512;
    printf("== tcgen05.mma M=128 K=16 fp16, %d back-to-back MMAs per issuing thread, %d SMs\n", reps, nsm);
    for (int percta : {1, 2}) {
        for (int issuers : {1}) {
            for (int ce : {0, 1, 2, 4, 8, 16}) {
                for (int N : {16, 128, 256}) {
                    if (issuers == 2 && N > 128) continue;
                    const size_t smem = percta == 1 ? 150 * 1024 : 60 * 1024;
                    const int grid = nsm * percta;
                    cudaMemset(out, 0, 1 << 20);
                    for (int distinct : {1, 2}) {
                    if (distinct == 2 && ce == 0) continue;
                    mma_probe<<<grid, 128, smem>>>(N, reps, ce, issuers, distinct, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h.data(), out, grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
                    double issue = 0, total = 0;
                    int n = 0;
                    for (int i = 0; i < grid * 2; ++i)
                        if (h[2 * i + 1] > 0) { issue += h[2 * i]; total += h[2 * i + 1]; ++n; }
                    printf("ctas/SM=%d issuers=%d waiter=%d commit_every=%2d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA per issuer "
                           "(floor N/2 = %d) -> per SM %.1f cyc/MMA\n",
                           percta, issuers, distinct - 1, ce, N, issue / n / reps, total / n / reps, N / 2,
                           total / n / reps / (percta * issuers));
                    }
                }
            }
        }
    }
    // ---- TMA streaming
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    const int total_rows = 28672;
    for (int pitch_el : {512}) {
        __half* buf;
        cudaMalloc(&buf, static_cast<size_t>(total_rows) * pitch_el * 2);
        cudaMemset(buf, 0, static_cast<size_t>(total_rows) * pitch_el * 2);
        for (int rows : std::vector<int>{}) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {static_cast<cuuint64_t>(pitch_el < 384 ? pitch_el : 384), static_cast<cuuint64_t>(total_rows)};
            cuuint64_t str[1] = {static_cast<cuuint64_t>(pitch_el) * 2};
            cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows)};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            for (int percta : {1, 2}) {
                for (int producers : {1, 2, 4})
                for (int stages : {1, 2, 3}) {
                    if (producers * stages * rows * 128 > (percta == 1 ? 190 : 100) * 1024) continue;
                    const size_t smem = (percta == 1 ? 190 : 100) * 1024;
                    const int grid = nsm * percta;
                    tma_probe<<<grid, 128, smem>>>(tm, rows, stages, 256, total_rows / rows, producers, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double tot = 0;
                    for (int i = 0; i < grid; ++i) tot += h[i];
                    const double cyc = tot / grid / 256 / producers;
                    printf("TMA pitch=%4d B box=%3d rows x128B ctas/SM=%d producers=%d stages=%d: %.0f cyc/box/CTA, %.1f B/clk/CTA, %.1f B/clk/SM, "
                           "chip %.0f B/clk\n",
                           pitch_el * 2, rows, percta, producers, stages, cyc, rows * 128 / cyc, rows * 128 / cyc * percta,
                           rows * 128 / cyc * grid);
                }
            }
        }
        cudaFree(buf);
    }
    return 0;
}
