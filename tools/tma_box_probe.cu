// Probe for the TMA lookup: small 3-D boxes (16 x 10 x 1 fp16) gathered from a [rows][H][W] tensor at arbitrary
// (also negative / overhanging) coordinates, issued by several lanes of a warp onto one mbarrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_box_probe tools/tma_box_probe.cu -lcuda
//   /tmp/tma_box_probe W H rows lanes iters blocks dyn        (dyn = 1: descriptor picked per lane from an array of 4)
// Prints: correctness of the zero fill against plain loads, boxes per microsecond, bytes/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mft_b200/csrc/ptx.cuh"

using namespace mftb;

struct Args {
    CUtensorMap tm[4];
    const __half* base;
    int W, H, rows, lanes, iters, dyn, rank, txbytes;
    unsigned long long* out;     // [0] mismatches, [1] checksum, [2] cycles (max over warps)
};

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Args P) {
    // dynamic shared memory only: with ANY static __shared__ variable in the kernel UTMALDG raises "illegal instruction" here
    extern __shared__ unsigned char dyn_raw[];
    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 127) & ~uintptr_t(127));
    unsigned char (*boxes)[16 * 512] = reinterpret_cast<unsigned char (*)[16 * 512]>(dyn);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn + 4 * 16 * 512);
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(&bar[wib], 1); fence_mbar_init(); }
    __syncwarp();
    const unsigned gw = blockIdx.x * 4 + wib;
    unsigned long long bad = 0, sum = 0;
    const long long t0 = clock64();
    for (int it = 0; it < P.iters; ++it) {
        const unsigned hsh = hash32(gw * 1315423911u + it * 2654435761u + lane * 97u);
        const int X0 = (static_cast<int>(hsh % (P.W + 24)) - 16) & ~7, Y0 = static_cast<int>((hsh >> 10) % (P.H + 12)) - 8;
        const int row = static_cast<int>((hsh >> 20) % P.rows);
        if (lane == 0) mbar_arrive_expect_tx(&bar[wib], P.lanes * static_cast<unsigned>(P.txbytes));
        __syncwarp();
        if (lane < P.lanes) tma_load_3d(boxes[wib] + lane * 512, P.dyn ? &P.tm[lane & 3] : &P.tm[0], &bar[wib], X0, Y0, row);
        if (!mbar_wait(&bar[wib], it & 1)) { bad += 1000000; break; }
        if (lane < P.lanes) {
            const __half* box = reinterpret_cast<const __half*>(boxes[wib] + lane * 512);
            const int bw = P.txbytes / 20;
            for (int r = 0; r < 10; ++r)
                for (int c = 0; c < bw; ++c) {
                    const int gx = X0 + c, gy = Y0 + r;
                    const bool in = gx >= 0 && gx < P.W && gy >= 0 && gy < P.H;
                    const unsigned short want = in ? __half_as_ushort(P.base[(static_cast<long>(row) * P.H + gy) * P.W + gx]) : 0;
                    const unsigned short got = __half_as_ushort(box[r * bw + c]);
                    bad += want != got;
                    sum += got;
                }
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    atomicAdd(&P.out[0], bad);
    atomicAdd(&P.out[1], sum);
    atomicMax(&P.out[2], static_cast<unsigned long long>(t1 - t0));
}

// throughput variant: no verification, D boxes per lane in flight
__global__ void __launch_bounds__(128) rate(const __grid_constant__ Args P) {
    // dynamic shared memory only: with ANY static __shared__ variable in the kernel UTMALDG raises "illegal instruction" here
    extern __shared__ unsigned char dyn_raw[];
    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 127) & ~uintptr_t(127));
    unsigned char (*boxes)[16 * 512] = reinterpret_cast<unsigned char (*)[16 * 512]>(dyn);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn + 4 * 16 * 512);
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(&bar[wib], 1); fence_mbar_init(); }
    __syncwarp();
    const unsigned gw = blockIdx.x * 4 + wib;
    unsigned long long sum = 0;
    const long long t0 = clock64();
    for (int it = 0; it < P.iters; ++it) {
        const unsigned hsh = hash32(gw * 1315423911u + it * 2654435761u + lane * 97u);
        const int X0 = (static_cast<int>(hsh % (P.W + 8)) - 8) & ~7, Y0 = static_cast<int>((hsh >> 10) % (P.H + 4)) - 4;
        const int row = static_cast<int>((hsh >> 20) % P.rows);
        if (lane == 0) mbar_arrive_expect_tx(&bar[wib], P.lanes * static_cast<unsigned>(P.txbytes));
        __syncwarp();
        if (lane < P.lanes) {
            if (P.rank == 3) tma_load_3d(boxes[wib] + (lane & 15) * 512, P.dyn ? &P.tm[lane & 3] : &P.tm[0], &bar[wib], X0, Y0, row);
            else tma_load_2d(boxes[wib] + (lane & 15) * 512, P.dyn ? &P.tm[lane & 3] : &P.tm[0], &bar[wib], X0, row * P.H + (Y0 < 0 ? 0 : Y0));
        }
        if (!mbar_wait(&bar[wib], it & 1)) break;
        sum += boxes[wib][lane * 4];
        __syncwarp();
    }
    const long long t1 = clock64();
    atomicAdd(&P.out[1], sum);
    atomicMax(&P.out[2], static_cast<unsigned long long>(t1 - t0));
}

// minimal: one elected lane of warp 0 loads one box into dynamic shared memory
__global__ void __launch_bounds__(128) minimal(const __grid_constant__ Args P) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) uint64_t bar1;
    if (threadIdx.x == 0) { mbar_init(&bar1, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar1, static_cast<unsigned>(P.txbytes));
            if (P.rank == 3) tma_load_3d(dsm, &P.tm[0], &bar1, 3, 5, 7);
            else tma_load_2d(dsm, &P.tm[0], &bar1, 3, 5);
        }
    }
    __syncthreads();
    const bool ok = mbar_wait(&bar1, 0);
    if (threadIdx.x == 0) { atomicAdd(&P.out[0], ok ? 0ull : 1000000ull); atomicAdd(&P.out[1], static_cast<unsigned long long>(dsm[0])); }
}

// the same with the descriptor (a) as a direct kernel parameter, (b) in global memory
__global__ void __launch_bounds__(256) minimal_direct(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, int use_global, int rank, int txbytes,
                                                      unsigned long long* out) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) uint64_t sbar;
    __shared__ __align__(128) unsigned char sbox[2048];
    uint64_t& bar1 = (rank & 16) ? sbar : *reinterpret_cast<uint64_t*>(dsm + 8192);          // barrier in static or dynamic shared memory
    unsigned char* dstp = (rank & 32) ? sbox : dsm;
    rank &= 15;
    const CUtensorMap* t = use_global ? gtm : &tm;
    if (threadIdx.x == 0) { mbar_init(&bar1, 1); fence_mbar_init(); tma_prefetch_desc(t); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar1, static_cast<unsigned>(txbytes & 0xffff));
            if (rank == 3) tma_load_3d(dstp, t, &bar1, txbytes >> 16, 5, 7);
            else tma_load_2d(dstp, t, &bar1, txbytes >> 16, 5);
        }
    }
    __syncthreads();
    const bool ok = mbar_wait(&bar1, 0);
    if (threadIdx.x == 0) { atomicAdd(&out[0], ok ? 0ull : 1000000ull); atomicAdd(&out[1], static_cast<unsigned long long>(dsm[0])); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 64, rows = argc > 3 ? atoi(argv[3]) : 28672;
    const int lanes = argc > 4 ? atoi(argv[4]) : 16, iters = argc > 5 ? atoi(argv[5]) : 8, blocks = argc > 6 ? atoi(argv[6]) : 1792;
    const int dyn = argc > 7 ? atoi(argv[7]) : 0, mode = argc > 8 ? atoi(argv[8]) : 0;
    const size_t n = static_cast<size_t>(rows) * H * W;
    std::vector<__half> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2half(static_cast<float>((i * 2654435761u >> 20) & 1023) * 0.125f + 1.0f);
    __half* d = nullptr;
    cudaMalloc(&d, n * 2);
    cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
    unsigned long long* out = nullptr;
    cudaMalloc(&out, 32);
    cudaMemset(out, 0, 32);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode entry point\n"); return 1; }
    Args a;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(W) * 2, static_cast<cuuint64_t>(W) * H * 2};
    auto envi = [](const char* k, int d) { return getenv(k) ? atoi(getenv(k)) : d; };
    const int rank = envi("PROBE_RANK", 3), swz = envi("PROBE_SWZ", 0);
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(envi("PROBE_BOX0", 16)), static_cast<cuuint32_t>(envi("PROBE_BOX1", 10)), 1}, es[3] = {1, 1, 1};
    if (rank == 2) { dims[1] = static_cast<cuuint64_t>(H) * rows; }
    const int promo = getenv("PROBE_PROMO") ? atoi(getenv("PROBE_PROMO")) : 0;
    for (int i = 0; i < 4; ++i) {
        CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(&a.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                         static_cast<CUtensorMapSwizzle>(swz), static_cast<CUtensorMapL2promotion>(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", static_cast<int>(r)); return 1; }
    }
    a.base = d; a.W = W; a.H = H; a.rows = rows; a.lanes = lanes; a.iters = iters; a.dyn = dyn; a.out = out; a.rank = rank; a.txbytes = box[0] * box[1] * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (mode == 3 || mode == 4) {
        CUtensorMap* gtm = nullptr;
        cudaMalloc(&gtm, 128);
        cudaMemcpy(gtm, &a.tm[0], 128, cudaMemcpyHostToDevice);
        for (int i = 0; i < 16; ++i) printf("%016llx%c", reinterpret_cast<const unsigned long long*>(&a.tm[0])[i], i % 4 == 3 ? '\n' : ' ');
        const int flags = envi("PROBE_FLAGS", 0), thr = envi("PROBE_THREADS", 256), dyn_bytes = envi("PROBE_DYN", 16384);
        minimal_direct<<<blocks, thr, dyn_bytes>>>(a.tm[0], gtm, mode == 4, rank | flags, a.txbytes | (envi("PROBE_C0", 3) << 16), out);
        printf("c0 %d ", envi("PROBE_C0", 3));
        printf("flags %d (16 = static barrier, 32 = static destination) threads %d dynamic smem %d: ", flags, thr, dyn_bytes);
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaMemcpy(&a.iters, out, 4, cudaMemcpyDeviceToHost);
        printf("minimal_direct (%s) rank %d box %u x %u swz %d: %s, timeouts %d\n", mode == 4 ? "global descriptor" : "param descriptor", rank, box[0], box[1], swz,
               cudaGetErrorString(e2), a.iters);
        return e2 == cudaSuccess ? 0 : 2;
    }
    if (mode == 2) {
        minimal<<<blocks, 128, 16384>>>(a);
        cudaError_t e2 = cudaDeviceSynchronize();
        cudaMemcpy(&a.iters, out, 4, cudaMemcpyDeviceToHost);
        printf("minimal rank %d box %u x %u swz %d: %s, timeouts %d\n", rank, box[0], box[1], swz, cudaGetErrorString(e2), a.iters);
        return e2 == cudaSuccess ? 0 : 2;
    }
    if (mode == 0) probe<<<blocks, 128, 4 * 16 * 512 + 32 + 128>>>(a); else rate<<<blocks, 128, 4 * 16 * 512 + 32 + 128>>>(a);      // warm-up
    cudaMemset(out, 0, 32);
    cudaEventRecord(e0);
    if (mode == 0) probe<<<blocks, 128, 4 * 16 * 512 + 32 + 128>>>(a); else rate<<<blocks, 128, 4 * 16 * 512 + 32 + 128>>>(a);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long res[4] = {0, 0, 0, 0};
    cudaMemcpy(res, out, 24, cudaMemcpyDeviceToHost);
    const double nbox = static_cast<double>(blocks) * 4 * iters * lanes;
    printf("W %d H %d rows %d lanes %d iters %d blocks %d dyn %d mode %s: %s  mismatches %llu  %.1f us  %.1f boxes/us  (%.0f boxes, %.1f cycles per box per SM)\n",
           W, H, rows, lanes, iters, blocks, dyn, mode ? "rate" : "verify", cudaGetErrorString(err), res[0], ms * 1e3, nbox / (ms * 1e3), nbox,
           ms * 1e-3 * 1.965e9 * 148 / nbox);
    return err == cudaSuccess ? 0 : 2;
}
