// Probe for the TMA lookup (csrc/kernels.cu: lookup_tma_kernel): small 3-D boxes (BW x 10 x 1 fp16) gathered from a
// [rows][H][W] tensor at arbitrary (also negative / overhanging) positions, several lanes of a warp issuing onto one mbarrier.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/tma_box_probe tools/tma_box_probe.cu -lcuda
//   tools/tma_box_probe W H rows lanes iters blocks mode [c0]
//     mode 0  verify: every box against plain loads with zero fill outside the map (box starts at multiples of 8 columns)
//     mode 1  rate:   boxes per microsecond with `lanes` boxes in flight per warp, 4 warps per block
//     mode 2  one box per lane whose innermost coordinate is c0: shows the alignment rule
//   environment: PROBE_BOX0 = box width in elements (default 24)
//
// What it established on a B200 (driver 580, CUDA 12.9), see DESIGN.md section 6b:
//   * the innermost box coordinate must be a multiple of 16 BYTES (fp16: 8 elements).  c0 = 0, 8, 16, 24 work; c0 = 1, 3, 4
//     raise "illegal instruction" at the UTMALDG although cuTensorMapEncodeTiled accepted the map.  The other coordinates
//     are free, negative and overhanging boxes are zero filled, a box may be larger than the tensor (8x8 map, 24x10 box);
//   * ~10 boxes/ns chip-wide (28 cycles per box per SM) with 16 boxes of 480 bytes in flight per warp and 32 warps per SM,
//     7 boxes/ns with 8 in flight; static or dynamic shared memory, descriptor picked per lane: no difference.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mft_b200/csrc/ptx.cuh"

using namespace mftb;

constexpr int kSlot = 512;                 // bytes per box slot (destinations are 128-byte aligned)

struct Args {
    CUtensorMap tm[4];
    const __half* base;
    int W, H, rows, lanes, iters, boxw, c0;
    unsigned long long* out;     // [0] mismatches / time-outs, [1] checksum
};

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ Args P) {
    __shared__ __align__(128) unsigned char boxes[4][16 * kSlot];
    __shared__ __align__(8) uint64_t bar[4];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(&bar[wib], 1); fence_mbar_init(); }
    __syncwarp();
    const unsigned gw = blockIdx.x * 4 + wib;
    const unsigned tx = static_cast<unsigned>(P.boxw) * 20u;
    unsigned long long bad = 0, sum = 0;
    for (int it = 0; it < P.iters; ++it) {
        const unsigned hsh = hash32(gw * 1315423911u + it * 2654435761u + lane * 97u);
        int X0 = (static_cast<int>(hsh % (P.W + 24)) - 16) & ~7;
        const int Y0 = static_cast<int>((hsh >> 10) % (P.H + 12)) - 8, row = static_cast<int>((hsh >> 20) % P.rows);
        if (MODE == 2) X0 = P.c0;
        if (lane == 0) mbar_arrive_expect_tx(&bar[wib], P.lanes * tx);
        __syncwarp();
        if (lane < P.lanes) tma_load_3d(boxes[wib] + (lane & 15) * kSlot, &P.tm[lane & 3], &bar[wib], X0, Y0, row);
        if (!mbar_wait(&bar[wib], it & 1)) { bad += 1000000; break; }
        if (MODE != 1 && lane < P.lanes && lane < 16) {
            const __half* box = reinterpret_cast<const __half*>(boxes[wib] + lane * kSlot);
            for (int r = 0; r < 10; ++r)
                for (int c = 0; c < P.boxw; ++c) {
                    const int gx = X0 + c, gy = Y0 + r;
                    const bool in = gx >= 0 && gx < P.W && gy >= 0 && gy < P.H;
                    const unsigned short want = in ? __half_as_ushort(P.base[(static_cast<long>(row) * P.H + gy) * P.W + gx]) : 0;
                    const unsigned short got = __half_as_ushort(box[r * P.boxw + c]);
                    bad += want != got;
                    sum += got;
                }
        } else {
            sum += boxes[wib][lane * 4];
        }
        __syncwarp();
    }
    atomicAdd(&P.out[0], bad);
    atomicAdd(&P.out[1], sum);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 64, rows = argc > 3 ? atoi(argv[3]) : 28672;
    const int lanes = argc > 4 ? atoi(argv[4]) : 16, iters = argc > 5 ? atoi(argv[5]) : 8, blocks = argc > 6 ? atoi(argv[6]) : 1184;
    const int mode = argc > 7 ? atoi(argv[7]) : 0, c0 = argc > 8 ? atoi(argv[8]) : 0;
    const int boxw = getenv("PROBE_BOX0") ? atoi(getenv("PROBE_BOX0")) : 24;
    if (mode != 1 && lanes > 16) { printf("verify modes: lanes <= 16\n"); return 1; }
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
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(W) * 2, static_cast<cuuint64_t>(W) * H * 2};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(boxw), 10, 1}, es[3] = {1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(&a.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", static_cast<int>(r)); return 1; }
    }
    a.base = d; a.W = W; a.H = H; a.rows = rows; a.lanes = lanes; a.iters = iters; a.boxw = boxw; a.c0 = c0; a.out = out;
    auto launch = [&]() {
        if (mode == 0) probe<0><<<blocks, 128>>>(a);
        else if (mode == 1) probe<1><<<blocks, 128>>>(a);
        else probe<2><<<blocks, 128>>>(a);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();                                                   // warm-up
    cudaMemset(out, 0, 32);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long res[2] = {0, 0};
    cudaMemcpy(res, out, 16, cudaMemcpyDeviceToHost);
    const double nbox = static_cast<double>(blocks) * 4 * iters * lanes;
    printf("map %dx%d rows %d box %dx10 lanes %d iters %d blocks %d mode %d c0 %d: %s, mismatches %llu, %.1f us, %.0f boxes/us, %.1f cycles per box per SM\n",
           W, H, rows, boxw, lanes, iters, blocks, mode, c0, cudaGetErrorString(err), res[0], ms * 1e3, nbox / (ms * 1e3), ms * 1e-3 * 1.965e9 * 148 / nbox);
    return err == cudaSuccess ? 0 : 2;
}
