// Bandwidth-bound kernels of the MFT hot path (everything that is not an implicit GEMM).
// Host-callable launchers; all pointers are device pointers; all launches go to `stream`; every launcher returns the
// launch's cudaError_t (checked by the engine).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace mftb {

constexpr int kMaxChains = 8;

struct ChainSelectArgs {
    const float* left[kMaxChains];   // template->left results, planar (4,H,W) = fx, fy, occlusion, sigma
    const float* right;              // left->current flows, planar (K,4,H,W)
    float* out;                      // selected template->current result, planar (4,H,W)
    uint8_t* index;                  // selected chain per pixel (H,W); may be nullptr
    int K, H, W;
    float occlusion_threshold;
};
cudaError_t launch_chain_select(const ChainSelectArgs& a, cudaStream_t stream);

// FlowOUTrackingResult.warp_backward (MFT/results.py:116-136): out[c,y,x] = bilinear(img[c], (x,y)+flow[:,y,x]),
// zeros outside, align_corners=True, same defined operation order as chain_select.  add_flow=1 turns it into
// FlowOUTrackingResult.chain (results.py:87-114) for C == 2: out = (p + S) - grid.
cudaError_t launch_warp_backward(const float* flow, const float* img, int C, int H, int W, int add_flow, float* out,
                          cudaStream_t stream);

// Bilinear point queries (MFT/results.py:138-188, MFT/utils/interpolation.py:76-94): out[c, i] =
// bilinear(field[c], points[i]) (+ points[i][c] when add_points, i.e. warp_forward_points).
cudaError_t launch_sample_points(const float* field, int C, int H, int W, const float* points_xy, int N, int add_points,
                          float* out, cudaStream_t stream);

// Forward splat (FlowOUTrackingResult.warp_forward -> interpolation.bilinear_splat, MFT/results.py:190-248,
// MFT/utils/interpolation.py:234-309): every (unmasked) source pixel adds img[y,x,:] * w to the four grid cells around
// (x, y) + flow with the reference's clamped bilinear weights; out = accum / counts where counts > 0, else 0 (or `border`).
// img / out: (H,W,C) float; mask: (H,W) uint8 or nullptr; counts: (H,W) float scratch.  Accumulation uses float atomics
// (order is not defined: results agree with the oracle to rounding, not bit for bit).
cudaError_t launch_warp_forward(const float* flow, const float* img, const uint8_t* mask, int C, int H, int W, int use_border,
                         float border, float* out, float* counts, cudaStream_t stream);

// uint8 BGR HWC frame -> fp16 im2col patches of the encoders' 7x7 stride-2 first conv,
// [ (Hp/2)*(Wp/2) ][152], k = (ky*7+kx)*3 + c (c: R,G,B), values 2*(v/255)-1; the frame is
// replicate-padded to Hp x Wp (pad_left/pad_top) first, the conv itself zero-pads.
cudaError_t launch_frame_patches(const uint8_t* bgr, int H, int W, int Hp, int Wp, int pad_left, int pad_top, __half* patches,
                          cudaStream_t stream);

// Instance norm over raw fp16 conv outputs [B][P][C].
cudaError_t launch_instnorm_stats(const __half* raw, int B, int P, int C, double* sums /*[B][2][C], zeroed here*/,
                           cudaStream_t stream);
// out = act((raw-mean)*rstd); if res != nullptr: out = relu(res + out).  act = relu if relu else identity.
cudaError_t launch_instnorm_apply(const __half* raw, const double* sums, int B, int P, int C, int relu, const __half* res,
                           __half* out, cudaStream_t stream);

struct PairSetup {
    const int* slots;                // device int[2*n_pairs]: (left, right) feature-slot per pair
    const __half* fmap_slots;        // [slot][Npx][256]
    const float* net_slots;          // [slot][Npx][128]
    const __half* inp_slots;         // [slot][Npx][128]
    __half* F1; __half* F2;          // [pair][Npx][256]
    float* h32;                      // [pair][Npx][128]
    __half* X;                       // [pair][Npx][512]  h | inp | motion | r*h
    float* coords1;                  // [pair][Npx][2]
    const float* init_flow;          // optional planar [pair][2][Npx] coarse flow added to the start coordinates (core/raft.py:153-154)
    int n_pairs, h, w;
};
cudaError_t launch_pair_setup(const PairSetup& a, cudaStream_t stream);

// 2x2 average pooling of the correlation volume over the target dims: L0 [rows][h*w] -> L1..L3.
// p1..p3: row pitch (elements) of the pooled levels (>= their width; the pad columns are never written and stay zero).
cudaError_t launch_corr_pool(const __half* L0, __half* L1, __half* L2, __half* L3, long rows, int h, int w, int p1, int p2, int p3,
                             cudaStream_t stream);

struct LookupArgs {
    const __half* lvl[4];            // pyramid levels [pair*Npx + n][h_l][pitch_l], fp16
    int pitch[4];                    // row pitch of each level in elements: w for level 0, (w >> l) rounded up to 8 below (zero pad)
    const float* coords1;            // [pair][Npx][2]
    __half* corr16;                  // [pair*Npx][328]  (324 used)
    __half* flowpatch16;             // [pair*Npx][104]  (98 used): 7x7x2 neighbourhood of the flow
    __half* X;                       // writes flow into X[:, 382:384]
    int n_pairs, h, w;
};
cudaError_t launch_lookup(const LookupArgs& a, cudaStream_t stream);

// The same lookup with the correlation windows fetched by the TMA unit: level l of the pyramid is a 3-D fp16 tensor
// (x: w_l, y: h_l, row: pair * Npx + n) with row pitch LookupArgs::pitch[l]; ONE 24 x 10 box per (pixel, level), starting at
// a multiple of 8 columns (the innermost TMA coordinate must be 16-byte aligned), lands in shared memory; out-of-map taps
// arrive as zeros (= grid_sample's zero padding): no per-element address arithmetic or bounds handling in the SM.  Needs
// every level's row pitch to be a multiple of 8 elements, i.e. a coarse width that is a multiple of 8 (the engine pads the
// pooled levels' rows).  Bit-identical to launch_lookup.
struct LookupTmaArgs {
    CUtensorMap tm[4];
    LookupArgs a;                    // a.lvl is not read; the other pointers start at row pix0
    int pix0;                        // first row of this launch in the tensor maps
    int* err_flag;                   // raised when a window never arrives (bounded wait)
};
cudaError_t launch_lookup_tma(const LookupTmaArgs& a, cudaStream_t stream);

struct OuPackArgs {
    const __half* X; const __half* corr16; const float* coords1; const float* delta32;
    __half* packed;                  // [pair*Npx][720]
    int n_pairs, h, w;
};
cudaError_t launch_ou_pack(const OuPackArgs& a, cudaStream_t stream);

struct UpsampleArgs {
    const float* mask32;             // [pair*Npx][576]
    const float* coords1;            // [pair*Npx][2]
    const float* ou32;               // [pair*Npx][4]: occlusion logit0, logit1, uncertainty, pad
    float* out;                      // planar (pairs,4,H,W): fx, fy, occlusion prob, sigma
    int n_pairs, h, w, H, W, pad_left, pad_top;
};
cudaError_t launch_upsample(const UpsampleArgs& a, cudaStream_t stream);

}  // namespace mftb
