/* mft_b200 -- C ABI of the B200-native MFT hot path (libmft_b200.so).
 *
 * This is the drop-in boundary: every entry point takes plain pointers and sizes (no torch
 * types), launches hand-written sm_100a kernels on the caller's CUDA stream and returns 0 on
 * success or a negative code (text via mftb200_last_error).  The reference is pure Python;
 * each function cites the reference interface it stands in for (paths relative to the
 * serycjon/MFT checkout).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Layouts (all device memory unless stated):
 *   frame        uint8  (H, W, 3)   BGR, host or device          -- what MFT.track() receives
 *   flow field   float  (4, H, W)   planar: flow_x, flow_y, occlusion in [0,1], sigma >= 0
 *                                    == FlowOUTrackingResult.{flow, occlusion, sigma} stacked
 *   chain index  uint8  (H, W)      position in the [inf, ascending delta] candidate order
 */
#ifndef MFT_B200_H
#define MFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mftb200_ctx mftb200_ctx;
typedef void* mftb200_stream;          /* cudaStream_t; NULL = default stream */

#define MFTB200_OK 0
#define MFTB200_ERR_ARG (-1)
#define MFTB200_ERR_CUDA (-2)
#define MFTB200_ERR_STATE (-3)
#define MFTB200_ERR_DEVICE_FLAG (-4)   /* a kernel reported a pipeline time-out */

#define MFTB200_NUM_LAYERS 47          /* packed conv layers, order documented in mft_b200/weights.py */
#define MFTB200_MAX_PAIRS 8

/* ---- lifetime ------------------------------------------------------------------------ */
/* Replaces RAFTWrapper.__init__ (MFT/raft.py:16-28): binds the calling thread's current CUDA
 * device; fails unless it is compute capability 10.x. */
int mftb200_create(mftb200_ctx** out);
void mftb200_destroy(mftb200_ctx* ctx);
const char* mftb200_last_error(const mftb200_ctx* ctx);   /* ctx may be NULL: creation errors */
const char* mftb200_version(void);

/* ---- weights (checkpoint consumed at MFT/raft.py:20-21) -------------------------------- */
/* One packed layer: w_f16 = host fp16 [cout_pad][ktot] (K contiguous, K = taps x cin padded to
 * 64), bias = host fp32 [bias_len], bias_len a multiple of 32 >= cout_pad. */
int mftb200_upload_layer(mftb200_ctx* ctx, int layer, const uint16_t* w_f16, const float* bias, int cout_pad,
                         int ktot, int bias_len);

/* ---- geometry --------------------------------------------------------------------------- */
/* Allocates workspace + builds TMA descriptors for H x W frames (any size >= 128; padded to a
 * multiple of 8 like InputPadder, MFT/RAFT/core/utils/utils.py:7-24), up to max_pairs (<= 8)
 * frame pairs per refine call, n_slots cached per-frame feature sets, `iters` GRU iterations
 * (flow_config.flow_iters).  All layers must be uploaded first. */
int mftb200_configure(mftb200_ctx* ctx, int H, int W, int max_pairs, int n_slots, int iters);

/* ---- per-frame encoders (RAFT.forward part 1: MFT/RAFT/core/raft.py:122-149) ------------ */
/* fnet + cnet of one frame, once, into feature slot `slot`.  bgr: (H,W,3) uint8; on_device=0
 * means a host pointer (copied with cudaMemcpyAsync on `stream`; pin it for overlap). */
int mftb200_encode_frame(mftb200_ctx* ctx, const uint8_t* bgr, int on_device, int slot, mftb200_stream stream);
/* 1 if `p` points into page-locked (pinned / registered) host memory, i.e. the frame copy of encode_frame is a true
 * asynchronous DMA and a binding need not stage the frame itself (the reference's `.cuda()` at MFT/raft.py:45 stages
 * pageable frames inside the driver); 0 for pageable or device memory. */
int mftb200_is_pinned_host(const void* p);

/* The feature-slot arrays themselves, for multi-GPU feature exchange (rank t % G encodes frame t, one
 * all-gather hands every rank the features): fmap = fp16 [n_slots][h*w][256], net = fp32 [n_slots][h*w][128] (tanh half
 * of cnet), inp = fp16 [n_slots][h*w][128] (relu half), h x w = padded H/8 x W/8; slot_bytes[3] = bytes of one slot in
 * each array.  Work enqueued on `stream` after this call sees every encode_frame issued before it complete (incl. the
 * context encoder that trails on the engine's own stream).  The reference has no counterpart (it re-runs the encoders
 * per pair, MFT/RAFT/core/raft.py:130-149). */
int mftb200_slot_buffers(mftb200_ctx* ctx, void** fmap, void** net, void** inp, size_t* slot_bytes, mftb200_stream stream);

/* ---- batched RAFT refinement (RAFT.forward part 2 + RAFTWrapper.compute_flow post-processing:
 * MFT/RAFT/core/raft.py:141-259, MFT/raft.py:56-62) ---------------------------------------- */
/* For each pair p: flow left_slots[p] -> right_slots[p].  out: device float (n_pairs,4,H,W). */
int mftb200_raft_refine(mftb200_ctx* ctx, int n_pairs, const int* left_slots, const int* right_slots, float* out,
                        mftb200_stream stream);
/* Same with a flow initialisation (RAFT.forward's flow_init, MFT/RAFT/core/raft.py:153-154; compute_flow's init_flow after
 * padding and downsample_flow_8, MFT/raft.py:49-53): init_flow = device float (n_pairs,2,h,w) at the coarse resolution
 * (h x w = padded H/8 x W/8), added to the start coordinates; NULL = none. */
int mftb200_raft_refine_init(mftb200_ctx* ctx, int n_pairs, const int* left_slots, const int* right_slots,
                             const float* init_flow, float* out, mftb200_stream stream);

/* ---- chaining + selection (chain_results + selection block + invalid mask:
 * MFT/MFT.py:114-142,233-239; MFT/results.py:87-136,250-265) ------------------------------ */
/* left[k]: device (4,H,W) template->left_k result; right: device (K,4,H,W) left_k->current
 * flows; out: device (4,H,W); index: device (H,W) uint8 or NULL.  Candidates must be ordered
 * [inf, ascending delta] (MFT.py:114).  Context-free: usable without create/configure. */
int mftb200_chain_select(int K, const float* const* left, const float* right, float occlusion_threshold, int H,
                         int W, float* out, uint8_t* index, mftb200_stream stream);

/* ---- FlowOUTrackingResult geometry (MFT/results.py:87-188) ------------------------------------ */
/* warp_backward: out (C,H,W) = bilinear sample of img (C,H,W) at grid + flow (2,H,W), zeros outside,
 * align_corners=True.  add_flow=1 with C == 2 gives FlowOUTrackingResult.chain (results.py:87-114). */
int mftb200_warp_backward(const float* flow, const float* img, int C, int H, int W, int add_flow, float* out,
                          mftb200_stream stream);
/* Point queries on a device field: out (C,N) = bilinear(field (C,H,W), points (N,2) xy).  add_points=1 adds
 * the point itself to channels 0,1 == warp_forward_points (results.py:138-157); add_points=0 == sample()
 * / interpolation.bilinear_sample (results.py:159-188, MFT/utils/interpolation.py:76-94). */
int mftb200_sample_points(const float* field, int C, int H, int W, const float* points_xy, int N, int add_points,
                          float* out, mftb200_stream stream);

/* Forward splat == FlowOUTrackingResult.warp_forward -> interpolation.bilinear_splat (MFT/results.py:190-248,
 * MFT/utils/interpolation.py:234-309; demo.py:139-144 propagates an edit with it): img (H,W,C) float is splatted to
 * grid + flow with the reference's clamped bilinear weights and normalised by the accumulated weight; cells that
 * receive nothing stay 0, or `border` when use_border != 0.  mask: (H,W) uint8 (0 = skip the source pixel) or NULL.
 * out: (H,W,C); counts: (H,W) float scratch (returns the accumulated weights).  Float atomics: equal to the
 * reference up to summation order. */
int mftb200_warp_forward(const float* flow, const float* img, const uint8_t* mask, int C, int H, int W, int use_border,
                         float border, float* out, float* counts, mftb200_stream stream);

/* ---- error state / diagnostics --------------------------------------------------------------------- */
/* Every pipeline wait inside the kernels is time bounded (2 s); a kernel that gives up raises a device flag instead of
 * hanging the GPU.  The reference has no counterpart (its kernels are ATen's): the flag surfaces as the exception a
 * CUDA error would raise under PyTorch (SURVEY 8b "Errors").
 *   mftb200_device_error_flag   synchronises the device and reads the flag: 0 = clean.
 *   mftb200_error_flag_async    enqueues a copy of the flag into a pinned mirror on `stream` (no synchronisation): call it
 *                               behind a frame's work, next to the result's device->host copy;
 *   mftb200_error_flag_poll     reads the mirror (valid once `stream` has been synchronised past the copy): 0 = clean.
 * After a non-zero flag the context refuses encode_frame / raft_refine (MFTB200_ERR_DEVICE_FLAG) until
 * mftb200_configure is called again: an aborted launch leaves the work queues of the persistent kernels undefined. */
int mftb200_device_error_flag(mftb200_ctx* ctx);
int mftb200_error_flag_async(mftb200_ctx* ctx, mftb200_stream stream);
int mftb200_error_flag_poll(mftb200_ctx* ctx);
/* Blocks the host until the newest encode_frame's host->device frame copy has left the caller's buffer (which may then
 * be refilled; the reference copies synchronously at MFT/raft.py:45). */
int mftb200_wait_frame_copied(mftb200_ctx* ctx);
/* Product options: "iters" = GRU iterations (flow_config.flow_iters); "defer_context" 0|1; "profile" 0|1 = per-launch
 * event timing (see mftb200_profile_fetch).  Test / tuning keys are listed in mft_b200/csrc/mft_b200_internal.h. */
int mftb200_set_option(mftb200_ctx* ctx, const char* key, int value);
/* Number of kernels launched by this context since creation (bench's gpu_launches). */
long long mftb200_launch_count(const mftb200_ctx* ctx);

/* With option "profile"=1 every launch step is bracketed by CUDA events on its stream.  Fetch (and
 * clear) the accumulated device time: index 0 = tensor-core conv/GEMM launches, 1 = the
 * bandwidth-bound kernels.  Used by bench.py for the live roofline numbers. */
int mftb200_profile_fetch(mftb200_ctx* ctx, double* ms_by_kind /*[2]*/, long long* steps_by_kind /*[2]*/);
/* Per-step event times in launch order (does not clear; call before mftb200_profile_fetch). */
int mftb200_profile_steps(mftb200_ctx* ctx, float* ms, int* kinds, int max_steps, int* n_steps);

#ifdef __cplusplus
}
#endif
#endif /* MFT_B200_H */
