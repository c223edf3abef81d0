/* mft_b200 -- test / tuning hooks of libmft_b200.so.  NOT part of the drop-in boundary (include/mft_b200.h): these
 * entry points exist for the unit tests (tests/test_gpu_kernels.py), the tuning tools (tools/) and stage-level parity
 * checks; a binding of the reference never calls them. */
#ifndef MFT_B200_INTERNAL_H
#define MFT_B200_INTERNAL_H

#include "../../include/mft_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mftb200_set_option test / tuning keys: "conv_impl" 0 = tcgen05 (product), 1 = SIMT cross-check kernel; "persist" 0 = one
 * launch per layer, 1 = one persistent launch per GRU iteration (default), 2 = one launch for all iterations;
 * "split_pairs"; "prog_timing"; "prog_tickets"; "corr_bulk_store"; "cluster" (0 = auto, 1|2|4|8) and "smem_cap_kib" =
 * conv-kernel knobs read by the next mftb200_configure. */

/* Process-wide conv-kernel tuning knobs (read when plans are built): "conv_v2" bit0 = use the 256-pixel haloed
 * kernel, bit1 = descriptor base-offset mode; "pdl"; "cluster"; "smem_cap_kib"; "prog_split_n" 0|1 = the 256-column
 * layers of the iteration program as two 128-column slices (twice the work items, half the per-item latency). */
int mftb200_set_global_option(const char* key, int value);
/* Named internal buffer for stage-level parity tests ("fmap_slots", "net_slots", "corr_l0", ...). */
int mftb200_debug_buffer(mftb200_ctx* ctx, const char* name, void** ptr, size_t* bytes);
/* Copies the first `bytes` bytes of a named internal buffer into caller device memory (syncs). */
int mftb200_debug_read(mftb200_ctx* ctx, const char* name, void* dst_device, size_t bytes);

/* Stand-alone convolution through the product kernel, for unit tests: x fp16 NHWC
 * (B,H,W,pitch) view of `cin` channels, packed weights as in upload_layer, taps = kh x kw
 * centred window, stride 1|2, output fp32 NHWC (B,Ho,Wo,cout_pad) = (acc + bias) [relu]. */
int mftb200_conv2d_test(const uint16_t* x_dev, int B, int H, int W, int pitch, int cin, const uint16_t* w_dev,
                        const float* bias_dev, int cout_pad, int n_tile, int kh, int kw, int stride, int relu,
                        float* out_dev, int impl, mftb200_stream stream);

/* Same launch repeated `reps` times with device timing (tuning aid): cluster / smem_cap_kib < 0 keep the
 * current setting, 0 = automatic; avg_ms receives the mean of launches 2..reps. */
int mftb200_conv2d_bench(const uint16_t* x_dev, int B, int H, int W, int pitch, int cin, const uint16_t* w_dev,
                         const float* bias_dev, int cout_pad, int n_tile, int kh, int kw, int stride, int relu,
                         float* out_dev, int impl, int cluster, int smem_cap_kib, int reps, float* avg_ms,
                         mftb200_stream stream);

/* As above; timing_dev (device int64 [n_ctas][16], may be NULL) receives per-CTA phase timestamps of the LAST
 * launch: clock64 at entry, set-up done, first stage landed, MMAs issued, accumulator ready, epilogue done,
 * exit; [7] = globaltimer ns at entry. */
int mftb200_conv2d_bench2(const uint16_t* x_dev, int B, int H, int W, int pitch, int cin, const uint16_t* w_dev,
                          const float* bias_dev, int cout_pad, int n_tile, int kh, int kw, int stride, int relu,
                          float* out_dev, int impl, int cluster, int smem_cap_kib, int reps, float* avg_ms,
                          long long* timing_dev, mftb200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MFT_B200_INTERNAL_H */
