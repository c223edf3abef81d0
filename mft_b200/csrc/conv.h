// Implicit-GEMM convolution on tcgen05 tensor cores: shared host/device definitions.
//
// One kernel serves every convolution of the RAFT-OU network and the all-pairs correlation:
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * Wt[cout, tap, cin]
// A   : activations, NHWC fp16 in HBM, viewed through a 4-D TMA tensor map (C, W, H, B); the
//       spatial shift of each tap is a coordinate offset of the TMA box, out-of-image taps are
//       zero-filled by TMA (== the convolution's zero padding), stride 2 = TMA element stride.
// Wt  : weights fp16 [cout][tap][cin padded to 64], K contiguous ("K-major"), 2-D tensor map.
// D   : 128 pixels x n_tile couts, fp32, accumulated in TMEM; fused epilogue per EpiMode.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace mftb {

enum EpiMode : int {
    EPI_F16 = 0,     // bias (+relu) (+residual add, relu) -> fp16 NHWC slice
    EPI_F32 = 1,     // (bias) * scale (+relu) -> fp32
    EPI_CNET = 2,    // cols <128: tanh -> fp32 net ; cols >=128: relu -> fp16 inp   (core/raft.py:146-149)
    EPI_GRU_ZR = 3,  // cols <128: z=sigmoid -> fp32 ; cols >=128: r=sigmoid, r*h -> fp16 (update.py:111-113)
    EPI_GRU_Q = 4,   // q=tanh ; h=(1-z)h+zq -> fp32 master + fp16 copy               (update.py:114)
    EPI_FLOW = 5,    // cols <2: delta_flow -> fp32 ; coords1 += delta                 (core/raft.py:184)
    EPI_NUM_MODES = 6
};

constexpr int kMaxTaps = 9;
constexpr int kTileM = 128;       // output pixels per CTA == TMEM lanes
constexpr int kChunkK = 64;       // fp16 channels per pipeline stage (= one 128-byte swizzle row)

struct ConvGeom {
    int H, W;                     // OUTPUT height / width
    int nbatch;                   // batch entries covered by this launch
    int b0;                       // first batch entry of this launch (sub-batches of the pair batch run as separate launches)
    int tile_h, tile_w;           // tile_h * tile_w == 128, tile_w a power of two
    int tile_w_log2;
    int tiles_x, tiles_y;
    int stride;                   // input step per output pixel (1 or 2)
    int ntaps, kchunks;           // K loop = ntaps * kchunks stages of 64 channels
    int kh, kw;                   // taps form a centred kh x kw window, row-major: tap = ky*kw + kx
    int n_tile, n_tiles;          // couts per CTA (multiple of 16, <= 256), CTAs along cout
    int b_rows_per_batch;         // B-matrix row offset per batch entry (correlation), 0 for weights
    int stages, tmem_cols;
    int cluster;                  // CTAs per cluster sharing (multicasting) the B operand: 1, 2, 4 or 8
    // variant 2 (256-pixel tile, haloed A operand shared by all taps): pitch / rows of the haloed box, stage counts
    int pxp, py, na, nb, base_off_mode;
};

struct ConvEpi {
    int relu;
    float scale;
    int n_valid;                  // couts that exist (columns >= n_valid are padding)
    const float* bias;            // [n_tiles*n_tile] or nullptr
    __half* out16; int out16_stride, out16_coff;
    float* out32;  int out32_stride, out32_coff;
    const __half* res16; int res_stride, res_coff;
    float* h32;                   // GRU hidden state master [pixel][128]
    float* z32;                   // GRU update gate scratch  [pixel][128]
    float* coords1;               // [pixel][2]
    float* delta32;               // [pixel][2]
    double* stats;                // optional [2][n_valid] per-channel sum / sum of squares of the (pre-activation) outputs,
                                  // accumulated with atomics: instance-norm statistics fused into the producing conv
    int tma_store;                // EPI_F32 only: the output goes out as 32 x 32 bulk tensor stores (ConvPlan::tmO)
    int* err_flag;
    long long* timing;            // optional per-CTA phase timestamps [cta][8] (tuning aid), nullptr in product runs
};

// Fully described launch (built once per layer at configure time).
struct ConvPlan {
    CUtensorMap tmA, tmB;
    CUtensorMap tmO;              // output matrix [rows][out32_stride] fp32 for the bulk-store epilogue (conv_plan_enable_tma_store)
    ConvGeom g;
    int variant;                  // 1 = 128-pixel tile, one A box per tap; 2 = 256-pixel tile with haloed A (see conv_tc.cu)
    CUtensorMap tmA2;
    ConvGeom g2;
    ConvEpi e;
    int mode;
    // raw views, used only by the SIMT cross-check kernel in tests
    const __half* a_base; int a_pitch, a_cin, in_H, in_W;
    const __half* b_base; int ktot;
};

struct TapList {
    int n, kh, kw;
};
TapList taps_rect(int kh, int kw);    // kh x kw window centred (odd sizes), row-major (ky, kx)

// Picks (tile_h, tile_w) with tile_h*tile_w == 128 minimising padded work for an H x W map.
void choose_tile(int H, int W, int* tile_h, int* tile_w);

// Fills tensor maps + geometry.  a_base: first channel of the A view; a_pitch: fp16 elements per
// pixel in HBM; a_cin: channels in the view (K beyond it reads as zero); in_H/in_W: INPUT dims.
// wt: [cout_pad][ntaps*kchunks*64] fp16.  Returns nullptr on success or an error string.
const char* conv_plan_init(ConvPlan* p, const __half* a_base, int a_pitch, int a_cin, int in_H, int in_W, int batch,
                           int stride, const TapList& taps, const __half* wt, int cout_pad, int n_tile,
                           int b_rows_per_batch, int force_tile_h, int force_tile_w);

// For an EPI_F32 plan whose epilogue fields are set: store the output with bulk tensor stores (one 32 x 32 fp32 block per
// instruction, clipped by TMA) instead of per-thread st.global.  Needs one-row tiles (GEMM view) and a dense row-major
// output.  Returns nullptr when enabled or the reason why not.
const char* conv_plan_enable_tma_store(ConvPlan* p, long rows);
// The all-pairs correlation plan (bulk-store enabled, 256 channels, 256-column slices) through the persistent kernel: resident
// source tile, streamed target slices, double-buffered accumulators.  Bit-identical to conv_launch of the same plan.
const char* corr_gemm_launch(const ConvPlan& p, int nbatch, int b0, cudaStream_t stream);
// Plain (un-swizzled, zero-filled, no L2 promotion) fp16 tensor map of `rank` dims; strides_bytes has rank - 1 entries.
const char* encode_tensor_map_plain(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                                    const unsigned long long* strides_bytes, const unsigned* box);

// Test / tuning override for the cluster size chosen by conv_plan_init (0 = automatic).
void conv_set_forced_cluster(int c);
// Variant-2 kernel on/off (default on) and its base-offset mode (tuning / bring-up).
void conv_set_v2(int on, int base_off_mode);
// Programmatic dependent launch on/off (default on).
void conv_set_pdl(int on);
// Caps the shared memory a CTA may use for pipeline stages (KiB, 0 = default 200).
void conv_set_smem_cap_kib(int kib);

// ------------------------------------------------------------------------------------------
// Persistent layer program: several dependent convolutions of the SAME spatial geometry executed by ONE launch of
// resident CTAs with tile-level dataflow between layers (conv_prog_kernel in conv_tc.cu).
// ------------------------------------------------------------------------------------------
constexpr int kMaxProgLayers = 16;

struct ProgLayer {
    CUtensorMap tmA, tmB;
    ConvGeom g;
    ConvEpi e;
    int mode;
    int dep0, dep1;               // program layers whose 3x3 tile neighbourhood (same batch entry) must be complete, -1 = none
    int succ[4];                  // layers that list this one as a dependency (filled by conv_prog_finish), -1 = none
    int n_dep;                    // number of valid dependencies (0 = root layer: its tiles are ready at launch)
    int kind;                     // 0 = convolution, 1 = correlation-pyramid lookup tile (no MMA; run by the epilogue warps)
    int iter_shift;               // 1: the dependencies are the PREVIOUS iteration's tiles (iteration 0 is ready at launch)
    int ry, rx;                   // dependency radius in tiles: the predecessor tiles within +-ry / +-rx must be complete
    int ny;                       // which n_tile-wide slice of the layer's output channels this program layer computes
    int pad_[3];
};

// Work distribution is a dataflow ready queue in global memory: a tile is pushed when the last of its predecessor
// tiles completes (per-tile arrival counters), CTAs pop tickets in push order -- no CTA ever holds a tile that is not
// ready to run.
struct ConvProgram {
    ProgLayer L[kMaxProgLayers];
    int n_layers;
    int tickets;                  // tiles a CTA may hold at once (scheduler run-ahead), 1..4; 0 = 4
    int static_order;             // 1: CTA c runs the items c, c + G, c + 2G, ... of the (iteration, layer, pair, tile) order and polls
                                  // each item's arrival counter itself -- no ready queue (no tail atomic, no queue round trip)
    int iters;                    // the whole layer sequence is repeated `iters` times (iterations overlap tile by tile)
    LookupArgs lk;                // operands of the lookup layer, if any
    int nbatch, b0;               // batch entries covered by this launch
    int max_batch;                // counter-array pitch
    int tiles_x, tiles_y;         // common to every layer
    unsigned epoch;               // tags the queue entries of THIS launch
    unsigned long long head_base, tail_base;   // values of *head / *tail when this launch starts
    unsigned long long* head;     // pop tickets (monotonic over launches)
    unsigned long long* tail;     // push tickets
    unsigned long long* queue;    // [queue_cap] entries (epoch << 32 | item), item = (layer * nbatch + batch) * tiles + tile
    int queue_cap;
    int* arrivals;                // [2][layer][max_batch][tiles_y*tiles_x]: set (epoch & 1) counts this launch, the other is cleared
    int* err_flag;
    long long* timing;            // optional [grid][16] per-CTA role timers (tuning aid)
};

// Appends plan `p` as the next layer of `prog` (checks the common geometry).  Returns nullptr or an error string.
// exact_halo: wait only for the predecessor tiles this layer's window actually reads (kh > 1: above / below, kw > 1:
// left / right); otherwise for the whole 3x3 neighbourhood (needed when iterations overlap inside one launch).
const char* conv_prog_add(ConvProgram* prog, const ConvPlan& p, int dep0, int dep1, bool exact_halo, int ny = 0);
// Appends a lookup layer (tile geometry of `like`) that depends on layer `dep` of the PREVIOUS iteration (a later
// layer of the sequence, wired up by conv_prog_finish).
const char* conv_prog_add_lookup(ConvProgram* prog, const ConvPlan& like, const LookupArgs& lk, int dep_prev_iter);
// Resolves the successor lists once all layers are added.
const char* conv_prog_finish(ConvProgram* prog);
// One launch for `iters` repetitions of the whole program; updates prog->epoch / bases for the next launch.
const char* conv_prog_launch(ConvProgram* prog, int nbatch, int b0, int iters, cudaStream_t stream);
// Work items (= tile launches folded into the program) of one iteration with `nbatch` batch entries.
long conv_prog_items(const ConvProgram& prog, int nbatch);

// Launch on `stream`; nbatch <= batch given at init.  use_simt=1 runs the SIMT cross-check kernel.
const char* conv_launch(const ConvPlan& p, int nbatch, cudaStream_t stream, int use_simt, int b0 = 0);

}  // namespace mftb
