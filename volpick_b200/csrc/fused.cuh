// Fused multi-layer tcgen05 conv chains (sm_100a): activations stay in shared memory between layers.
#pragma once

#include <vector>

#include "tcconv.cuh"

namespace vp {

constexpr int FZ_MAX_TERMS = 32;   // tcgen05.mma per M tile of one layer (taps x channel pairs x {1|3} split terms)
constexpr int FZ_MAX_TILES = 8;    // M tiles (128 rows) of one layer per work item
constexpr int FZ_MAX_LAYERS = 4;
constexpr int FZ_NPIPE = 2;        // independent work-item pipelines per CTA (each: loader, issuer, 4 epilogue warps)
constexpr int FZ_NSLOT = 1;        // input buffers per pipeline: the slot is free again once the first layer's MMAs have retired, long
                                   // before the item ends, so ONE slot still lets the loader run a whole item ahead (and leaves the
                                   // shared memory for longer items: 53 instead of 47 rows = 6 instead of 7 items per blinded window)
constexpr int FZ_NBUF = 4;         // TMEM accumulator buffers per pipeline
constexpr int FZ_NCOLS = 64;       // TMEM columns per accumulator (max MMA N of the chain)
// warps: NPIPE loaders, NPIPE issuers, 4 NPIPE layer-epilogue warps, 4 NPIPE head warps
constexpr int FZ_THREADS = 32 * (2 * FZ_NPIPE + 4 * FZ_NPIPE + 4 * FZ_NPIPE);

// One x2-up-sampling conv layer (polyphase form, see tcconv.cu) inside a fused chain.  All row indices are
// relative to the work item: global row at level k = c_k * tile_index + relative row.
struct FzLayer {
    int cin8;      // input planes of 8 channels
    int nout;      // MMA N = 2 * coutp
    int coutp;     // output channels per phase (multiple of 8)
    int n_tiles;   // M tiles per work item
    int n_terms;   // MMAs per tile (without the bias MMA)
    int in_off;    // input buffer: byte offset inside the pipeline's arena (layer 0: slot 0)
    int in_rows;   // rows per input plane
    int in_pitch;  // plane pitch of the input buffer in 16-byte rows (>= in_rows; layer 0: padded for conflict-free cp.async stores)
    int out_off;   // output buffer byte offset inside the pipeline's arena
    int out_rows;  // rows per output plane / valid rows of the fp32 planar buffer
    int out_rp;    // fp32 planar output: row pitch in floats
    int lvl;       // level of the layer's input: its rows are (375-level rows << lvl)
    int a_row0;    // input-buffer row read by (tile 0, lane 0, tap 0)
    int s_lo;      // relative input row of (tile 0, lane 0)
    int c_in;      // input-level rows per tile index
    int out_lo;    // relative row of output-buffer row 0 (even)
    int T_out;     // valid global rows of the output level: rows outside [0, T_out) are conv zero padding
    int out_kind;  // 0: 16-bit planes for the next MMA layer, 1: fp32 planar [c][row] for the head
    int w_off;     // resident weights: byte offset in dynamic shared memory
    uint8_t dep[FZ_MAX_TILES][2];  // producer steps this tile waits for, as distances back in the step sequence (0: none)
};

// Compile-time MMA schedules of the decoder tail (decoder.convs.3-6 in polyphase form): MMA N, taps, 16-channel
// pairs.  decb_build() checks them against the host-built layers.
constexpr int FZ_DEC_NOUT[4] = {64, 32, 32, 16};
constexpr int FZ_DEC_NTAPS[4] = {5, 5, 5, 7};
constexpr int FZ_DEC_NQ[4] = {2, 2, 1, 1};
// f16x3 only: layers whose hi / lo weight splits are stacked along the MMA N (A_hi is then read once for the two
// terms A_hi W_hi and A_hi W_lo: 2 instead of 3 A-tile reads per K step; the MMAs here are bound by the
// shared-memory operand reads, 4 KB of A per MMA against 0.5 - 2 KB of B).  Needs 2 N <= FZ_NCOLS accumulator columns.
constexpr bool FZ_DEC_STACK[4] = {false, true, true, true};

// Decoder tail: decoder.convs.3-6 + sigmoid(conv k11) head for the three EQTransformer decoders.
struct FzDecB {
    FzLayer L[FZ_MAX_LAYERS];
    int n_layers, steps_per_item;
    const uint16_t *x;  // [split][group][B][T0][cin0] channel-last 16-bit
    long long x_split, x_gs;
    int T0, cin0, in_lo0, c0, in_slot_bytes;
    int tiles_per_seq, B;
    int dbg;             // profiling aid (env VP_DECB_DBG): 1 no MMAs, 2 no layer epilogues, 4 no head, 8 no input loads
    int row_off0;        // 375-level row of tile 0 (> 0 when the leading output samples are blinded and need not be computed)
    int pipe_stride;     // bytes of one pipeline's arena (in[2] | X | Y)
    int blob_off;        // resident weight blob: smem byte offset, bytes per group
    int blob_bytes;
    int bias_off;        // fp32 biases [layer][FZ_NCOLS] inside the blob: smem byte offset
    const uint16_t *blob;  // device: [group][blob_bytes / 2]
    float head_w[3][88];  // [group][c * 11 + k]
    float head_b[3];
    int head_in_off, head_rp, head_swz, head_wait_layer, W, L_out;
    float *y;  // (B, 3, L_out) probabilities
    int smem_bytes;
};

struct DecBPlan {
    FzDecB p;
    int split = 0, opt = 0;          // head outputs per thread
    std::vector<uint16_t> blob;      // host copy until uploaded
    uint16_t *d_blob = nullptr;
    bool ready = false;
};

// dec: the 7 decoder TcLayers (groups = 3, host weight blocks still present); m: rows of the 375-sample
// level per work item.
int decb_build(DecBPlan &plan, const TcLayer *dec, int split, int m, const float (*head_w)[88], const float *head_b);
int decb_upload(DecBPlan &plan);
void decb_free(DecBPlan &plan);
// keep_lo / keep_hi: only output samples [keep_lo, keep_hi) of every window have to be computed (annotate's blinding
// discards the rest): tiles that lie entirely outside are skipped and their part of y is left untouched.
int decb_launch(const DecBPlan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, float *y, int keep_lo,
                int keep_hi, cudaStream_t s);

// Decoder tail, second generation (fused_dec2.cu): the 1500- and 3000-sample levels live in tensor memory and are read by
// tcgen05.mma as its A operand; a work item is 128 rows of the 375-sample level, of which 120 carry valid outputs.
struct FzDecB2 {
    const uint16_t *x;  // [split][group][B][375][32] channel-last 16-bit
    long long x_split, x_gs;
    int B, tiles_per_seq, row_off0, row_hi;  // rows [row_off0, row_hi) of the 375-sample level carry kept output samples
    int dbg;  // env VP_DECB_DBG (timing only, results are wrong): 4 no head, 32 the earlier issue order, 16 drain decoder.convs.5 before the next item's first accumulator
    int T0, L_out;
    int in_off, s1_off, head_off, xch_off, blob_off, blob_bytes;  // shared-memory byte offsets
    int w0_off, w1_off, w2_off, w3_off;
    const uint16_t *blob;  // device: [group][blob_bytes / 2]
    float *y;              // (B, 3, L_out) probabilities
    int smem_bytes;
    // fp32 biases per decoder, read through the constant bank (kernel parameter) so that the epilogues on the hand-over chain do
    // not wait for the shared-memory pipe: decoder.convs.3 [64] | .4 [16] | .5 [32] | .6 [16]
    float bias_c[3][128];
    float head_c[3][100];  // head weights [8][12] (11 taps + pad) and the bias at [96], per decoder
};
struct DecB2Plan {
    FzDecB2 p;
    int split = 0;
    std::vector<uint16_t> blob;
    uint16_t *d_blob = nullptr;
    bool ready = false;
};
// dec: the 7 decoder TcLayers (groups = 3, host weight blocks still present); w4 / b4: raw (16, 32, 7) weights and (16) biases of
// decoder.convs.4 per decoder (folded over two samples here)
int decb2_build(DecB2Plan &plan, const TcLayer *dec, const float *const *w4, const float *const *b4, int split, const float (*head_w)[88],
                const float *head_b);
int decb2_upload(DecB2Plan &plan);
void decb2_free(DecB2Plan &plan);
int decb2_launch(const DecB2Plan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, float *y, int keep_lo, int keep_hi,
                 cudaStream_t s);

// Encoder front (fused_enc.cu): encoder.convs.1-3 (Conv1d + ReLU + MaxPool1d(2) each) in one kernel, the 1500- and 750-sample
// levels in tensor memory; a work item is 128 rows of the 375-sample level, of which 94 carry valid outputs.
struct FzEncA {
    const uint16_t *x;  // [split][B][3000][8] channel-last 16-bit (encoder.convs.0 output)
    long long x_split;
    uint16_t *y;        // [split][B][375][32] channel-last 16-bit (encoder.convs.4 input)
    long long y_split;
    int B, tiles_per_seq, T0;
    int in_off, xch_off, blob_off, blob_bytes, w1_off, w2_off, w3_off;  // shared-memory byte offsets
    const uint16_t *blob;  // device
    int smem_bytes;
    float bias_c[64];  // convs.1 [16] | convs.2 [16] | convs.3 [32], read through the constant bank
};
struct EncAPlan {
    FzEncA p;
    int split = 0;
    std::vector<uint16_t> blob;
    uint16_t *d_blob = nullptr;
    bool ready = false;
};
// enc2 / enc3: the TcLayers of encoder.convs.2 / .3 (host weight blocks still present); w1 / b1: raw (16, 8, 9) weights and (16)
// biases of encoder.convs.1 (folded over eight samples here)
int enca_build(EncAPlan &plan, const TcLayer &enc2, const TcLayer &enc3, const float *w1, const float *b1, int split);
int enca_upload(EncAPlan &plan);
void enca_free(EncAPlan &plan);
int enca_launch(const EncAPlan &plan, const uint16_t *x, long long x_split, int B, uint16_t *y, long long y_split, cudaStream_t s);

// res-CNN stack: the 14 convs of res_cnn_stack.members.0-6 in one persistent launch (fused_res.cu).
constexpr int RS_MAX_LAYERS = 14;
struct ResLayerP {
    const uint16_t *x;  // [split][NS][T][64] channel-last 16-bit operand
    uint16_t *y;        // 16-bit output, same layout: relu(v * psc + psh) when affine, else v
    const uint16_t *w;  // tcconv weight blocks [ntaps * 4][split][2][64][8]
    const float *bias, *psc, *psh;  // [64]
    float *res;         // fp32 residual stream [NS][T][64] added to the conv output (nullptr: none)
    int ntaps;          // 3: 'same' conv; 2: one zero on the right
    int affine;         // post-affine + ReLU on the 16-bit output
    int write_res;      // store the sum back to res (in place)
};
struct ResStackP {
    ResLayerP l[RS_MAX_LAYERS];
    int n_layers, NS, T, fmt16;
    long long split16;  // elements between the hi and lo planes of x / y
};
int resstack_launch(const ResStackP &p, int split, cudaStream_t s);

// The same stack with every activation on chip (fused_res2.cu): the 16-bit operand of a tile lives in shared memory through
// all 14 layers (each epilogue overwrites it in place), the fp32 residual stream lives in TMEM (conv2 accumulates onto it).
struct ResStack2P {
    const uint16_t *x;    // relu(bn1_0(x0)) as the 16-bit operand [split][NS][T][64] (written by the last encoder stage)
    const float *xres;    // x0: fp32 residual stream [NS][T][64]
    uint16_t *y;          // stack output as 16-bit operand [split][NS][T][64]
    long long split16;    // elements between the hi and lo planes of x / y
    const uint16_t *w[RS_MAX_LAYERS];  // tcconv weight blocks [ntaps * 4][split][2][64][8] of conv1 / conv2 of block l / 2
    int ntaps[RS_MAX_LAYERS];
    const float *par;     // device [14][3][64]: bias, scale, shift of the epilogue of layer l (see resstack2_params)
    int NS, T, fmt16;
};
// Host: epilogue parameters.  b1 / b2: conv biases [7][64]; n1 / n2: folded pre-activation BatchNorm (scale, shift) [7][64].
//   conv1 of block i: out = relu((acc + b1_i) * n2_i.scale + n2_i.shift)
//   conv2 of block i: acc holds x0 + sum_j conv2_j (biases excluded); out = relu((acc + B_i) * n1_{i+1}.scale + n1_{i+1}.shift),
//                     B_i = sum_{j <= i} b2_j; the last block writes acc + B_6 without the affine
void resstack2_params(const float *const *b1, const float *const *b2, const float *const *n1s, const float *const *n1h,
                      const float *const *n2s, const float *const *n2h, std::vector<float> &par);
int resstack2_launch(const ResStack2P &p, int split, cudaStream_t s);

// Decoder middle: decoder.convs.1 + decoder.convs.2 (fused_deca.cu).
struct FzDecA {
    const uint16_t *x;  // [split][group][B][94][64] channel-last 16-bit (decoder.convs.0 output)
    long long x_split, x_gs;
    uint16_t *y;        // [split][group][B][375][32] channel-last 16-bit (input of the decoder tail)
    long long y_split, y_gs;
    const uint16_t *blob;  // device: per group { convs.1 blocks | convs.2 polyphase blocks | fp32 biases [128 + 64] }
    const float *fixw;     // device: per group [2][32][64] = w[:, :, 4], w[:, :, 3] of decoder.convs.2 (crop correction)
    int blob_off, blob_bytes, w1_off, w2_off, bias_off;  // shared-memory byte offsets
    int B, smem_bytes;
    float bias_c[3][192];  // per decoder: convs.1 [128] | convs.2 [64]; read through the constant bank (the shared-memory pipe is
                           // saturated by the MMAs' operand reads while the epilogues run: a broadcast LDS waits ~170 cycles there)
};

struct DecAPlan {
    FzDecA p;
    int split = 0;
    std::vector<uint16_t> blob;
    std::vector<float> fixw;
    uint16_t *d_blob = nullptr;
    float *d_fixw = nullptr;
    bool ready = false;
};
// dec1: decoder.convs.1 polyphase layer; dec2p: decoder.convs.2 built as a polyphase layer WITHOUT the crop;
// w2: the three raw (32, 64, 5) weight tensors of decoder.convs.2
int deca_build(DecAPlan &plan, const TcLayer &dec1, const TcLayer &dec2p, int split, const float *const *w2);
int deca_upload(DecAPlan &plan);
void deca_free(DecAPlan &plan);
int deca_launch(const DecAPlan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, uint16_t *y, long long y_split,
                long long y_gs, cudaStream_t s);

}  // namespace vp
