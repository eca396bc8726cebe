// Fused multi-layer tcgen05 conv chains (sm_100a): activations stay in shared memory between layers.
#pragma once

#include "tcconv.cuh"

namespace vp {

constexpr int FZ_MAX_TERMS = 32;   // tcgen05.mma per M tile of one layer (taps x channel pairs x {1|3} split terms)
constexpr int FZ_MAX_TILES = 8;    // M tiles (128 rows) of one layer per work item
constexpr int FZ_MAX_LAYERS = 4;
constexpr int FZ_NBUF = 4;         // TMEM accumulator buffers (steps in flight)
constexpr int FZ_THREADS = 320;    // warp 0 loader, warp 1 MMA issuer, warps 2-9 two epilogue groups

// One x2-up-sampling conv layer (polyphase form, see tcconv.cu) inside a fused chain.  All row indices are
// relative to the work item: global row at level k = c_k * tile_index + relative row.
struct FzLayer {
    int cin8;      // input planes of 8 channels
    int nout;      // MMA N = 2 * coutp
    int coutp;     // output channels per phase (multiple of 8)
    int n_tiles;   // M tiles per work item
    int n_terms;   // MMAs per tile
    int in_off;    // input buffer: byte offset in dynamic shared memory (layer 0: slot 0)
    int in_rows;   // rows per input plane (= plane pitch in 16-byte units)
    int out_off;   // output buffer byte offset
    int out_rows;  // rows per output plane / valid rows of the fp32 planar buffer
    int out_rp;    // fp32 planar output: row pitch in floats
    int s_lo;      // relative input row of (tile 0, lane 0)
    int c_in;      // input-level rows per tile index
    int out_lo;    // relative row of output-buffer row 0 (even)
    int T_out;     // valid global rows of the output level: rows outside [0, T_out) are conv zero padding
    int out_kind;  // 0: 16-bit planes for the next MMA layer, 1: fp32 planar [c][row] for the head
    int w_off;     // resident weights: byte offset in shared memory
    int w_bytes;
    int bias_soff;     // float offset into the shared bias array
    long long w_goff;  // weights in global memory: element offset of group 0, group stride
    long long w_gs;
    int b_goff, b_gs;  // bias in global memory (floats)
    uint32_t term_a[FZ_MAX_TERMS], term_b[FZ_MAX_TERMS];
    uint8_t dep[FZ_MAX_TILES][2];  // producer steps this tile waits for, as distances back in the step sequence (0: none)
};

// Decoder tail: decoder.convs.3-6 + sigmoid(conv k11) head for the three EQTransformer decoders.
struct FzDecB {
    FzLayer L[FZ_MAX_LAYERS];
    int n_layers, steps_per_item;
    const uint16_t *x;  // [split][group][B][T0][cin0] channel-last 16-bit
    long long x_split, x_gs;
    int T0, cin0, in_lo0, c0, in_slot_bytes;
    int tiles_per_seq, B;
    const uint16_t *w;
    const float *bias;
    float head_w[3][88];  // [group][c * 11 + k]
    float head_b[3];
    int head_in_off, head_rp, head_row0, W, L_out;
    float *y;  // (B, 3, L_out) probabilities
    size_t smem_bytes;
};

struct TcLayer;
// dec: the 7 decoder TcLayers (groups = 3) of one precision set; m: rows of the 375-sample level per work item.
int decb_build(FzDecB &p, const TcLayer *dec, int split, int m, const float (*head_w)[88], const float *head_b);
int decb_launch(const FzDecB &plan, int split, const uint16_t *x, long long x_split, long long x_gs, int B,
                const uint16_t *w_dev, const float *b_dev, float *y, cudaStream_t s);

}  // namespace vp
