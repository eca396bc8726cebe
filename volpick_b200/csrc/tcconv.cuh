// tcgen05 / TMEM implicit-GEMM Conv1d layer (sm_100a) -- shared declarations.
#pragma once

#include <vector>

#include "common.cuh"

namespace vp {

constexpr int TC_MAX_MMA = 64;
constexpr int TC_MAX_TERMS = 3 * TC_MAX_MMA;

struct TcMma {
    int a_row;    // first staged row of the A operand for this MMA (tap offset)
    int a_plane;  // first 8-channel plane of the A operand
    int a_rowk;   // 1: the two K halves are consecutive ROWS (taps-in-K, 8-channel inputs); 0: consecutive planes
    int b_block;  // weight block index
};

// Kernel parameters of one tensor-core conv layer launch (see tcconv.cu for the data layouts).
struct TcP {
    const uint16_t *x;  // channel-last 16-bit activations [SPLIT][G][NS][x_pitch][CIN_A]; row (seq, u) lives at seq * x_pitch + x_roff + u
    int64_t x_split, x_gs;
    int x_pitch, x_roff;
    // optional second source (channel concatenation without a copy: PhaseNet skip | up-sampled): planes cinA8.. of the
    // staged tile come from x2, [SPLIT][NS][x2_pitch][CIN - CIN_A]
    const uint16_t *x2;
    int64_t x2_split;
    int x2_pitch, x2_roff, cinA8;
    int T_in, T_eff, ups, Tp, NS, row0, n_rows, cin8, n_stages;
    // plane pitch of the staged tile in 16-byte rows (>= n_rows).  The producers copy one staged row per thread, plane after
    // plane (16-byte pieces at the row-pitch lane stride).  A coalesced mapping (consecutive lanes = consecutive planes of one
    // row, conflict-free padded pitch) cut the LDGSTS shared-memory wavefronts 4-7x but was measured SLOWER (EQT tcconv class
    // 6.29 -> 6.36 ms, PhaseNet 3.31 -> 3.43 ms): the per-piece index arithmetic costs more issue slots than the wavefronts.
    int pitch;
    int ca;  // 1: producers use cp.async.ca (the sector of a 16-byte piece stays in L1 for the next plane of the row)
    int st256;  // 1: the 16-bit output rows, group and split offsets are 32-byte aligned -> adjacent 8-channel groups leave as one 256-bit store
    const uint16_t *w;  // [G][n_blocks][SPLIT][2][NOUT][8]
    int64_t w_gs;
    int n_blocks;
    const float *bias;  // [G][NOUT]
    int64_t b_gs;
    // MMA schedule of one tile, one entry per tcgen05.mma: low words of the A / B shared-memory
    // descriptors relative to the stage base / weight base (16-byte units | LBO << 16)
    int n_terms;
    uint32_t term_a[TC_MAX_TERMS], term_b[TC_MAX_TERMS];
    int fmt16;  // 0: fp16, 1: bf16
    int dbg;    // profiling aid (env VP_TC_DBG): 1 skip A loads, 2 skip MMAs, 4 skip epilogue body
    int act, pool, ph, cout, coutp, T_valid, T_out;
    int out_fmt;  // 0: channel-last 16-bit [SPLIT][G][NS][y_pitch][cout_cl], row (seq, t) at seq * y_pitch + y_roff + t;
                  // 1: fp32 (seq stride y_ss, channel stride y_cs); 2: fp32 like 1 after the fused 1x1 conv (8 -> 3) + softmax head;
                  // 3: none (only the fp32 row-major second output y32 is written)
    void *y;
    int64_t y_split, y_gs, y_ss, y_cs;
    int cout_cl, y_pitch, y_roff;
    float head_w[24], head_b[3];  // out_fmt 2: PhaseNet `out` conv, [class][channel]
    // time folded into the MMA N (fold > 1): accumulator column n holds channel n % fold_c of sample fold * row + n / fold_c - fold_o;
    // samples outside [0, fold_T) are written as zeros (they are the next layer's conv padding) / not written (fp32 outputs)
    int fold, fold_c, fold_o, fold_T;
    // epilogue extras (single group, direct, un-pooled layers: the res-CNN stack)
    const float *post_scale, *post_shift;  // [NOUT]: 16-bit output = relu(v * scale + shift) (pre-activation BN + ReLU of the next conv)
    const float *res;                      // fp32 row-major [NS][T_out][cout] added to v
    float *y32;                            // second output: v as fp32 row-major [NS][T_out][cout]
};

// Host-side description of one layer: packed weight blocks + MMA schedule.
struct TcLayer {
    int cin = 0, cout = 0, k = 0;
    int nout = 0;      // MMA N (phases * padded cout)
    int ph = 1;        // 1: direct, 2: polyphase x2 up-sampling folded into the weights
    int ups = 1;       // loader-side x2 nearest up-sampling (direct form)
    int crop = 0;      // samples dropped at the end of the up-sampled signal
    int split = 2;     // 2: fp16 hi/lo operands, 3 MMAs per K step (fp32-equivalent); 1: single bf16 pass
    int row0 = 0, halo = 0, n_blocks = 0, groups = 1;
    int sched_taps = 0, sched_nq = 0;  // taps (tap pairs when sched_nq == 0) and 16-channel pairs of the MMA schedule
    std::vector<TcMma> mma;
    std::vector<uint16_t> blocks;  // [G][n_blocks][split][2][nout][8]
    std::vector<float> bias;       // [G][nout]
    int64_t w_off = -1, b_off = -1;  // offsets (bytes / floats) once uploaded
};

enum { TC_DIRECT = 0, TC_POLYPHASE = 1, TC_DIRECT_UPS = 2 };

// weights: per group (cout, cin, k) fp32 with BatchNorm already folded; bias per group (cout) or nullptr
int tc_build_layer(TcLayer &L, int mode, int cin, int cout, int k, int crop, int split, int groups,
                   const float *const *weights, const float *const *bias, int pad_left = -1);

// 'same' Conv1d (odd k <= 9) folded over 4 time steps: (4 cout, 4 cin, 3) row-conv weights + replicated bias (tcconv.cu)
void tc_fold4_same(const float *W, const float *bias, int cout, int cin, int k, std::vector<float> &wf, std::vector<float> &bf);

struct TcIO {
    const uint16_t *x;
    int64_t x_split, x_gs;
    int T_in, NS;
    const uint16_t *w_dev;
    const float *b_dev;
    int act, pool;
    int out_fmt;
    void *y;
    int64_t y_split, y_gs, y_ss, y_cs;
    int cout_cl;
    const float *post_scale = nullptr, *post_shift = nullptr, *res = nullptr;
    float *y32 = nullptr;
    // storage overrides (0 / negative: dense defaults): rows per sequence and first row of the input / output buffers,
    // output rows to store per sequence, second input source, fused softmax head
    int x_pitch = 0, x_roff = 0, y_pitch = 0, y_roff = 0, T_valid = 0;
    const uint16_t *x2 = nullptr;
    int64_t x2_split = 0;
    int x2_pitch = 0, x2_roff = 0, cin_a = 0;  // cin_a: channels taken from x when x2 is set
    const float *head_w = nullptr, *head_b = nullptr;  // HOST pointers, out_fmt 2
    int fold = 1, fold_c = 0, fold_o = 0, fold_T = 0;   // output time folding (see TcP)
    // 1: the layer is a 'same' conv folded over 4 samples (64 columns = 4 samples x 16 channels) followed by ReLU + MaxPool1d(2):
    // the epilogue pools inside the accumulator row and writes [T / 4][2 x 16] = [T / 2][16] (tc_epilogue_foldpool)
    int foldpool = 0;
    // 1: stage the input with cp.async.ca (sectors kept in L1 between the two 16-byte pieces of a row).  Measured: PhaseNet
    // tcconv class 3.30 -> 3.21 ms per station-day, EQTransformer 6.23 -> 6.30 ms: set by the PhaseNet forward only.
    int ca = 0;
};
int tc_out_len(const TcLayer &L, int T_in, int pool);
int tc_launch(const TcLayer &L, const TcIO &io, cudaStream_t s);

// fp32 channel-first (NS, C, T) [strides] -> channel-last 16-bit [SPLIT][NS][T][C8]
int launch_pack_cl16(const float *x, int64_t x_ss, int64_t x_cs, int NS, int C, int T, int split, uint16_t *y,
                     int64_t y_split, int c8, cudaStream_t s);

}  // namespace vp
