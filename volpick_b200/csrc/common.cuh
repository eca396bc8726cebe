// Shared helpers for the volpick_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/volpick_b200.h"

namespace vp {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Per-kernel-class CUDA-event timing (vp_kernel_timing): off by default, one branch per launch when off.
enum KClass {
    KC_SLICE = 0, KC_SLICE_ENC0, KC_TCCONV, KC_DECA, KC_DECB, KC_CONV_F32, KC_CONVT_F32, KC_LSTM, KC_ATTN, KC_PACK,
    KC_STACK, KC_TRIM, KC_PICK, KC_FILTER, KC_RESSTACK, KC_COUNT
};
extern thread_local bool g_ktimer_on;
void ktimer_mark(int cls, cudaStream_t s, bool end);
struct KTimer {  // scoped: events on the launching stream around everything launched while it lives
    int cls;
    cudaStream_t s;
    KTimer(int c, cudaStream_t st) : cls(c), s(st) {
        if (g_ktimer_on) ktimer_mark(cls, s, false);
    }
    ~KTimer() {
        if (g_ktimer_on) ktimer_mark(cls, s, true);
    }
};

#define VP_CUDA_CHECK(expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            vp::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return VP_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

#define VP_LAUNCH_CHECK()                                                                       \
    do {                                                                                        \
        vp::count_launch();                                                                     \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            vp::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return VP_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

#define VP_REQUIRE(cond, code, ...)                                                             \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            vp::set_error(__VA_ARGS__);                                                         \
            return (code);                                                                      \
        }                                                                                       \
    } while (0)

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Opt-in of `bytes` of dynamic shared memory for `kernel` on the CURRENT device.  cudaFuncSetAttribute is per function and
// per device: the record of what has been granted is keyed by (kernel, device) and guarded by a mutex, so handles on
// several devices in one process and several host threads on one handle are fine.  Returns VP_OK / VP_ERR_CUDA.
int ensure_dyn_smem(const void *kernel, size_t bytes);
// Multiprocessors of the current device (cached per device; 148 on a B200).
int device_sm_count();

// ---- plan structures shared between model.cu and the kernel translation units -------------

struct ConvP {
    const float *x;   // (B, CIN, Lin) with batch stride x_bs (floats), group stride x_gs
    int64_t x_bs, x_gs;
    const float *w;   // per group [CIN][K][COUTP] (BatchNorm scale folded, COUT zero-padded)
    int64_t w_gs;
    const float *bias;  // per group [COUTP]
    int64_t b_gs;
    const float *pre_scale;  // [CIN] input affine + ReLU (pre-activation BatchNorm), or nullptr
    const float *pre_shift;
    const float *res;  // residual (B, COUT, Lout), batch stride r_bs, or nullptr
    int64_t r_bs;
    float *y;  // (B, COUT, Lout) with batch stride y_bs, group stride y_gs
    int64_t y_bs, y_gs;
    int Lin;         // stored input length
    int Lin_eff;     // input length seen by the conv (after x2 nearest up-sampling and crop)
    int Lconv;       // conv output positions
    int Lout;        // stored output length (after max-pool 2)
    int pad_left;    // zeros on the left of the (up-sampled) input
    int cout_store;  // real output channels
};

typedef int (*conv_launch_fn)(const ConvP &p, int B, int G, cudaStream_t s);

// key: compile-time shape of a conv instance
struct ConvKey {
    int cin, coutp, k, stride, ups, pool, act, pre, res;
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_SOFTMAX3 = 3 };

conv_launch_fn find_conv_fp32(const ConvKey &key, int Lconv);

struct ConvTP {
    const float *x;  // (B, CIN, Lin)
    int64_t x_bs;
    const float *w;     // [CIN][7][COUT] folded
    const float *bias;  // [COUT]
    float *y;           // (B, *, Lout) channel offset already applied
    int64_t y_bs;
    int Lin, Lout, shift;  // y[t] = relu(convT(x)[t + shift] + bias)
};
int launch_convt_fp32(int cin, int cout, const ConvTP &p, int B, cudaStream_t s);

// K1 fused with encoder.convs.0 (+ ReLU + MaxPool1d(2)) for the tensor-core path (prepost.cu)
int launch_slice_enc0(const void *trace, int dtype, int64_t ch_stride, const int64_t *starts, int64_t nw, int L, int scope,
                      int taper, const float *w_host, const float *b_host, int split, uint16_t *out, int64_t out_split,
                      cudaStream_t s, int k = 11, int out_pitch = 0);

struct LstmP {
    const float *x;  // (B, CIN, T), group stride x_gs
    int64_t x_bs, x_gs;
    const float *w_ih;  // per group, per direction: [CIN][16 units][4 gates]
    const float *w_hh;  // per group, per direction: [16 k][16 units][4 gates]
    const float *bias;  // per group, per direction: [16 units][4 gates]  (b_ih + b_hh)
    int64_t w_gs_ih, w_gs_hh, w_gs_b;
    float *y;  // (B, 16*ndir, T), group stride y_gs
    int64_t y_bs, y_gs;
    int T, ndir, B;
    // cin == 0: the input projection W_ih x + b was computed by a tensor-core GEMM: proj[b][t][dir][16 units][4 gates] fp32
    const float *proj = nullptr;
};
int launch_lstm(int cin, const LstmP &p, int G, cudaStream_t s);

struct AttnP {
    const float *x;  // (B, 16, T)
    int64_t x_bs, x_gs;
    const float *w;  // packed parameter block per group (see model.cu pack_attention)
    int64_t w_gs;
    float *y;
    int64_t y_bs, y_gs;
    int T, B;
    int width;  // 0: full attention, else band width (3)
    int mode;   // 0: transformer (attention + LN + FF + LN), 1: attention only
};
int launch_attention(const AttnP &p, int G, cudaStream_t s);
// layout of the packed attention / transformer parameter block (floats)
enum {
    AW_WT = 0,              // [16][32]
    AW_WX = 512,            // [16][32]
    AW_BH = 1024,           // [32]
    AW_WA = 1056,           // [32]
    AW_BA = 1088,           // [1] (+3 pad)
    AW_G1 = 1092,           // [16]
    AW_B1 = 1108,           // [16]
    AW_L1W = 1124,          // [128][16]   lin1.weight
    AW_L1B = 1124 + 2048,   // [128]
    AW_L2W = 3300,          // [128][16]   lin2.weight transposed
    AW_L2B = 3300 + 2048,   // [16]
    AW_G2 = 5364,           // [16]
    AW_B2 = 5380,           // [16]
    AW_SIZE = 5396
};

}  // namespace vp
