// fp32 ("exact mode") Conv1d / ConvTranspose1d kernels on the CUDA cores, fused epilogues.
//
// Replaces, for the volpick EQTransformer / PhaseNet forwards:
//   F.conv1d + bias + BatchNorm1d(eval, folded on the host) + ReLU / sigmoid / channel softmax,
//   nn.Upsample(x2, nearest) [+ crop] fused into the input gather, MaxPool1d(2) with the -1e10
//   right pad fused into the store, pre-activation BatchNorm+ReLU (ResCNNBlock) fused into the
//   input gather, residual add fused into the store, F.conv_transpose1d(k7, s4) + crop + skip
//   concat written straight into the concat buffer.
//
// Mapping: one CTA = one window x one tile of 32*TT output positions x all output channels.
// A warp owns TCO output channels (weights are warp-uniform -> shared-memory broadcasts), a lane
// owns TT output positions strided by 32 (conflict-free activation reads).  Input channels are
// streamed through shared memory in chunks of CCH.
#include "common.cuh"

namespace vp {

template <int CIN, int COUTP, int K, int STRIDE, int UPS, int POOL, int ACT, int PRE, int RES, int TCO, int TT,
          int CCH>
__global__ void __launch_bounds__(32 * (COUTP / TCO)) conv1d_f32_kernel(const ConvP p) {
    constexpr int NW = COUTP / TCO;
    constexpr int NT = 32 * NW;
    constexpr int TILE = 32 * TT;
    constexpr int XT = (TILE - 1) * STRIDE + K;
    constexpr int XS_FLOATS = (CCH * XT + 3) & ~3;
    static_assert(COUTP % TCO == 0 && TCO % 4 == 0, "channel tiling");
    static_assert(CIN % CCH == 0, "CCH must divide CIN");
    static_assert(ACT != ACT_SOFTMAX3 || (NW == 1 && TCO == 4), "softmax needs all channels in one thread");

    extern __shared__ __align__(16) float smem[];
    float *xs = smem;              // [CCH][XT]
    float *ws = smem + XS_FLOATS;  // [CCH][K][COUTP]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, g = blockIdx.z;
    const int t0 = blockIdx.x * TILE;
    const int co0 = warp * TCO;
    const float *xb = p.x + (int64_t)g * p.x_gs + (int64_t)b * p.x_bs;
    const float *wg = p.w + (int64_t)g * p.w_gs;
    const int in0 = t0 * STRIDE - p.pad_left;

    float acc[TCO][TT];
#pragma unroll
    for (int q = 0; q < TCO; ++q)
#pragma unroll
        for (int j = 0; j < TT; ++j) acc[q][j] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += CCH) {
        if (c0 > 0) __syncthreads();
        for (int idx = tid; idx < CCH * XT; idx += NT) {
            const int c = idx / XT;
            const int pp = idx - c * XT;
            const int u = in0 + pp;
            float v = 0.f;
            if (u >= 0 && u < p.Lin_eff) {
                const int src = (UPS == 2) ? (u >> 1) : u;
                v = __ldg(xb + (int64_t)(c0 + c) * p.Lin + src);
                if (PRE) v = fmaxf(fmaf(v, __ldg(p.pre_scale + c0 + c), __ldg(p.pre_shift + c0 + c)), 0.f);
            }
            xs[idx] = v;
        }
        {
            const float4 *src = reinterpret_cast<const float4 *>(wg + (int64_t)c0 * K * COUTP);
            float4 *dst = reinterpret_cast<float4 *>(ws);
            for (int idx = tid; idx < CCH * K * COUTP / 4; idx += NT) dst[idx] = __ldg(src + idx);
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CCH; ++c) {
            const float *xr = xs + c * XT + lane * STRIDE;
            const float *wr = ws + c * K * COUTP + co0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float wv[TCO];
#pragma unroll
                for (int q4 = 0; q4 < TCO / 4; ++q4) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wr + k * COUTP + q4 * 4);
                    wv[q4 * 4 + 0] = w4.x;
                    wv[q4 * 4 + 1] = w4.y;
                    wv[q4 * 4 + 2] = w4.z;
                    wv[q4 * 4 + 3] = w4.w;
                }
#pragma unroll
                for (int j = 0; j < TT; ++j) {
                    const float xv = xr[j * 32 * STRIDE + k];
#pragma unroll
                    for (int q = 0; q < TCO; ++q) acc[q][j] = fmaf(wv[q], xv, acc[q][j]);
                }
            }
        }
    }

    // ---- epilogue -------------------------------------------------------------------------
    float *yb = p.y + (int64_t)g * p.y_gs + (int64_t)b * p.y_bs;
    const float *bias = p.bias + (int64_t)g * p.b_gs + co0;
    if (ACT == ACT_SOFTMAX3) {
#pragma unroll
        for (int j = 0; j < TT; ++j) {
            const int t = t0 + lane + 32 * j;
            const float v0 = acc[0][j] + bias[0], v1 = acc[1][j] + bias[1], v2 = acc[2][j] + bias[2];
            const float m = fmaxf(v0, fmaxf(v1, v2));
            const float e0 = expf(v0 - m), e1 = expf(v1 - m), e2 = expf(v2 - m);
            const float s = e0 + e1 + e2;
            if (t < p.Lout) {
                yb[t] = e0 / s;
                yb[(int64_t)p.Lout + t] = e1 / s;
                yb[2 * (int64_t)p.Lout + t] = e2 / s;
            }
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < TCO; ++q) {
        const int co = co0 + q;
        const float bq = __ldg(bias + q);
        const bool co_ok = co < p.cout_store;
#pragma unroll
        for (int j = 0; j < TT; ++j) {
            const int t = t0 + lane + 32 * j;
            float v = acc[q][j] + bq;
            if (RES) {
                if (co_ok && t < p.Lout) v += __ldg(p.res + (int64_t)b * p.r_bs + (int64_t)co * p.Lout + t);
            }
            if (ACT == ACT_RELU) v = fmaxf(v, 0.f);
            if (ACT == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
            if (POOL == 2) {
                if (t >= p.Lconv) v = -1e10f;  // SeisBench Encoder: right pad with -1e10 before MaxPool1d(2)
                const float o = __shfl_xor_sync(0xffffffffu, v, 1);
                v = fmaxf(v, o);
                const int tp = t >> 1;
                if (!(lane & 1) && co_ok && tp < p.Lout) yb[(int64_t)co * p.Lout + tp] = v;
            } else {
                if (co_ok && t < p.Lout) yb[(int64_t)co * p.Lout + t] = v;
            }
        }
    }
}

template <int CIN, int COUTP, int K, int STRIDE, int UPS, int POOL, int ACT, int PRE, int RES, int TCO, int TT,
          int CCH>
static int launch_conv(const ConvP &p, int B, int G, cudaStream_t s) {
    constexpr int NW = COUTP / TCO;
    constexpr int TILE = 32 * TT;
    constexpr int XT = (TILE - 1) * STRIDE + K;
    constexpr int XS_FLOATS = (CCH * XT + 3) & ~3;
    constexpr size_t SMEM = (size_t)(XS_FLOATS + CCH * K * COUTP) * sizeof(float);
    auto kern = conv1d_f32_kernel<CIN, COUTP, K, STRIDE, UPS, POOL, ACT, PRE, RES, TCO, TT, CCH>;
    if (int rc = ensure_dyn_smem((const void *)kern, SMEM)) return rc;
    dim3 grid((p.Lconv + TILE - 1) / TILE, B, G);
    KTimer kt(KC_CONV_F32, s);
    kern<<<grid, 32 * NW, SMEM, s>>>(p);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

struct ConvEntry {
    ConvKey key;
    conv_launch_fn fn;
};

#define REG(CIN, COUTP, K, S, U, P, ACT, PRE, RES, TCO, TT, CCH) \
    { {CIN, COUTP, K, S, U, P, ACT, PRE, RES}, &launch_conv<CIN, COUTP, K, S, U, P, ACT, PRE, RES, TCO, TT, CCH> }

static const ConvEntry kConvTable[] = {
    // ---- EQTransformer encoder: relu(conv) -> maxpool2
    REG(3, 8, 11, 1, 1, 2, ACT_RELU, 0, 0, 4, 8, 3),
    REG(8, 16, 9, 1, 1, 2, ACT_RELU, 0, 0, 8, 8, 8),
    REG(16, 16, 7, 1, 1, 2, ACT_RELU, 0, 0, 8, 8, 16),
    REG(16, 32, 7, 1, 1, 2, ACT_RELU, 0, 0, 8, 8, 16),
    REG(32, 32, 5, 1, 1, 2, ACT_RELU, 0, 0, 8, 4, 16),
    REG(32, 64, 5, 1, 1, 2, ACT_RELU, 0, 0, 8, 6, 16),
    REG(64, 64, 3, 1, 1, 2, ACT_RELU, 0, 0, 8, 3, 16),
    // ---- res-CNN: conv(relu(bn(x))) [+ x]
    REG(64, 64, 3, 1, 1, 1, ACT_NONE, 1, 0, 8, 2, 16),
    REG(64, 64, 3, 1, 1, 1, ACT_NONE, 1, 1, 8, 2, 16),
    REG(64, 64, 2, 1, 1, 1, ACT_NONE, 1, 0, 8, 2, 16),
    REG(64, 64, 2, 1, 1, 1, ACT_NONE, 1, 1, 8, 2, 16),
    // ---- BiLSTM block 1x1 conv (+ folded BN)
    REG(32, 16, 1, 1, 1, 1, ACT_NONE, 0, 0, 4, 2, 32),
    // ---- decoders: relu(conv(upsample2(x)))
    REG(16, 64, 3, 1, 2, 1, ACT_RELU, 0, 0, 8, 3, 16),
    REG(64, 64, 5, 1, 2, 1, ACT_RELU, 0, 0, 8, 6, 16),
    REG(64, 32, 5, 1, 2, 1, ACT_RELU, 0, 0, 8, 4, 16),
    REG(32, 32, 7, 1, 2, 1, ACT_RELU, 0, 0, 8, 8, 16),
    REG(32, 16, 7, 1, 2, 1, ACT_RELU, 0, 0, 8, 8, 16),
    REG(16, 16, 9, 1, 2, 1, ACT_RELU, 0, 0, 8, 8, 16),
    REG(16, 8, 11, 1, 2, 1, ACT_RELU, 0, 0, 4, 8, 16),
    // ---- heads: sigmoid(conv k11), one real output channel
    REG(8, 4, 11, 1, 1, 1, ACT_SIGMOID, 0, 0, 4, 8, 8),
    // ---- PhaseNet
    REG(3, 8, 7, 1, 1, 1, ACT_RELU, 0, 0, 4, 8, 3),
    REG(8, 8, 7, 1, 1, 1, ACT_RELU, 0, 0, 4, 8, 8),
    REG(8, 8, 7, 4, 1, 1, ACT_RELU, 0, 0, 4, 4, 8),
    REG(8, 16, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 8, 8),
    REG(16, 16, 7, 4, 1, 1, ACT_RELU, 0, 0, 8, 2, 16),
    REG(16, 32, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 6, 16),
    REG(32, 32, 7, 4, 1, 1, ACT_RELU, 0, 0, 8, 2, 16),
    REG(32, 64, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 2, 16),
    REG(64, 64, 7, 4, 1, 1, ACT_RELU, 0, 0, 8, 1, 16),
    REG(64, 128, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 1, 16),
    REG(128, 64, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 2, 16),
    REG(64, 32, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 6, 16),
    REG(32, 16, 7, 1, 1, 1, ACT_RELU, 0, 0, 8, 8, 16),
    REG(16, 8, 7, 1, 1, 1, ACT_RELU, 0, 0, 4, 8, 16),
    REG(8, 4, 1, 1, 1, 1, ACT_SOFTMAX3, 0, 0, 4, 8, 8),
};

conv_launch_fn find_conv_fp32(const ConvKey &k, int /*Lconv*/) {
    for (const ConvEntry &e : kConvTable) {
        const ConvKey &a = e.key;
        if (a.cin == k.cin && a.coutp == k.coutp && a.k == k.k && a.stride == k.stride && a.ups == k.ups &&
            a.pool == k.pool && a.act == k.act && a.pre == k.pre && a.res == k.res)
            return e.fn;
    }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose1d(k=7, stride=4, no bias) + folded BN + ReLU + crop, written into the concat
// buffer (PhaseNet up branch).  out[t] uses taps k = ph and k = ph + 4 (ph = (t+shift) & 3).
template <int CIN, int COUT, int CCH>
__global__ void __launch_bounds__(128) convt_k7s4_kernel(const ConvTP p) {
    __shared__ __align__(16) float ws[CCH * 7 * COUT];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t = blockIdx.x * 128 + tid;
    const int u = t + p.shift;
    const int i0 = u >> 2, ph = u & 3;
    const bool a_ok = (t < p.Lout) && (i0 < p.Lin);
    const bool b_ok = (t < p.Lout) && (ph <= 2) && (i0 >= 1) && (i0 - 1 < p.Lin);
    const int kb = (ph <= 2) ? ph + 4 : ph;
    const float *xb = p.x + (int64_t)b * p.x_bs;

    float acc[COUT];
#pragma unroll
    for (int q = 0; q < COUT; ++q) acc[q] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += CCH) {
        if (c0 > 0) __syncthreads();
        {
            const float4 *src = reinterpret_cast<const float4 *>(p.w + (int64_t)c0 * 7 * COUT);
            float4 *dst = reinterpret_cast<float4 *>(ws);
            for (int idx = tid; idx < CCH * 7 * COUT / 4; idx += 128) dst[idx] = __ldg(src + idx);
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CCH; ++c) {
            const float xa = a_ok ? __ldg(xb + (int64_t)(c0 + c) * p.Lin + i0) : 0.f;
            const float xv = b_ok ? __ldg(xb + (int64_t)(c0 + c) * p.Lin + i0 - 1) : 0.f;
            const float *wa = ws + (c * 7 + ph) * COUT;
            const float *wb = ws + (c * 7 + kb) * COUT;
#pragma unroll
            for (int q4 = 0; q4 < COUT / 4; ++q4) {
                const float4 a4 = *reinterpret_cast<const float4 *>(wa + q4 * 4);
                const float4 b4 = *reinterpret_cast<const float4 *>(wb + q4 * 4);
                acc[q4 * 4 + 0] = fmaf(a4.x, xa, fmaf(b4.x, xv, acc[q4 * 4 + 0]));
                acc[q4 * 4 + 1] = fmaf(a4.y, xa, fmaf(b4.y, xv, acc[q4 * 4 + 1]));
                acc[q4 * 4 + 2] = fmaf(a4.z, xa, fmaf(b4.z, xv, acc[q4 * 4 + 2]));
                acc[q4 * 4 + 3] = fmaf(a4.w, xa, fmaf(b4.w, xv, acc[q4 * 4 + 3]));
            }
        }
    }
    if (t < p.Lout) {
        float *yb = p.y + (int64_t)b * p.y_bs;
#pragma unroll
        for (int q = 0; q < COUT; ++q) yb[(int64_t)q * p.Lout + t] = fmaxf(acc[q] + __ldg(p.bias + q), 0.f);
    }
}

template <int CIN, int COUT, int CCH>
static int launch_convt(const ConvTP &p, int B, cudaStream_t s) {
    dim3 grid((p.Lout + 127) / 128, B);
    KTimer kt(KC_CONVT_F32, s);
    convt_k7s4_kernel<CIN, COUT, CCH><<<grid, 128, 0, s>>>(p);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

int launch_convt_fp32(int cin, int cout, const ConvTP &p, int B, cudaStream_t s) {
    if (cin == 128 && cout == 64) return launch_convt<128, 64, 8>(p, B, s);
    if (cin == 64 && cout == 32) return launch_convt<64, 32, 16>(p, B, s);
    if (cin == 32 && cout == 16) return launch_convt<32, 16, 32>(p, B, s);
    if (cin == 16 && cout == 8) return launch_convt<16, 8, 16>(p, B, s);
    set_error("no ConvTranspose1d instance for %d -> %d", cin, cout);
    return VP_ERR_UNSUPPORTED;
}

}  // namespace vp
