// Sequence kernels of the EQTransformer bottleneck (T = 47): LSTM recurrences and the additive
// self-attention / transformer blocks.  fp32; exp / tanh / sigmoid through ex2.approx + rcp.approx in
// cancellation-free forms (absolute error about 2e-7, see ex2_approx below).
//
// Replaces nn.LSTM(in, 16[, bidirectional]) inside BiLSTMBlock / pick_lstms, SeqSelfAttention,
// LayerNormalization, FeedForward and Transformer of seisbench/models/eqtransformer.py
// (weight shapes from
// /root/reference/Final_models/volpick/eqtransformer/volpick.pt.v1).
#include <cstdlib>

#include "common.cuh"

namespace vp {

// ---------------------------------------------------------------------------------------------
// LSTM: 16 lanes = the 16 hidden units of one (window, direction) sequence; a 128-thread CTA runs
// 8 sequences.  W_hh lives in registers (4 gates x 16), W_ih in shared memory as float4 per
// (input channel, unit) = the 4 gates; h is exchanged with width-16 shuffles.  (Measured alternatives of round 2, all slower or
// neutral: W_hh in shared memory -- 56 instead of 128 registers, twice the warps per SM -- 2.05 vs 1.55 ms per station-day; the
// gate FMAs as packed FFMA2: 1.59 ms; W_ih of the 16-channel layers in registers too -- no LDS.128 of W_ih, a third of the shared-memory
// wavefronts per step, but 168 registers and three CTAs per SM -- 1.85 ms: the recurrence needs its 16 warps per SM; h staged in
// shared memory as [unit][t] and written out as one contiguous run per sequence instead of a strided store per step -- 1.52 ms, unchanged.)
// Transcendentals on the SFU with fp32-level ABSOLUTE accuracy (about 2e-7): ex2.approx and rcp.approx are good
// to ~2 ulp; the forms below never cancel (tanh = 1 - 2 / (e^{2x} + 1), sigmoid = 1 / (1 + e^{-x})) and saturate
// correctly (ex2 -> +inf gives rcp -> 0).  They replace expf / tanhf (about 20 instructions each) with 4-5.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanhf_(float x) { return fmaf(-2.f, rcp_approx(1.f + ex2_approx(2.8853900817779268f * x)), 1.f); }

template <int CIN>
__global__ void __launch_bounds__(128) lstm_kernel(const LstmP p) {
    extern __shared__ __align__(16) float smem[];
    float4 *wih = reinterpret_cast<float4 *>(smem);  // [ndir][CIN][16]
    const int g = blockIdx.y;
    const int tid = threadIdx.x;
    const int ndir = p.ndir;
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.w_ih + (int64_t)g * p.w_gs_ih);
        for (int i = tid; i < ndir * CIN * 16; i += 128) wih[i] = __ldg(src + i);
    }
    __syncthreads();

    const int j = tid & 15;
    const int64_t seq = (int64_t)blockIdx.x * 8 + (tid >> 4);
    const int64_t nseq = (int64_t)p.B * ndir;
    const bool active = seq < nseq;
    const int64_t b = active ? seq / ndir : 0;
    const int dir = active ? (int)(seq % ndir) : 0;
    const unsigned hmask = 0xffffu << (tid & 16);

    float4 whh[16];
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.w_hh + (int64_t)g * p.w_gs_hh) + dir * 256;
#pragma unroll
        for (int k = 0; k < 16; ++k) whh[k] = __ldg(src + k * 16 + j);
    }
    const float4 bias = CIN == 0 ? make_float4(0.f, 0.f, 0.f, 0.f)
                                 : __ldg(reinterpret_cast<const float4 *>(p.bias + (int64_t)g * p.w_gs_b) + dir * 16 + j);
    const float4 *wi = wih + dir * CIN * 16 + j;
    const float *xb = CIN == 0 ? nullptr : p.x + (int64_t)g * p.x_gs + b * p.x_bs;
    const float4 *pj = CIN == 0 ? reinterpret_cast<const float4 *>(p.proj) + ((b * p.T) * ndir + dir) * 16 + j : nullptr;
    float *yb = p.y + (int64_t)g * p.y_gs + b * p.y_bs + (int64_t)(dir * 16 + j) * p.T;
    const int T = p.T;

    // input projection of one time step: independent of the recurrence, computed one step ahead so that its
    // loads and FMAs overlap the shuffle / transcendental chain of the current step.  (Blocking it over 4 steps to
    // save W_ih reads was measured SLOWER: 168 registers, one CTA fewer per SM, 1.31 -> 1.66 ms per station-day.)
    auto in_proj = [&](int t) {
        if constexpr (CIN == 0) return __ldg(pj + (int64_t)t * ndir * 16);  // bias included by the GEMM
        float4 a0 = bias, a1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int ci = 0; ci < CIN; ci += 2) {
            const float x0 = __ldg(xb + (int64_t)ci * T + t), x1 = __ldg(xb + (int64_t)(ci + 1) * T + t);
            const float4 w0 = wi[ci * 16], w1 = wi[(ci + 1) * 16];
            a0.x = fmaf(w0.x, x0, a0.x), a0.y = fmaf(w0.y, x0, a0.y), a0.z = fmaf(w0.z, x0, a0.z), a0.w = fmaf(w0.w, x0, a0.w);
            a1.x = fmaf(w1.x, x1, a1.x), a1.y = fmaf(w1.y, x1, a1.y), a1.z = fmaf(w1.z, x1, a1.z), a1.w = fmaf(w1.w, x1, a1.w);
        }
        return make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    };

    float h = 0.f, c = 0.f;
    float4 ain = in_proj(dir ? (T - 1) : 0);
    for (int s = 0; s < T; ++s) {
        const int t = dir ? (T - 1 - s) : s;
        float4 a = ain, a2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const float h0 = __shfl_sync(hmask, h, k, 16), h1 = __shfl_sync(hmask, h, k + 1, 16);
            a.x = fmaf(whh[k].x, h0, a.x), a.y = fmaf(whh[k].y, h0, a.y), a.z = fmaf(whh[k].z, h0, a.z), a.w = fmaf(whh[k].w, h0, a.w);
            a2.x = fmaf(whh[k + 1].x, h1, a2.x), a2.y = fmaf(whh[k + 1].y, h1, a2.y), a2.z = fmaf(whh[k + 1].z, h1, a2.z),
            a2.w = fmaf(whh[k + 1].w, h1, a2.w);
        }
        if (s + 1 < T) ain = in_proj(dir ? (T - 2 - s) : (s + 1));
        const float ig = sigmoidf_(a.x + a2.x), fg = sigmoidf_(a.y + a2.y), gg = tanhf_(a.z + a2.z), og = sigmoidf_(a.w + a2.w);
        c = fmaf(fg, c, ig * gg);
        h = og * tanhf_(c);
        if (active) yb[t] = h;
    }
}

int launch_lstm(int cin, const LstmP &p, int G, cudaStream_t s) {
    const int64_t nseq = (int64_t)p.B * p.ndir;
    dim3 grid((unsigned)((nseq + 7) / 8), G);
    const size_t smem = (size_t)p.ndir * cin * 16 * sizeof(float4);
    KTimer kt(KC_LSTM, s);
    if (cin == 0 && p.proj != nullptr) {
        lstm_kernel<0><<<grid, 128, 0, s>>>(p);
    } else if (cin == 64) {
        lstm_kernel<64><<<grid, 128, smem, s>>>(p);
    } else if (cin == 16) {
        lstm_kernel<16><<<grid, 128, smem, s>>>(p);
    } else {
        set_error("no LSTM instance for %d input channels", cin);
        return VP_ERR_UNSUPPORTED;
    }
    VP_LAUNCH_CHECK();
    return VP_OK;
}

// ---------------------------------------------------------------------------------------------
// Additive self-attention (+ optional transformer tail).  One thread = one query time step of one
// window; a 240-thread CTA handles 5 windows in 48-thread slots (T = 47: 98 % of the lanes work; 64-thread slots left
// 27 % of them idle) and loads the 21.6 KB parameter block once for the five.
constexpr int AT_MAXT = 48;   // T <= 48 (T = 47 for 6000-sample windows)
constexpr int AT_XP = 20;     // pitch of Xs rows (time-major x; float4 aligned: the weighted sum reads a row as two LDS.128 per lane)
constexpr int AT_KP = 36;     // pitch of the k-projection rows (float4 aligned)
constexpr int AT_OFF_K = (AW_SIZE + 3) & ~3;                      // 16-byte aligned (float4 reads)
constexpr int AT_WPC = 5;     // windows per CTA: 240 threads x 127 registers -> two CTAs (10 windows) per SM, register-file bound
constexpr int AT_NT = AT_WPC * AT_MAXT;
constexpr int AT_OFF_X = AT_OFF_K + AT_WPC * AT_MAXT * AT_KP;
constexpr int AT_SMEM_FLOATS = AT_OFF_X + AT_WPC * AT_MAXT * AT_XP;

__global__ void __launch_bounds__(AT_NT, 2) attention_kernel(const AttnP p) {
    extern __shared__ __align__(16) float at_smem[];
    float *wsm = at_smem;                                               // [AW_SIZE]
    float(*Ks)[AT_MAXT * AT_KP] = reinterpret_cast<float(*)[AT_MAXT * AT_KP]>(at_smem + AT_OFF_K);
    float(*Xs)[AT_MAXT * AT_XP] = reinterpret_cast<float(*)[AT_MAXT * AT_XP]>(at_smem + AT_OFF_X);

    const int tid = threadIdx.x;
    const int g = blockIdx.y;
    const int T = p.T;
    {
        const float *src = p.w + (int64_t)g * p.w_gs;
        const int n = (p.mode == 0) ? AW_SIZE : AW_G1;
        for (int i = tid; i < n; i += AT_NT) {
            const float v = __ldg(src + i);
            wsm[i] = (i >= AW_WA && i < AW_WA + 32) ? -2.f * v : v;  // Wa is only used as -2 Wa (see the emission loop)
        }
    }
    const int wl = tid / AT_MAXT;      // window slot in the CTA
    const int i = tid - wl * AT_MAXT;  // query time step
    const int64_t b = (int64_t)blockIdx.x * AT_WPC + wl;
    const bool bval = b < p.B;
    const float *xb = p.x + (int64_t)g * p.x_gs + (bval ? b : 0) * p.x_bs;
    // transpose-load x (16, T) -> Xs[t][c]
    for (int idx = i; idx < 16 * T; idx += AT_MAXT) {
        const int c = idx / T, t = idx - c * T;
        Xs[wl][t * AT_XP + c] = bval ? __ldg(xb + idx) : 0.f;
    }
    __syncthreads();

    const bool act = bval && i < T;
    float x[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) x[c] = act ? Xs[wl][i * AT_XP + c] : 0.f;

    // Emissions e_ij = Wa . tanh(q_i + k_j + bh) + ba.  With E(z) = 2^(2 log2(e) z) = e^{2z}:
    //     tanh(a + b) = 1 - 2 / (E(a) E(b) + 1),
    // so the T x T x 32 grid costs one FFMA + one MUFU.RCP + one FFMA per element, and only the 2 x T x 32
    // exponentials E(q_i + bh), E(k_j) go through ex2.  The exponents are clamped to +-63 (|q|, |k| <= 21.8, far
    // beyond tanh saturation for bounded LayerNorm / LSTM inputs) so that the product never overflows or hits 0 * inf.
    constexpr float kTwoLog2e = 2.8853900817779268f;
    float q[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) q[u] = wsm[AW_BH + u];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int u = 0; u < 32; ++u) q[u] = fmaf(x[c], wsm[AW_WT + c * 32 + u], q[u]);
    }
#pragma unroll
    for (int u = 0; u < 32; ++u) q[u] = ex2_approx(fminf(fmaxf(kTwoLog2e * q[u], -63.f), 63.f));
    if (i < T) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            float kv = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) kv = fmaf(x[c], wsm[AW_WX + c * 32 + u], kv);
            Ks[wl][i * AT_KP + u] = ex2_approx(fminf(fmaxf(kTwoLog2e * kv, -63.f), 63.f));
        }
    }
    __syncthreads();

    // softmax over j with the row max over the FULL row, band mask applied after the exp, denominator + 1e-5
    // (SeqSelfAttention, original_compatible=False) -- evaluated online: the running maximum rescales the partial sums, so
    // the T x T emissions are never stored (9.4 KB of shared memory per window less: twice the windows per SM).
    float emax = -INFINITY;
    float e0 = wsm[AW_BA];
#pragma unroll
    for (int u = 0; u < 32; ++u) e0 = fmaf(-0.5f, wsm[AW_WA + u], e0);  // ba + sum_u Wa_u; each term then adds -2 Wa_u / (E E + 1)
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    float ssum = 0.f;
    const int half = p.width / 2;
    for (int jj = 0; jj < T; ++jj) {
        const float4 *kr = reinterpret_cast<const float4 *>(&Ks[wl][jj * AT_KP]);
        float ea = e0, eb = 0.f;
#pragma unroll
        for (int u4 = 0; u4 < 8; u4 += 2) {
            const float4 k4 = kr[u4], k5 = kr[u4 + 1];
            ea = fmaf(wsm[AW_WA + u4 * 4 + 0], rcp_approx(fmaf(q[u4 * 4 + 0], k4.x, 1.f)), ea);
            eb = fmaf(wsm[AW_WA + u4 * 4 + 1], rcp_approx(fmaf(q[u4 * 4 + 1], k4.y, 1.f)), eb);
            ea = fmaf(wsm[AW_WA + u4 * 4 + 2], rcp_approx(fmaf(q[u4 * 4 + 2], k4.z, 1.f)), ea);
            eb = fmaf(wsm[AW_WA + u4 * 4 + 3], rcp_approx(fmaf(q[u4 * 4 + 3], k4.w, 1.f)), eb);
            ea = fmaf(wsm[AW_WA + u4 * 4 + 4], rcp_approx(fmaf(q[u4 * 4 + 4], k5.x, 1.f)), ea);
            eb = fmaf(wsm[AW_WA + u4 * 4 + 5], rcp_approx(fmaf(q[u4 * 4 + 5], k5.y, 1.f)), eb);
            ea = fmaf(wsm[AW_WA + u4 * 4 + 6], rcp_approx(fmaf(q[u4 * 4 + 6], k5.z, 1.f)), ea);
            eb = fmaf(wsm[AW_WA + u4 * 4 + 7], rcp_approx(fmaf(q[u4 * 4 + 7], k5.w, 1.f)), eb);
        }
        const float e = ea + eb;
        if (e > emax) {  // new row maximum: bring the partial sums to the new reference
            const float sc = ex2_approx(1.4426950408889634f * (emax - e));  // exp(-inf) = 0 on the first step
            ssum *= sc;
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] *= sc;
            emax = e;
        }
        bool in_band = true;
        if (p.width > 0) {
            const int lower = jj - half;
            in_band = lower <= i && i < lower + p.width;
        }
        if (in_band) {
            const float w = ex2_approx(1.4426950408889634f * (e - emax));
            ssum += w;
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = fmaf(w, Xs[wl][jj * AT_XP + c], v[c]);
        }
    }
    const float inv = 1.f / (ssum + 1e-5f);
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] *= inv;

    float *yb = p.y + (int64_t)g * p.y_gs + (bval ? b : 0) * p.y_bs;
    if (p.mode == 1) {
        if (act) {
#pragma unroll
            for (int c = 0; c < 16; ++c) yb[c * T + i] = v[c];
        }
        return;
    }
    // transformer tail: y = LN(x + attn); out = LN(y + lin2(relu(lin1(y))))
    float y[16];
    {
        float mean = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            y[c] = x[c] + v[c];
            mean += y[c];
        }
        mean *= (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float d = y[c] - mean;
            var = fmaf(d, d, var);
        }
        var = var * (1.f / 16.f) + 1e-14f;
        const float sd = sqrtf(var);
#pragma unroll
        for (int c = 0; c < 16; ++c) y[c] = (y[c] - mean) / sd * wsm[AW_G1 + c] + wsm[AW_B1 + c];
    }
    float o[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = wsm[AW_L2B + c];
#pragma unroll 4
    for (int m = 0; m < 128; ++m) {
        const float4 *w1 = reinterpret_cast<const float4 *>(&wsm[AW_L1W + m * 16]);
        float hsum = wsm[AW_L1B + m];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 w = w1[c4];
            hsum = fmaf(w.x, y[c4 * 4 + 0], hsum);
            hsum = fmaf(w.y, y[c4 * 4 + 1], hsum);
            hsum = fmaf(w.z, y[c4 * 4 + 2], hsum);
            hsum = fmaf(w.w, y[c4 * 4 + 3], hsum);
        }
        hsum = fmaxf(hsum, 0.f);
        const float4 *w2 = reinterpret_cast<const float4 *>(&wsm[AW_L2W + m * 16]);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 w = w2[c4];
            o[c4 * 4 + 0] = fmaf(w.x, hsum, o[c4 * 4 + 0]);
            o[c4 * 4 + 1] = fmaf(w.y, hsum, o[c4 * 4 + 1]);
            o[c4 * 4 + 2] = fmaf(w.z, hsum, o[c4 * 4 + 2]);
            o[c4 * 4 + 3] = fmaf(w.w, hsum, o[c4 * 4 + 3]);
        }
    }
    {
        float mean = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            o[c] += y[c];
            mean += o[c];
        }
        mean *= (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float d = o[c] - mean;
            var = fmaf(d, d, var);
        }
        var = var * (1.f / 16.f) + 1e-14f;
        const float sd = sqrtf(var);
        if (act) {
#pragma unroll
            for (int c = 0; c < 16; ++c) yb[c * T + i] = (o[c] - mean) / sd * wsm[AW_G2 + c] + wsm[AW_B2 + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same block with TWO threads per query time step (adjacent lanes): each holds 16 of the 32 attention units, 8 of the 16
// channels of the weighted sum and 64 of the 128 hidden units of the feed-forward layer; the halves meet in one shuffle per
// emission (both lanes then take the same softmax path on bit-identical values: a + b == b + a), 8 for the LayerNorm input and 16 for
// the feed-forward output.  attention_kernel needs 127 registers (q[32], Wa[32], v[16] live in the T x T loop): 15 warps per
// SM, issue slots 55 % busy, MUFU 36 % (ncu, profiles/r02e) -- latency-bound.  Half the state per thread fits 64 registers:
// twice the warps for the same work.
// ncu source view of this kernel (transformer launch, 4096 windows, round 2 capture r02i): 43 % of the stall samples in the T x T loop --
// 18 MUFU per 87 instructions, and tools/mufu_probe.cu measures 16 MUFU lanes / clk / SM on B200, so that loop is MUFU-bound (36 XU
// cycles against 22 issue cycles per warp iteration) -- and 36 % in the feed-forward loop (two wavefronts per LDS.128: the lanes of a
// pair read different weights).  Measured and left out (2.25 vs 2.24 ms per station-day, bit-compatible): the feed-forward layer on its
// own thread mapping (thread = query x half of the hidden units, one weight address per warp, LayerNorm rows and partial outputs staged
// through the dead K / x tiles), together with two reciprocals from one MUFU.RCP (1/x = y rcp(xy), 24 instead of 15.7 elements / clk /
// SM in the probe) behind a CTA-wide vote that every exponent is within +-31 so that xy stays normal.  The unchanged time says the
// vote does not pass on the real weights (not verified further); a clamp per element instead would cost the issue slots the MUFU saves.
constexpr int A2_TPW = 2 * AT_MAXT;      // threads per window
constexpr int A2_NT = AT_WPC * A2_TPW;   // 480

__global__ void __launch_bounds__(A2_NT, 2) attention2_kernel(const AttnP p) {
    extern __shared__ __align__(16) float at_smem[];
    float *wsm = at_smem;
    float(*Ks)[AT_MAXT * AT_KP] = reinterpret_cast<float(*)[AT_MAXT * AT_KP]>(at_smem + AT_OFF_K);
    float(*Xs)[AT_MAXT * AT_XP] = reinterpret_cast<float(*)[AT_MAXT * AT_XP]>(at_smem + AT_OFF_X);

    const int tid = threadIdx.x;
    const int g = blockIdx.y;
    const int T = p.T;
    {
        const float *src = p.w + (int64_t)g * p.w_gs;
        const int n = (p.mode == 0) ? AW_SIZE : AW_G1;
        for (int i = tid; i < n; i += A2_NT) {
            const float v = __ldg(src + i);
            wsm[i] = (i >= AW_WA && i < AW_WA + 32) ? -2.f * v : v;  // Wa is only used as -2 Wa
        }
    }
    const int wl = tid / A2_TPW;
    const int r = tid - wl * A2_TPW;
    const int i = r >> 1;    // query time step
    const int hs = r & 1;    // which half of the units / channels / hidden units
    const int64_t b = (int64_t)blockIdx.x * AT_WPC + wl;
    const bool bval = b < p.B;
    const float *xb = p.x + (int64_t)g * p.x_gs + (bval ? b : 0) * p.x_bs;
    for (int idx = r; idx < 16 * T; idx += A2_TPW) {
        const int c = idx / T, t = idx - c * T;
        Xs[wl][t * AT_XP + c] = bval ? __ldg(xb + idx) : 0.f;
    }
    __syncthreads();

    const bool act = bval && i < T;
    const float *xr = &Xs[wl][(i < T ? i : 0) * AT_XP];
    constexpr float kTwoLog2e = 2.8853900817779268f;
    float q[16];
    {
        float kv[16];
        const float4 *bh4 = reinterpret_cast<const float4 *>(wsm + AW_BH + hs * 16);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
            const float4 t4 = bh4[u4];
            q[u4 * 4 + 0] = t4.x, q[u4 * 4 + 1] = t4.y, q[u4 * 4 + 2] = t4.z, q[u4 * 4 + 3] = t4.w;
            kv[u4 * 4 + 0] = kv[u4 * 4 + 1] = kv[u4 * 4 + 2] = kv[u4 * 4 + 3] = 0.f;
        }
#pragma unroll 4
        for (int c = 0; c < 16; ++c) {
            const float xc = act ? xr[c] : 0.f;
            const float4 *wt4 = reinterpret_cast<const float4 *>(wsm + AW_WT + c * 32 + hs * 16);
            const float4 *wx4 = reinterpret_cast<const float4 *>(wsm + AW_WX + c * 32 + hs * 16);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const float4 a = wt4[u4], k = wx4[u4];
                q[u4 * 4 + 0] = fmaf(xc, a.x, q[u4 * 4 + 0]), q[u4 * 4 + 1] = fmaf(xc, a.y, q[u4 * 4 + 1]);
                q[u4 * 4 + 2] = fmaf(xc, a.z, q[u4 * 4 + 2]), q[u4 * 4 + 3] = fmaf(xc, a.w, q[u4 * 4 + 3]);
                kv[u4 * 4 + 0] = fmaf(xc, k.x, kv[u4 * 4 + 0]), kv[u4 * 4 + 1] = fmaf(xc, k.y, kv[u4 * 4 + 1]);
                kv[u4 * 4 + 2] = fmaf(xc, k.z, kv[u4 * 4 + 2]), kv[u4 * 4 + 3] = fmaf(xc, k.w, kv[u4 * 4 + 3]);
            }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) q[u] = ex2_approx(fminf(fmaxf(kTwoLog2e * q[u], -63.f), 63.f));
        if (i < T) {
            float4 *kd = reinterpret_cast<float4 *>(&Ks[wl][i * AT_KP + hs * 16]);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4)
                kd[u4] = make_float4(ex2_approx(fminf(fmaxf(kTwoLog2e * kv[u4 * 4 + 0], -63.f), 63.f)),
                                     ex2_approx(fminf(fmaxf(kTwoLog2e * kv[u4 * 4 + 1], -63.f), 63.f)),
                                     ex2_approx(fminf(fmaxf(kTwoLog2e * kv[u4 * 4 + 2], -63.f), 63.f)),
                                     ex2_approx(fminf(fmaxf(kTwoLog2e * kv[u4 * 4 + 3], -63.f), 63.f)));
        }
    }
    __syncthreads();

    float e0 = wsm[AW_BA];
#pragma unroll 8
    for (int u = 0; u < 32; ++u) e0 = fmaf(-0.5f, wsm[AW_WA + u], e0);  // ba + sum_u Wa_u
    float wa[16];
    {
        const float4 *wa4 = reinterpret_cast<const float4 *>(wsm + AW_WA + hs * 16);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
            const float4 t4 = wa4[u4];
            wa[u4 * 4 + 0] = t4.x, wa[u4 * 4 + 1] = t4.y, wa[u4 * 4 + 2] = t4.z, wa[u4 * 4 + 3] = t4.w;
        }
    }
    float emax = -INFINITY, ssum = 0.f;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = 0.f;
    const int half = p.width / 2;
    for (int jj = 0; jj < T; ++jj) {
        const float4 *kr = reinterpret_cast<const float4 *>(&Ks[wl][jj * AT_KP + hs * 16]);
        float ea = 0.f, eb = 0.f;
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
            const float4 k4 = kr[u4];
            ea = fmaf(wa[u4 * 4 + 0], rcp_approx(fmaf(q[u4 * 4 + 0], k4.x, 1.f)), ea);
            eb = fmaf(wa[u4 * 4 + 1], rcp_approx(fmaf(q[u4 * 4 + 1], k4.y, 1.f)), eb);
            ea = fmaf(wa[u4 * 4 + 2], rcp_approx(fmaf(q[u4 * 4 + 2], k4.z, 1.f)), ea);
            eb = fmaf(wa[u4 * 4 + 3], rcp_approx(fmaf(q[u4 * 4 + 3], k4.w, 1.f)), eb);
        }
        const float mine = ea + eb;
        const float e = (mine + __shfl_xor_sync(0xffffffffu, mine, 1)) + e0;  // the same bits in both lanes of the pair
        if (e > emax) {
            const float sc = ex2_approx(1.4426950408889634f * (emax - e));
            ssum *= sc;
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] *= sc;
            emax = e;
        }
        bool in_band = true;
        if (p.width > 0) {
            const int lower = jj - half;
            in_band = lower <= i && i < lower + p.width;
        }
        if (in_band) {
            const float w = ex2_approx(1.4426950408889634f * (e - emax));
            ssum += w;
            const float4 *xj = reinterpret_cast<const float4 *>(&Xs[wl][jj * AT_XP + hs * 8]);
            const float4 x0 = xj[0], x1 = xj[1];
            v[0] = fmaf(w, x0.x, v[0]), v[1] = fmaf(w, x0.y, v[1]), v[2] = fmaf(w, x0.z, v[2]), v[3] = fmaf(w, x0.w, v[3]);
            v[4] = fmaf(w, x1.x, v[4]), v[5] = fmaf(w, x1.y, v[5]), v[6] = fmaf(w, x1.z, v[6]), v[7] = fmaf(w, x1.w, v[7]);
        }
    }
    const float inv = 1.f / (ssum + 1e-5f);
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] *= inv;

    float *yb = p.y + (int64_t)g * p.y_gs + (bval ? b : 0) * p.y_bs + (int64_t)(hs * 8) * T + i;
    if (p.mode == 1) {
        if (act) {
#pragma unroll
            for (int c = 0; c < 8; ++c) yb[c * T] = v[c];
        }
        return;
    }
    // transformer tail: y = LN(x + attn); out = LN(y + lin2(relu(lin1(y)))); LayerNorm sums over the pair
    float y[16];  // [0, 8): this lane's channels, [8, 16): the partner's
    {
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            y[c] = (act ? xr[hs * 8 + c] : 0.f) + v[c];
            sum += y[c];
        }
        const float mean = (sum + __shfl_xor_sync(0xffffffffu, sum, 1)) * (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float d = y[c] - mean;
            var = fmaf(d, d, var);
        }
        var = (var + __shfl_xor_sync(0xffffffffu, var, 1)) * (1.f / 16.f) + 1e-14f;
        const float sd = sqrtf(var);
#pragma unroll
        for (int c = 0; c < 8; ++c) y[c] = (y[c] - mean) / sd * wsm[AW_G1 + hs * 8 + c] + wsm[AW_B1 + hs * 8 + c];
#pragma unroll
        for (int c = 0; c < 8; ++c) y[8 + c] = __shfl_xor_sync(0xffffffffu, y[c], 1);
    }
    float o[16];  // same order as y: [own 8 | partner's 8]
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = 0.f;
    const int own = hs * 8, oth = 8 - own;
#pragma unroll 2
    for (int mm = 0; mm < 64; ++mm) {
        const int m = hs * 64 + mm;
        const float4 *w1o = reinterpret_cast<const float4 *>(&wsm[AW_L1W + m * 16 + own]);
        const float4 *w1p = reinterpret_cast<const float4 *>(&wsm[AW_L1W + m * 16 + oth]);
        float hsum = wsm[AW_L1B + m];
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
            const float4 a = w1o[c4], bq = w1p[c4];
            hsum = fmaf(a.x, y[c4 * 4 + 0], hsum), hsum = fmaf(a.y, y[c4 * 4 + 1], hsum);
            hsum = fmaf(a.z, y[c4 * 4 + 2], hsum), hsum = fmaf(a.w, y[c4 * 4 + 3], hsum);
            hsum = fmaf(bq.x, y[8 + c4 * 4 + 0], hsum), hsum = fmaf(bq.y, y[8 + c4 * 4 + 1], hsum);
            hsum = fmaf(bq.z, y[8 + c4 * 4 + 2], hsum), hsum = fmaf(bq.w, y[8 + c4 * 4 + 3], hsum);
        }
        hsum = fmaxf(hsum, 0.f);
        const float4 *w2o = reinterpret_cast<const float4 *>(&wsm[AW_L2W + m * 16 + own]);
        const float4 *w2p = reinterpret_cast<const float4 *>(&wsm[AW_L2W + m * 16 + oth]);
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
            const float4 a = w2o[c4], bq = w2p[c4];
            o[c4 * 4 + 0] = fmaf(a.x, hsum, o[c4 * 4 + 0]), o[c4 * 4 + 1] = fmaf(a.y, hsum, o[c4 * 4 + 1]);
            o[c4 * 4 + 2] = fmaf(a.z, hsum, o[c4 * 4 + 2]), o[c4 * 4 + 3] = fmaf(a.w, hsum, o[c4 * 4 + 3]);
            o[8 + c4 * 4 + 0] = fmaf(bq.x, hsum, o[8 + c4 * 4 + 0]), o[8 + c4 * 4 + 1] = fmaf(bq.y, hsum, o[8 + c4 * 4 + 1]);
            o[8 + c4 * 4 + 2] = fmaf(bq.z, hsum, o[8 + c4 * 4 + 2]), o[8 + c4 * 4 + 3] = fmaf(bq.w, hsum, o[8 + c4 * 4 + 3]);
        }
    }
    {
        // this lane's 8 channels: own partial + the partner's partial for them (its "partner's 8")
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float theirs = __shfl_xor_sync(0xffffffffu, o[8 + c], 1);
            o[c] = ((o[c] + theirs) + wsm[AW_L2B + own + c]) + y[c];
            sum += o[c];
        }
        const float mean = (sum + __shfl_xor_sync(0xffffffffu, sum, 1)) * (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float d = o[c] - mean;
            var = fmaf(d, d, var);
        }
        var = (var + __shfl_xor_sync(0xffffffffu, var, 1)) * (1.f / 16.f) + 1e-14f;
        const float sd = sqrtf(var);
        if (act) {
#pragma unroll
            for (int c = 0; c < 8; ++c) yb[c * T] = (o[c] - mean) / sd * wsm[AW_G2 + own + c] + wsm[AW_B2 + own + c];
        }
    }
}

int launch_attention(const AttnP &p, int G, cudaStream_t s) {
    if (p.T > AT_MAXT) {
        set_error("attention kernel supports T <= %d (got %d)", AT_MAXT, p.T);
        return VP_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)((p.B + AT_WPC - 1) / AT_WPC), G);
    constexpr size_t smem = AT_SMEM_FLOATS * sizeof(float);
    const char *v1 = getenv("VP_ATTN_V1");  // A/B aid: one thread per query (read per launch)
    if (v1 && atoi(v1) != 0) {
        if (int rc = ensure_dyn_smem((const void *)attention_kernel, smem)) return rc;
        KTimer kt(KC_ATTN, s);
        attention_kernel<<<grid, AT_NT, smem, s>>>(p);
        VP_LAUNCH_CHECK();
        return VP_OK;
    }
    if (int rc = ensure_dyn_smem((const void *)attention2_kernel, smem)) return rc;
    KTimer kt(KC_ATTN, s);
    attention2_kernel<<<grid, A2_NT, smem, s>>>(p);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

}  // namespace vp
