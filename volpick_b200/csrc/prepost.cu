// HBM-bound kernels on either side of the network:
//   K1  window slicer + demean + peak normalisation (+ EQTransformer cosine taper)
//       replaces WaveformModel._cut_fragments_array + annotate_batch_pre      
//   K9  overlap-add stacking with blinding ("avg" = np.nanmean order, "max" = np.nanmax)
//       replaces annotate_batch_post + WaveformModel._reassemble_blocks_array
//   K10 hysteresis trigger + first-argmax peak extraction
//       replaces obspy trigger_onset + np.argmax as restated in the reference at
//       /root/reference/volpick/model/eval_taks0.py:46-56                     
//   plus the NaN-trim bounds of WaveformModel._trim_nan                       
#include <math_constants.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

// ------------------------------------------------------------------------------------------ K1
struct Taper {
    float t[6];
};
// 1 / sum_i (i - (L - 1) / 2)^2 = 12 / (L (L^2 - 1)): the denominator of the least-squares slope (norm_detrend)
static inline float detrend_inv_tt(int L) { return L > 1 ? (float)(12.0 / ((double)L * ((double)L * (double)L - 1.0))) : 0.f; }

template <typename T>
__device__ __forceinline__ float ld_as_float(const T *p) {
    return (float)__ldg(p);
}

__device__ __forceinline__ float block_reduce_sum3(float3 &v, float *red /*[3*32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    }
    __syncthreads();  // protect red[] reuse
    if (lane == 0) {
        red[warp] = v.x;
        red[32 + warp] = v.y;
        red[64 + warp] = v.z;
    }
    __syncthreads();
    float3 r = make_float3(0.f, 0.f, 0.f);
    for (int w = 0; w < nw; ++w) {  // fixed order: deterministic
        r.x += red[w];
        r.y += red[32 + w];
        r.z += red[64 + w];
    }
    v = r;
    return 0.f;
}

// Window means in double: counts with a large DC offset (1e6 counts, +-100 of signal) lose the signal's low bits in an fp32
// sum; SeisBench 0.4 takes the mean in NumPy's float64 (int32 / float64 traces), torch's CPU mean accumulates in double.
__device__ __forceinline__ void block_reduce_sum3d(double (&v)[3], double *red /*[3*32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    }
    __syncthreads();  // protect red[] reuse
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) red[32 * c + warp] = v[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double r = 0.0;
        for (int w = 0; w < nw; ++w) r += red[32 * c + w];  // fixed order: deterministic
        v[c] = r;
    }
}

__device__ __forceinline__ void block_reduce_max3(float3 &v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x = fmaxf(v.x, __shfl_xor_sync(0xffffffffu, v.x, o));
        v.y = fmaxf(v.y, __shfl_xor_sync(0xffffffffu, v.y, o));
        v.z = fmaxf(v.z, __shfl_xor_sync(0xffffffffu, v.z, o));
    }
    __syncthreads();
    if (lane == 0) {
        red[warp] = v.x;
        red[32 + warp] = v.y;
        red[64 + warp] = v.z;
    }
    __syncthreads();
    float3 r = make_float3(0.f, 0.f, 0.f);
    for (int w = 0; w < nw; ++w) {
        r.x = fmaxf(r.x, red[w]);
        r.y = fmaxf(r.y, red[32 + w]);
        r.z = fmaxf(r.z, red[64 + w]);
    }
    v = r;
}

// One CTA (NT threads) per window, all three channels; each thread keeps PT samples per channel in
// registers so the trace is read once (coalesced 128 B per warp request) and the window written once.
template <typename Tin, int PT, int NT>
__global__ void __launch_bounds__(NT) slice_normalize_kernel(const Tin *__restrict__ trace, int64_t ch_stride,
                                                             const int64_t *__restrict__ starts, int L,
                                                             int peak_scope, int flags, float inv_tt, Taper tap,
                                                             float *__restrict__ out) {
    __shared__ double redd[96];
    float *red = reinterpret_cast<float *>(redd);
    const int64_t w = blockIdx.x;
    const int64_t s = __ldg(starts + w);
    const int tid = threadIdx.x;
    float v[3][PT];
    double sum[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int idx = tid + j * NT;
        const bool ok = idx < L;
        v[0][j] = ok ? ld_as_float(trace + s + idx) : 0.f;
        v[1][j] = ok ? ld_as_float(trace + ch_stride + s + idx) : 0.f;
        v[2][j] = ok ? ld_as_float(trace + 2 * ch_stride + s + idx) : 0.f;
        sum[0] += (double)v[0][j];
        sum[1] += (double)v[1][j];
        sum[2] += (double)v[2][j];
    }
    block_reduce_sum3d(sum, redd);
    const float3 mean = make_float3((float)(sum[0] / (double)L), (float)(sum[1] / (double)L), (float)(sum[2] / (double)L));
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        v[0][j] -= mean.x;
        v[1][j] -= mean.y;
        v[2][j] -= mean.z;
    }
    if (flags & VP_PRE_DETREND) {  // norm_detrend: least-squares line of the demeaned window (scipy.signal.detrend)
        const float c0 = 0.5f * (float)(L - 1);
        double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int idx = tid + j * NT;
            if (idx < L) {
                const double t = (double)((float)idx - c0);
                sx += t * (double)v[0][j];
                sy += t * (double)v[1][j];
                sz += t * (double)v[2][j];
            }
        }
        float3 st = make_float3((float)sx, (float)sy, (float)sz);
        block_reduce_sum3(st, red);
        const float3 beta = make_float3(st.x * inv_tt, st.y * inv_tt, st.z * inv_tt);
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const float t = (float)(tid + j * NT) - c0;
            v[0][j] -= beta.x * t;
            v[1][j] -= beta.y * t;
            v[2][j] -= beta.z * t;
        }
    }
    float3 pk = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int idx = tid + j * NT;
        if (idx < L) {
            pk.x = fmaxf(pk.x, fabsf(v[0][j]));
            pk.y = fmaxf(pk.y, fabsf(v[1][j]));
            pk.z = fmaxf(pk.z, fabsf(v[2][j]));
        }
    }
    block_reduce_max3(pk, red);
    if (peak_scope == VP_PEAK_PER_WINDOW) {
        const float m = fmaxf(pk.x, fmaxf(pk.y, pk.z));
        pk = make_float3(m, m, m);
    }
    const float3 den = make_float3(pk.x + 1e-10f, pk.y + 1e-10f, pk.z + 1e-10f);
    float *ob = out + w * 3 * (int64_t)L;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int idx = tid + j * NT;
        if (idx < L) {
            float tp = 1.f;
            if (flags & VP_PRE_TAPER) {
                if (idx < 6) tp = tap.t[idx];
                if (idx >= L - 6) tp = tap.t[L - 1 - idx];
            }
            ob[idx] = (v[0][j] / den.x) * tp;
            ob[(int64_t)L + idx] = (v[1][j] / den.y) * tp;
            ob[2 * (int64_t)L + idx] = (v[2][j] / den.z) * tp;
        }
    }
}


// ------------------------------------------------------------------------------------------ K1, TMA-staged
// The stand-alone slicer as a persistent kernel: a window's three components are staged in shared memory by three bulk
// copies (cp.async.bulk, UBLKCP) issued by one thread two windows ahead, so the 72 KB of a window are in flight while the
// CTA normalises the previous one; the normalised window leaves as 16-byte stores.  (One CTA per window with 36 scalar
// loads per thread reached 27 % of the HBM peak: every window paid load latency -> two block reductions -> store in series.)
// A component's window starts at an arbitrary sample: the copy starts at the 16-byte boundary below it and the samples are
// read at the remaining offset.  Arithmetic (summation order included) = slice_normalize_kernel's: same results bit for bit.
constexpr int K1T_NT = 512, K1T_STAGES = 2;

template <typename Tin>
__global__ void __launch_bounds__(K1T_NT, 1) slice_tma_kernel(const Tin *__restrict__ trace, int64_t ch_stride, int64_t n_alloc,
                                                              const int64_t *__restrict__ starts, int64_t n_windows, int L, int pitch,
                                                              int peak_scope, int flags, float inv_tt, Taper tap, float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t k1_smem[];
    __shared__ __align__(8) uint64_t full[K1T_STAGES], empty[K1T_STAGES];
    __shared__ double redd[96];
    __shared__ int s_off[K1T_STAGES][3];
    float *red = reinterpret_cast<float *>(redd);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < K1T_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], K1T_NT);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t sbase = smem_u32(k1_smem);
    const int64_t first = blockIdx.x, step = gridDim.x;

    auto issue = [&](int64_t w, int st) {  // thread 0: the three components of window w -> stage st
        const int64_t s0 = __ldg(starts + w);
        uint32_t bytes = 0;
        int offs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t a = c * ch_stride + s0, a_al = a & ~(int64_t)3;
            offs[c] = (int)(a - a_al);
            int64_t n_el = (offs[c] + L + 3) & ~3;
            if (a_al + n_el > n_alloc) n_el = (n_alloc - a_al) & ~(int64_t)3;  // never read past the record: the tail is loaded below
            s_off[st][c] = offs[c] | ((int)n_el << 2);
            bytes += (uint32_t)n_el * 4u;
        }
        mbar_arrive_expect_tx(&full[st], bytes);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t a_al = (c * ch_stride + s0) & ~(int64_t)3;
            const uint32_t n_el = (uint32_t)(s_off[st][c] >> 2);
            if (n_el) bulk_load(sbase + (uint32_t)((st * 3 + c) * pitch) * 4u, trace + a_al, n_el * 4u, &full[st]);
        }
    };

    int n = 0;
    if (tid == 0) {
        for (int k = 0; k < K1T_STAGES; ++k)
            if (first + k * step < n_windows) issue(first + k * step, k);
    }
    for (int64_t w = first; w < n_windows; w += step, ++n) {
        const int st = n % K1T_STAGES;
        const uint32_t ph = (uint32_t)(n / K1T_STAGES) & 1u;
        mbar_wait(&full[st], ph);
        const int64_t s0 = __ldg(starts + w);
        const Tin *xs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int so = s_off[st][c];
            const int off = so & 3, n_el = so >> 2;
            Tin *base = reinterpret_cast<Tin *>(k1_smem) + (size_t)(st * 3 + c) * pitch;
            // samples the bulk copy could not bring without reading past the record end (last window of the last component only)
            for (int i = n_el - off + tid; i < L; i += K1T_NT)
                if (i >= 0) base[off + i] = __ldg(trace + c * ch_stride + s0 + i);
            xs[c] = base + off;
        }
        __syncthreads();
        double sum[3] = {0.0, 0.0, 0.0};
        constexpr int PT = 12;  // the per-thread sample order of slice_normalize_kernel<Tin, 12, 512>
        float v[3][PT];
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int idx = tid + j * K1T_NT;
            const bool ok = idx < L;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c][j] = ok ? (float)xs[c][idx] : 0.f;
                sum[c] += (double)v[c][j];
            }
        }
        // the stage is free as soon as every thread has its samples in registers: release it and refill it two windows ahead
        mbar_arrive(&empty[st]);
        if (tid == 0 && w + K1T_STAGES * step < n_windows) {
            mbar_wait(&empty[st], ph);
            issue(w + K1T_STAGES * step, st);
        }
        block_reduce_sum3d(sum, redd);
        const float3 mean = make_float3((float)(sum[0] / (double)L), (float)(sum[1] / (double)L), (float)(sum[2] / (double)L));
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            v[0][j] -= mean.x;
            v[1][j] -= mean.y;
            v[2][j] -= mean.z;
        }
        if (flags & VP_PRE_DETREND) {
            const float c0 = 0.5f * (float)(L - 1);
            double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
            for (int j = 0; j < PT; ++j) {
                const int idx = tid + j * K1T_NT;
                if (idx < L) {
                    const double t = (double)((float)idx - c0);
                    sx += t * (double)v[0][j];
                    sy += t * (double)v[1][j];
                    sz += t * (double)v[2][j];
                }
            }
            float3 stt = make_float3((float)sx, (float)sy, (float)sz);
            block_reduce_sum3(stt, red);
            const float3 beta = make_float3(stt.x * inv_tt, stt.y * inv_tt, stt.z * inv_tt);
#pragma unroll
            for (int j = 0; j < PT; ++j) {
                const float t = (float)(tid + j * K1T_NT) - c0;
                v[0][j] -= beta.x * t;
                v[1][j] -= beta.y * t;
                v[2][j] -= beta.z * t;
            }
        }
        float3 pk = make_float3(0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            if (tid + j * K1T_NT < L) {
                pk.x = fmaxf(pk.x, fabsf(v[0][j]));
                pk.y = fmaxf(pk.y, fabsf(v[1][j]));
                pk.z = fmaxf(pk.z, fabsf(v[2][j]));
            }
        }
        block_reduce_max3(pk, red);
        if (peak_scope == VP_PEAK_PER_WINDOW) {
            const float m = fmaxf(pk.x, fmaxf(pk.y, pk.z));
            pk = make_float3(m, m, m);
        }
        const float3 den = make_float3(pk.x + 1e-10f, pk.y + 1e-10f, pk.z + 1e-10f);
        float *ob = out + w * 3 * (int64_t)L;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const int idx = tid + j * K1T_NT;
            if (idx < L) {
                float tp = 1.f;
                if (flags & VP_PRE_TAPER) {
                    if (idx < 6) tp = tap.t[idx];
                    if (idx >= L - 6) tp = tap.t[L - 1 - idx];
                }
                ob[idx] = (v[0][j] / den.x) * tp;
                ob[(int64_t)L + idx] = (v[1][j] / den.y) * tp;
                ob[2 * (int64_t)L + idx] = (v[2][j] / den.z) * tp;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ K1 + encoder.convs.0
// Tensor-core path of the EQTransformer: the window slicer / normaliser fused with the first encoder stage
// (Conv1d 3 -> 8, k = 11, 'same') + ReLU + MaxPool1d(2), written as the 16-bit channel-last operand of
// encoder.convs.1 ([split][window][L / 2][8]; fp16 hi/lo or bf16).  With 3 input channels the implicit GEMM
// keeps 5 % of the tensor pipe busy (K padded 3 -> 8, N padded 8 -> 16) and needs a 32-byte-per-sample packed
// copy of the windows first; on the CUDA cores the stage is 264 fp32 FMAs per sample straight from the
// normalised window in shared memory, and the fp32 windows never reach HBM.
// Normalisation arithmetic (summation order included) is that of slice_normalize_kernel<Tin, 12, 512>.
struct Enc0W {
    float w[8 * 3 * 11];  // [co][ci][k]
    float b[8];
};
constexpr int E0_NT = 512, E0_PT = 12, E0_HALO = 5;

// POOL = true: EQTransformer (k = 11, ReLU, MaxPool1d(2)); POOL = false: PhaseNet `inc` (k = 7, folded BatchNorm, ReLU),
// one thread per output sample.
template <typename Tin, int SPLIT, bool POOL>
__global__ void __launch_bounds__(E0_NT, 2) slice_enc0_kernel(const Tin *__restrict__ trace, int64_t ch_stride,
                                                              const int64_t *__restrict__ starts, int L, int peak_scope,
                                                              int flags, float inv_tt, Taper tap, const __grid_constant__ Enc0W wt,
                                                              uint16_t *__restrict__ out, int64_t out_split, int out_pitch) {
    extern __shared__ __align__(16) float e0_smem[];
    __shared__ double redd[96];
    float *red = reinterpret_cast<float *>(redd);
    const int XP = L + 12;  // xs[c][i + 5] = normalised sample i; zero halo on both sides
    float *xs = e0_smem;
    const int64_t w = blockIdx.x;
    const int64_t s = __ldg(starts + w);
    const int tid = threadIdx.x;
    double sum[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < E0_PT; ++j) {
        const int idx = tid + j * E0_NT;
        const bool ok = idx < L;
        const float v0 = ok ? ld_as_float(trace + s + idx) : 0.f;
        const float v1 = ok ? ld_as_float(trace + ch_stride + s + idx) : 0.f;
        const float v2 = ok ? ld_as_float(trace + 2 * ch_stride + s + idx) : 0.f;
        sum[0] += (double)v0;
        sum[1] += (double)v1;
        sum[2] += (double)v2;
        if (ok) {
            xs[idx + E0_HALO] = v0;
            xs[XP + idx + E0_HALO] = v1;
            xs[2 * XP + idx + E0_HALO] = v2;
        }
    }
    block_reduce_sum3d(sum, redd);
    const float3 mean = make_float3((float)(sum[0] / (double)L), (float)(sum[1] / (double)L), (float)(sum[2] / (double)L));
    float3 beta = make_float3(0.f, 0.f, 0.f);
    const float c0 = 0.5f * (float)(L - 1);
    if (flags & VP_PRE_DETREND) {  // same arithmetic as slice_normalize_kernel
        double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
        for (int j = 0; j < E0_PT; ++j) {
            const int idx = tid + j * E0_NT;
            if (idx < L) {
                const double t = (double)((float)idx - c0);
                sx += t * (double)(xs[idx + E0_HALO] - mean.x);
                sy += t * (double)(xs[XP + idx + E0_HALO] - mean.y);
                sz += t * (double)(xs[2 * XP + idx + E0_HALO] - mean.z);
            }
        }
        float3 st = make_float3((float)sx, (float)sy, (float)sz);
        block_reduce_sum3(st, red);
        beta = make_float3(st.x * inv_tt, st.y * inv_tt, st.z * inv_tt);
    }
    float3 pk = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < E0_PT; ++j) {
        const int idx = tid + j * E0_NT;
        if (idx < L) {  // each thread re-reads exactly what it wrote
            float v0 = xs[idx + E0_HALO] - mean.x, v1 = xs[XP + idx + E0_HALO] - mean.y, v2 = xs[2 * XP + idx + E0_HALO] - mean.z;
            if (flags & VP_PRE_DETREND) {
                const float t = (float)idx - c0;
                v0 -= beta.x * t;
                v1 -= beta.y * t;
                v2 -= beta.z * t;
            }
            xs[idx + E0_HALO] = v0;
            xs[XP + idx + E0_HALO] = v1;
            xs[2 * XP + idx + E0_HALO] = v2;
            pk.x = fmaxf(pk.x, fabsf(v0));
            pk.y = fmaxf(pk.y, fabsf(v1));
            pk.z = fmaxf(pk.z, fabsf(v2));
        }
    }
    block_reduce_max3(pk, red);
    if (peak_scope == VP_PEAK_PER_WINDOW) {
        const float m = fmaxf(pk.x, fmaxf(pk.y, pk.z));
        pk = make_float3(m, m, m);
    }
    const float3 den = make_float3(pk.x + 1e-10f, pk.y + 1e-10f, pk.z + 1e-10f);
#pragma unroll
    for (int j = 0; j < E0_PT; ++j) {
        const int idx = tid + j * E0_NT;
        if (idx < L) {
            float tp = 1.f;
            if (flags & VP_PRE_TAPER) {
                if (idx < 6) tp = tap.t[idx];
                if (idx >= L - 6) tp = tap.t[L - 1 - idx];
            }
            xs[idx + E0_HALO] = (xs[idx + E0_HALO] / den.x) * tp;
            xs[XP + idx + E0_HALO] = (xs[XP + idx + E0_HALO] / den.y) * tp;
            xs[2 * XP + idx + E0_HALO] = (xs[2 * XP + idx + E0_HALO] / den.z) * tp;
        }
    }
    if (tid < 3 * 12) {  // zero padding of the 'same' conv: 5 samples before, 7 after (one spare for the float2 reads)
        const int c = tid / 12, q = tid - c * 12;
        xs[c * XP + (q < E0_HALO ? q : L + q)] = 0.f;
    }
    __syncthreads();

    if constexpr (!POOL) {
        // PhaseNet: conv (3 -> 8, k = 7, 'same') + ReLU, weights as [co][ci][7]
        for (int p = tid; p < L; p += E0_NT) {
            float a0[8];
#pragma unroll
            for (int co = 0; co < 8; ++co) a0[co] = wt.b[co];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const float x = xs[c * XP + p + E0_HALO - 3 + q];
#pragma unroll
                    for (int co = 0; co < 8; ++co) a0[co] = fmaf(x, wt.w[(co * 3 + c) * 7 + q], a0[co]);
                }
            float v[8];
#pragma unroll
            for (int co = 0; co < 8; ++co) v[co] = fmaxf(a0[co], 0.f);
            uint4 hi, lo;
            pack8_split16<SPLIT>(v, hi, lo);
            uint16_t *yb = out + (w * out_pitch + p) * 8;
            *reinterpret_cast<uint4 *>(yb) = hi;
            if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + out_split) = lo;
        }
        for (int p = L + tid; p < out_pitch; p += E0_NT) {  // rows past the window inside the output pitch: conv zero padding
            uint16_t *yb = out + (w * out_pitch + p) * 8;
            *reinterpret_cast<uint4 *>(yb) = make_uint4(0u, 0u, 0u, 0u);
            if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + out_split) = make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }
    // conv + ReLU + pool: one thread = one pooled output sample p (conv outputs 2p, 2p + 1), all 8 channels
    const int LP = L >> 1;
    for (int p = tid; p < LP; p += E0_NT) {
        float in[3][12];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const float2 v2 = *reinterpret_cast<const float2 *>(xs + c * XP + 2 * p + 2 * q);
                in[c][2 * q] = v2.x;
                in[c][2 * q + 1] = v2.y;
            }
        float a0[8], a1[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) a0[co] = a1[co] = wt.b[co];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const float x = in[c][q];
#pragma unroll
                for (int co = 0; co < 8; ++co) {
                    if (q < 11) a0[co] = fmaf(x, wt.w[(co * 3 + c) * 11 + q], a0[co]);
                    if (q >= 1) a1[co] = fmaf(x, wt.w[(co * 3 + c) * 11 + q - 1], a1[co]);
                }
            }
        float v[8];
#pragma unroll
        for (int co = 0; co < 8; ++co) v[co] = fmaxf(fmaxf(a0[co], a1[co]), 0.f);
        uint4 hi, lo;
        pack8_split16<SPLIT>(v, hi, lo);
        uint16_t *yb = out + (w * LP + p) * 8;
        *reinterpret_cast<uint4 *>(yb) = hi;
        if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + out_split) = lo;
    }
}

int launch_slice_enc0(const void *trace, int dtype, int64_t ch_stride, const int64_t *starts, int64_t nw, int L, int scope,
                      int taper, const float *w_host /*(8,3,k)*/, const float *b_host /*(8)*/, int split, uint16_t *out,
                      int64_t out_split, cudaStream_t s, int k, int out_pitch) {
    VP_REQUIRE((k == 11 && L % 2 == 0 || k == 7) && L <= E0_PT * E0_NT, VP_ERR_UNSUPPORTED,
               "fused slicer + first conv: window length %d / kernel size %d unsupported", L, k);
    if (nw == 0) return VP_OK;
    Taper tap;
    for (int i = 0; i < 6; ++i) {
        const double a = M_PI + (M_PI * i) / 5.0;
        tap.t[i] = (float)(0.5 * (1.0 + std::cos(a)));
    }
    const float inv_tt = detrend_inv_tt(L);
    Enc0W wt;
    std::memset(&wt, 0, sizeof(wt));
    std::memcpy(wt.w, w_host, sizeof(float) * 8 * 3 * k);
    std::memcpy(wt.b, b_host, sizeof(wt.b));
    const size_t smem = (size_t)3 * (L + 12) * sizeof(float);
#define VP_E0_LAUNCH(T, S)                                                                                                  \
    do {                                                                                                                    \
        auto kern = k == 11 ? slice_enc0_kernel<T, S, true> : slice_enc0_kernel<T, S, false>;                               \
        if (int rc = ensure_dyn_smem((const void *)kern, (size_t)3 * (E0_PT * E0_NT + 12) * 4)) return rc;                  \
        kern<<<(unsigned)nw, E0_NT, smem, s>>>((const T *)trace, ch_stride, starts, L, scope, taper, inv_tt, tap, wt, out, out_split, \
                                               out_pitch > 0 ? out_pitch : L);                                              \
    } while (0)
    KTimer kt(KC_SLICE_ENC0, s);
    if (dtype == VP_DTYPE_F32 && split == 2) VP_E0_LAUNCH(float, 2);
    else if (dtype == VP_DTYPE_F32) VP_E0_LAUNCH(float, 1);
    else if (split == 2) VP_E0_LAUNCH(int32_t, 2);
    else VP_E0_LAUNCH(int32_t, 1);
#undef VP_E0_LAUNCH
    VP_LAUNCH_CHECK();
    return VP_OK;
}

template <typename Tin>
static int launch_slice(const Tin *trace, int64_t ch_stride, int64_t n_alloc, const int64_t *starts, int64_t nw, int L, int scope,
                        int taper, float *out, cudaStream_t s) {
    Taper tap;
    for (int i = 0; i < 6; ++i) {  // 0.5 * (1 + cos(linspace(pi, 2 pi, 6))) in double, rounded once
        const double a = M_PI + (M_PI * i) / 5.0;
        tap.t[i] = (float)(0.5 * (1.0 + std::cos(a)));
    }
    const float inv_tt = detrend_inv_tt(L);
    KTimer kt(KC_SLICE, s);
    static const bool tma_off = getenv("VP_K1_TMA") && atoi(getenv("VP_K1_TMA")) == 0;  // A/B aid
    if (!tma_off && L > 4096 && L <= 12 * K1T_NT && (reinterpret_cast<uintptr_t>(trace) & 15) == 0) {  // shorter windows: the one-CTA-per-window kernel below is faster (PhaseNet: 0.13 vs 0.19 ms)
        // persistent TMA-staged slicer: 2 stages x 3 components x (L + 8) samples of shared memory, one CTA per SM
        const int pitch = (L + 3 + 3 + 4) & ~3;
        const size_t smem = (size_t)K1T_STAGES * 3 * pitch * 4;
        auto kern = slice_tma_kernel<Tin>;
        if (int rc = ensure_dyn_smem((const void *)kern, smem)) return rc;
        const unsigned grid = (unsigned)std::min<int64_t>(nw, device_sm_count());
        kern<<<grid, K1T_NT, smem, s>>>(trace, ch_stride, n_alloc, starts, nw, L, pitch, scope, taper, inv_tt, tap, out);
        VP_LAUNCH_CHECK();
        return VP_OK;
    }
    if (L <= 6 * 512) {
        slice_normalize_kernel<Tin, 6, 512><<<(unsigned)nw, 512, 0, s>>>(trace, ch_stride, starts, L, scope, taper, inv_tt, tap, out);
    } else if (L <= 12 * 512) {
        slice_normalize_kernel<Tin, 12, 512><<<(unsigned)nw, 512, 0, s>>>(trace, ch_stride, starts, L, scope, taper, inv_tt, tap, out);
    } else if (L <= 24 * 1024) {
        slice_normalize_kernel<Tin, 24, 1024><<<(unsigned)nw, 1024, 0, s>>>(trace, ch_stride, starts, L, scope, taper, inv_tt, tap, out);
    } else {
        set_error("window length %d not supported by the slicer (max %d)", L, 24 * 1024);
        return VP_ERR_UNSUPPORTED;
    }
    VP_LAUNCH_CHECK();
    return VP_OK;
}

// ------------------------------------------------------------------------------------------ K9
constexpr int STACK_MAXCOV = 64;

// NumPy pairwise_sum order for n values (numpy/_core/src/umath/loops_utils.h.src), n <= 128.
template <int MAXN, bool EXACT>
__device__ __forceinline__ float np_pairwise_sum(const float (&a)[MAXN], int n_rt) {
    const int n = EXACT ? MAXN : n_rt;
    if (n < 8) {
        float res = -0.0f;
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
            if (i < n) res += a[i];
        return res;
    }
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = a[i < MAXN ? i : 0];
    const int nblk = n - (n % 8);
#pragma unroll
    for (int i = 8; i + 8 <= MAXN; i += 8) {
        if (i < nblk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        }
    }
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
    for (int i = 8; i < MAXN; ++i)
        if (i >= nblk && i < n) res += a[i];
    return res;
}

// One thread per output sample n (all labels).  Windows covering n form a contiguous index range
// because starts are sorted ascending; slot = i % coverage and later windows overwrite the slot,
// exactly like the reference's NaN-buffer assignment.  EXACT: coverage == COV at compile time
// (3 and 13 are the volpick configurations), so the slot array stays in registers.
template <int COV, bool EXACT>
__global__ void __launch_bounds__(256) stack_kernel(const float *__restrict__ y, const int64_t *__restrict__ starts,
                                                    int64_t nwin, int L, int nlab, int cov_rt, int64_t b0, int64_t b1,
                                                    int mode, float *__restrict__ out, int64_t pred_len) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= pred_len) return;
    const int cov = EXACT ? COV : cov_rt;
    // last window with start <= n (upper_bound - 1); starts is small and L1/L2 resident
    int64_t lo = 0, up = nwin;
    while (lo < up) {
        const int64_t mid = (lo + up) >> 1;
        if (__ldg(starts + mid) <= n) lo = mid + 1; else up = mid;
    }
    const int64_t hi = lo - 1;
    int64_t first = hi + 1;
    while (first > 0 && __ldg(starts + first - 1) + L > n) --first;

    for (int c = 0; c < nlab; ++c) {
        float slot[COV];
#pragma unroll
        for (int k = 0; k < COV; ++k) slot[k] = CUDART_NAN_F;
        for (int64_t i = first; i <= hi; ++i) {
            const int64_t off = n - __ldg(starts + i);
            if (off >= L) continue;
            const float val =
                (off < b0 || off >= L - b1) ? CUDART_NAN_F : __ldg(y + (i * nlab + c) * (int64_t)L + off);
            const int sl = (int)(i % cov);
#pragma unroll
            for (int k = 0; k < COV; ++k)
                if (k == sl) slot[k] = val;
        }
        float res;
        if (mode == VP_STACK_AVG) {
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < COV; ++k) {
                if (k < cov && !isnan(slot[k])) ++cnt; else slot[k] = 0.f;
            }
            const float tot = 0.0f + np_pairwise_sum<COV, EXACT>(slot, cov);
            res = cnt ? __fdiv_rn(tot, (float)cnt) : CUDART_NAN_F;
        } else {
            res = CUDART_NAN_F;
#pragma unroll
            for (int k = 0; k < COV; ++k)
                if (k < cov && !isnan(slot[k]) && (isnan(res) || slot[k] > res)) res = slot[k];
        }
        out[(int64_t)c * pred_len + n] = res;
    }
}

// Vectorised K9 for the compile-time coverages (3, 13): one thread = 4 consecutive output samples, all labels.
//  * window range [first, hi] of the group from a closed-form guess (n / stride) that is then VERIFIED against
//    starts[] and corrected, so any sorted starts array gives the same result as the binary search above;
//  * slots are visited in slot order k = i % COV (consecutive windows map to distinct slots), which is the
//    order np.nanmean sums in; a group covered by more than COV windows (the reference would overwrite slots)
//    takes the generic per-sample path;
//  * 16-byte loads of y when the window offset is 4-aligned (regular strides), scalar loads otherwise (tail window).
template <int COV>
__global__ void __launch_bounds__(256) stack4_kernel(const float *__restrict__ y, const int64_t *__restrict__ starts,
                                                     int nwin, int L, int nlab, int stride, int b0, int b1, int mode,
                                                     float *__restrict__ out, int64_t pred_len) {
    const int64_t n0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (n0 >= pred_len) return;
    const int64_t n3 = (n0 + 3 < pred_len) ? n0 + 3 : pred_len - 1;
    // hi = last window with start <= n3; first = first window with start + L > n0
    int hi = (int)((n3 / stride < nwin - 1) ? n3 / stride : nwin - 1);
    while (hi + 1 < nwin && __ldg(starts + hi + 1) <= n3) ++hi;
    while (hi >= 0 && __ldg(starts + hi) > n3) --hi;
    int first = (int)((n0 - L + 1 > 0) ? (n0 - L + stride) / stride : 0);
    if (first > hi + 1) first = hi + 1;
    while (first > 0 && __ldg(starts + first - 1) + L > n0) --first;
    while (first <= hi && __ldg(starts + first) + L <= n0) ++first;
    const bool vec_out = ((pred_len & 3) == 0) && (n0 + 3 < pred_len);
    if (hi - first + 1 > COV) {
        // more covering windows than slots: reproduce the reference's slot overwrite per sample
        for (int e = 0; e < 4 && n0 + e < pred_len; ++e) {
            const int64_t n = n0 + e;
            for (int c = 0; c < nlab; ++c) {
                float slot[COV];
#pragma unroll
                for (int k = 0; k < COV; ++k) slot[k] = CUDART_NAN_F;
                for (int i = first; i <= hi; ++i) {
                    const int64_t off = n - __ldg(starts + i);
                    if (off < 0 || off >= L) continue;
                    const float val = (off < b0 || off >= L - b1) ? CUDART_NAN_F : __ldg(y + ((int64_t)i * nlab + c) * L + off);
                    const int sl = i % COV;
#pragma unroll
                    for (int k = 0; k < COV; ++k)
                        if (k == sl) slot[k] = val;
                }
                float res;
                if (mode == VP_STACK_AVG) {
                    int cnt = 0;
#pragma unroll
                    for (int k = 0; k < COV; ++k) {
                        if (!isnan(slot[k])) ++cnt; else slot[k] = 0.f;
                    }
                    const float tot = 0.0f + np_pairwise_sum<COV, true>(slot, COV);
                    res = cnt ? __fdiv_rn(tot, (float)cnt) : CUDART_NAN_F;
                } else {
                    res = CUDART_NAN_F;
#pragma unroll
                    for (int k = 0; k < COV; ++k)
                        if (!isnan(slot[k]) && (isnan(res) || slot[k] > res)) res = slot[k];
                }
                out[(int64_t)c * pred_len + n] = res;
            }
        }
        return;
    }
    const int fm = first % COV;
    // per slot: window index and offset of sample n0 inside it (same for all labels)
    int widx[COV];
    int woff[COV];
#pragma unroll
    for (int k = 0; k < COV; ++k) {
        const int i = first + ((k - fm + COV) % COV);
        widx[k] = (i <= hi) ? i : -1;
        woff[k] = (i <= hi) ? (int)(n0 - __ldg(starts + i)) : 0;
    }
    for (int c = 0; c < nlab; ++c) {
        float slot[4][COV];
#pragma unroll
        for (int k = 0; k < COV; ++k) {
            float v[4] = {CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F};
            if (widx[k] >= 0) {
                const int off = woff[k];
                const float *row = y + ((int64_t)widx[k] * nlab + c) * L;
                // the 4 samples that are inside the window and outside the blinded margins
                if (off >= b0 && off + 3 < L - b1 && ((off & 3) == 0) && ((L & 3) == 0)) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(row + off));
                    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int o = off + e;
                        if (o >= b0 && o < L - b1 && o >= 0 && o < L) v[e] = __ldg(row + o);
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) slot[e][k] = v[e];
        }
        float res[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (mode == VP_STACK_AVG) {
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < COV; ++k) {
                    if (!isnan(slot[e][k])) ++cnt; else slot[e][k] = 0.f;
                }
                const float tot = 0.0f + np_pairwise_sum<COV, true>(slot[e], COV);
                res[e] = cnt ? __fdiv_rn(tot, (float)cnt) : CUDART_NAN_F;
            } else {
                float r = CUDART_NAN_F;
#pragma unroll
                for (int k = 0; k < COV; ++k)
                    if (!isnan(slot[e][k]) && (isnan(r) || slot[e][k] > r)) r = slot[e][k];
                res[e] = r;
            }
        }
        float *ob = out + (int64_t)c * pred_len + n0;
        if (vec_out) {
            *reinterpret_cast<float4 *>(ob) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
            for (int e = 0; e < 4 && n0 + e < pred_len; ++e) ob[e] = res[e];
        }
    }
}

// ------------------------------------------------------------------------------------------ trim
__global__ void nan_bounds_init_kernel(int64_t *bounds, int nlab, int64_t pred_len) {
    const int i = threadIdx.x;
    if (i < nlab) {
        bounds[2 * i] = pred_len;
        bounds[2 * i + 1] = -1;
    }
}

__global__ void __launch_bounds__(256) nan_bounds_kernel(const float *__restrict__ a, int nlab, int64_t pred_len,
                                                         int64_t *bounds) {
    const int c = blockIdx.y;
    int64_t lo = pred_len, hi = -1;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < pred_len; n += (int64_t)gridDim.x * blockDim.x) {
        if (!isnan(__ldg(a + (int64_t)c * pred_len + n))) {
            if (n < lo) lo = n;
            if (n > hi) hi = n;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int64_t l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        const int64_t h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo < pred_len) atomicMin(reinterpret_cast<long long *>(bounds + 2 * c), (long long)lo);
        if (hi >= 0) atomicMax(reinterpret_cast<long long *>(bounds + 2 * c + 1), (long long)hi);
    }
}

// ------------------------------------------------------------------------------------------ K10
// trigger_onset + first argmax (+ _trim_nan bounds) in ONE pass over the annotation: a CTA stages a tile of
// PK_TILE samples of one label in shared memory with aligned 16-byte loads, turns it into three bit masks
// (x > thr_off, x > thr_on, !isnan) with warp ballots, finds the run starts of the tile from the mask words and lets
// one warp resolve each run out of shared memory (end = next zero bit, onset = first set bit of the on-mask, first
// maximum by a strided scan).  Only a run that leaves the tile is continued from global memory.
constexpr int PK_TILE = 4096;
constexpr int PK_NT = 256;
constexpr int PK_WORDS = PK_TILE / 32;
constexpr int PK_MAXLAB = 4;

struct PickLabels {
    const float *x[PK_MAXLAB];
    float thr_on[PK_MAXLAB], thr_off[PK_MAXLAB];
    int label[PK_MAXLAB];
    int pick[PK_MAXLAB];  // 0: only the NaN bounds of this label are wanted
    int bound_slot[PK_MAXLAB];
};
// Window mode (vp_pick_windows; the reference's evaluate() path): blockIdx.y = window * n_pick + entry; the trace of an
// entry is the slice [lo, hi) of row (window, label[entry]) of y (n_windows, n_labels, L); trigger indices are relative
// to lo and the trigger label is window * n_labels + label.
struct PickWin {
    const float *y;  // nullptr: trace mode
    int64_t L;
    int n_labels, n_pick;
    int64_t win0;            // absolute index of the first window of this launch
    const int64_t *borders;  // device (n_windows, 2) = [lo, hi) per window, or nullptr for the whole window
};

// Continues a run over global memory from `base`, PK_WALK * 32 samples per round trip (the loads of a round are
// independent; a run that leaves its tile is latency bound: 32 samples per trip made a 2000-sample detection run the
// critical path of the whole launch).
constexpr int PK_WALK = 8;
__device__ __forceinline__ void pick_walk_global(const float *__restrict__ x, int64_t n, int64_t base, float thr_on,
                                                 float thr_off, int lane, int64_t &on, int64_t &pk, float &best,
                                                 int64_t &end) {
    for (;; base += 32 * PK_WALK) {
        float vv[PK_WALK];
#pragma unroll
        for (int k = 0; k < PK_WALK; ++k) {
            const int64_t idx = base + 32 * k + lane;
            vv[k] = (idx < n) ? __ldg(x + idx) : CUDART_NAN_F;
        }
#pragma unroll
        for (int k = 0; k < PK_WALK; ++k) {
            const int64_t b0 = base + 32 * k;
            const int64_t idx = b0 + lane;
            const float v = vv[k];
            const bool above = v > thr_off;  // NaN compares false
            const unsigned not_above = __ballot_sync(0xffffffffu, !above);
            const int nvalid = not_above ? (__ffs(not_above) - 1) : 32;
            const bool in_run = lane < nvalid;
            if (on < 0) {
                const unsigned m_on = __ballot_sync(0xffffffffu, in_run && v > thr_on);
                if (m_on) on = b0 + (__ffs(m_on) - 1);
            }
            if (on >= 0) {
                float bv = (in_run && idx >= on) ? v : -CUDART_INF_F;
                int64_t bi = idx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) {
                        bv = ov;
                        bi = oi;
                    }
                }
                if (bv > best) {  // strict: the earliest chunk keeps ties -> first maximum
                    best = bv;
                    pk = bi;
                }
            }
            if (nvalid < 32) {
                end = b0 + nvalid - 1;
                return;
            }
        }
    }
}

__global__ void __launch_bounds__(PK_NT) pick_tile_kernel(const PickLabels P, const PickWin W, int64_t n,
                                                          vp_trigger *__restrict__ picks, int64_t pick_cap,
                                                          unsigned long long *__restrict__ npicks,
                                                          int64_t *__restrict__ bounds) {
    __shared__ __align__(16) float sv[PK_TILE];
    __shared__ uint32_t m_off[PK_WORDS];
    __shared__ uint16_t run_start[PK_TILE / 2];
    __shared__ int n_runs, s_lo, s_hi, s_prev, s_prev_ok, s_next_ok;
    int li = W.y != nullptr ? 0 : (int)blockIdx.y;
    const float *__restrict__ x = P.x[li];
    int out_label = P.label[li];
    if (W.y != nullptr) {  // window mode: this CTA's trace is a slice of one (window, label) row
        const int64_t win = blockIdx.y / W.n_pick;
        li = (int)(blockIdx.y - win * W.n_pick);
        int64_t lo = 0, hi = W.L;
        if (W.borders != nullptr) {
            lo = min(max(__ldg(W.borders + 2 * win), (int64_t)0), W.L);
            hi = min(max(__ldg(W.borders + 2 * win + 1), lo), W.L);
        }
        x = W.y + (win * W.n_labels + P.label[li]) * W.L + lo;
        n = hi - lo;
        out_label = (int)((W.win0 + win) * W.n_labels + P.label[li]);
        if ((int64_t)blockIdx.x * PK_TILE - 3 >= n) return;  // tile beyond the slice (uniform for the CTA)
    }
    const float thr_on = P.thr_on[li], thr_off = P.thr_off[li];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = (int)((reinterpret_cast<uintptr_t>(x) >> 2) & 3);  // elements past a 16-byte boundary
    const int64_t t0 = (int64_t)blockIdx.x * PK_TILE - shift;
    if (tid == 0) {
        n_runs = 0;
        s_lo = PK_TILE;
        s_hi = -1;
        const float pv = (t0 > 0 && t0 - 1 < n) ? __ldg(x + t0 - 1) : CUDART_NAN_F;
        s_prev = pv > thr_off ? 1 : 0;
        s_prev_ok = isnan(pv) ? 0 : 1;
    }
    if (tid == 32) s_next_ok = (t0 + PK_TILE < n && !isnan(__ldg(x + t0 + PK_TILE))) ? 1 : 0;
    int n_ok = 0;  // non-NaN samples loaded by this thread
    {
        // all loads of the tile are issued before the first use: with the bounds test inside the loop the compiler kept one
        // 16-byte load in flight per thread (four serial DRAM round trips per CTA, 24 % of the HBM peak)
        constexpr int NIT = PK_TILE / (4 * PK_NT);
        float4 v[NIT];
        if (t0 >= 0 && t0 + PK_TILE <= n) {  // interior tile (uniform for the CTA)
#pragma unroll
            for (int it = 0; it < NIT; ++it) v[it] = __ldg(reinterpret_cast<const float4 *>(x + t0 + (it * PK_NT + tid) * 4));
        } else {
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int64_t g = t0 + (it * PK_NT + tid) * 4;
                v[it].x = (g >= 0 && g < n) ? __ldg(x + g) : CUDART_NAN_F;
                v[it].y = (g + 1 >= 0 && g + 1 < n) ? __ldg(x + g + 1) : CUDART_NAN_F;
                v[it].z = (g + 2 >= 0 && g + 2 < n) ? __ldg(x + g + 2) : CUDART_NAN_F;
                v[it].w = (g + 3 >= 0 && g + 3 < n) ? __ldg(x + g + 3) : CUDART_NAN_F;
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            n_ok += (v[it].x == v[it].x) + (v[it].y == v[it].y) + (v[it].z == v[it].z) + (v[it].w == v[it].w);
            *reinterpret_cast<float4 *>(sv + (it * PK_NT + tid) * 4) = v[it];
        }
    }
    // tile-wide census: all samples valid (the common case), none, or mixed (only then the NaN positions are searched)
    const int all_ok = __syncthreads_and(n_ok == PK_TILE / PK_NT);
    const int none_ok = all_ok ? 0 : __syncthreads_and(n_ok == 0);
    const bool want_pick = P.pick[li] != 0;
    if (none_ok) return;  // NaN only: no run, no bound
    int lo = PK_TILE, hi = -1;
#pragma unroll 4
    for (int j = 0; j < PK_WORDS / (PK_NT / 32); ++j) {
        const int word = warp * (PK_WORDS / (PK_NT / 32)) + j;
        const float v = sv[word * 32 + lane];
        if (want_pick) {
            const unsigned b_off = __ballot_sync(0xffffffffu, v > thr_off);  // NaN compares false
            if (lane == 0) m_off[word] = b_off;
        }
        if (!all_ok && bounds != nullptr) {
            const unsigned b_ok = __ballot_sync(0xffffffffu, v == v);
            if (b_ok) {
                lo = min(lo, word * 32 + __ffs(b_ok) - 1);
                hi = max(hi, word * 32 + 31 - __clz(b_ok));
            }
        }
    }
    if (bounds != nullptr) {
        if (all_ok) {
            if (tid == 0) {
                s_lo = 0;
                s_hi = PK_TILE - 1;
            }
        } else if (lane == 0 && hi >= 0) {
            atomicMin(&s_lo, lo);
            atomicMax(&s_hi, hi);
        }
    }
    __syncthreads();
    if (bounds != nullptr && tid == 0 && s_hi >= 0) {
        // the first (last) non-NaN sample of the label follows (precedes) a NaN or the edge of the trace: a tile whose
        // valid samples continue into its neighbours cannot hold it, so interior tiles issue no same-address atomics
        const int slot = P.bound_slot[li];
        if (s_lo > 0 || !s_prev_ok) atomicMin(reinterpret_cast<long long *>(bounds + 2 * slot), (long long)(t0 + s_lo));
        if (s_hi < PK_TILE - 1 || !s_next_ok)
            atomicMax(reinterpret_cast<long long *>(bounds + 2 * slot + 1), (long long)(t0 + s_hi));
    }
    if (!want_pick) return;
    if (tid < PK_WORDS) {
        const uint32_t m = m_off[tid];
        const uint32_t carry = tid > 0 ? (m_off[tid - 1] >> 31) : (uint32_t)s_prev;
        uint32_t st = m & ~((m << 1) | carry);
        while (st) {
            const int b = __ffs(st) - 1;
            st &= st - 1;
            run_start[atomicAdd(&n_runs, 1)] = (uint16_t)(tid * 32 + b);
        }
    }
    __syncthreads();
    const int nr = n_runs;
    for (int r = warp; r < nr; r += PK_NT / 32) {
        const int s = run_start[r];
        const int w = s >> 5;
        const uint32_t from_s = 0xffffffffu << (s & 31);
        // last sample of the run inside the tile
        int end_in = PK_TILE - 1;
        bool leaves = true;
        for (int wb = w; wb < PK_WORDS; wb += 32) {
            const int wi = wb + lane;
            uint32_t inv = 0;
            if (wi < PK_WORDS) {
                inv = ~m_off[wi];
                if (wi == w) inv &= from_s;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, inv != 0);
            if (bal) {
                const int l = __ffs(bal) - 1;
                const uint32_t iv = __shfl_sync(0xffffffffu, inv, l);
                end_in = (wb + l) * 32 + __ffs(iv) - 2;
                leaves = false;
                break;
            }
        }
        // first sample above thr_on and the first maximum from there on, one pass over [s, end_in] in shared memory
        int on_in = -1;
        float bv = -CUDART_INF_F;
        int bi = 0x7fffffff;
        for (int base = s; base <= end_in; base += 32) {
            const int i = base + lane;
            const float v = (i <= end_in) ? sv[i] : -CUDART_INF_F;
            if (on_in < 0) {
                const unsigned m_on = __ballot_sync(0xffffffffu, v > thr_on);
                if (m_on) on_in = base + __ffs(m_on) - 1;
            }
            if (on_in >= 0 && i >= on_in && v > bv) {  // strict: the earliest index of a lane keeps ties
                bv = v;
                bi = i;
            }
        }
        if (on_in >= 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) {
                    bv = ov;
                    bi = oi;
                }
            }
        }
        int64_t on = on_in >= 0 ? t0 + on_in : -1, pk = on_in >= 0 ? t0 + bi : -1, end = t0 + end_in;
        float best = bv;
        if (leaves) pick_walk_global(x, n, t0 + PK_TILE, thr_on, thr_off, lane, on, pk, best, end);
        if (on >= 0 && lane == 0) {
            const unsigned long long slot = atomicAdd(npicks, 1ULL);
            if ((int64_t)slot < pick_cap) {
                vp_trigger q;
                q.s0 = on;
                q.s1 = end;
                q.s_peak = pk;
                q.value = best;
                q.label = out_label;
                picks[slot] = q;
            }
        }
    }
}

}  // namespace vp

// ============================================================================================ C ABI
using namespace vp;

extern "C" int vp_slice_normalize(const void *trace, int dtype, int64_t n_samples, int64_t ch_stride,
                                  const int64_t *starts, int64_t n_windows, int64_t in_samples, int peak_scope,
                                  int taper, float *out, void *stream) {
    VP_REQUIRE(trace && starts && out, VP_ERR_ARG, "vp_slice_normalize: null pointer");
    VP_REQUIRE(n_windows >= 0 && in_samples > 12 && in_samples <= n_samples, VP_ERR_ARG,
               "vp_slice_normalize: bad sizes (n=%lld, L=%lld, windows=%lld)", (long long)n_samples,
               (long long)in_samples, (long long)n_windows);
    VP_REQUIRE(peak_scope == VP_PEAK_PER_CHANNEL || peak_scope == VP_PEAK_PER_WINDOW, VP_ERR_ARG,
               "vp_slice_normalize: bad peak_scope %d", peak_scope);
    if (n_windows == 0) return VP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == VP_DTYPE_F32)
        return launch_slice<float>((const float *)trace, ch_stride, 2 * ch_stride + n_samples, starts, n_windows, (int)in_samples,
                                   peak_scope, taper, out, s);
    if (dtype == VP_DTYPE_I32)
        return launch_slice<int32_t>((const int32_t *)trace, ch_stride, 2 * ch_stride + n_samples, starts, n_windows,
                                     (int)in_samples, peak_scope, taper, out, s);
    set_error("vp_slice_normalize: unknown dtype %d", dtype);
    return VP_ERR_ARG;
}

extern "C" int64_t vp_coverage(int64_t L, int64_t overlap) {
    const int64_t stride = L - overlap;
    if (stride <= 0) return -1;
    // int(np.ceil(L / stride + 1)) evaluated like NumPy does, in double
    return (int64_t)std::ceil((double)L / (double)stride + 1.0);
}

extern "C" int vp_stack(const float *y, const int64_t *starts, int64_t n_windows, int64_t in_samples, int n_labels,
                        int64_t overlap, int64_t blind0, int64_t blind1, int mode, float *out, int64_t pred_len,
                        void *stream) {
    VP_REQUIRE(y && starts && out, VP_ERR_ARG, "vp_stack: null pointer");
    VP_REQUIRE(mode == VP_STACK_AVG || mode == VP_STACK_MAX, VP_ERR_ARG,
               "Stacking method %d unknown. Known methods are: 'avg' (0), 'max' (1)", mode);
    const int64_t cov = vp_coverage(in_samples, overlap);
    VP_REQUIRE(cov > 0, VP_ERR_ARG, "vp_stack: overlap %lld >= window length %lld", (long long)overlap,
               (long long)in_samples);
    VP_REQUIRE(cov <= STACK_MAXCOV, VP_ERR_UNSUPPORTED, "vp_stack: coverage %lld > %d not supported",
               (long long)cov, STACK_MAXCOV);
    VP_REQUIRE(blind0 >= 0 && blind1 >= 0, VP_ERR_ARG, "vp_stack: negative blinding");
    if (n_windows == 0 || pred_len == 0) return VP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((pred_len + 255) / 256);
#define VP_STACK_LAUNCH(COV, EXACT)                                                                              \
    stack_kernel<COV, EXACT><<<grid, 256, 0, s>>>(y, starts, n_windows, (int)in_samples, n_labels, (int)cov, blind0, \
                                                  blind1, mode, out, pred_len)
    const int64_t stride = in_samples - overlap;
    const bool small = n_windows < (1LL << 30) && in_samples < (1LL << 30) && blind0 <= in_samples && blind1 <= in_samples;
    const unsigned grid4 = (unsigned)((pred_len + 1023) / 1024);
    KTimer kt(KC_STACK, s);
    static const bool scalar_only = getenv("VP_STACK_SCALAR") && atoi(getenv("VP_STACK_SCALAR")) != 0;  // debugging aid
    if (cov == 3 && small && !scalar_only)
        stack4_kernel<3><<<grid4, 256, 0, s>>>(y, starts, (int)n_windows, (int)in_samples, n_labels, (int)stride, (int)blind0,
                                               (int)blind1, mode, out, pred_len);
    else if (cov == 13 && small && !scalar_only)
        stack4_kernel<13><<<grid4, 256, 0, s>>>(y, starts, (int)n_windows, (int)in_samples, n_labels, (int)stride, (int)blind0,
                                                (int)blind1, mode, out, pred_len);
    else if (cov == 3) VP_STACK_LAUNCH(3, true);
    else if (cov == 13) VP_STACK_LAUNCH(13, true);
    else if (cov <= 16) VP_STACK_LAUNCH(16, false);
    else VP_STACK_LAUNCH(STACK_MAXCOV, false);
#undef VP_STACK_LAUNCH
    VP_LAUNCH_CHECK();
    return VP_OK;
}

extern "C" int vp_nan_bounds(const float *annotation, int n_labels, int64_t pred_len, int64_t *bounds, void *stream) {
    VP_REQUIRE(annotation && bounds && n_labels > 0 && n_labels <= 32, VP_ERR_ARG, "vp_nan_bounds: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    KTimer kt(KC_TRIM, s);
    nan_bounds_init_kernel<<<1, 32, 0, s>>>(bounds, n_labels, pred_len);
    VP_LAUNCH_CHECK();
    if (pred_len > 0) {
        dim3 grid((unsigned)std::min<int64_t>((pred_len + 255) / 256, (int64_t)device_sm_count() * 8), n_labels);
        nan_bounds_kernel<<<grid, 256, 0, s>>>(annotation, n_labels, pred_len, bounds);
        VP_LAUNCH_CHECK();
    }
    return VP_OK;
}

extern "C" int64_t vp_pick_scratch_bytes(int64_t n_samples) {
    (void)n_samples;  // the single-pass kernel keeps its run list in shared memory; the argument is kept for the ABI
    return 256;
}

static int launch_pick_tiles(const PickLabels &P, int n_lab, int64_t n, vp_trigger *picks, int64_t capacity,
                             int64_t *count, int64_t *bounds, cudaStream_t s) {
    // +1 tile: the tile grid of a label starts up to 3 samples before its first sample (16-byte alignment)
    const unsigned tiles = (unsigned)((n + 3 + PK_TILE - 1) / PK_TILE);
    KTimer kt(KC_PICK, s);
    PickWin W = {};
    pick_tile_kernel<<<dim3(tiles, n_lab), PK_NT, 0, s>>>(P, W, n, picks, capacity, (unsigned long long *)count, bounds);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

extern "C" int vp_pick(const float *trace, int64_t n_samples, float thr_on, float thr_off, int label, vp_trigger *picks,
                       int64_t capacity, int64_t *count, void *scratch, int64_t scratch_bytes, void *stream) {
    VP_REQUIRE(trace && picks && count, VP_ERR_ARG, "vp_pick: null pointer");
    (void)scratch;
    (void)scratch_bytes;
    if (n_samples <= 0) return VP_OK;
    PickLabels P = {};
    P.x[0] = trace;
    P.thr_on[0] = thr_on;
    P.thr_off[0] = thr_off;
    P.label[0] = label;
    P.pick[0] = 1;
    return launch_pick_tiles(P, 1, n_samples, picks, capacity, count, nullptr, (cudaStream_t)stream);
}

extern "C" int vp_pick_labels(const float *annotation, int n_labels, int64_t pred_len, const float *thr_on,
                              const float *thr_off, vp_trigger *picks, int64_t capacity, int64_t *count, int64_t *bounds,
                              void *stream) {
    VP_REQUIRE(annotation && thr_on && thr_off && count, VP_ERR_ARG, "vp_pick_labels: null pointer");
    VP_REQUIRE(n_labels > 0 && n_labels <= PK_MAXLAB, VP_ERR_ARG, "vp_pick_labels: n_labels %d outside [1, %d]", n_labels,
               PK_MAXLAB);
    VP_REQUIRE(capacity == 0 || picks, VP_ERR_ARG, "vp_pick_labels: pick buffer missing");
    cudaStream_t s = (cudaStream_t)stream;
    if (bounds) {
        nan_bounds_init_kernel<<<1, 32, 0, s>>>(bounds, n_labels, pred_len);
        VP_LAUNCH_CHECK();
    }
    if (pred_len <= 0) return VP_OK;
    PickLabels P = {};
    int nl = 0;
    for (int c = 0; c < n_labels; ++c) {
        const bool pick = thr_on[c] > 0.f && capacity > 0;  // <= 0 or NaN: no picks for that label
        if (!pick && !bounds) continue;
        P.x[nl] = annotation + (int64_t)c * pred_len;
        P.thr_on[nl] = thr_on[c];
        P.thr_off[nl] = thr_off[c];
        P.label[nl] = c;
        P.pick[nl] = pick ? 1 : 0;
        P.bound_slot[nl] = c;
        ++nl;
    }
    if (nl == 0) return VP_OK;
    return launch_pick_tiles(P, nl, pred_len, picks, capacity, count, bounds, s);
}

extern "C" int vp_pick_windows(const float *y, int64_t n_windows, int n_labels, int64_t in_samples, const int64_t *borders,
                               const float *thr_on, const float *thr_off, vp_trigger *picks, int64_t capacity, int64_t *count,
                               void *stream) {
    VP_REQUIRE(y && thr_on && thr_off && picks && count, VP_ERR_ARG, "vp_pick_windows: null pointer");
    VP_REQUIRE(n_labels > 0 && n_labels <= PK_MAXLAB && in_samples > 0 && n_windows >= 0, VP_ERR_ARG,
               "vp_pick_windows: bad sizes (windows %lld, labels %d, samples %lld)", (long long)n_windows, n_labels, (long long)in_samples);
    VP_REQUIRE(n_windows * n_labels < (1LL << 31), VP_ERR_ARG, "vp_pick_windows: window * label index exceeds the trigger label field");
    PickLabels P = {};
    PickWin W = {};
    int np = 0;
    for (int c = 0; c < n_labels; ++c) {
        if (!(thr_on[c] > 0.f)) continue;  // <= 0 or NaN: the label is not picked
        P.thr_on[np] = thr_on[c];
        P.thr_off[np] = thr_off[c];
        P.label[np] = c;
        P.pick[np] = 1;
        ++np;
    }
    if (np == 0 || n_windows == 0 || capacity <= 0) return VP_OK;
    W.y = y;
    W.L = in_samples;
    W.n_labels = n_labels;
    W.n_pick = np;
    W.borders = borders;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned tiles = (unsigned)((in_samples + 3 + PK_TILE - 1) / PK_TILE);
    KTimer kt(KC_PICK, s);
    const int64_t max_win = 65535 / np;  // grid.y limit
    for (int64_t w0 = 0; w0 < n_windows; w0 += max_win) {
        const int64_t nw = std::min(max_win, n_windows - w0);
        PickWin Wc = W;
        Wc.y = y + w0 * n_labels * in_samples;
        Wc.borders = borders ? borders + 2 * w0 : nullptr;
        Wc.win0 = w0;
        pick_tile_kernel<<<dim3(tiles, (unsigned)(nw * np)), PK_NT, 0, s>>>(P, Wc, in_samples, picks, capacity, (unsigned long long *)count, nullptr);
        VP_LAUNCH_CHECK();
    }
    return VP_OK;
}
