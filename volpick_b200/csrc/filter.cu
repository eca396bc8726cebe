// Stream pre-filter on the device: cascaded second-order sections (scipy.signal.sosfilt semantics, float64 arithmetic),
// optionally zero-phase (forward pass, then the same filter over the reversed signal) -- what
// obspy.signal.filter.{highpass,lowpass,bandpass,bandstop} run when a SeisBench model carries filter_args / filter_kwargs
// (reference use: model_training/test_onephase.ipynb cell 43, volpick/data/utils.py:702-704).  The host designs the
// sections (scipy.signal.iirfilter + zpk2sos, as ObsPy does); this file only evaluates them.
//
// A biquad is a linear recurrence, so a record of N samples is cut into blocks of FL_BLOCK samples:
//   1. every block runs the recurrence from a ZERO state (one thread per block and channel, direct form II transposed
//      exactly as sosfilt) and stores its local output and its final state;
//   2. one thread per channel carries the true block start states along the record:  s_{k+1} = A^B s_k + e_k;
//   3. every sample adds the homogeneous response of its block's start state:  y[n] += h0[j] s.z0 + h1[j] s.z1,
//      with h0 / h1 = responses to the unit states, computed once per section on the host.
// Sections run one after the other on a float64 work buffer; the last pass writes float32.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace vp {

constexpr int FL_BLOCK = 2048;
constexpr int FL_MAX_SECTIONS = 8;

struct Biquad {
    double b0, b1, b2, a1, a2;
};

template <typename Tin>
__device__ __forceinline__ double fl_load(const Tin *p) {
    return (double)__ldg(p);
}

// pass 1: zero-state response of every block.  rev: the pass runs over the reversed signal (index n stands for N-1-n).
template <typename Tin>
__global__ void __launch_bounds__(128) iir_block_kernel(const Tin *__restrict__ x, int64_t x_cs, double *__restrict__ y, int64_t n,
                                                        int n_blocks, int n_ch, Biquad q, int rev, double *__restrict__ end_state) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_blocks * n_ch) return;
    const int c = (int)(t / n_blocks), k = (int)(t - (int64_t)c * n_blocks);
    const int64_t i0 = (int64_t)k * FL_BLOCK, i1 = min(i0 + FL_BLOCK, n);
    const Tin *xc = x + (int64_t)c * x_cs;
    double *yc = y + (int64_t)c * n;
    double z0 = 0.0, z1 = 0.0;
    for (int64_t i = i0; i < i1; ++i) {
        const int64_t idx = rev ? (n - 1 - i) : i;
        const double xv = fl_load(xc + idx);
        const double yv = q.b0 * xv + z0;  // scipy _sosfilt: direct form II transposed, this operation order
        z0 = q.b1 * xv - q.a1 * yv + z1;
        z1 = q.b2 * xv - q.a2 * yv;
        yc[idx] = yv;
    }
    end_state[2 * t] = z0;
    end_state[2 * t + 1] = z1;
}

// pass 2: start state of every block (in place: end_state[k] <- state at the START of block k)
__global__ void iir_carry_kernel(double *__restrict__ state, int n_blocks, int n_ch, double m00, double m01, double m10, double m11) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) return;
    double *s = state + 2 * (int64_t)c * n_blocks;
    double z0 = 0.0, z1 = 0.0;
    for (int k = 0; k < n_blocks; ++k) {
        const double e0 = s[2 * k], e1 = s[2 * k + 1];
        s[2 * k] = z0;
        s[2 * k + 1] = z1;
        const double n0 = m00 * z0 + m01 * z1 + e0;
        const double n1 = m10 * z0 + m11 * z1 + e1;
        z0 = n0;
        z1 = n1;
    }
}

// pass 3: add the homogeneous response of the block start state; optionally emit float32
__global__ void __launch_bounds__(256) iir_fix_kernel(double *__restrict__ y, int64_t n, int n_blocks, int n_ch,
                                                      const double *__restrict__ h /*[2][FL_BLOCK]*/, const double *__restrict__ state,
                                                      int rev, float *__restrict__ out32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // position in pass order
    const int c = blockIdx.y;
    if (i >= n) return;
    const int k = (int)(i / FL_BLOCK), j = (int)(i - (int64_t)k * FL_BLOCK);
    const double *s = state + 2 * ((int64_t)c * n_blocks + k);
    const int64_t idx = rev ? (n - 1 - i) : i;
    const double v = y[(int64_t)c * n + idx] + (__ldg(h + j) * s[0] + __ldg(h + FL_BLOCK + j) * s[1]);
    if (out32) out32[(int64_t)c * n + idx] = (float)v;
    else y[(int64_t)c * n + idx] = v;
}

}  // namespace vp

using namespace vp;

extern "C" int64_t vp_sosfilt_workspace_bytes(int64_t n_samples, int n_channels, int n_sections) {
    if (n_samples < 0 || n_channels <= 0 || n_sections <= 0 || n_sections > FL_MAX_SECTIONS) return VP_ERR_ARG;
    const int64_t n_blocks = (n_samples + FL_BLOCK - 1) / FL_BLOCK;
    return align_up(2 * n_channels * n_samples * 8, 256)        // two float64 work buffers (ping-pong between sections)
           + align_up(2 * n_channels * n_blocks * 8, 256)        // block states
           + align_up((int64_t)n_sections * 2 * FL_BLOCK * 8, 256);  // homogeneous responses
}

extern "C" int vp_sosfilt(const void *x, int dtype, int64_t n, int64_t ch_stride, int n_channels, const double *sos, int n_sections,
                          int zerophase, float *y, void *workspace, int64_t workspace_bytes, void *stream) {
    VP_REQUIRE(x && sos && y && workspace, VP_ERR_ARG, "vp_sosfilt: null pointer");
    VP_REQUIRE(dtype == VP_DTYPE_F32 || dtype == VP_DTYPE_I32, VP_ERR_ARG, "vp_sosfilt: unknown dtype %d", dtype);
    VP_REQUIRE(n_channels > 0 && n_channels <= 65535 && n_sections > 0 && n_sections <= FL_MAX_SECTIONS, VP_ERR_ARG,
               "vp_sosfilt: %d channels / %d sections unsupported", n_channels, n_sections);
    const int64_t need = vp_sosfilt_workspace_bytes(n, n_channels, n_sections);
    VP_REQUIRE(workspace_bytes >= need, VP_ERR_WORKSPACE, "vp_sosfilt: workspace too small (%lld < %lld bytes)",
               (long long)workspace_bytes, (long long)need);
    if (n == 0) return VP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int n_blocks = (int)((n + FL_BLOCK - 1) / FL_BLOCK);
    char *ws = (char *)workspace;
    double *buf[2];
    buf[0] = (double *)ws;
    buf[1] = buf[0] + (int64_t)n_channels * n;
    double *state = (double *)(ws + align_up(2 * n_channels * n * 8, 256));
    double *d_h = (double *)((char *)state + align_up(2 * (int64_t)n_channels * n_blocks * 8, 256));
    // per section: normalised coefficients, homogeneous responses h0 / h1 and the block transition matrix A^B
    std::vector<double> h((size_t)n_sections * 2 * FL_BLOCK);
    std::vector<Biquad> q(n_sections);
    std::vector<double> M((size_t)n_sections * 4);
    for (int i = 0; i < n_sections; ++i) {
        const double *c = sos + 6 * i;
        VP_REQUIRE(c[3] != 0.0 && std::isfinite(c[3]), VP_ERR_ARG, "vp_sosfilt: section %d has a0 = %g", i, c[3]);
        q[i] = Biquad{c[0] / c[3], c[1] / c[3], c[2] / c[3], c[4] / c[3], c[5] / c[3]};
        for (int e = 0; e < 2; ++e) {
            double z0 = e == 0 ? 1.0 : 0.0, z1 = e == 0 ? 0.0 : 1.0;
            for (int j = 0; j < FL_BLOCK; ++j) {
                const double yv = z0;  // x = 0
                h[((size_t)i * 2 + e) * FL_BLOCK + j] = yv;
                const double nz0 = -q[i].a1 * yv + z1, nz1 = -q[i].a2 * yv;
                z0 = nz0;
                z1 = nz1;
            }
            M[(size_t)i * 4 + 0 + e] = z0;  // column e of A^B
            M[(size_t)i * 4 + 2 + e] = z1;
        }
    }
    VP_CUDA_CHECK(cudaMemcpyAsync(d_h, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s));  // pageable: staged before return
    const int n_pass = zerophase ? 2 : 1;
    const int total = n_pass * n_sections;
    const int64_t threads1 = (int64_t)n_blocks * n_channels;
    int step = 0;
    const double *src64 = nullptr;
    for (int pass = 0; pass < n_pass; ++pass)
        for (int i = 0; i < n_sections; ++i, ++step) {
            double *dst = buf[step & 1];
            const bool last = step == total - 1;
            {
                KTimer kt(KC_FILTER, s);
                const unsigned g1 = (unsigned)((threads1 + 127) / 128);
                if (step == 0) {
                    if (dtype == VP_DTYPE_F32)
                        iir_block_kernel<float><<<g1, 128, 0, s>>>((const float *)x, ch_stride, dst, n, n_blocks, n_channels, q[i], pass, state);
                    else
                        iir_block_kernel<int32_t><<<g1, 128, 0, s>>>((const int32_t *)x, ch_stride, dst, n, n_blocks, n_channels, q[i], pass, state);
                } else {
                    iir_block_kernel<double><<<g1, 128, 0, s>>>(src64, n, dst, n, n_blocks, n_channels, q[i], pass, state);
                }
                VP_LAUNCH_CHECK();
                iir_carry_kernel<<<(n_channels + 31) / 32, 32, 0, s>>>(state, n_blocks, n_channels, M[(size_t)i * 4 + 0], M[(size_t)i * 4 + 1],
                                                                      M[(size_t)i * 4 + 2], M[(size_t)i * 4 + 3]);
                VP_LAUNCH_CHECK();
                dim3 g3((unsigned)((n + 255) / 256), n_channels);
                iir_fix_kernel<<<g3, 256, 0, s>>>(dst, n, n_blocks, n_channels, d_h + (size_t)i * 2 * FL_BLOCK, state, pass, last ? y : nullptr);
                VP_LAUNCH_CHECK();
            }
            src64 = dst;
        }
    return VP_OK;
}
