// tcgen05 / TMEM implicit-GEMM Conv1d (sm_100a): the tensor-core path of the network forward.
//
// Replaces F.conv1d (+ folded BatchNorm, bias, ReLU / sigmoid, MaxPool1d(2), nn.Upsample(x2) + crop) of
// seisbench/models/eqtransformer.py Encoder / Decoder / heads (SURVEY.md Appendix A, K2 in section 2c).
//
// GEMM view.  Activations are channel-last 16-bit rows [seq][t][C].  All sequences of a launch are laid
// on one virtual time axis with pitch Tp >= T + halo per sequence (rows between sequences read as
// zero), so M tiles of 128 rows are dense even for T = 47.  For one tile the rows [m0+row0, m0+row0+
// 128+halo) are staged in shared memory as 8-channel planes  sA[split][plane][row][8]  (16-byte rows).
// A tap is then nothing but a row offset of the SAME staged tile:
//     D[128 x N] += A_tap[128 x 16] * W_tap[16 x N]
// with A_tap described by a no-swizzle K-major UMMA descriptor whose start address is shifted by
// tap * 16 bytes (SBO = 128 B between 8-row groups, LBO = plane pitch between the two K halves; for
// 8-channel inputs the two K halves are two consecutive taps, LBO = 16 B).  No im2col is ever built.
// x2 nearest up-sampling is folded into the weights (polyphase): row s produces outputs 2s and 2s+1
// as N = 2*C_out columns from ceil((k+1)/2)+1 taps on the un-upsampled input.
//
// Precision.  split = 2: operands are fp16 hi/lo pairs (x = hi + lo, 22 significant bits) and every K
// step issues hi*hi + hi*lo + lo*hi into the fp32 TMEM accumulator -> fp32-equivalent results on the
// tensor cores ("f16x3" mode).  split = 1: one bf16 pass.
//
// Execution: persistent warp-specialised CTAs (producers -> cp.async ring -> one tcgen05 issuer -> TMEM
// double buffer -> four epilogue warps), see tcconv_kernel below.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tc_ptx.cuh"
#include "tcconv.cuh"

namespace vp {

// ------------------------------------------------------------------------------------------ the layer kernel
// Persistent, warp-specialised CTA of 32 * (EW + 4) threads:
//   warps 0..EW-1   epilogue: TMEM lane quarter (warp & 3), column half (warp >> 2 when EW == 8) -> registers ->
//                   bias / residual / act / pool -> global
//   warp  EW        MMA issuer (one elected lane) + TMEM allocation (2 accumulator buffers)
//   warps EW+1..+3  producers: cp.async (zero-fill) of the A tiles into an NSTAGE ring
// Barriers: full[stage] (96 producer arrivals via cp.async.mbarrier.arrive.noinc), empty[stage]
// (tcgen05.commit), accf[acc] (tcgen05.commit), acce[acc] (32 * EW epilogue arrivals).  The weights and the bias of the
// layer are loaded once per CTA and stay resident in shared memory.
// EW = 8 is used when only one or two CTAs fit an SM (wide N / large resident weights): with four epilogue
// warps the TMEM -> convert -> store chain of a 128 x 128 tile was the measured bottleneck (dec1: 245 us with,
// 33 us without the epilogue).
constexpr int TC_PRODUCERS = 96;
constexpr int TC_MAX_STAGES = 4;

template <int SPLIT>
__device__ __forceinline__ void tc_pack8(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        if (SPLIT == 2) {
            const __half2 hh = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
            h[i] = *reinterpret_cast<const uint32_t *>(&hh);
            l[i] = *reinterpret_cast<const uint32_t *>(&ll);
        } else {
            const __nv_bfloat162 bb = __floats2bfloat162_rn(a, b);
            h[i] = *reinterpret_cast<const uint32_t *>(&bb);
            l[i] = 0u;
        }
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Epilogue of one 128-row tile with the layer's options fixed at compile time (FLAGS bits: 1 MaxPool1d(2),
// 2 polyphase (two output rows per accumulator row), 4 res-CNN extras = fp32 residual in / fp32 row-major out /
// pre-activation BatchNorm + ReLU of the next conv).  16-bit channel-last output, ReLU unless the extras carry
// the activation, NOUT == phases * cout (no padding columns).  The run-time-flag epilogue in the kernel body had
// ~1000 instructions per warp and tile, most of them flag tests, re-derived indices and dead output formats; the
// measured cost was ~22 K cycles per 128 x 128 tile, ten times the MMA time.
template <int NOUT, int SPLIT, int EW, int FLAGS>
__device__ __forceinline__ void tc_epilogue_fixed(const TcP &p, uint32_t trow, int half, int lane, int seq, int srow, bool row_ok,
                                                  int g, const float *s_bias, const float *s_psc, const float *s_psh) {
    constexpr bool POOL = (FLAGS & 1) != 0, PH2 = (FLAGS & 2) != 0, EXTRA = (FLAGS & 4) != 0;
    constexpr int NHALF = EW / 4, COLS = NOUT / NHALF, COUT = PH2 ? NOUT / 2 : NOUT;
    constexpr int RC = COLS < 32 ? COLS : 32;  // accumulator columns in flight per thread
    static_assert(COLS % 8 == 0 && COLS % RC == 0, "column split");
    const bool st = POOL ? (row_ok && !(lane & 1)) : row_ok;
    const int t0 = POOL ? (srow >> 1) : (PH2 ? 2 * srow : srow);
    const int64_t orow0 = (int64_t)seq * p.T_out + t0;
    uint16_t *y16 = reinterpret_cast<uint16_t *>(p.y) + (int64_t)g * p.y_gs + orow0 * COUT;
    uint4 pend_hi = make_uint4(0u, 0u, 0u, 0u), pend_lo = pend_hi;  // even 8-channel group waiting for its neighbour
#pragma unroll
    for (int rc = 0; rc < COLS; rc += RC) {
        const int nb = half * COLS + rc;
        uint32_t r[RC];
#pragma unroll
        for (int c = 0; c < RC; c += 16) {
            if constexpr (RC >= 16) {
                uint32_t(&r16)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[c]);
                tmem_ld16_nowait(trow + (uint32_t)(nb + c), r16);
            } else {
                uint32_t(&r8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&r[c]);
                tmem_ld8_nowait(trow + (uint32_t)(nb + c), r8);
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < RC; g8 += 8) {
            const int n0 = nb + g8;
            const int phi = (PH2 && n0 >= COUT) ? 1 : 0;
            const int c0 = n0 - phi * COUT;
            const bool ok = st && (t0 + phi) < p.T_out;
            float w8[8];
            const float4 b0 = *reinterpret_cast<const float4 *>(&s_bias[n0]), b1 = *reinterpret_cast<const float4 *>(&s_bias[n0 + 4]);
            w8[0] = __uint_as_float(r[g8 + 0]) + b0.x, w8[1] = __uint_as_float(r[g8 + 1]) + b0.y;
            w8[2] = __uint_as_float(r[g8 + 2]) + b0.z, w8[3] = __uint_as_float(r[g8 + 3]) + b0.w;
            w8[4] = __uint_as_float(r[g8 + 4]) + b1.x, w8[5] = __uint_as_float(r[g8 + 5]) + b1.y;
            w8[6] = __uint_as_float(r[g8 + 6]) + b1.z, w8[7] = __uint_as_float(r[g8 + 7]) + b1.w;
            if constexpr (EXTRA) {
                if (p.res != nullptr && ok) {  // residual stream, fp32 row-major [seq][t][cout]; may alias y32 (in place)
                    const float4 *rp = reinterpret_cast<const float4 *>(p.res + orow0 * COUT + c0);
                    const float4 r0 = rp[0], r1 = rp[1];
                    w8[0] += r0.x, w8[1] += r0.y, w8[2] += r0.z, w8[3] += r0.w;
                    w8[4] += r1.x, w8[5] += r1.y, w8[6] += r1.z, w8[7] += r1.w;
                }
            }
            if (!EXTRA || p.act == ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) w8[i] = fmaxf(w8[i], 0.f);
            }
            if constexpr (POOL) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float mine = row_ok ? w8[i] : -1e10f;  // SeisBench pads odd lengths with -1e10 before MaxPool1d(2)
                    w8[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
                }
            }
            if constexpr (EXTRA) {
                if (p.y32 != nullptr && ok) {
                    float4 *yp = reinterpret_cast<float4 *>(p.y32 + orow0 * COUT + c0);
                    yp[0] = make_float4(w8[0], w8[1], w8[2], w8[3]);
                    yp[1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
                }
                if (p.post_scale != nullptr) {
                    const float4 s0 = *reinterpret_cast<const float4 *>(&s_psc[n0]), s1 = *reinterpret_cast<const float4 *>(&s_psc[n0 + 4]);
                    const float4 h0 = *reinterpret_cast<const float4 *>(&s_psh[n0]), h1 = *reinterpret_cast<const float4 *>(&s_psh[n0 + 4]);
                    w8[0] = fmaxf(fmaf(w8[0], s0.x, h0.x), 0.f), w8[1] = fmaxf(fmaf(w8[1], s0.y, h0.y), 0.f);
                    w8[2] = fmaxf(fmaf(w8[2], s0.z, h0.z), 0.f), w8[3] = fmaxf(fmaf(w8[3], s0.w, h0.w), 0.f);
                    w8[4] = fmaxf(fmaf(w8[4], s1.x, h1.x), 0.f), w8[5] = fmaxf(fmaf(w8[5], s1.y, h1.y), 0.f);
                    w8[6] = fmaxf(fmaf(w8[6], s1.z, h1.z), 0.f), w8[7] = fmaxf(fmaf(w8[7], s1.w, h1.w), 0.f);
                }
            }
            uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
            if (ok) pack8_split16<SPLIT>(w8, hi, lo);
            uint16_t *yb = y16 + phi * COUT + c0;
            if constexpr (RC >= 16 && COUT % 16 == 0) {
                // two adjacent 8-channel groups (same row, same phase, same `ok`) leave as one 32-byte store per split:
                // whole sectors, half the L1 / L2 transactions of 16-byte pieces at a row-pitch lane stride
                if ((g8 & 8) == 0) {
                    pend_hi = hi;
                    pend_lo = lo;
                } else if (ok) {
                    if (p.st256) {
                        st_global_256(yb - 8, pend_hi, hi);
                        if (SPLIT == 2) st_global_256(yb - 8 + p.y_split, pend_lo, lo);
                    } else {
                        *reinterpret_cast<uint4 *>(yb - 8) = pend_hi;
                        *reinterpret_cast<uint4 *>(yb) = hi;
                        if (SPLIT == 2) {
                            *reinterpret_cast<uint4 *>(yb - 8 + p.y_split) = pend_lo;
                            *reinterpret_cast<uint4 *>(yb + p.y_split) = lo;
                        }
                    }
                }
            } else if (ok) {
                *reinterpret_cast<uint4 *>(yb) = hi;
                if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + p.y_split) = lo;
            }
        }
    }
}

// FLAGS == 8: 'same' Conv1d on the [T / 4][4 C_in] view of the channel-last input (time folded into the MMA K / N: column
// = sample * 16 + channel, 4 samples x 16 output channels per accumulator row) + ReLU + MaxPool1d(2).  The pool pairs
// (samples 2j, 2j + 1) live in the same accumulator row, so no shuffle is needed; a thread writes the 16 pooled channels
// of pooled sample j as 32 contiguous bytes per split of the [T / 4][2 x 16] = [T / 2][16] output.  T % 4 == 0: a row is
// entirely inside or outside the sequence.  EW = 4: both pooled samples per warp; EW = 8: pooled sample = column half.
template <int SPLIT, int EW>
__device__ __forceinline__ void tc_epilogue_foldpool(const TcP &p, uint32_t trow, int half, int seq, int srow, bool row_ok,
                                                     const float *s_bias) {
    static_assert(EW == 4 || EW == 8, "fold + pool epilogue: one or two warps per TMEM lane quarter");
    constexpr int PER = EW == 4 ? 2 : 1;
    uint16_t *y16 = reinterpret_cast<uint16_t *>(p.y) + ((int64_t)seq * p.T_out + srow) * 32;
#pragma unroll
    for (int jj = 0; jj < PER; ++jj) {
        const int j = half * PER + jj;
        uint32_t r[32];
        {
            uint32_t(&r0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
            uint32_t(&r1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[16]);
            tmem_ld16_nowait(trow + (uint32_t)(32 * j), r0);
            tmem_ld16_nowait(trow + (uint32_t)(32 * j + 16), r1);
        }
        tmem_ld_wait();
        if (!row_ok) continue;
        float w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 ba = *reinterpret_cast<const float4 *>(&s_bias[32 * j + 4 * q]);
            const float4 bb = *reinterpret_cast<const float4 *>(&s_bias[32 * j + 16 + 4 * q]);
            w[4 * q + 0] = fmaxf(fmaxf(__uint_as_float(r[4 * q + 0]) + ba.x, __uint_as_float(r[16 + 4 * q + 0]) + bb.x), 0.f);
            w[4 * q + 1] = fmaxf(fmaxf(__uint_as_float(r[4 * q + 1]) + ba.y, __uint_as_float(r[16 + 4 * q + 1]) + bb.y), 0.f);
            w[4 * q + 2] = fmaxf(fmaxf(__uint_as_float(r[4 * q + 2]) + ba.z, __uint_as_float(r[16 + 4 * q + 2]) + bb.z), 0.f);
            w[4 * q + 3] = fmaxf(fmaxf(__uint_as_float(r[4 * q + 3]) + ba.w, __uint_as_float(r[16 + 4 * q + 3]) + bb.w), 0.f);
        }
        uint4 h0, l0, h1, l1;
        pack8_split16<SPLIT>(&w[0], h0, l0);
        pack8_split16<SPLIT>(&w[8], h1, l1);
        uint16_t *yb = y16 + 16 * j;
        if (p.st256) {
            st_global_256(yb, h0, h1);
            if (SPLIT == 2) st_global_256(yb + p.y_split, l0, l1);
        } else {
            *reinterpret_cast<uint4 *>(yb) = h0;
            *reinterpret_cast<uint4 *>(yb + 8) = h1;
            if (SPLIT == 2) {
                *reinterpret_cast<uint4 *>(yb + p.y_split) = l0;
                *reinterpret_cast<uint4 *>(yb + 8 + p.y_split) = l1;
            }
        }
    }
}

// NTAPS > 0: the tile's MMA schedule is fixed at compile time (umma_conv_tile); NTAPS == 0: generic
// instance that walks the host-built schedule table (any layer shape; slower issue).
// FLAGS >= 0: epilogue options fixed at compile time (tc_epilogue_fixed); FLAGS < 0: run-time options (any layer).
template <int NOUT, int SPLIT, int NTAPS, int NQ, int EW, int FLAGS>
__global__ void __launch_bounds__(32 * (EW + 4), EW == 16 ? 1 : EW == 8 ? 2 : 4) tcconv_kernel(const __grid_constant__ TcP p) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], accf_bar[2], acce_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[NOUT], s_psc[NOUT], s_psh[NOUT];
    constexpr int NCOLS = NOUT < 32 ? 32 : NOUT;
    constexpr int THREADS = 32 * (EW + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const int n_rows = p.n_rows, cin8 = p.cin8, nstage = p.n_stages;
    const uint32_t PL = (uint32_t)p.pitch * 16u;  // bytes per 8-channel plane
    const uint32_t a_bytes = ((uint32_t)SPLIT * cin8 * PL + 127u) & ~127u;
    const uint32_t w_bytes = ((uint32_t)p.n_blocks * SPLIT * 2 * NOUT * 16u + 127u) & ~127u;
    const uint32_t sB_u = smem_u32(tc_smem);
    const uint32_t sA_u = sB_u + w_bytes;
    const int n_tiles = (int)(((int64_t)p.NS * p.Tp + 127) / 128);

    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&full_bar[i], TC_PRODUCERS);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accf_bar[i], 1);
            mbar_init(&acce_bar[i], 32 * EW);
        }
        fence_barrier_init();
    }
    if (warp == EW) tmem_alloc(&tmem_base_s, 2 * NCOLS);
    {   // resident weights, bias, optional post-affine of the 16-bit output
        const int wpieces = p.n_blocks * SPLIT * 2 * NOUT;
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.w + (int64_t)g * p.w_gs);
        for (int idx = tid; idx < wpieces; idx += THREADS) cp_async16(sB_u + (uint32_t)idx * 16u, wg + idx, 16u);
        for (int idx = tid; idx < NOUT; idx += THREADS) {
            s_bias[idx] = __ldg(p.bias + (int64_t)g * p.b_gs + idx);
            s_psc[idx] = p.post_scale ? __ldg(p.post_scale + idx) : 1.f;
            s_psh[idx] = p.post_shift ? __ldg(p.post_shift + idx) : 0.f;
        }
        cp_async_wait_all();
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp > EW) {
        // ================= producers =================
        const int ptid = tid - 32 * (EW + 1);
        const int cinA8 = p.cinA8;
        const int CA = cinA8 * 8, CB = (cin8 - cinA8) * 8;
        const uint16_t *xg = p.x + (int64_t)g * p.x_gs;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            const int m0 = tile * 128;
            const uint32_t sbase = sA_u + (uint32_t)stage * a_bytes;
            for (int r = ptid; r < n_rows && !(p.dbg & 1); r += TC_PRODUCERS) {
                const int v = m0 + p.row0 + r;
                bool valid = v >= 0;
                int seq = 0, u = 0;
                if (valid) {
                    seq = v / p.Tp;
                    u = v - seq * p.Tp;
                    valid = (seq < p.NS) && (u < p.T_eff);
                }
                const int srow = (p.ups == 2) ? (u >> 1) : u;
                const uint16_t *src = valid ? (xg + ((int64_t)seq * p.x_pitch + p.x_roff + srow) * CA) : xg;
                const uint16_t *src2 = (valid && CB) ? (p.x2 + ((int64_t)seq * p.x2_pitch + p.x2_roff + srow) * CB) : xg;
                const uint32_t nb = valid ? 16u : 0u;
#pragma unroll
                for (int s = 0; s < SPLIT; ++s) {
                    const uint16_t *ss = valid ? (src + (int64_t)s * p.x_split) : xg;
                    const uint16_t *s2 = (valid && CB) ? (src2 + (int64_t)s * p.x2_split) : xg;
                    if (p.ca) {
                        for (int c = 0; c < cinA8; ++c)
                            cp_async16_ca(sbase + (uint32_t)(s * cin8 + c) * PL + (uint32_t)r * 16u, ss + c * 8, nb);
                        for (int c = cinA8; c < cin8; ++c)
                            cp_async16_ca(sbase + (uint32_t)(s * cin8 + c) * PL + (uint32_t)r * 16u, s2 + (c - cinA8) * 8, nb);
                    } else {
                        for (int c = 0; c < cinA8; ++c)
                            cp_async16(sbase + (uint32_t)(s * cin8 + c) * PL + (uint32_t)r * 16u, ss + c * 8, nb);
                        for (int c = cinA8; c < cin8; ++c)
                            cp_async16(sbase + (uint32_t)(s * cin8 + c) * PL + (uint32_t)r * 16u, s2 + (c - cinA8) * 8, nb);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(&full_bar[stage]);
            if (++stage == nstage) {
                stage = 0;
                phase ^= 1u;
            }
        }
        cp_async_wait_all();
    } else if (warp == EW) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers);
        // one elected lane issues.
        const uint32_t idesc = umma_idesc(NOUT, p.fmt16);
        const uint32_t sB16 = sB_u >> 4;
        const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1 (Blackwell), SBO = 128 B
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const int n_terms = p.n_terms;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&acce_bar[acc], acc_phase ^ 1u);
            mbar_wait(&full_bar[stage], phase);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t sA16 = (sA_u + (uint32_t)stage * a_bytes) >> 4;
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NCOLS);
            if (elect_one()) {
                if constexpr (NTAPS > 0) {
                    if (p.dbg & 2) umma_f16(d_tmem, desc_hi | (uint64_t)(sA16 + p.term_a[0]), desc_hi | (uint64_t)(sB16 + p.term_b[0]), idesc, 0u); else {
                    umma_conv_tile<NOUT, SPLIT, NTAPS, NQ>(d_tmem, sA16, (uint32_t)p.pitch, sB16, idesc, 0u);
                    }
                } else {
                    umma_f16(d_tmem, desc_hi | (uint64_t)(sA16 + p.term_a[0]), desc_hi | (uint64_t)(sB16 + p.term_b[0]), idesc, 0u);
#pragma unroll 4
                    for (int i = 1; i < ((p.dbg & 2) ? 1 : n_terms); ++i)
                        umma_f16(d_tmem, desc_hi | (uint64_t)(sA16 + p.term_a[i]), desc_hi | (uint64_t)(sB16 + p.term_b[i]), idesc, 1u);
                }
                umma_commit(&empty_bar[stage]);
                umma_commit(&accf_bar[acc]);
            }
            __syncwarp();
            if (++stage == nstage) {
                stage = 0;
                phase ^= 1u;
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ================= epilogue =================
        constexpr int NHALF = EW / 4;                  // column splits per lane quarter
        constexpr int COLS = NOUT / NHALF;             // columns per epilogue warp
        constexpr int CH = COLS >= 16 ? 16 : 8;        // columns per TMEM load
        const int quarter = warp & 3, half = warp >> 2;
        const int coutp = p.coutp;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&accf_bar[acc], acc_phase);
            tc_fence_after();
            const int v = tile * 128 + quarter * 32 + lane;
            const int seq = v / p.Tp;
            const int srow = v - seq * p.Tp;
            const bool row_ok = (seq < p.NS) && (srow < p.T_valid);
            const uint32_t trow = tmem_base + (uint32_t)(acc * NCOLS) + ((uint32_t)(quarter * 32) << 16);
            // output row bases (polyphase: row 2 srow + phi; pool: row srow >> 1, written by the even lane)
            const bool st = (p.pool == 2) ? (row_ok && !(lane & 1)) : row_ok;
            const int t0 = (p.pool == 2) ? (srow >> 1) : p.ph * srow;
            const int64_t orow0 = (int64_t)seq * p.y_pitch + p.y_roff + t0;
            if constexpr (FLAGS == 8) {
                if (!(p.dbg & 4)) tc_epilogue_foldpool<SPLIT, EW>(p, trow, half, seq, srow, row_ok, s_bias);
            } else if constexpr (FLAGS >= 0) {
                if (!(p.dbg & 4)) tc_epilogue_fixed<NOUT, SPLIT, EW, FLAGS>(p, trow, half, lane, seq, srow, row_ok, g, s_bias, s_psc, s_psh);
            } else if (!(p.dbg & 4)) {
#pragma unroll
                for (int cc = 0; cc < COLS; cc += CH) {
                    const int nb = half * COLS + cc;  // first accumulator column of this chunk
                    float a[CH];
                    {
                        uint32_t r[CH];
                        if constexpr (CH == 16) {
                            uint32_t(&r16)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
                            tmem_ld16_nowait(trow + (uint32_t)nb, r16);
                        } else {
                            uint32_t(&r8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&r[0]);
                            tmem_ld8_nowait(trow + (uint32_t)nb, r8);
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < CH; ++i) a[i] = __uint_as_float(r[i]);
                    }
                    uint16_t *pend = nullptr;  // 16-bit output: even 8-channel group waiting for its neighbour
                    uint4 pend_hi = make_uint4(0u, 0u, 0u, 0u), pend_lo = pend_hi;
#pragma unroll
                    for (int g8 = 0; g8 < CH; g8 += 8) {
                        const int n0 = nb + g8;
                        const int phi = (n0 >= coutp) ? 1 : 0;
                        const int c0 = n0 - phi * coutp;
                        if (phi >= p.ph || c0 >= p.cout) continue;  // padding columns (warp-uniform)
                        float *w8 = &a[g8];
                        const float4 b0 = *reinterpret_cast<const float4 *>(&s_bias[n0]), b1 = *reinterpret_cast<const float4 *>(&s_bias[n0 + 4]);
                        w8[0] += b0.x, w8[1] += b0.y, w8[2] += b0.z, w8[3] += b0.w;
                        w8[4] += b1.x, w8[5] += b1.y, w8[6] += b1.z, w8[7] += b1.w;
                        int t_out = t0 + phi;
                        bool ok = st && t_out < p.T_out;
                        bool fold_zero = false;  // folded output sample outside the sequence: 16-bit rows get zeros
                        if (p.fold > 1) {
                            t_out = p.fold * srow + n0 / p.fold_c - p.fold_o;
                            fold_zero = t_out < 0 || t_out >= p.fold_T;
                        }
                        if (p.res != nullptr && ok) {  // residual stream, fp32 row-major [seq][t][cout]
                            const float4 *rp = reinterpret_cast<const float4 *>(p.res + (orow0 + phi) * p.cout + c0);
                            const float4 r0 = rp[0], r1 = rp[1];  // plain loads: y32 may alias res (in-place residual stream)
                            w8[0] += r0.x, w8[1] += r0.y, w8[2] += r0.z, w8[3] += r0.w;
                            w8[4] += r1.x, w8[5] += r1.y, w8[6] += r1.z, w8[7] += r1.w;
                        }
                        if (p.act == ACT_RELU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = fmaxf(w8[i], 0.f);
                        } else if (p.act == ACT_SIGMOID) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = 1.f / (1.f + expf(-w8[i]));
                        }
                        if (p.pool == 2) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float mine = row_ok ? w8[i] : -1e10f;  // SeisBench pads odd lengths with -1e10 before MaxPool1d(2)
                                w8[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
                            }
                        }
                        if (p.y32 != nullptr && ok) {  // second output: the layer's fp32 value, row-major [seq][t][cout]
                            float4 *yp = reinterpret_cast<float4 *>(p.y32 + (orow0 + phi) * p.cout + c0);
                            yp[0] = make_float4(w8[0], w8[1], w8[2], w8[3]);
                            yp[1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
                        }
                        if (p.post_scale != nullptr) {  // pre-activation BatchNorm + ReLU of the next conv, applied to the main output
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = fmaxf(fmaf(w8[i], s_psc[n0 + i], s_psh[n0 + i]), 0.f);
                        }
                        if (!ok || p.out_fmt == 3) continue;  // out_fmt 3: the fp32 second output was the only one
                        if (fold_zero) {
                            if (p.out_fmt != 0) continue;
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = 0.f;
                        }
                        if (p.out_fmt == 2) {  // PhaseNet head: 1x1 conv (8 -> 3) + softmax over the classes, fp32 (seq, 3, T)
                            float z[3];
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                float a0 = p.head_b[j];
#pragma unroll
                                for (int i = 0; i < 8; ++i) a0 = fmaf(p.head_w[j * 8 + i], w8[i], a0);
                                z[j] = a0;
                            }
                            const float zm = fmaxf(z[0], fmaxf(z[1], z[2]));
                            const float e0 = expf(z[0] - zm), e1 = expf(z[1] - zm), e2 = expf(z[2] - zm);
                            const float inv = 1.f / (e0 + e1 + e2);
                            float *yb = reinterpret_cast<float *>(p.y) + (int64_t)seq * p.y_ss + t_out;
                            yb[0] = e0 * inv;
                            yb[p.y_cs] = e1 * inv;
                            yb[2 * p.y_cs] = e2 * inv;
                        } else if (p.out_fmt == 0) {
                            uint4 hi, lo;
                            tc_pack8<SPLIT>(w8, hi, lo);
                            uint16_t *yb = reinterpret_cast<uint16_t *>(p.y) + (int64_t)g * p.y_gs + (orow0 + phi) * p.cout_cl + c0;
                            if (CH == 16 && g8 == 0 && p.st256) {  // wait for the neighbouring group: one 32-byte store when it follows
                                pend = yb;
                                pend_hi = hi;
                                pend_lo = lo;
                            } else if (pend != nullptr) {
                                st_pair16(pend, pend_hi, yb, hi);
                                if (SPLIT == 2) st_pair16(pend + p.y_split, pend_lo, yb + p.y_split, lo);
                                pend = nullptr;
                            } else {
                                *reinterpret_cast<uint4 *>(yb) = hi;
                                if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + p.y_split) = lo;
                            }
                        } else {
                            float *yb = reinterpret_cast<float *>(p.y) + (int64_t)g * p.y_gs + (int64_t)seq * p.y_ss + t_out;
                            const int cb = p.fold > 1 ? c0 % p.fold_c : c0;  // folded: channel inside the sample
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (c0 + i < p.cout) yb[(int64_t)(cb + i) * p.y_cs] = w8[i];
                        }
                    }
                    if (pend != nullptr) {  // the neighbouring group was a padding / masked column group
                        *reinterpret_cast<uint4 *>(pend) = pend_hi;
                        if (SPLIT == 2) *reinterpret_cast<uint4 *>(pend + p.y_split) = pend_lo;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acce_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW) tmem_dealloc(tmem_base, 2 * NCOLS);
}

// ------------------------------------------------------------------------------------------ pack kernel
// fp32 channel-first (NS, C, T) -> channel-last 16-bit rows [split][NS][T][c8*8] (zero padded channels)
__global__ void __launch_bounds__(128) pack_cl16_kernel(const float *__restrict__ x, int64_t x_ss, int64_t x_cs, int NS, int C,
                                                        int T, int split, int fmt16, uint16_t *__restrict__ y, int64_t y_split,
                                                        int c8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (seq, t)
    if (i >= (int64_t)NS * T) return;
    const int64_t seq = i / T;
    const int t = (int)(i - seq * T);
    const float *xb = x + seq * x_ss + t;
    for (int q = 0; q < c8; ++q) {
        uint16_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = q * 8 + e;
            const float v = (c < C) ? __ldg(xb + (int64_t)c * x_cs) : 0.f;
            split16(v, fmt16, split, hi[e], lo[e]);
        }
        uint16_t *yb = y + (i * c8 + q) * 8;
        uint4 h4, l4;
        h4.x = hi[0] | ((uint32_t)hi[1] << 16);
        h4.y = hi[2] | ((uint32_t)hi[3] << 16);
        h4.z = hi[4] | ((uint32_t)hi[5] << 16);
        h4.w = hi[6] | ((uint32_t)hi[7] << 16);
        *reinterpret_cast<uint4 *>(yb) = h4;
        if (split == 2) {
            l4.x = lo[0] | ((uint32_t)lo[1] << 16);
            l4.y = lo[2] | ((uint32_t)lo[3] << 16);
            l4.z = lo[4] | ((uint32_t)lo[5] << 16);
            l4.w = lo[6] | ((uint32_t)lo[7] << 16);
            *reinterpret_cast<uint4 *>(yb + y_split) = l4;
        }
    }
}

int launch_pack_cl16(const float *x, int64_t x_ss, int64_t x_cs, int NS, int C, int T, int split, uint16_t *y, int64_t y_split,
                     int c8, cudaStream_t s) {
    const int64_t n = (int64_t)NS * T;
    if (n == 0) return VP_OK;
    KTimer kt(KC_PACK, s);
    pack_cl16_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(x, x_ss, x_cs, NS, C, T, split, split == 2 ? 0 : 1, y, y_split, c8);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

// channel-last 16-bit rows [split][NS][T][C] -> fp32 channel-first (NS, C, T) (layer-level tests of 16-bit outputs)
__global__ void unpack_cl16_kernel(const uint16_t *__restrict__ x, int64_t x_split, int split, int fmt16, int NS, int C, int T,
                                   float *__restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // ((seq, t), c)
    if (i >= (int64_t)NS * T * C) return;
    const int c = (int)(i % C);
    const int64_t st = i / C, seq = st / T;
    const int t = (int)(st - seq * T);
    float v;
    if (fmt16 == 0) {
        v = __half2float(__ushort_as_half(x[i]));
        if (split == 2) v += __half2float(__ushort_as_half(x[i + x_split]));
    } else {
        v = __bfloat162float(__ushort_as_bfloat16(x[i]));
    }
    y[(seq * C + c) * T + t] = v;
}

// ------------------------------------------------------------------------------------------ host: layer builder
static inline int floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

static void to16(float w, int split, uint16_t &hi, uint16_t &lo) {
    if (split == 2) {
        const __half h = __float2half_rn(w);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(w - __half2float(h)));
    } else {
        hi = __bfloat16_as_ushort(__float2bfloat16_rn(w));
        lo = 0;
    }
}

// 'same' Conv1d (odd k <= 9, pad (k - 1) / 2) with 4 time steps folded into the channels: on the [T / 4][4 C] view of a
// channel-last buffer (the same bytes) it is a 3-tap row conv (row pad 1) from 4 cin to 4 cout columns; column = sample *
// cout + channel.  The 8 / 16-channel encoder layers are bound by the shared-memory reads of the A operand (4 KB per
// tcgen05.mma whatever its N), so 2.5x / 1.75x fewer, wider MMAs win.
void tc_fold4_same(const float *W, const float *bias, int cout, int cin, int k, std::vector<float> &wf, std::vector<float> &bf) {
    const int F = 4, P = (k - 1) / 2;
    wf.assign((size_t)F * cout * F * cin * 3, 0.f);
    bf.assign((size_t)F * cout, 0.f);
    for (int q = 0; q < F; ++q)
        for (int co = 0; co < cout; ++co) {
            bf[(size_t)q * cout + co] = bias ? bias[co] : 0.f;
            for (int j = 0; j < k; ++j) {
                const int pos = q + j - P;  // input sample relative to 4 * row
                const int d = pos >= 0 ? pos / F : -((-pos + F - 1) / F), lane = pos - d * F;  // row tap d in {-1, 0, 1}
                for (int ci = 0; ci < cin; ++ci)
                    wf[(((size_t)q * cout + co) * (F * cin) + lane * cin + ci) * 3 + (d + 1)] = W[((size_t)co * cin + ci) * k + j];
            }
        }
}

int tc_build_layer(TcLayer &L, int mode, int cin, int cout, int k, int crop, int split, int groups,
                   const float *const *weights, const float *const *bias, int pad_left) {
    VP_REQUIRE(cin % 8 == 0 && (cin == 8 || cin % 16 == 0), VP_ERR_UNSUPPORTED, "tc conv: cin %d unsupported", cin);
    VP_REQUIRE(!(mode == TC_POLYPHASE && (cin == 8 || crop != 0)), VP_ERR_UNSUPPORTED, "tc conv: polyphase needs cin>=16, no crop");
    L = TcLayer();
    L.cin = cin;
    L.cout = cout;
    L.k = k;
    L.split = split;
    L.groups = groups;
    L.crop = crop;
    L.ph = (mode == TC_POLYPHASE) ? 2 : 1;
    L.ups = (mode == TC_DIRECT_UPS) ? 2 : 1;
    const int coutp = (cout + 7) / 8 * 8;
    int nout = L.ph * coutp;
    nout = nout <= 16 ? 16 : nout <= 32 ? 32 : nout <= 64 ? 64 : 128;
    VP_REQUIRE(L.ph * coutp <= 128, VP_ERR_UNSUPPORTED, "tc conv: %d output columns exceed 128", L.ph * coutp);
    L.nout = nout;
    const int p = (pad_left >= 0) ? pad_left : k / 2;  // zeros on the left of the input ('same' conv: k / 2)
    VP_REQUIRE(pad_left < 0 || mode == TC_DIRECT, VP_ERR_UNSUPPORTED, "tc conv: explicit left padding needs the direct form");
    const int nq = cin / 16;
    // effective taps: weff[phi][tap][co][ci]
    int ntaps, o_min;
    if (mode == TC_POLYPHASE) {
        o_min = floordiv2(0 - p);
        const int o_max = floordiv2(1 + (k - 1) - p);
        ntaps = o_max - o_min + 1;
    } else {
        o_min = -p;
        ntaps = k;
    }
    std::vector<float> weff((size_t)L.ph * ntaps * cout * cin);
    const int npairs = (ntaps + 1) / 2;
    L.n_blocks = (cin == 8) ? npairs : ntaps * nq;
    L.row0 = o_min;
    L.halo = (cin == 8) ? (2 * npairs - 1) : (ntaps - 1);
    VP_REQUIRE(L.n_blocks <= TC_MAX_MMA, VP_ERR_UNSUPPORTED, "tc conv: %d MMAs per tile exceed %d", L.n_blocks, TC_MAX_MMA);
    const size_t blk = (size_t)split * 2 * nout * 8;
    L.blocks.assign((size_t)groups * L.n_blocks * blk, 0);
    L.bias.assign((size_t)groups * nout, 0.f);
    for (int g = 0; g < groups; ++g) {
        const float *W = weights[g];
        std::fill(weff.begin(), weff.end(), 0.f);
        for (int phi = 0; phi < L.ph; ++phi)
            for (int kk = 0; kk < k; ++kk) {
                const int o = (mode == TC_POLYPHASE) ? floordiv2(phi + kk - p) : (kk - p);
                const int j = o - o_min;
                for (int co = 0; co < cout; ++co)
                    for (int ci = 0; ci < cin; ++ci)
                        weff[(((size_t)phi * ntaps + j) * cout + co) * cin + ci] += W[((size_t)co * cin + ci) * k + kk];
            }
        uint16_t *B = L.blocks.data() + (size_t)g * L.n_blocks * blk;
        auto put = [&](int block, int kh, int n, int e, float w) {
            uint16_t hi, lo;
            to16(w, split, hi, lo);
            B[(size_t)block * blk + ((size_t)(0 * 2 + kh) * nout + n) * 8 + e] = hi;
            if (split == 2) B[(size_t)block * blk + ((size_t)(1 * 2 + kh) * nout + n) * 8 + e] = lo;
        };
        for (int phi = 0; phi < L.ph; ++phi)
            for (int co = 0; co < cout; ++co) {
                const int n = phi * coutp + co;
                if (cin == 8) {
                    for (int jp = 0; jp < npairs; ++jp)
                        for (int kh = 0; kh < 2; ++kh) {
                            const int j = 2 * jp + kh;
                            if (j >= ntaps) continue;
                            for (int e = 0; e < 8; ++e) put(jp, kh, n, e, weff[(((size_t)phi * ntaps + j) * cout + co) * cin + e]);
                        }
                } else {
                    for (int j = 0; j < ntaps; ++j)
                        for (int q = 0; q < nq; ++q)
                            for (int kh = 0; kh < 2; ++kh)
                                for (int e = 0; e < 8; ++e)
                                    put(j * nq + q, kh, n, e, weff[(((size_t)phi * ntaps + j) * cout + co) * cin + 16 * q + 8 * kh + e]);
                }
                L.bias[(size_t)g * nout + n] = (bias && bias[g]) ? bias[g][co] : 0.f;
            }
    }
    L.sched_taps = (cin == 8) ? npairs : ntaps;
    L.sched_nq = (cin == 8) ? 0 : nq;
    L.mma.clear();
    if (cin == 8) {
        for (int jp = 0; jp < npairs; ++jp) L.mma.push_back(TcMma{2 * jp, 0, 1, jp});
    } else {
        for (int j = 0; j < ntaps; ++j)
            for (int q = 0; q < nq; ++q) L.mma.push_back(TcMma{j, 2 * q, 0, j * nq + q});
    }
    return VP_OK;
}

int tc_out_len(const TcLayer &L, int T_in, int pool) {
    const int T_eff = (L.ups == 2) ? 2 * T_in - L.crop : T_in;
    const int T_conv = (L.ph == 2) ? 2 * T_in : T_eff;
    return pool == 2 ? (T_conv + 1) / 2 : T_conv;
}

// CTAs per SM for a layer: as many (<= 4) as leave every CTA a >= 2-stage ring and 2 TMEM accumulators.
constexpr int tc_occupancy(size_t w_bytes, size_t a_bytes, int ncols2) {
    for (int o = 4; o > 1; --o) {
        const size_t share = (size_t)224 * 1024 / o - 3072;  // per CTA: 1 KB reserved by the driver + static shared memory
        if (o * ncols2 > 512 || w_bytes + 2 * a_bytes > share) continue;
        return o;
    }
    return 1;
}
constexpr size_t tc_up128(size_t v) { return (v + 127) & ~(size_t)127; }
// CTAs per SM of a compile-time layer shape; epilogue warps: 16 / 8 / 4 for 1 / 2 / more CTAs per SM
constexpr int tc_epi_warps_for(int occ, int nout) { return occ == 1 && nout >= 64 ? 16 : occ <= 2 ? 8 : 4; }
template <int NOUT, int SPLIT, int NTAPS, int NQ>
constexpr int tc_occ_fixed() {
    constexpr int blocks = NQ == 0 ? NTAPS : NTAPS * NQ;
    constexpr int cin8 = NQ == 0 ? 1 : 2 * NQ;
    constexpr int rows = 128 + (NQ == 0 ? 2 * NTAPS - 1 : NTAPS - 1);
    return tc_occupancy(tc_up128((size_t)blocks * SPLIT * 2 * NOUT * 16), tc_up128((size_t)SPLIT * cin8 * rows * 16),
                        2 * (NOUT < 32 ? 32 : NOUT));
}
template <int NOUT, int SPLIT, int NTAPS, int NQ>
constexpr int tc_epi_warps() {
    return tc_epi_warps_for(tc_occ_fixed<NOUT, SPLIT, NTAPS, NQ>(), NOUT);
}

template <int NOUT, int SPLIT, int NTAPS, int NQ, int EW, int FLAGS>
static int launch_tc(const TcP &p, dim3 grid, size_t smem, cudaStream_t s) {
    auto kern = tcconv_kernel<NOUT, SPLIT, NTAPS, NQ, EW, FLAGS>;
    if (int rc = ensure_dyn_smem((const void *)kern, smem)) return rc;
    KTimer kt(KC_TCCONV, s);
    // (programmatic dependent launch -- the prologue of a layer under the tail of the previous one -- was measured neutral for
    // both networks: PhaseNet 287 vs 290 station-days/s; not kept)
    kern<<<grid, 32 * (EW + 4), smem, s>>>(p);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

int tc_launch(const TcLayer &L, const TcIO &io, cudaStream_t s) {
    TcP p;
    std::memset(&p, 0, sizeof(p));
    p.x = io.x;
    p.x_split = io.x_split;
    p.x_gs = io.x_gs;
    p.T_in = io.T_in;
    p.x_pitch = io.x_pitch > 0 ? io.x_pitch : io.T_in;
    p.x_roff = io.x_roff;
    p.x2 = io.x2;
    p.x2_split = io.x2_split;
    p.x2_pitch = io.x2_pitch;
    p.x2_roff = io.x2_roff;
    p.cinA8 = io.x2 ? io.cin_a / 8 : L.cin / 8;
    VP_REQUIRE(!io.x2 || (io.cin_a % 8 == 0 && io.cin_a > 0 && io.cin_a < L.cin && L.cin != 8 && L.ups == 1), VP_ERR_UNSUPPORTED,
               "tc conv: bad two-source split (%d of %d channels)", io.cin_a, L.cin);
    p.ups = L.ups;
    p.T_eff = (L.ups == 2) ? 2 * io.T_in - L.crop : io.T_in;
    const int left = -L.row0, right = L.halo + L.row0;
    int Tp = p.T_eff + std::max(left, right);
    Tp += Tp & 1;
    p.Tp = Tp;
    p.NS = io.NS;
    p.row0 = L.row0;
    p.n_rows = 128 + L.halo;
    p.cin8 = L.cin / 8;
    p.pitch = p.n_rows;
    {
        static const int ca = getenv("VP_TC_CA") ? atoi(getenv("VP_TC_CA")) : 1;  // debugging aid
        p.ca = (ca && io.ca && p.cin8 >= 2) ? 1 : 0;  // rows of >= 32 bytes: two 16-byte pieces per sector
    }
    p.w = io.w_dev;
    p.w_gs = (int64_t)L.n_blocks * L.split * 2 * L.nout * 8;
    p.n_blocks = L.n_blocks;
    p.bias = io.b_dev;
    p.b_gs = L.nout;
    {
        const uint32_t PL16 = (uint32_t)p.pitch;  // plane pitch in 16-byte units
        const int nterm = (L.split == 2) ? 3 : 1;
        p.n_terms = 0;
        for (const TcMma &e : L.mma)
            for (int t = 0; t < nterm; ++t) {
                const int sa = (t == 2) ? 1 : 0;  // hi*hi, hi*lo, lo*hi
                const int sb = (t == 1) ? 1 : 0;
                const uint32_t a_off = (uint32_t)((sa * p.cin8 + e.a_plane) * p.pitch + e.a_row);
                const uint32_t lbo16 = e.a_rowk ? 1u : PL16;
                const uint32_t b_off = (uint32_t)((e.b_block * L.split + sb) * 2 * L.nout);
                p.term_a[p.n_terms] = a_off | (lbo16 << 16);
                p.term_b[p.n_terms] = b_off | ((uint32_t)L.nout << 16);
                ++p.n_terms;
            }
    }
    p.fmt16 = (L.split == 2) ? 0 : 1;
    {
        const char *e = getenv("VP_TC_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    p.act = io.act;
    p.pool = io.pool;
    p.ph = L.ph;
    p.cout = L.cout;
    p.coutp = (L.cout + 7) / 8 * 8;
    p.T_valid = (L.ph == 2) ? io.T_in : p.T_eff;
    p.T_out = tc_out_len(L, io.T_in, io.pool);
    if (io.T_valid > 0) {  // explicit number of output rows per sequence ('valid' convs of the reshaped stride-4 layers)
        VP_REQUIRE(L.ph == 1 && io.pool != 2 && io.T_valid <= Tp, VP_ERR_UNSUPPORTED, "tc conv: T_valid override needs a direct un-pooled layer");
        p.T_valid = io.T_valid;
        p.T_out = io.T_valid;
    }
    p.y_pitch = io.y_pitch > 0 ? io.y_pitch : p.T_out;
    p.y_roff = io.y_roff;
    if (io.out_fmt == 2) {
        VP_REQUIRE(io.head_w && io.head_b && (L.cout == 8 || (io.fold > 1 && io.fold_c == 8)) && L.ph == 1, VP_ERR_UNSUPPORTED,
                   "tc conv: the softmax head needs an 8-channel direct layer");
        std::memcpy(p.head_w, io.head_w, sizeof(p.head_w));
        std::memcpy(p.head_b, io.head_b, sizeof(p.head_b));
    }
    p.fold = io.fold > 1 ? io.fold : 1;
    p.fold_c = io.fold_c;
    p.fold_o = io.fold_o;
    p.fold_T = io.fold_T;
    VP_REQUIRE(p.fold == 1 || (L.ph == 1 && io.pool != 2 && io.fold_c % 8 == 0 && io.fold_c > 0 && L.cout == io.fold * io.fold_c &&
                               !io.res && !io.y32 && !io.post_scale),
               VP_ERR_UNSUPPORTED, "tc conv: folded output needs a plain direct layer with fold * fold_c output columns");
    const bool custom = p.fold > 1 || io.x2 || p.x_pitch != io.T_in || p.x_roff != 0 || p.y_pitch != p.T_out || p.y_roff != 0 || io.T_valid > 0 ||
                        io.out_fmt == 2;
    p.out_fmt = io.out_fmt;
    static const bool st256_off = getenv("VP_TC_ST256") && atoi(getenv("VP_TC_ST256")) == 0;  // debugging aid
    p.st256 = (!st256_off && io.out_fmt == 0 && reinterpret_cast<uintptr_t>(io.y) % 32 == 0 && (io.y_split * 2) % 32 == 0 && (io.y_gs * 2) % 32 == 0 &&
               L.cout % 16 == 0)
                  ? 1
                  : 0;
    p.y = io.y;
    p.y_split = io.y_split;
    p.y_gs = io.y_gs;
    p.y_ss = io.y_ss;
    p.y_cs = io.y_cs;
    p.cout_cl = io.cout_cl;
    p.post_scale = io.post_scale;
    p.post_shift = io.post_shift;
    p.res = io.res;
    p.y32 = io.y32;
    VP_REQUIRE(!(io.res || io.y32 || io.post_scale) || (L.ph == 1 && L.cout % 8 == 0 && L.groups == 1 && (!io.res || io.pool == 1)),
               VP_ERR_UNSUPPORTED, "tc conv: residual / post-affine / fp32 second output need a direct single-group layer");
    VP_REQUIRE(!(io.pool == 2 && L.ph == 2), VP_ERR_UNSUPPORTED, "tc conv: pooling with polyphase output is not supported");
    const int64_t rows = (int64_t)io.NS * Tp;
    const int64_t n_tiles = (rows + 127) / 128;
    const size_t a_bytes = ((size_t)L.split * p.cin8 * p.pitch * 16 + 127) & ~(size_t)127;
    const size_t w_bytes = ((size_t)L.n_blocks * L.split * 2 * L.nout * 16 + 127) & ~(size_t)127;
    // ring depth / residency: two CTAs per SM when a >= 2-stage ring fits in half the shared memory
    // ring depth / residency: several CTAs per SM (one MMA issuer each) when >= 2 stages fit in the share
    const size_t kFull = 224 * 1024;
    const int ncols2 = 2 * (L.nout < 32 ? 32 : L.nout);
    const int occ = tc_occupancy(w_bytes, a_bytes, ncols2);
    const size_t share = kFull / occ - 3072;
    VP_REQUIRE(w_bytes + a_bytes <= share, VP_ERR_UNSUPPORTED, "tc conv: %zu bytes of shared memory exceed the SM", w_bytes + a_bytes);
    const int stages = (int)std::min<size_t>(TC_MAX_STAGES, (share - w_bytes) / a_bytes);
    p.n_stages = stages;
    const size_t smem = w_bytes + (size_t)stages * a_bytes;
    int64_t ctas = ((int64_t)device_sm_count() * occ + L.groups - 1) / L.groups;
    ctas = std::max<int64_t>(1, std::min<int64_t>(ctas, n_tiles));
    dim3 grid((unsigned)ctas, L.groups);
    // compile-time MMA schedules for the layer shapes of the EQTransformer (N, taps | tap pairs, channel pairs)
    static const bool generic_only = getenv("VP_TC_GENERIC") && atoi(getenv("VP_TC_GENERIC")) != 0;
    const int st = L.sched_taps, sq = L.sched_nq;
    // epilogue options of this launch; the compile-time epilogue needs the common case: 16-bit output, ReLU (or the
    // res-CNN extras carrying the activation), no padding columns
    const bool extra = io.res || io.y32 || io.post_scale;
    const bool plain = !custom && io.out_fmt == 0 && L.nout == L.ph * L.cout && io.cout_cl == L.cout &&
                       (extra ? (io.act == ACT_RELU || io.act == ACT_NONE) : io.act == ACT_RELU);
    int flags = plain ? ((io.pool == 2 ? 1 : 0) | (L.ph == 2 ? 2 : 0) | (extra ? 4 : 0)) : -1;
    if (io.foldpool) {
        VP_REQUIRE(plain && flags == 0 && L.nout == 64 && st == 3 && (sq == 2 || sq == 4), VP_ERR_UNSUPPORTED,
                   "tc conv: fold + pool epilogue needs a plain 64-column 3-tap layer (N=%d, %d taps, %d pairs)", L.nout, st, sq);
        flags = 8;
        // ReLU + MaxPool inside the accumulator row; EW = 4 at >= 3 CTAs per SM (32 -> 4 x 16), 8 otherwise (64 -> 4 x 16)
        if (sq == 2) {
            if (L.split == 2) return launch_tc<64, 2, 3, 2, 4, 8>(p, grid, smem, s);
            return launch_tc<64, 1, 3, 2, 4, 8>(p, grid, smem, s);
        }
        if (L.split == 2) return launch_tc<64, 2, 3, 4, 8, 8>(p, grid, smem, s);
        return launch_tc<64, 1, 3, 4, 8, 8>(p, grid, smem, s);
    }
#define VP_TC_FIXED(N, T, Q, F)                                                                                  \
    if (!generic_only && L.nout == N && st == T && sq == Q && flags == F) {                                      \
        if (L.split == 2) return launch_tc<N, 2, T, Q, tc_epi_warps<N, 2, T, Q>(), F>(p, grid, smem, s);         \
        return launch_tc<N, 1, T, Q, tc_epi_warps<N, 1, T, Q>(), F>(p, grid, smem, s);                           \
    }
    VP_TC_FIXED(16, 5, 0, 1);   // encoder.convs.1 (8 -> 16, k9) + pool
    VP_TC_FIXED(16, 7, 1, 1);   // encoder.convs.2 (16 -> 16, k7) + pool
    VP_TC_FIXED(32, 7, 1, 1);   // encoder.convs.3 (16 -> 32, k7) + pool
    VP_TC_FIXED(32, 5, 2, 1);   // encoder.convs.4 (32 -> 32, k5) + pool
    VP_TC_FIXED(64, 5, 2, 1);   // encoder.convs.5 (32 -> 64, k5) + pool
    VP_TC_FIXED(64, 3, 4, 5);   // encoder.convs.6 (64 -> 64, k3) + pool -> residual stream + relu(bn(x))
    VP_TC_FIXED(64, 3, 4, 4);   // res-CNN k3 convs
    VP_TC_FIXED(64, 2, 4, 4);   // res-CNN k2 convs
    VP_TC_FIXED(128, 3, 1, 2);  // decoder.convs.0 polyphase (16 -> 2 x 64)
    VP_TC_FIXED(128, 3, 4, 2);  // decoder.convs.1 polyphase (64 -> 2 x 64)
    VP_TC_FIXED(32, 5, 4, 0);   // decoder.convs.2 (64 -> 32, k5, loader-side up-sampling)
#undef VP_TC_FIXED
    // (Compile-time MMA schedules with the run-time epilogue for PhaseNet's k = 7 concat convs and folded level-0 layers were
    // measured neutral, 3.41 vs 3.43 ms per station-day: their issue rate is not the limit.  They run the generic instances.)
    // generic instances with 8 epilogue warps (two per TMEM lane quarter) when at most two CTAs fit an SM: with four, the
    // TMEM -> convert -> store chain of a wide tile is the serial bottleneck of the CTA
    static const bool ew8_off = getenv("VP_TC_EW8") && atoi(getenv("VP_TC_EW8")) == 0;
#define VP_TC_CASE8(N, S) \
    if (!ew8_off && occ <= 2 && L.nout == N && L.split == S) return launch_tc<N, S, 0, 0, 8, -1>(p, grid, smem, s)
    VP_TC_CASE8(32, 2);
    VP_TC_CASE8(64, 2);
    VP_TC_CASE8(128, 2);
    VP_TC_CASE8(32, 1);
    VP_TC_CASE8(64, 1);
    VP_TC_CASE8(128, 1);
#undef VP_TC_CASE8
#define VP_TC_CASE(N, S) \
    if (L.nout == N && L.split == S) return launch_tc<N, S, 0, 0, 4, -1>(p, grid, smem, s)
    VP_TC_CASE(16, 2);
    VP_TC_CASE(32, 2);
    VP_TC_CASE(64, 2);
    VP_TC_CASE(128, 2);
    VP_TC_CASE(16, 1);
    VP_TC_CASE(32, 1);
    VP_TC_CASE(64, 1);
    VP_TC_CASE(128, 1);
#undef VP_TC_CASE
    set_error("tc conv: no instance for N=%d split=%d", L.nout, L.split);
    return VP_ERR_UNSUPPORTED;
}

}  // namespace vp

// ============================================================================================ debug C ABI
using namespace vp;

// Layer-level parity hook: one Conv1d through the tensor-core path.
//   x: device fp32 (NS, CIN, T_in); w_host: (COUT, CIN, K) fp32 on the HOST; y: device fp32 (NS, COUT, T_out)
//   mode: 0 direct 'same' conv, 1 x2 nearest up-sampling folded into the weights (polyphase),
//         2 direct conv on the x2 up-sampled (loader-side) input minus `crop` trailing samples,
//         3 the direct 'same' conv + ReLU + MaxPool1d(2) on the time-folded [T / 4][4 C] view (encoder.convs.1 / .2: 8 / 16 -> 16)
//   precision: VP_PREC_F16X3 | VP_PREC_BF16
extern "C" VP_API int vp_tcconv_debug(const float *x, int NS, int CIN, int T_in, const float *w_host, const float *bias_host,
                                      int COUT, int K, int mode, int crop, int act, int pool, int precision, float *y,
                                      void *stream) {
    VP_REQUIRE(x && w_host && y, VP_ERR_ARG, "vp_tcconv_debug: null pointer");
    VP_REQUIRE(precision == VP_PREC_F16X3 || precision == VP_PREC_BF16, VP_ERR_ARG, "vp_tcconv_debug: precision must be f16x3 or bf16");
    cudaStream_t s = (cudaStream_t)stream;
    const int split = precision == VP_PREC_F16X3 ? 2 : 1;
    const int c8 = (CIN + 7) / 8;
    const int cin_p = c8 * 8;
    // zero-pad the input channels to a multiple of 8 on the weight side as well
    std::vector<float> wp((size_t)COUT * cin_p * K, 0.f);
    for (int co = 0; co < COUT; ++co)
        for (int ci = 0; ci < CIN; ++ci)
            for (int kk = 0; kk < K; ++kk) wp[((size_t)co * cin_p + ci) * K + kk] = w_host[((size_t)co * CIN + ci) * K + kk];
    TcLayer L;
    const float *wl[1] = {wp.data()};
    const float *bl[1] = {bias_host};
    int rc;
    const bool foldpool = mode == 3;  // 'same' conv on the [T / 4][4 C] view + ReLU + MaxPool1d(2) inside the accumulator row
    std::vector<float> wf, bf;
    if (foldpool) {
        VP_REQUIRE(COUT == 16 && (cin_p == 8 || cin_p == 16) && T_in % 4 == 0 && pool == 2 && act == ACT_RELU && (K & 1) && K <= 9,
                   VP_ERR_UNSUPPORTED, "vp_tcconv_debug: mode 3 is the folded encoder layer (8 / 16 -> 16 channels, ReLU, pool 2, T %% 4 == 0)");
        tc_fold4_same(wp.data(), bias_host, COUT, cin_p, K, wf, bf);
        const float *wfl[1] = {wf.data()}, *bfl[1] = {bf.data()};
        rc = tc_build_layer(L, TC_DIRECT, 4 * cin_p, 4 * COUT, 3, 0, split, 1, wfl, bfl, 1);
    } else {
        rc = tc_build_layer(L, mode, cin_p, COUT, K, crop, split, 1, wl, bl);
    }
    if (rc != VP_OK) return rc;
    uint16_t *d_x = nullptr, *d_w = nullptr;
    float *d_b = nullptr;
    const int64_t xs = (int64_t)NS * T_in * cin_p;
    VP_CUDA_CHECK(cudaMalloc(&d_x, (size_t)split * xs * 2 + 64));
    VP_CUDA_CHECK(cudaMalloc(&d_w, L.blocks.size() * 2));
    VP_CUDA_CHECK(cudaMalloc(&d_b, L.bias.size() * 4));
    VP_CUDA_CHECK(cudaMemcpyAsync(d_w, L.blocks.data(), L.blocks.size() * 2, cudaMemcpyHostToDevice, s));
    VP_CUDA_CHECK(cudaMemcpyAsync(d_b, L.bias.data(), L.bias.size() * 4, cudaMemcpyHostToDevice, s));
    rc = launch_pack_cl16(x, (int64_t)CIN * T_in, T_in, NS, CIN, T_in, split, d_x, xs, c8, s);
    if (rc == VP_OK) {
        TcIO io;
        io.x = d_x;
        io.x_split = xs;
        io.x_gs = 0;
        io.T_in = T_in;
        io.NS = NS;
        io.w_dev = d_w;
        io.b_dev = d_b;
        io.act = act;
        io.pool = pool;
        io.out_fmt = 1;
        io.y = y;
        io.y_split = 0;
        io.y_gs = 0;
        const int T_out = tc_out_len(L, T_in, pool);
        io.y_ss = (int64_t)COUT * T_out;
        io.y_cs = T_out;
        io.cout_cl = 0;
        if (foldpool) {  // 16-bit channel-last output [split][NS][T / 4][2 x 16] = [NS][T / 2][16], unpacked to fp32 (NS, 16, T / 2)
            const int64_t ys = (int64_t)NS * (T_in / 2) * 16;
            uint16_t *d_y16 = nullptr;
            VP_CUDA_CHECK(cudaMalloc(&d_y16, (size_t)split * ys * 2 + 64));
            io.T_in = T_in / 4;
            io.pool = 1;
            io.foldpool = 1;
            io.out_fmt = 0;
            io.y = d_y16;
            io.y_split = ys;
            io.y_ss = 0;
            io.y_cs = 0;
            io.cout_cl = L.cout;
            rc = tc_launch(L, io, s);
            if (rc == VP_OK) {
                const int64_t n = ys;
                unpack_cl16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_y16, ys, split, split == 2 ? 0 : 1, NS, 16, T_in / 2, y);
            }
            cudaStreamSynchronize(s);
            cudaFree(d_y16);
        } else {
            rc = tc_launch(L, io, s);
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(d_x);
    cudaFree(d_w);
    cudaFree(d_b);
    if (rc == VP_OK && e != cudaSuccess) {
        set_error("vp_tcconv_debug: %s", cudaGetErrorString(e));
        return VP_ERR_CUDA;
    }
    return rc;
}


// Micro-benchmark of one tensor-core layer on synthetic data (device time per launch, CUDA events).
extern "C" VP_API int vp_tcconv_bench(int NS, int CIN, int T_in, int COUT, int K, int mode, int crop, int pool, int precision,
                                      int out_fmt, int iters, float *ms_out) {
    VP_REQUIRE(ms_out && iters > 0, VP_ERR_ARG, "vp_tcconv_bench: bad argument");
    const int split = precision == VP_PREC_F16X3 ? 2 : 1;
    const int c8 = (CIN + 7) / 8, cin_p = c8 * 8;
    std::vector<float> w((size_t)COUT * cin_p * K), b(COUT, 0.1f);
    for (size_t i = 0; i < w.size(); ++i) w[i] = 0.01f * (float)((int)(i * 2654435761u % 200) - 100) / 100.f;
    TcLayer L;
    const float *wl[1] = {w.data()}, *bl[1] = {b.data()};
    int rc = tc_build_layer(L, mode, cin_p, COUT, K, crop, split, 1, wl, bl);
    if (rc != VP_OK) return rc;
    const int T_out = tc_out_len(L, T_in, pool);
    const int cout_cl = (COUT + 7) / 8 * 8;
    uint16_t *d_x = nullptr, *d_w = nullptr, *d_y = nullptr;
    float *d_b = nullptr;
    const int64_t xs = (int64_t)NS * T_in * cin_p, ys = (int64_t)NS * T_out * cout_cl;
    VP_CUDA_CHECK(cudaMalloc(&d_x, (size_t)split * xs * 2 + 64));
    VP_CUDA_CHECK(cudaMalloc(&d_y, (size_t)std::max<int64_t>(split * ys * 2, ys * 4) + 64));
    VP_CUDA_CHECK(cudaMalloc(&d_w, L.blocks.size() * 2));
    VP_CUDA_CHECK(cudaMalloc(&d_b, L.bias.size() * 4));
    VP_CUDA_CHECK(cudaMemset(d_x, 0, (size_t)split * xs * 2));
    VP_CUDA_CHECK(cudaMemcpy(d_w, L.blocks.data(), L.blocks.size() * 2, cudaMemcpyHostToDevice));
    VP_CUDA_CHECK(cudaMemcpy(d_b, L.bias.data(), L.bias.size() * 4, cudaMemcpyHostToDevice));
    TcIO io;
    io.x = d_x;
    io.x_split = xs;
    io.x_gs = 0;
    io.T_in = T_in;
    io.NS = NS;
    io.w_dev = d_w;
    io.b_dev = d_b;
    io.act = ACT_RELU;
    io.pool = pool;
    io.out_fmt = out_fmt;
    io.y = d_y;
    io.y_split = ys;
    io.y_gs = 0;
    io.y_ss = (int64_t)COUT * T_out;
    io.y_cs = T_out;
    io.cout_cl = cout_cl;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2 && rc == VP_OK; ++i) rc = tc_launch(L, io, 0);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters && rc == VP_OK; ++i) rc = tc_launch(L, io, 0);
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_x);
    cudaFree(d_y);
    cudaFree(d_w);
    cudaFree(d_b);
    if (rc == VP_OK && e != cudaSuccess) {
        set_error("vp_tcconv_bench: %s", cudaGetErrorString(e));
        return VP_ERR_CUDA;
    }
    return rc;
}
