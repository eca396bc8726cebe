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
// Persistent, warp-specialised CTA (256 threads):
//   warps 0-3  epilogue: TMEM lane quarter -> registers -> bias/act/pool -> global
//   warp  4    MMA issuer (one elected lane) + TMEM allocation (2 accumulator buffers)
//   warps 5-7  producers: cp.async (zero-fill) of the A tiles into an NSTAGE ring
// Barriers: full[stage] (96 producer arrivals via cp.async.mbarrier.arrive.noinc), empty[stage]
// (tcgen05.commit), accf[acc] (tcgen05.commit), acce[acc] (128 epilogue arrivals).  The weights of the
// layer are loaded once per CTA and stay resident in shared memory.
constexpr int TC_THREADS = 256;
constexpr int TC_PRODUCERS = 96;
constexpr int TC_MAX_STAGES = 4;


template <int NOUT, int SPLIT>
__global__ void __launch_bounds__(TC_THREADS) tcconv_kernel(const __grid_constant__ TcP p) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], accf_bar[2], acce_bar[2];
    __shared__ uint32_t tmem_base_s;
    constexpr int NCOLS = NOUT < 32 ? 32 : NOUT;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const int n_rows = p.n_rows, cin8 = p.cin8, nstage = p.n_stages;
    const uint32_t PL = (uint32_t)n_rows * 16u;  // bytes per 8-channel plane
    const uint32_t a_bytes = ((uint32_t)SPLIT * cin8 * PL + 127u) & ~127u;
    const uint32_t w_bytes = ((uint32_t)p.n_blocks * SPLIT * 2 * NOUT * 16u + 127u) & ~127u;
    const uint32_t sB_u = smem_u32(tc_smem);
    const uint32_t sA_u = sB_u + w_bytes;
    const int64_t n_tiles = ((int64_t)p.NS * p.Tp + 127) / 128;

    if (tid == 0) {
        for (int i = 0; i < nstage; ++i) {
            mbar_init(&full_bar[i], TC_PRODUCERS);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&accf_bar[i], 1);
            mbar_init(&acce_bar[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_s, 2 * NCOLS);
    {   // resident weights
        const int wpieces = p.n_blocks * SPLIT * 2 * NOUT;
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.w + (int64_t)g * p.w_gs);
        for (int idx = tid; idx < wpieces; idx += TC_THREADS) cp_async16(sB_u + (uint32_t)idx * 16u, wg + idx, 16u);
        cp_async_wait_all();
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= 5) {
        // ================= producers =================
        const int ptid = tid - 160;
        const int CIN = cin8 * 8;
        const uint16_t *xg = p.x + (int64_t)g * p.x_gs;
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            const int64_t m0 = tile * 128;
            const uint32_t sbase = sA_u + (uint32_t)stage * a_bytes;
            for (int r = ptid; r < n_rows && !(p.dbg & 1); r += TC_PRODUCERS) {
                const int64_t v = m0 + p.row0 + r;
                bool valid = v >= 0;
                int64_t seq = 0;
                int u = 0;
                if (valid) {
                    seq = v / p.Tp;
                    u = (int)(v - seq * p.Tp);
                    valid = (seq < p.NS) && (u < p.T_eff);
                }
                const int srow = (p.ups == 2) ? (u >> 1) : u;
                const uint16_t *src = valid ? (xg + (seq * p.T_in + srow) * CIN) : xg;
                const uint32_t nb = valid ? 16u : 0u;
#pragma unroll
                for (int s = 0; s < SPLIT; ++s) {
                    const uint16_t *ss = valid ? (src + (int64_t)s * p.x_split) : xg;
                    for (int c = 0; c < cin8; ++c)
                        cp_async16(sbase + (uint32_t)((s * cin8 + c) * n_rows + r) * 16u, ss + c * 8, nb);
                }
            }
            cp_async_mbar_arrive_noinc(&full_bar[stage]);
            if (++stage == nstage) {
                stage = 0;
                phase ^= 1u;
            }
        }
        cp_async_wait_all();
    } else if (warp == 4) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers);
        // one elected lane issues.  Per MMA only two adds remain: the schedule was precomputed on the host.
        const uint32_t idesc = umma_idesc(NOUT, p.fmt16);
        const uint32_t sB16 = sB_u >> 4;
        const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1 (Blackwell), SBO = 128 B
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const int n_terms = p.n_terms;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&acce_bar[acc], acc_phase ^ 1u);
            mbar_wait(&full_bar[stage], phase);
            fence_proxy_async();
            tc_fence_after();
            const uint32_t sA16 = (sA_u + (uint32_t)stage * a_bytes) >> 4;
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NCOLS);
            if (elect_one()) {
                umma_f16(d_tmem, desc_hi | (uint64_t)(sA16 + p.term_a[0]), desc_hi | (uint64_t)(sB16 + p.term_b[0]), idesc, 0u);
#pragma unroll 4
                for (int i = 1; i < ((p.dbg & 2) ? 1 : n_terms); ++i)
                    umma_f16(d_tmem, desc_hi | (uint64_t)(sA16 + p.term_a[i]), desc_hi | (uint64_t)(sB16 + p.term_b[i]), idesc, 1u);
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
        const float *bias = p.bias + (int64_t)g * p.b_gs;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&accf_bar[acc], acc_phase);
            tc_fence_after();
            const int64_t v = tile * 128 + tid;
            const int64_t seq = v / p.Tp;
            const int srow = (int)(v - seq * p.Tp);
            const bool row_ok = (seq < p.NS) && (srow < p.T_valid);
            const uint32_t trow = tmem_base + (uint32_t)(acc * NCOLS) + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int n0 = 0; n0 < ((p.dbg & 4) ? 0 : NOUT); n0 += 8) {
                const int phi = n0 / p.coutp;
                const int c0 = n0 - phi * p.coutp;
                if (phi >= p.ph || c0 >= p.cout) continue;  // padding columns (warp-uniform)
                float a[8];
                tmem_ld8(trow + (uint32_t)n0, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t = a[i] + __ldg(bias + n0 + i);
                    if (p.act == ACT_RELU) t = fmaxf(t, 0.f);
                    if (p.act == ACT_SIGMOID) t = 1.f / (1.f + expf(-t));
                    a[i] = t;
                }
                int t_out = p.ph * srow + phi;
                bool st = row_ok;
                if (p.pool == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float mine = row_ok ? a[i] : -1e10f;  // SeisBench pads odd lengths with -1e10 before MaxPool1d(2)
                        a[i] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
                    }
                    st = row_ok && !(tid & 1);
                    t_out = srow >> 1;
                }
                if (!st || t_out >= p.T_out) continue;
                if (p.out_fmt == 0) {
                    uint16_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split16(a[i], p.fmt16, SPLIT, hi[i], lo[i]);
                    uint16_t *yb = reinterpret_cast<uint16_t *>(p.y) + (int64_t)g * p.y_gs + (seq * p.T_out + t_out) * p.cout_cl + c0;
                    uint4 ph4, pl4;
                    ph4.x = hi[0] | ((uint32_t)hi[1] << 16);
                    ph4.y = hi[2] | ((uint32_t)hi[3] << 16);
                    ph4.z = hi[4] | ((uint32_t)hi[5] << 16);
                    ph4.w = hi[6] | ((uint32_t)hi[7] << 16);
                    *reinterpret_cast<uint4 *>(yb) = ph4;
                    if (SPLIT == 2) {
                        pl4.x = lo[0] | ((uint32_t)lo[1] << 16);
                        pl4.y = lo[2] | ((uint32_t)lo[3] << 16);
                        pl4.z = lo[4] | ((uint32_t)lo[5] << 16);
                        pl4.w = lo[6] | ((uint32_t)lo[7] << 16);
                        *reinterpret_cast<uint4 *>(yb + p.y_split) = pl4;
                    }
                } else {
                    float *yb = reinterpret_cast<float *>(p.y) + (int64_t)g * p.y_gs + seq * p.y_ss + t_out;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c0 + i < p.cout) yb[(int64_t)(c0 + i) * p.y_cs] = a[i];
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
    if (warp == 4) tmem_dealloc(tmem_base, 2 * NCOLS);
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
    pack_cl16_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(x, x_ss, x_cs, NS, C, T, split, split == 2 ? 0 : 1, y, y_split, c8);
    VP_LAUNCH_CHECK();
    return VP_OK;
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

int tc_build_layer(TcLayer &L, int mode, int cin, int cout, int k, int crop, int split, int groups,
                   const float *const *weights, const float *const *bias) {
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
    const int p = k / 2;
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

template <int NOUT, int SPLIT>
static int launch_tc(const TcP &p, dim3 grid, size_t smem, cudaStream_t s) {
    auto kern = tcconv_kernel<NOUT, SPLIT>;
    static size_t attr = 0;
    if (smem > attr) {
        VP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    kern<<<grid, TC_THREADS, smem, s>>>(p);
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
    p.w = io.w_dev;
    p.w_gs = (int64_t)L.n_blocks * L.split * 2 * L.nout * 8;
    p.n_blocks = L.n_blocks;
    p.bias = io.b_dev;
    p.b_gs = L.nout;
    {
        const uint32_t PL16 = (uint32_t)p.n_rows;  // plane pitch in 16-byte units
        const int nterm = (L.split == 2) ? 3 : 1;
        p.n_terms = 0;
        for (const TcMma &e : L.mma)
            for (int t = 0; t < nterm; ++t) {
                const int sa = (t == 2) ? 1 : 0;  // hi*hi, hi*lo, lo*hi
                const int sb = (t == 1) ? 1 : 0;
                const uint32_t a_off = (uint32_t)((sa * p.cin8 + e.a_plane) * p.n_rows + e.a_row);
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
    p.out_fmt = io.out_fmt;
    p.y = io.y;
    p.y_split = io.y_split;
    p.y_gs = io.y_gs;
    p.y_ss = io.y_ss;
    p.y_cs = io.y_cs;
    p.cout_cl = io.cout_cl;
    VP_REQUIRE(!(io.pool == 2 && L.ph == 2), VP_ERR_UNSUPPORTED, "tc conv: pooling with polyphase output is not supported");
    const int64_t rows = (int64_t)io.NS * Tp;
    const int64_t n_tiles = (rows + 127) / 128;
    const size_t a_bytes = ((size_t)L.split * p.cin8 * p.n_rows * 16 + 127) & ~(size_t)127;
    const size_t w_bytes = ((size_t)L.n_blocks * L.split * 2 * L.nout * 16 + 127) & ~(size_t)127;
    // ring depth / residency: two CTAs per SM when a >= 2-stage ring fits in half the shared memory
    // ring depth / residency: several CTAs per SM (one MMA issuer each) when >= 2 stages fit in the share
    const size_t kFull = 224 * 1024;
    const int ncols2 = 2 * (L.nout < 32 ? 32 : L.nout);
    int occ = 1, stages = 0;
    for (int o = 4; o >= 1; --o) {
        const size_t share = kFull / o - 1024;  // 1 KB per CTA reserved by the driver
        if (o > 1 && (o * ncols2 > 512 || w_bytes + 2 * a_bytes > share)) continue;
        VP_REQUIRE(w_bytes + a_bytes <= share, VP_ERR_UNSUPPORTED, "tc conv: %zu bytes of shared memory exceed the SM", w_bytes + a_bytes);
        occ = o;
        stages = (int)std::min<size_t>(TC_MAX_STAGES, (share - w_bytes) / a_bytes);
        break;
    }
    p.n_stages = stages;
    const size_t smem = w_bytes + (size_t)stages * a_bytes;
    int64_t ctas = (148 * occ + L.groups - 1) / L.groups;
    ctas = std::max<int64_t>(1, std::min<int64_t>(ctas, n_tiles));
    dim3 grid((unsigned)ctas, L.groups);
#define VP_TC_CASE(N, S) \
    if (L.nout == N && L.split == S) return launch_tc<N, S>(p, grid, smem, s)
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
//         2 direct conv on the x2 up-sampled (loader-side) input minus `crop` trailing samples
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
    int rc = tc_build_layer(L, mode, cin_p, COUT, K, crop, split, 1, wl, bl);
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
        rc = tc_launch(L, io, s);
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
