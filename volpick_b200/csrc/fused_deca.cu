// Fused decoder middle on tcgen05 (sm_100a): decoder.convs.1 -> decoder.convs.2 in ONE kernel.
//
// Replaces, for the three EQTransformer decoders (decoder_d, pick_decoders.0/1; SURVEY.md Appendix A,
// seisbench/models/eqtransformer.py Decoder), two nn.Upsample(x2) + F.conv1d + ReLU stages
//     (B, 64, 94) --up x2, conv k5--> (B, 64, 188) --up x2, drop last, conv k5--> (B, 32, 375)
// that the layer-by-layer path ran as two kernels with 48 KB per (window, decoder) of 16-bit activations written
// to and re-read from HBM in 16-byte pieces (the measured bottleneck of both: their epilogues / loaders, not
// their MMAs).  One work item = one (window, decoder) sequence; the 188-row intermediate lives in shared memory.
//
// Both layers run in the polyphase form of tcconv.cu (x2 up-sampling folded into the weights: row s yields
// outputs 2s, 2s + 1 from taps s - 1, s, s + 1).  decoder.convs.2 crops the up-sampled signal to 375 samples
// BEFORE the conv, i.e. u[375] is zero padding instead of x[187]; polyphase assumes u[375] = x[187], which only
// touches outputs 373 (through w[:, :, 4]) and 374 (through w[:, :, 3]).  The two 32 x 64 matrix-vector products
// are subtracted in fp32 on the CUDA cores (512 threads x 8 FMAs per work item).
//
// CTA = 18 warps: loader, tcgen05 issuer, 16 epilogue warps (TMEM lane quarter = warp % 4, four column slices).
// Weights of both layers (147 KB in f16x3) stay resident in shared memory; one input slot and the intermediate
// fill the rest, so a CTA runs ONE work item at a time and overlaps the next item's first layer with the
// epilogues of the current one.
#include <algorithm>
#include <cstring>

#include "fused.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

constexpr int DA_T0 = 94, DA_T1 = 188, DA_T2 = 375;
constexpr int DA_IN_ROWS = DA_T0 + 2, DA_MID_ROWS = DA_T1 + 2;  // one zero row before and after the sequence
// Plane pitch of the input slot in 16-byte rows.  The slot is filled by ONE TMA box per item (round 2): the 5-D tensor map of
// the channel-last input, box [8 channels][96 rows from t = -1][1 sequence][8 planes][split], lands as [split][plane][96 rows][16 B]
// -- the UMMA layout -- and the two rows outside the sequence (t = -1, 94) arrive as zeros: they are the conv's padding.  (Round 1
// copied 1,504 16-byte cp.async pieces per item, ~1,500 shared-memory wavefronts against 192 for the box; the odd pitch 97 that
// kept those stores conflict-free is no longer needed.)
constexpr int DA_IN_PITCH = DA_IN_ROWS;

struct FzDecAK {
    alignas(64) CUtensorMap x_map;  // (8 channels, t, group * B + window, 8-channel plane, split) over p.x
    FzDecA p;
};
constexpr int DA_EW = 16;
constexpr int DA_THREADS = 32 * (2 + DA_EW);

template <int SPLIT>
__device__ __forceinline__ void da_pack8(const float *v, uint4 &hi, uint4 &lo) {
    pack8_split16<SPLIT>(v, hi, lo);
}

template <int SPLIT>
__global__ void __launch_bounds__(DA_THREADS, 1) deca_kernel(const __grid_constant__ FzDecAK K) {
    const FzDecA &p = K.p;
    extern __shared__ __align__(128) uint8_t da_smem[];
    __shared__ __align__(8) uint64_t in_full, in_free, acc_full[3], done_bar[3];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_fix[2][32];  // crop correction of outputs 373 / 374

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const uint32_t sbase = smem_u32(da_smem);
    constexpr uint32_t IN_BYTES = SPLIT * 8 * DA_IN_PITCH * 16, MID_BYTES = SPLIT * 8 * DA_MID_ROWS * 16;
    uint8_t *s_mid = da_smem + IN_BYTES;  // the input slot [0, IN_BYTES) is written by TMA only
    const float *s_bias = p.bias_c[g];  // [128] dec1, [64] dec2: constant bank (kernel parameter)

    if (tid == 0) {
        mbar_init(&in_full, 1);
        tma_prefetch_desc(&K.x_map);
        mbar_init(&in_free, 1);
        for (int i = 0; i < 3; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&done_bar[i], DA_EW);
        }
        fence_barrier_init();
    }
    // accumulators: convs.1 at column 0 (128 columns, + 128 for the stacked W_lo half in f16x3), convs.2 tile t at 256 + 128 t (64 + 64)
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    {   // resident weights + biases; zero rows around the sequences (never written again)
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.blob + (long long)g * (p.blob_bytes / 2));
        for (int idx = tid; idx < p.blob_bytes / 16; idx += DA_THREADS) cp_async16(sbase + p.blob_off + idx * 16, wg + idx, 16u);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int idx = tid; idx < SPLIT * 8 * 2; idx += DA_THREADS) {
            const int pl = idx >> 1, e = idx & 1;
            *reinterpret_cast<uint4 *>(s_mid + ((size_t)pl * DA_MID_ROWS + (e ? DA_MID_ROWS - 1 : 0)) * 16) = z;
        }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int n_items = p.B;

    if (warp == 0) {
        // ================= loader: one TMA box per item (rows -1 .. 94 of all planes and splits) =================
        if (lane == 0) {
            int n = 0;
            for (int b = blockIdx.x; b < n_items; b += gridDim.x, ++n) {
                mbar_wait(&in_free, (n & 1) ^ 1);
                mbar_arrive_expect_tx(&in_full, IN_BYTES);
                tma_load_5d(sbase, &K.x_map, &in_full, 0, -1, g * n_items + b, 0, 0);
            }
        }
    } else if (warp == 1) {
        // ================= tcgen05 issuer =================
        const uint32_t fmt = SPLIT == 2 ? 0 : 1;
        const uint32_t in16 = sbase >> 4, mid16 = (sbase + IN_BYTES) >> 4;
        const uint32_t w1 = (sbase + p.w1_off) >> 4, w2 = (sbase + p.w2_off) >> 4;
        int n = 0;
        for (int b = blockIdx.x; b < n_items; b += gridDim.x, ++n) {
            const uint32_t par = n & 1;
            // ---- decoder.convs.1: rows s = 0 .. 127 (94 valid), N = 2 x 64
            mbar_wait(&done_bar[0], par ^ 1);  // accumulator 0 drained (previous item)
            mbar_wait(&in_full, par);
            fence_proxy_async();
            tc_fence_after();
            if (elect_one()) {
                // f16x3: W_hi and W_lo stacked along N (A_hi is read once for A_hi W_hi and A_hi W_lo: two instead of three 4 KB
                // A-tile reads per K step; the shared-memory operand reads bound this kernel), the epilogue adds the two halves
                if constexpr (SPLIT == 2) umma_conv_tile_stacked<128, 3, 4>(tmem_base, in16, DA_IN_PITCH, w1, umma_idesc(256, 0), umma_idesc(128, 0), 0u);
                else umma_conv_tile<128, SPLIT, 3, 4>(tmem_base, in16, DA_IN_PITCH, w1, umma_idesc(128, fmt), 0u);
                umma_commit(&acc_full[0]);
                umma_commit(&in_free);
            }
            __syncwarp();
            // ---- decoder.convs.2: two tiles of 128 rows (188 valid), N = 2 x 32; needs the whole intermediate
            mbar_wait(&done_bar[0], par);
            fence_proxy_async();
            tc_fence_after();
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&done_bar[1 + t], par ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    if constexpr (SPLIT == 2)
                        umma_conv_tile_stacked<64, 3, 4>(tmem_base + 256u + 128u * t, mid16 + 128u * t, DA_MID_ROWS, w2, umma_idesc(128, 0), umma_idesc(64, 0), 0u);
                    else umma_conv_tile<64, SPLIT, 3, 4>(tmem_base + 256u + 128u * t, mid16 + 128u * t, DA_MID_ROWS, w2, umma_idesc(64, fmt), 0u);
                    umma_commit(&acc_full[1 + t]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue warps =================
        const int e = warp - 2;
        const int lq = warp & 3;   // TMEM lane quarter this warp may read
        const int cs = e >> 2;     // column slice 0 .. 3
        const int r = lq * 32 + lane;
        const int etid = e * 32 + lane;  // 0 .. 511
        const float *fixw = p.fixw + (long long)g * 2 * 32 * 64;
        uint16_t *yg = p.y + (long long)g * p.y_gs;
        int n = 0;
        for (int b = blockIdx.x; b < n_items; b += gridDim.x, ++n) {
            const uint32_t par = n & 1;
            // ---- decoder.convs.1 -> intermediate rows 1 + 2s, 2 + 2s; this warp: channels [16 cs, 16 cs + 16) of both phases
            mbar_wait(&acc_full[0], par);
            tc_fence_after();
            if (lq < 3) {  // rows 96 .. 127 do not exist
                const uint32_t tacc = tmem_base + ((uint32_t)(lq * 32) << 16);
                uint32_t r0[16], r1[16];
                tmem_ld16_nowait(tacc + (uint32_t)(16 * cs), r0);
                tmem_ld16_nowait(tacc + (uint32_t)(64 + 16 * cs), r1);
                tmem_ld_wait();
                if constexpr (SPLIT == 2) {  // + the stacked half A_hi W_lo
                    uint32_t q0[16], q1[16];
                    tmem_ld16_nowait(tacc + (uint32_t)(128 + 16 * cs), q0);
                    tmem_ld16_nowait(tacc + (uint32_t)(192 + 16 * cs), q1);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        r0[i] = __float_as_uint(__uint_as_float(r0[i]) + __uint_as_float(q0[i]));
                        r1[i] = __float_as_uint(__uint_as_float(r1[i]) + __uint_as_float(q1[i]));
                    }
                }
                if (r < DA_T0) {
                    float v0[16], v1[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v0[i] = fmaxf(__uint_as_float(r0[i]) + s_bias[16 * cs + i], 0.f);
                        v1[i] = fmaxf(__uint_as_float(r1[i]) + s_bias[64 + 16 * cs + i], 0.f);
                    }
                    // lanes with bit 2 set store phase 1 first: a quarter warp covers 8 distinct 16-byte bank groups
                    const int sel = (lane >> 2) & 1;
                    uint8_t *dA = s_mid + (size_t)(1 + 2 * r + sel) * 16, *dB = s_mid + (size_t)(2 + 2 * r - sel) * 16;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int pl = 2 * cs + h;
                        uint4 h0, l0, h1, l1;
                        da_pack8<SPLIT>(&v0[8 * h], h0, l0);
                        da_pack8<SPLIT>(&v1[8 * h], h1, l1);
                        *reinterpret_cast<uint4 *>(dA + (size_t)pl * DA_MID_ROWS * 16) = sel ? h1 : h0;
                        *reinterpret_cast<uint4 *>(dB + (size_t)pl * DA_MID_ROWS * 16) = sel ? h0 : h1;
                        if (SPLIT == 2) {
                            *reinterpret_cast<uint4 *>(dA + (size_t)(8 + pl) * DA_MID_ROWS * 16) = sel ? l1 : l0;
                            *reinterpret_cast<uint4 *>(dB + (size_t)(8 + pl) * DA_MID_ROWS * 16) = sel ? l0 : l1;
                        }
                    }
                }
            }
            // ---- crop correction: fix[0][co] = sum_ci w[co][ci][4] x187[ci], fix[1][co] = sum_ci w[co][ci][3] x187[ci]
            named_bar_sync(1, 32 * DA_EW);  // intermediate row 187 written; previous item's readers of s_fix are done
            {
                const int which = etid >> 8, co = (etid >> 3) & 31, sl = etid & 7;
                const uint4 xh = *reinterpret_cast<const uint4 *>(s_mid + ((size_t)sl * DA_MID_ROWS + DA_T1) * 16);
                const __half2 *hh = reinterpret_cast<const __half2 *>(&xh);
                const __nv_bfloat162 *bb = reinterpret_cast<const __nv_bfloat162 *>(&xh);
                float x8[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = SPLIT == 2 ? __half22float2(hh[i]) : __bfloat1622float2(bb[i]);
                    x8[2 * i] = f.x;
                    x8[2 * i + 1] = f.y;
                }
                if (SPLIT == 2) {
                    const uint4 xl = *reinterpret_cast<const uint4 *>(s_mid + ((size_t)(8 + sl) * DA_MID_ROWS + DA_T1) * 16);
                    const __half2 *ll = reinterpret_cast<const __half2 *>(&xl);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(ll[i]);
                        x8[2 * i] += f.x;
                        x8[2 * i + 1] += f.y;
                    }
                }
                const float4 *wp = reinterpret_cast<const float4 *>(fixw + ((size_t)which * 32 + co) * 64 + sl * 8);
                const float4 wa = __ldg(wp), wb = __ldg(wp + 1);
                float acc = wa.x * x8[0];
                acc = fmaf(wa.y, x8[1], acc), acc = fmaf(wa.z, x8[2], acc), acc = fmaf(wa.w, x8[3], acc);
                acc = fmaf(wb.x, x8[4], acc), acc = fmaf(wb.y, x8[5], acc), acc = fmaf(wb.z, x8[6], acc), acc = fmaf(wb.w, x8[7], acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (sl == 0) s_fix[which][co] = acc;
            }
            fence_proxy_async();  // generic-proxy writes of the intermediate -> visible to the tensor-core proxy
            tc_fence_before();
            named_bar_sync(1, 32 * DA_EW);  // s_fix complete
            if (lane == 0) mbar_arrive(&done_bar[0]);
            // ---- decoder.convs.2 -> global (B, 375, 32) 16-bit; this warp: channels [16 hc, 16 hc + 16) of phase ph, i.e. 32
            // contiguous bytes of one output row per thread and split (one 256-bit store instead of two 16-byte pieces)
            const int ph = cs >> 1, hc = cs & 1;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&acc_full[1 + t], par);
                tc_fence_after();
                const int s = 128 * t + r;
                if (t == 0 || lq < 2) {  // rows 192 .. 255 do not exist
                    const uint32_t tacc = tmem_base + 256u + 128u * t + ((uint32_t)(lq * 32) << 16);
                    uint32_t r0[16];
                    tmem_ld16_nowait(tacc + (uint32_t)(32 * ph + 16 * hc), r0);
                    if constexpr (SPLIT == 2) {
                        uint32_t q0[16];
                        tmem_ld16_nowait(tacc + (uint32_t)(64 + 32 * ph + 16 * hc), q0);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r0[i] = __float_as_uint(__uint_as_float(r0[i]) + __uint_as_float(q0[i]));
                    } else {
                        tmem_ld_wait();
                    }
                    if (s < DA_T1 && 2 * s + ph < DA_T2) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float a = __uint_as_float(r0[i]);
                            if (ph == 1 && s == DA_T1 - 2) a -= s_fix[0][16 * hc + i];  // output 373 = row 186, phase 1
                            if (ph == 0 && s == DA_T1 - 1) a -= s_fix[1][16 * hc + i];  // output 374 = row 187, phase 0
                            v[i] = fmaxf(a + s_bias[128 + 32 * ph + 16 * hc + i], 0.f);
                        }
                        uint4 h0, l0, h1, l1;
                        da_pack8<SPLIT>(&v[0], h0, l0);
                        da_pack8<SPLIT>(&v[8], h1, l1);
                        uint16_t *yb = yg + ((long long)b * DA_T2 + 2 * s + ph) * 32 + 16 * hc;
                        st_pair16(yb, h0, yb + 8, h1);
                        if (SPLIT == 2) st_pair16(yb + p.y_split, l0, yb + 8 + p.y_split, l1);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&done_bar[1 + t]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ host
int deca_build(DecAPlan &plan, const TcLayer &dec1, const TcLayer &dec2p, int split, const float *const *w2 /*3 x (32, 64, 5)*/) {
    FzDecA &p = plan.p;
    std::memset(&p, 0, sizeof(p));
    plan.split = split;
    plan.ready = false;
    const int G = 3;
    VP_REQUIRE(dec1.ph == 2 && dec1.nout == 128 && dec1.sched_taps == 3 && dec1.sched_nq == 4 && dec1.row0 == -1 && dec1.groups == G,
               VP_ERR_UNSUPPORTED, "deca: decoder.convs.1 is not the compiled (N 128, 3 taps, 4 pairs) polyphase layer");
    VP_REQUIRE(dec2p.ph == 2 && dec2p.nout == 64 && dec2p.sched_taps == 3 && dec2p.sched_nq == 4 && dec2p.row0 == -1 && dec2p.groups == G,
               VP_ERR_UNSUPPORTED, "deca: decoder.convs.2 is not the compiled (N 64, 3 taps, 4 pairs) polyphase layer");
    VP_REQUIRE(!dec1.blocks.empty() && !dec2p.blocks.empty(), VP_ERR_ARG, "deca: host weight blocks are gone");
    auto up128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    const size_t in_bytes = (size_t)split * 8 * DA_IN_PITCH * 16, mid_bytes = (size_t)split * 8 * DA_MID_ROWS * 16;
    const size_t w1_bytes = up128((size_t)dec1.n_blocks * split * 2 * 128 * 16), w2_bytes = up128((size_t)dec2p.n_blocks * split * 2 * 64 * 16);
    const size_t bias_bytes = up128((128 + 64) * sizeof(float));
    p.blob_off = (int)up128(in_bytes + mid_bytes);
    p.w1_off = p.blob_off;
    p.w2_off = p.blob_off + (int)w1_bytes;
    p.bias_off = p.w2_off + (int)w2_bytes;
    p.blob_bytes = (int)(w1_bytes + w2_bytes + bias_bytes);
    p.smem_bytes = p.blob_off + p.blob_bytes;
    VP_REQUIRE(p.smem_bytes <= 226 * 1024, VP_ERR_UNSUPPORTED, "deca: %d bytes of shared memory", p.smem_bytes);
    plan.blob.assign((size_t)G * p.blob_bytes / 2, 0);
    plan.fixw.assign((size_t)G * 2 * 32 * 64, 0.f);
    const size_t e1 = (size_t)dec1.n_blocks * split * 2 * 128 * 8, e2 = (size_t)dec2p.n_blocks * split * 2 * 64 * 8;
    for (int g = 0; g < G; ++g) {
        uint16_t *dst = plan.blob.data() + (size_t)g * p.blob_bytes / 2;
        // TcLayer blocks: [block][split][k-half][nout][8]; the stacked f16x3 schedule wants [block][k-half][split * nout + n][8]
        auto put = [&](uint16_t *d, const TcLayer &TL, size_t elems) {
            const uint16_t *src = TL.blocks.data() + (size_t)g * elems;
            if (split != 2) {
                std::memcpy(d, src, elems * sizeof(uint16_t));
                return;
            }
            const size_t blk = (size_t)2 * 2 * TL.nout * 8;
            for (int b = 0; b < TL.n_blocks; ++b)
                for (int sp = 0; sp < 2; ++sp)
                    for (int kh = 0; kh < 2; ++kh)
                        std::memcpy(d + (size_t)b * blk + ((size_t)kh * 2 * TL.nout + (size_t)sp * TL.nout) * 8,
                                    src + (size_t)b * blk + ((size_t)sp * 2 + kh) * TL.nout * 8, (size_t)TL.nout * 8 * sizeof(uint16_t));
        };
        put(dst, dec1, e1);
        put(dst + w1_bytes / 2, dec2p, e2);
        float *bd = reinterpret_cast<float *>(dst + (w1_bytes + w2_bytes) / 2);
        for (int n = 0; n < 128; ++n) bd[n] = p.bias_c[g][n] = dec1.bias[(size_t)g * 128 + n];
        for (int n = 0; n < 64; ++n) bd[128 + n] = p.bias_c[g][128 + n] = dec2p.bias[(size_t)g * 64 + n];
        for (int co = 0; co < 32; ++co)
            for (int ci = 0; ci < 64; ++ci) {
                plan.fixw[(((size_t)g * 2 + 0) * 32 + co) * 64 + ci] = w2[g][((size_t)co * 64 + ci) * 5 + 4];
                plan.fixw[(((size_t)g * 2 + 1) * 32 + co) * 64 + ci] = w2[g][((size_t)co * 64 + ci) * 5 + 3];
            }
    }
    return VP_OK;
}

int deca_upload(DecAPlan &plan) {
    VP_CUDA_CHECK(cudaMalloc(&plan.d_blob, plan.blob.size() * sizeof(uint16_t)));
    VP_CUDA_CHECK(cudaMemcpy(plan.d_blob, plan.blob.data(), plan.blob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    VP_CUDA_CHECK(cudaMalloc(&plan.d_fixw, plan.fixw.size() * sizeof(float)));
    VP_CUDA_CHECK(cudaMemcpy(plan.d_fixw, plan.fixw.data(), plan.fixw.size() * sizeof(float), cudaMemcpyHostToDevice));
    plan.blob.clear();
    plan.blob.shrink_to_fit();
    plan.ready = true;
    return VP_OK;
}

void deca_free(DecAPlan &plan) {
    if (plan.d_blob) cudaFree(plan.d_blob);
    if (plan.d_fixw) cudaFree(plan.d_fixw);
    plan.d_blob = nullptr;
    plan.d_fixw = nullptr;
    plan.ready = false;
}

template <int SPLIT>
static int deca_launch_t(const FzDecA &p, dim3 grid, cudaStream_t s) {
    auto kern = deca_kernel<SPLIT>;
    FzDecAK K;
    K.p = p;
    {
        VP_REQUIRE(p.x_gs == (long long)p.B * DA_T0 * 64 && (p.x_split * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(p.x) % 16 == 0,
                   VP_ERR_ARG, "deca: the input groups must be contiguous ([split][group][B][94][64]) and 16-byte aligned");
        const uint64_t dims[5] = {8, (uint64_t)DA_T0, (uint64_t)3 * p.B, 8, (uint64_t)SPLIT};
        const uint64_t strides[4] = {128, (uint64_t)DA_T0 * 128, 16, (uint64_t)p.x_split * 2};
        const uint32_t box[5] = {8, (uint32_t)DA_IN_ROWS, 1, 8, (uint32_t)SPLIT};
        if (int rc = tma_encode_u16(&K.x_map, p.x, 5, dims, strides, box)) return rc;
    }
    if (int rc = ensure_dyn_smem((const void *)kern, (size_t)p.smem_bytes)) return rc;
    KTimer kt(KC_DECA, s);
    kern<<<grid, DA_THREADS, p.smem_bytes, s>>>(K);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

int deca_launch(const DecAPlan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, uint16_t *y, long long y_split,
                long long y_gs, cudaStream_t s) {
    VP_REQUIRE(plan.ready, VP_ERR_UNSUPPORTED, "deca: plan not uploaded");
    if (B == 0) return VP_OK;
    FzDecA p = plan.p;
    p.x = x;
    p.x_split = x_split;
    p.x_gs = x_gs;
    p.y = y;
    p.y_split = y_split;
    p.y_gs = y_gs;
    p.B = B;
    p.blob = plan.d_blob;
    p.fixw = plan.d_fixw;
    dim3 grid((unsigned)std::min(std::max(device_sm_count() / 3, 1), B), 3);  // one CTA per SM over the 3 decoders
    return plan.split == 2 ? deca_launch_t<2>(p, grid, s) : deca_launch_t<1>(p, grid, s);
}

}  // namespace vp
