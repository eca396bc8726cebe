// Fused res-CNN stack on tcgen05 (sm_100a): the 14 Conv1d (64 -> 64, k = 3 / right-padded k = 2) of
// res_cnn_stack.members.0-6 (seisbench/models/eqtransformer.py ResCNNBlock) in ONE persistent
// launch instead of 14.
//
// At T = 47 a layer of a 4096-window chunk is ~10 M-tiles per SM: the layer-by-layer path spent ~50 us per launch on
// ~6 us of MMAs (weight load, TMEM allocation, pipeline fill / drain and the launch gap dominate).  Here M tiles are
// aligned with the sequences (pitch 64 rows: two 47-sample sequences per 128-row tile, the rows in between are the convs'
// zero padding), so a tile's input at layer l + 1 is exactly the tile's own output at layer l.  Each CTA keeps the same
// tiles through all layers (layer-outer, tile-inner loop): the only synchronisation between layers is CTA-local
// (__threadfence + __syncthreads), the weights of layer l + 1 stream into a second shared-memory buffer while layer l
// computes, and barriers / TMEM / pipeline state live across the whole stack.
//
// Per layer the data flow is that of tcconv_kernel (producers -> cp.async ring -> one tcgen05 issuer -> TMEM double
// buffer -> 16 epilogue warps); activations travel between layers through L2-resident global buffers:
//   conv1:  x16 = relu(bn1(x))  ->  y16 = relu(bn2(conv1(x16)))
//   conv2:  y16                 ->  x (fp32 residual stream, in place) += conv2(y16);  x16 = relu(bn1_next(x))
#include <cstdlib>
#include <cstring>

#include "fused.cuh"
#include "tc_ptx.cuh"

namespace vp {

constexpr int RS_EW = 16;                       // epilogue warps
constexpr int RS_THREADS = 32 * (RS_EW + 4);    // + issuer + 3 producer warps
constexpr int RS_PRODUCERS = 96;
constexpr int RS_ROWS = 130;                    // staged rows per tile: in-tile rows -1 .. 128
constexpr int RS_NST = 3;                       // A-tile ring depth
constexpr int RS_N = 64;                        // channels in = out = MMA N
constexpr int RS_NACC = 4;                      // TMEM accumulator buffers (the issuer runs up to 4 tiles ahead of the epilogue)

template <int SPLIT>
__device__ __forceinline__ void rs_epilogue(const ResLayerP &L, int T, int64_t y_split, uint32_t trow, int half, int seq, int srow,
                                            bool row_ok, const float4 (&rres)[4], const float *s_bias, const float *s_psc,
                                            const float *s_psh) {
    constexpr int COLS = RS_N / (RS_EW / 4);  // 16 columns per epilogue warp
    const int64_t orow = (int64_t)seq * T + srow;
    const int nb = half * COLS;
    uint32_t r[COLS];
    {
        uint32_t(&r16)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
        tmem_ld16_nowait(trow + (uint32_t)nb, r16);
    }
    tmem_ld_wait();
#pragma unroll
    for (int g8 = 0; g8 < COLS; g8 += 8) {
        const int n0 = nb + g8;
        float w8[8];
        const float4 b0 = *reinterpret_cast<const float4 *>(&s_bias[n0]), b1 = *reinterpret_cast<const float4 *>(&s_bias[n0 + 4]);
        w8[0] = __uint_as_float(r[g8 + 0]) + b0.x, w8[1] = __uint_as_float(r[g8 + 1]) + b0.y;
        w8[2] = __uint_as_float(r[g8 + 2]) + b0.z, w8[3] = __uint_as_float(r[g8 + 3]) + b0.w;
        w8[4] = __uint_as_float(r[g8 + 4]) + b1.x, w8[5] = __uint_as_float(r[g8 + 5]) + b1.y;
        w8[6] = __uint_as_float(r[g8 + 6]) + b1.z, w8[7] = __uint_as_float(r[g8 + 7]) + b1.w;
        if (!row_ok) continue;
        if (L.res != nullptr) {  // fp32 residual stream [seq][t][64], updated in place by the thread that owns the row
            float4 *rp = reinterpret_cast<float4 *>(L.res + orow * RS_N + n0);
            const float4 r0 = rres[g8 / 4], r1 = rres[g8 / 4 + 1];  // loaded before the accumulator wait
            w8[0] += r0.x, w8[1] += r0.y, w8[2] += r0.z, w8[3] += r0.w;
            w8[4] += r1.x, w8[5] += r1.y, w8[6] += r1.z, w8[7] += r1.w;
            if (L.write_res) {
                rp[0] = make_float4(w8[0], w8[1], w8[2], w8[3]);
                rp[1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
            }
        }
        if (L.affine) {  // pre-activation BatchNorm + ReLU of the next conv
            const float4 s0 = *reinterpret_cast<const float4 *>(&s_psc[n0]), s1 = *reinterpret_cast<const float4 *>(&s_psc[n0 + 4]);
            const float4 h0 = *reinterpret_cast<const float4 *>(&s_psh[n0]), h1 = *reinterpret_cast<const float4 *>(&s_psh[n0 + 4]);
            w8[0] = fmaxf(fmaf(w8[0], s0.x, h0.x), 0.f), w8[1] = fmaxf(fmaf(w8[1], s0.y, h0.y), 0.f);
            w8[2] = fmaxf(fmaf(w8[2], s0.z, h0.z), 0.f), w8[3] = fmaxf(fmaf(w8[3], s0.w, h0.w), 0.f);
            w8[4] = fmaxf(fmaf(w8[4], s1.x, h1.x), 0.f), w8[5] = fmaxf(fmaf(w8[5], s1.y, h1.y), 0.f);
            w8[6] = fmaxf(fmaf(w8[6], s1.z, h1.z), 0.f), w8[7] = fmaxf(fmaf(w8[7], s1.w, h1.w), 0.f);
        }
        uint4 hi, lo;
        pack8_split16<SPLIT>(w8, hi, lo);
        uint16_t *yb = L.y + orow * RS_N + n0;
        *reinterpret_cast<uint4 *>(yb) = hi;
        if (SPLIT == 2) *reinterpret_cast<uint4 *>(yb + y_split) = lo;
    }
}

template <int SPLIT>
__global__ void __launch_bounds__(RS_THREADS, 1) resstack_kernel(const __grid_constant__ ResStackP P) {
    extern __shared__ __align__(128) uint8_t rs_smem[];
    __shared__ __align__(8) uint64_t full_bar[RS_NST], empty_bar[RS_NST], accf_bar[RS_NACC], acce_bar[RS_NACC];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[2][RS_N], s_psc[2][RS_N], s_psh[2][RS_N];
    constexpr uint32_t W_BYTES = 12u * SPLIT * 2 * RS_N * 16;                          // weights of a k = 3 layer
    constexpr uint32_t A_BYTES = ((uint32_t)SPLIT * 8 * RS_ROWS * 16 + 127u) & ~127u;  // one staged tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sW_u = smem_u32(rs_smem);
    const uint32_t sA_u = sW_u + 2 * W_BYTES;
    const int n_tiles = (P.NS + 1) >> 1;
    const int T = P.T;

    if (tid == 0) {
        for (int i = 0; i < RS_NST; ++i) {
            mbar_init(&full_bar[i], RS_PRODUCERS);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < RS_NACC; ++i) {
            mbar_init(&accf_bar[i], 1);
            mbar_init(&acce_bar[i], 32 * RS_EW);
        }
        fence_barrier_init();
    }
    if (warp == RS_EW) tmem_alloc(&tmem_base_s, RS_NACC * RS_N);
    auto load_layer = [&](int l) {  // weights, bias and post-affine of layer l -> buffer l & 1 (asynchronous for the weights)
        const ResLayerP &L = P.l[l];
        const int pieces = L.ntaps * 4 * SPLIT * 2 * RS_N;
        const uint4 *wg = reinterpret_cast<const uint4 *>(L.w);
        const uint32_t dst = sW_u + (uint32_t)(l & 1) * W_BYTES;
        for (int idx = tid; idx < pieces; idx += RS_THREADS) cp_async16(dst + (uint32_t)idx * 16u, wg + idx, 16u);
        if (tid < RS_N) {
            s_bias[l & 1][tid] = __ldg(L.bias + tid);
            s_psc[l & 1][tid] = L.affine ? __ldg(L.psc + tid) : 1.f;
            s_psh[l & 1][tid] = L.affine ? __ldg(L.psh + tid) : 0.f;
        }
    };
    load_layer(0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // pipeline state of this thread's role, alive across the layers
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int l = 0; l < P.n_layers; ++l) {
        const ResLayerP &L = P.l[l];
        // weights of layer l have landed; the previous layer's outputs (global memory, written by this CTA) are visible
        cp_async_wait_all();
        fence_proxy_async();
        __syncthreads();
        if (l + 1 < P.n_layers) load_layer(l + 1);  // layer l - 1 (the previous user of that buffer) has retired
        const uint32_t sB_u = sW_u + (uint32_t)(l & 1) * W_BYTES;
        if (warp > RS_EW) {
            // ================= producers: rows -1 .. 128 of the tile (two sequences at pitch 64) =================
            const int ptid = tid - 32 * (RS_EW + 1);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                const uint32_t sbase = sA_u + (uint32_t)stage * A_BYTES;
                for (int r = ptid; r < RS_ROWS; r += RS_PRODUCERS) {
                    const int v = r - 1;
                    const int seq = 2 * tile + (v >> 6), u = v & 63;
                    const bool valid = v >= 0 && v < 128 && seq < P.NS && u < T;
                    const uint16_t *src = valid ? L.x + ((int64_t)seq * T + u) * RS_N : L.x;
                    const uint32_t nb = valid ? 16u : 0u;
#pragma unroll
                    for (int s = 0; s < SPLIT; ++s) {
                        const uint16_t *ss = valid ? src + (int64_t)s * P.split16 : L.x;
#pragma unroll
                        for (int c = 0; c < 8; ++c) cp_async16(sbase + (uint32_t)((s * 8 + c) * RS_ROWS + r) * 16u, ss + c * 8, nb);
                    }
                }
                cp_async_mbar_arrive_noinc(&full_bar[stage]);
                if (++stage == RS_NST) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        } else if (warp == RS_EW) {
            // ================= MMA issuer =================
            const uint32_t idesc = umma_idesc(RS_N, P.fmt16);
            const uint32_t sB16 = sB_u >> 4;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&acce_bar[acc], acc_phase ^ 1u);
                mbar_wait(&full_bar[stage], phase);
                fence_proxy_async();
                tc_fence_after();
                const uint32_t sA16 = (sA_u + (uint32_t)stage * A_BYTES) >> 4;
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * RS_N);
                if (elect_one()) {
                    // k = 3 ('same'): tap j reads staged rows j ..; k = 2 (one zero on the right): rows j + 1 ..
                    if (L.ntaps == 3) umma_conv_tile<RS_N, SPLIT, 3, 4>(d_tmem, sA16, RS_ROWS, sB16, idesc, 0u);
                    else umma_conv_tile<RS_N, SPLIT, 2, 4>(d_tmem, sA16 + 1u, RS_ROWS, sB16, idesc, 0u);
                    umma_commit(&empty_bar[stage]);
                    umma_commit(&accf_bar[acc]);
                }
                __syncwarp();
                if (++stage == RS_NST) {
                    stage = 0;
                    phase ^= 1u;
                }
                if (++acc == RS_NACC) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        } else {
            // ================= epilogue (16 warps: TMEM lane quarter = warp & 3, 16-column slice = warp >> 2) =================
            const int quarter = warp & 3, half = warp >> 2;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int v = quarter * 32 + lane;
                const int seq = 2 * tile + (v >> 6), srow = v & 63;
                const bool row_ok = seq < P.NS && srow < T;
                float4 rres[4];  // this thread's 16 residual values: in flight while the accumulator is still being computed
                if (L.res != nullptr && row_ok) {
                    const float4 *rp = reinterpret_cast<const float4 *>(L.res + ((int64_t)seq * T + srow) * RS_N + half * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) rres[i] = rp[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) rres[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(&accf_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t trow = tmem_base + (uint32_t)(acc * RS_N) + ((uint32_t)(quarter * 32) << 16);
                rs_epilogue<SPLIT>(L, T, P.split16, trow, half, seq, srow, row_ok, rres, s_bias[l & 1], s_psc[l & 1], s_psh[l & 1]);
                tc_fence_before();
                mbar_arrive(&acce_bar[acc]);
                if (++acc == RS_NACC) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
            __threadfence();  // this layer's outputs before the CTA barrier that releases the next layer's loads
        }
    }
    cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == RS_EW) tmem_dealloc(tmem_base, RS_NACC * RS_N);
}

int resstack_launch(const ResStackP &p, int split, cudaStream_t s) {
    VP_REQUIRE(p.n_layers > 0 && p.n_layers <= RS_MAX_LAYERS && p.T <= 47 + 16 && p.T > 0, VP_ERR_UNSUPPORTED,
               "res-CNN stack: %d layers, T = %d unsupported", p.n_layers, p.T);
    VP_REQUIRE(p.T + 1 <= 64, VP_ERR_UNSUPPORTED, "res-CNN stack: sequences of %d rows do not fit the 64-row pitch", p.T);
    if (p.NS == 0) return VP_OK;
    const size_t w_bytes = (size_t)12 * split * 2 * RS_N * 16;
    const size_t a_bytes = ((size_t)split * 8 * RS_ROWS * 16 + 127) & ~(size_t)127;
    const size_t smem = 2 * w_bytes + RS_NST * a_bytes;
    const int n_tiles = (p.NS + 1) / 2;
    const unsigned grid = (unsigned)std::min(148, n_tiles);
    KTimer kt(KC_TCCONV, s);
    if (split == 2) {
        static bool attr = false;
        if (!attr) {
            VP_CUDA_CHECK(cudaFuncSetAttribute(resstack_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        resstack_kernel<2><<<grid, RS_THREADS, smem, s>>>(p);
    } else {
        static bool attr = false;
        if (!attr) {
            VP_CUDA_CHECK(cudaFuncSetAttribute(resstack_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        resstack_kernel<1><<<grid, RS_THREADS, smem, s>>>(p);
    }
    VP_LAUNCH_CHECK();
    return VP_OK;
}

}  // namespace vp
