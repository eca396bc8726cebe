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
constexpr int RS_PITCH = 131;                   // plane pitch in 16-byte rows (odd: planes start in different bank groups)
constexpr int RS_NST = 3;                       // A-tile ring depth
constexpr int RS_N = 64;                        // channels in = out = MMA N
constexpr int RS_NACC = 4;                      // TMEM accumulator buffers (the issuer runs up to 4 tiles ahead of the epilogue)

// The fields of a layer the epilogue needs, copied to registers once per layer (indexing the kernel-parameter array with
// the run-time layer number inside the tile loop costs a constant load with a long scoreboard wait per use).
struct RsEpi {
    uint16_t *y;
    float *res;
    int affine, write_res;
};

// One 16-column round of a row: accumulator + bias (+ residual, stored back) -> [BN + ReLU] -> 16-bit split store.
template <int SPLIT>
__device__ __forceinline__ void rs_epilogue(const RsEpi &L, int T, int64_t y_split, uint32_t trow, int half, int seq, int srow,
                                            bool row_ok, const float4 (&rres)[4], const float *s_bias, const float *s_psc,
                                            const float *s_psh) {
    const int64_t orow = (int64_t)seq * T + srow;
    const int nb = half * 16;
    uint32_t r[16];
    tmem_ld16_nowait(trow + (uint32_t)nb, r);
    tmem_ld_wait();
    if (!row_ok) return;
    float w[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4 *>(&s_bias[nb + 4 * q]);
        w[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b4.x, w[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
        w[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z, w[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
    }
    if (L.res != nullptr) {  // fp32 residual stream [seq][t][64], updated in place by the thread that owns the row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            w[4 * q + 0] += rres[q].x, w[4 * q + 1] += rres[q].y, w[4 * q + 2] += rres[q].z, w[4 * q + 3] += rres[q].w;
        }
        if (L.write_res) {
            float *rp = L.res + orow * RS_N + nb;
            st_global_256(rp, &w[0]);
            st_global_256(rp + 8, &w[8]);
        }
    }
    if (L.affine) {  // pre-activation BatchNorm + ReLU of the next conv
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 s4 = *reinterpret_cast<const float4 *>(&s_psc[nb + 4 * q]), h4 = *reinterpret_cast<const float4 *>(&s_psh[nb + 4 * q]);
            w[4 * q + 0] = fmaxf(fmaf(w[4 * q + 0], s4.x, h4.x), 0.f), w[4 * q + 1] = fmaxf(fmaf(w[4 * q + 1], s4.y, h4.y), 0.f);
            w[4 * q + 2] = fmaxf(fmaf(w[4 * q + 2], s4.z, h4.z), 0.f), w[4 * q + 3] = fmaxf(fmaf(w[4 * q + 3], s4.w, h4.w), 0.f);
        }
    }
    uint4 hi0, lo0, hi1, lo1;
    pack8_split16<SPLIT>(&w[0], hi0, lo0);
    pack8_split16<SPLIT>(&w[8], hi1, lo1);
    uint16_t *yb = L.y + orow * RS_N + nb;
    st_global_256(yb, hi0, hi1);
    if (SPLIT == 2) st_global_256(yb + y_split, lo0, lo1);
}

template <int SPLIT>
__global__ void __launch_bounds__(RS_THREADS, 1) resstack_kernel(const __grid_constant__ ResStackP P) {
    extern __shared__ __align__(128) uint8_t rs_smem[];
    __shared__ __align__(8) uint64_t full_bar[RS_NST], empty_bar[RS_NST], accf_bar[RS_NACC], acce_bar[RS_NACC];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[2][RS_N], s_psc[2][RS_N], s_psh[2][RS_N];
    constexpr uint32_t W_BYTES = 12u * SPLIT * 2 * RS_N * 16;                          // weights of a k = 3 layer
    constexpr uint32_t A_BYTES = ((uint32_t)SPLIT * 8 * RS_PITCH * 16 + 127u) & ~127u;  // one staged tile

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
            mbar_init(&acce_bar[i], 32 * 4);  // the four warps of the group that owns the accumulator
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
    int stage = 0, acc = 0, nseq = 0;
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
                // one staged row per thread, plane after plane.  (8 lanes per row = one coalesced 128-byte global row, with the
                // odd plane pitch conflict-free in shared memory, cut the LDGSTS wavefronts 6x but measured 0.6 % slower:
                // the per-piece index arithmetic costs more than the wavefronts.)
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
                        for (int c = 0; c < 8; ++c) cp_async16(sbase + (uint32_t)((s * 8 + c) * RS_PITCH + r) * 16u, ss + c * 8, nb);
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
                    if (L.ntaps == 3) umma_conv_tile<RS_N, SPLIT, 3, 4>(d_tmem, sA16, RS_PITCH, sB16, idesc, 0u);
                    else umma_conv_tile<RS_N, SPLIT, 2, 4>(d_tmem, sA16 + 1u, RS_PITCH, sB16, idesc, 0u);
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
            // ================= epilogue: 4 groups of 4 warps (TMEM lane quarter = warp & 3) =================
            // Group g owns accumulator buffer g, i.e. every 4th tile of this CTA, and converts all 64 columns of it in four
            // 16-column rounds.  (With the 16 warps on ONE tile, 16 columns each, the CTA converted a single tile at a
            // time and the chain accumulator wait -> TMEM load -> residual (global) -> stores ran at ~7400 cycles per
            // tile against ~1400 cycles of MMAs; four tiles in flight overlap those latencies.)
            const int quarter = warp & 3, grp = warp >> 2;
            RsEpi E;
            E.y = L.y;
            E.res = L.res;
            E.affine = L.affine;
            E.write_res = L.write_res;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++nseq) {
                if ((nseq & 3) != grp) continue;
                const int v = quarter * 32 + lane;
                const int seq = 2 * tile + (v >> 6), srow = v & 63;
                const bool row_ok = seq < P.NS && srow < T;
                const bool has_res = E.res != nullptr && row_ok;
                const float *rp = E.res + ((int64_t)seq * T + srow) * RS_N;
                float4 cur[4], nxt[4];  // residual values of the current / next 16-column round
#pragma unroll
                for (int i = 0; i < 4; ++i) cur[i] = make_float4(0.f, 0.f, 0.f, 0.f), nxt[i] = cur[i];
                if (has_res) {
                    ld_global_256(rp, cur[0], cur[1]);
                    ld_global_256(rp + 8, cur[2], cur[3]);
                }
                mbar_wait(&accf_bar[grp], acc_phase);
                tc_fence_after();
                const uint32_t trow = tmem_base + (uint32_t)(grp * RS_N) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3 && has_res) {
                        ld_global_256(rp + 16 * (c + 1), nxt[0], nxt[1]);
                        ld_global_256(rp + 16 * (c + 1) + 8, nxt[2], nxt[3]);
                    }
                    rs_epilogue<SPLIT>(E, T, P.split16, trow, c, seq, srow, row_ok, cur, s_bias[l & 1], s_psc[l & 1], s_psh[l & 1]);
                    if (c < 3) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
                    }
                }
                tc_fence_before();
                mbar_arrive(&acce_bar[grp]);
                acc_phase ^= 1u;
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
    for (int l = 0; l < p.n_layers; ++l)  // the epilogue moves 32-byte pieces (256-bit loads / stores)
        VP_REQUIRE(reinterpret_cast<uintptr_t>(p.l[l].y) % 32 == 0 && reinterpret_cast<uintptr_t>(p.l[l].res) % 32 == 0 &&
                       (p.split16 * 2) % 32 == 0,
                   VP_ERR_ARG, "res-CNN stack: layer %d buffers are not 32-byte aligned", l);
    const size_t w_bytes = (size_t)12 * split * 2 * RS_N * 16;
    const size_t a_bytes = ((size_t)split * 8 * RS_PITCH * 16 + 127) & ~(size_t)127;
    const size_t smem = 2 * w_bytes + RS_NST * a_bytes;
    const int n_tiles = (p.NS + 1) / 2;
    const unsigned grid = (unsigned)std::min(device_sm_count(), n_tiles);
    KTimer kt(KC_RESSTACK, s);
    if (split == 2) {
        if (int rc = ensure_dyn_smem((const void *)resstack_kernel<2>, smem)) return rc;
        resstack_kernel<2><<<grid, RS_THREADS, smem, s>>>(p);
    } else {
        if (int rc = ensure_dyn_smem((const void *)resstack_kernel<1>, smem)) return rc;
        resstack_kernel<1><<<grid, RS_THREADS, smem, s>>>(p);
    }
    VP_LAUNCH_CHECK();
    return VP_OK;
}

}  // namespace vp
