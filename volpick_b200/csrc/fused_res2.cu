// Res-CNN stack with every activation on chip (sm_100a): the 14 Conv1d (64 -> 64, k = 3 / right-padded k = 2) of
// res_cnn_stack.members.0-6 (seisbench/models/eqtransformer.py ResCNNBlock).
//
// fused_res.cu runs the 14 layers in one launch but every layer round-trips its activations through global memory
// (ncu: 1.6 GB of DRAM traffic per 4096-window launch against 150 MB of input + output).  Here a tile (two 47-sample
// sequences at a 64-row pitch = 128 MMA rows) stays on the SM through all 14 layers:
//   * the 16-bit operand (fp16 hi / lo planes, or bf16) of the tile is ONE TMA box: a 5-D tensor map over the
//     channel-last buffer [split][seq][t][64] delivers [split][8-channel plane][2 seq x 64 rows][8] -- exactly the
//     K-major no-swizzle UMMA layout, plane pitch 128 rows -- and fills rows 47 .. 63 of each sequence with zeros
//     (out-of-range rows of the map), which are the convs' zero padding: tap -1 of a sequence's first row reads the
//     previous plane's last row, tap +1 of its last row reads row 47;
//   * every epilogue overwrites the tile's operand IN PLACE (the layer's MMAs have retired when its accumulator is
//     full), so the next layer's A operand is already where its descriptors point;
//   * the fp32 residual stream lives in TMEM: it is loaded once (tcgen05.st), every conv2 ACCUMULATES onto it (the first
//     MMA of the layer does not overwrite), and the conv2 biases are added when it is read (cumulative sums, host side).
// Weights (48 KB per k = 3 layer) stream through a double buffer with one cp.async.bulk per layer; the layer loop is
// outermost over a group of R2_SLOTS tiles so that a layer's weights are fetched once per group.
// Warps: 0 producer (TMA), 1 tcgen05 issuer, 2 .. 13 epilogue (one group of four per tile slot; TMEM lane quarter = warp % 4).
#include <cstdlib>
#include <cstring>

#include "fused.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

constexpr int R2_SLOTS = 3;
constexpr int R2_THREADS = 32 * (2 + 4 * R2_SLOTS);
constexpr int R2_L = 14;
constexpr int R2_N = 64;
constexpr int R2_PAD = 128;  // zero bytes in front of slot 0 (row -1 of its first plane)

struct ResStack2K {
    alignas(64) CUtensorMap x_map;
    ResStack2P p;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int SPLIT>
__global__ void __launch_bounds__(R2_THREADS, 1) resstack2_kernel(const __grid_constant__ ResStack2K K) {
    extern __shared__ __align__(128) uint8_t r2_smem[];
    __shared__ __align__(8) uint64_t w_full[2], w_free[2], a_tma[R2_SLOTS], a_epi[R2_SLOTS], a_free[R2_SLOTS], acc_full[R2_SLOTS];
    __shared__ uint32_t tmem_base_s;
    constexpr uint32_t A_BYTES = 16384u * SPLIT;                  // one tile: [split][8 planes][128 rows][16 B]
    constexpr uint32_t W_BYTES = 12u * SPLIT * 2 * R2_N * 16;     // weights of a k = 3 layer
    const ResStack2P &P = K.p;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t s0 = smem_u32(r2_smem);
    const uint32_t sA = s0 + R2_PAD, sW = sA + R2_SLOTS * A_BYTES, sPar = sW + 2 * W_BYTES;
    float *par = reinterpret_cast<float *>(r2_smem + R2_PAD + R2_SLOTS * A_BYTES + 2 * W_BYTES);
    const int n_tiles = (P.NS + 1) >> 1;
    // this CTA's contiguous range of tiles, walked in groups of R2_SLOTS (the last group may be partial)
    const int tile_lo = (int)((long long)blockIdx.x * n_tiles / gridDim.x), tile_hi = (int)((long long)(blockIdx.x + 1) * n_tiles / gridDim.x);
    const int n_groups = (tile_hi - tile_lo + R2_SLOTS - 1) / R2_SLOTS;
    const int T = P.T;
    (void)sPar;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_free[i], 1);
        }
        for (int i = 0; i < R2_SLOTS; ++i) {
            mbar_init(&a_tma[i], 1);
            mbar_init(&a_epi[i], 4);
            mbar_init(&a_free[i], 1);
            mbar_init(&acc_full[i], 1);
        }
        fence_barrier_init();
        tma_prefetch_desc(&K.x_map);
    }
    if (tid < R2_PAD / 4) reinterpret_cast<uint32_t *>(r2_smem)[tid] = 0u;
    for (int i = tid; i < R2_L * 3 * R2_N; i += R2_THREADS) par[i] = __ldg(P.par + i);
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ================= producer: operand tiles (one TMA box each) and the layers' weights (one bulk copy each)
        if (lane == 0) {
            for (int n = 0; n < n_groups; ++n) {
                for (int slot = 0; slot < R2_SLOTS; ++slot) {
                    const int tile = tile_lo + n * R2_SLOTS + slot;
                    if (tile >= tile_hi) break;
                    if (n > 0) mbar_wait(&a_free[slot], (uint32_t)(n - 1) & 1u);  // layer 13 of the previous tile in this slot has retired
                    mbar_arrive_expect_tx(&a_tma[slot], A_BYTES);
                    tma_load_5d(sA + (uint32_t)slot * A_BYTES, &K.x_map, &a_tma[slot], 0, 0, 2 * tile, 0, 0);
                }
                for (int l = 0; l < R2_L; ++l) {
                    const int b = l & 1, u = n * (R2_L / 2) + (l >> 1);
                    if (u > 0) mbar_wait(&w_free[b], (uint32_t)(u - 1) & 1u);
                    const uint32_t bytes = (uint32_t)P.ntaps[l] * 4u * SPLIT * 2 * R2_N * 16;
                    mbar_arrive_expect_tx(&w_full[b], bytes);
                    bulk_load(sW + (uint32_t)b * W_BYTES, P.w[l], bytes, &w_full[b]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= tcgen05 issuer: layer-outer, slot-inner
        const uint32_t idesc = umma_idesc(R2_N, P.fmt16);
        for (int n = 0; n < n_groups; ++n) {
            for (int l = 0; l < R2_L; ++l) {
                const int b = l & 1, u = n * (R2_L / 2) + (l >> 1);
                mbar_wait(&w_full[b], (uint32_t)u & 1u);
                const uint32_t w16 = (sW + (uint32_t)b * W_BYTES) >> 4;
                const int ntaps = P.ntaps[l];
                for (int slot = 0; slot < R2_SLOTS; ++slot) {
                    if (tile_lo + n * R2_SLOTS + slot >= tile_hi) break;
                    if (l == 0) mbar_wait(&a_tma[slot], (uint32_t)n & 1u);
                    mbar_wait(&a_epi[slot], (uint32_t)l & 1u);  // phase n * 14 + l: residual stored (l = 0) / previous layer's operand written
                    fence_proxy_async();
                    tc_fence_after();
                    const uint32_t a16 = (sA + (uint32_t)slot * A_BYTES) >> 4;
                    // conv1 (even layers) overwrites accumulator C; conv2 accumulates onto the residual stream R
                    const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 128 + ((l & 1) ? 0 : 64));
                    if (elect_one()) {
                        // k = 3 ('same'): tap j reads rows j - 1 ..; k = 2 (one zero on the right): rows j ..
                        if (ntaps == 3) umma_conv_tile<R2_N, SPLIT, 3, 4>(d_tmem, a16 - 1u, 128u, w16, idesc, (uint32_t)(l & 1));
                        else umma_conv_tile<R2_N, SPLIT, 2, 4>(d_tmem, a16, 128u, w16, idesc, (uint32_t)(l & 1));
                        umma_commit(&acc_full[slot]);
                        if (l == R2_L - 1) umma_commit(&a_free[slot]);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&w_free[b]);
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: group g owns tile slot g
        const int g = (warp - 2) >> 2, q = warp & 3;
        const int v = q * 32 + lane;  // row of the tile
        const uint32_t tR = tmem_base + (uint32_t)(g * 128) + ((uint32_t)(q * 32) << 16), tC = tR + 64u;
        uint8_t *aslot = r2_smem + R2_PAD + (size_t)g * A_BYTES;
        for (int n = 0; n < n_groups; ++n) {
            const int tile = tile_lo + n * R2_SLOTS + g;
            if (tile >= tile_hi) break;
            const int seq = 2 * tile + (v >> 6), srow = v & 63;
            const bool row_ok = seq < P.NS && srow < T;
            const int64_t orow = (int64_t)seq * T + srow;
            {   // residual stream of the tile -> TMEM (the previous tile's last epilogue has read it: program order)
                const float *rp = P.xres + orow * R2_N;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[16];
                    if (row_ok) {
                        float4 a0, a1, a2, a3;
                        ld_global_256(rp + 16 * c, a0, a1);
                        ld_global_256(rp + 16 * c + 8, a2, a3);
                        x[0] = a0.x, x[1] = a0.y, x[2] = a0.z, x[3] = a0.w, x[4] = a1.x, x[5] = a1.y, x[6] = a1.z, x[7] = a1.w;
                        x[8] = a2.x, x[9] = a2.y, x[10] = a2.z, x[11] = a2.w, x[12] = a3.x, x[13] = a3.y, x[14] = a3.z, x[15] = a3.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) x[i] = 0.f;
                    }
                    tmem_st16(tR + (uint32_t)(16 * c), x);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_epi[g]);
            }
            for (int l = 0; l < R2_L; ++l) {
                mbar_wait(&acc_full[g], (uint32_t)l & 1u);  // phase n * 14 + l
                tc_fence_after();
                const uint32_t tacc = (l & 1) ? tR : tC;
                const float *pb = par + l * 3 * R2_N;
                const bool last = l == R2_L - 1;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int nb = 16 * c;
                    uint32_t r[16];
                    tmem_ld16_nowait(tacc + (uint32_t)nb, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float w[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(pb + nb + 4 * i);
                        w[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + b4.x, w[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y;
                        w[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z, w[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w;
                    }
                    if (!last) {  // pre-activation BatchNorm + ReLU of the next conv
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 s4 = *reinterpret_cast<const float4 *>(pb + R2_N + nb + 4 * i),
                                         h4 = *reinterpret_cast<const float4 *>(pb + 2 * R2_N + nb + 4 * i);
                            w[4 * i + 0] = fmaxf(fmaf(w[4 * i + 0], s4.x, h4.x), 0.f), w[4 * i + 1] = fmaxf(fmaf(w[4 * i + 1], s4.y, h4.y), 0.f);
                            w[4 * i + 2] = fmaxf(fmaf(w[4 * i + 2], s4.z, h4.z), 0.f), w[4 * i + 3] = fmaxf(fmaf(w[4 * i + 3], s4.w, h4.w), 0.f);
                        }
                    }
                    uint4 hi0, lo0, hi1, lo1;
                    pack8_split16<SPLIT>(&w[0], hi0, lo0);
                    pack8_split16<SPLIT>(&w[8], hi1, lo1);
                    if (!last) {  // in place: planes 2c, 2c + 1 of the tile's operand
                        uint8_t *d = aslot + ((size_t)(2 * c) * 128 + v) * 16;
                        *reinterpret_cast<uint4 *>(d) = hi0;
                        *reinterpret_cast<uint4 *>(d + 2048) = hi1;
                        if (SPLIT == 2) {
                            *reinterpret_cast<uint4 *>(d + 16384) = lo0;
                            *reinterpret_cast<uint4 *>(d + 16384 + 2048) = lo1;
                        }
                    } else {
                        uint16_t *yb = P.y + orow * R2_N + nb;
                        st_global_256(yb, hi0, hi1);
                        if (SPLIT == 2) st_global_256(yb + P.split16, lo0, lo1);
                    }
                }
                tc_fence_before();
                if (!last) {
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_epi[g]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

void resstack2_params(const float *const *b1, const float *const *b2, const float *const *n1s, const float *const *n1h,
                      const float *const *n2s, const float *const *n2h, std::vector<float> &par) {
    par.assign((size_t)R2_L * 3 * R2_N, 0.f);
    double cum[R2_N] = {0.0};
    for (int i = 0; i < 7; ++i) {
        float *a = par.data() + (size_t)(2 * i) * 3 * R2_N, *b = a + 3 * R2_N;
        for (int c = 0; c < R2_N; ++c) {
            a[c] = b1[i][c];
            a[R2_N + c] = n2s[i][c];
            a[2 * R2_N + c] = n2h[i][c];
            cum[c] += (double)b2[i][c];
            b[c] = (float)cum[c];
            b[R2_N + c] = i < 6 ? n1s[i + 1][c] : 1.f;
            b[2 * R2_N + c] = i < 6 ? n1h[i + 1][c] : 0.f;
        }
    }
}

int resstack2_launch(const ResStack2P &p, int split, cudaStream_t s) {
    VP_REQUIRE(p.T > 0 && p.T + 1 <= 64, VP_ERR_UNSUPPORTED, "res-CNN stack: sequences of %d rows do not fit the 64-row pitch", p.T);
    if (p.NS == 0) return VP_OK;
    VP_REQUIRE(reinterpret_cast<uintptr_t>(p.x) % 16 == 0 && reinterpret_cast<uintptr_t>(p.y) % 32 == 0 &&
                   reinterpret_cast<uintptr_t>(p.xres) % 32 == 0 && (p.split16 * 2) % 32 == 0,
               VP_ERR_ARG, "res-CNN stack: buffers are not 32-byte aligned");
    for (int l = 0; l < R2_L; ++l)
        VP_REQUIRE(reinterpret_cast<uintptr_t>(p.w[l]) % 16 == 0 && (p.ntaps[l] == 2 || p.ntaps[l] == 3), VP_ERR_ARG,
                   "res-CNN stack: layer %d weights", l);
    ResStack2K K;
    K.p = p;
    {   // (8 channels, t, seq, plane, split): the box [8][64][2][8][split] is one tile in UMMA order
        const uint64_t dims[5] = {8, (uint64_t)p.T, (uint64_t)p.NS, 8, (uint64_t)split};
        const uint64_t strides[4] = {128, (uint64_t)p.T * 128, 16, (uint64_t)p.split16 * 2};
        const uint32_t box[5] = {8, 64, 2, 8, (uint32_t)split};
        if (int rc = tma_encode_u16(&K.x_map, p.x, 5, dims, strides, box)) return rc;
    }
    const size_t a_bytes = (size_t)16384 * split, w_bytes = (size_t)12 * split * 2 * R2_N * 16;
    const size_t smem = R2_PAD + R2_SLOTS * a_bytes + 2 * w_bytes + (size_t)R2_L * 3 * R2_N * sizeof(float);
    const int n_tiles = (p.NS + 1) / 2;
    const unsigned grid = (unsigned)std::min(device_sm_count(), (n_tiles + R2_SLOTS - 1) / R2_SLOTS);
    KTimer kt(KC_RESSTACK, s);
    if (split == 2) {
        if (int rc = ensure_dyn_smem((const void *)resstack2_kernel<2>, smem)) return rc;
        resstack2_kernel<2><<<grid, R2_THREADS, smem, s>>>(K);
    } else {
        if (int rc = ensure_dyn_smem((const void *)resstack2_kernel<1>, smem)) return rc;
        resstack2_kernel<1><<<grid, R2_THREADS, smem, s>>>(K);
    }
    VP_LAUNCH_CHECK();
    return VP_OK;
}

// ------------------------------------------------------------------------------------------ tensor maps (host)
int tma_encode_u16(CUtensorMap *map, const void *gaddr, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                   const uint32_t *box) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;  // process-wide and immutable once resolved (the driver entry point does not depend on the device)
    if (fn == nullptr) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        VP_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        VP_REQUIRE(qres == cudaDriverEntryPointSuccess && ptr != nullptr, VP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    VP_REQUIRE(rank >= 1 && rank <= 5, VP_ERR_ARG, "tensor map rank %d", rank);
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void *>(gaddr), gdim, gstr, bdim, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VP_REQUIRE(r == CUDA_SUCCESS, VP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VP_OK;
}

}  // namespace vp
