// Fused encoder front (sm_100a): encoder.convs.1 -> .2 -> .3 of the EQTransformer, each Conv1d('same') + ReLU + MaxPool1d(2), in ONE
// kernel with the 1500- and 750-sample levels in TENSOR MEMORY (the layout idea of fused_dec2.cu run backwards).
//
// Replaces three layer-by-layer launches of tcconv.cu (SeisBench eqtransformer.py Encoder stages 1-3; restated at oracle/nets.py
// Encoder) that were bound by HBM: per window they moved 192 + 144 + 96 KB of 16-bit hi / lo activations at 45 - 60 % of the HBM
// peak; fused, the chain reads the 96 KB of the 3000-sample level and writes the 48 KB of the 375-sample level.
//
// A work item = 128 consecutive rows R0 .. R0 + 127 of the 375-sample level of one window; TMEM lane r holds what feeds row
// R0 + r: 8 samples of the 3000 level (input), 4 of the 1500 level, 2 of the 750 level.  MaxPool1d(2) pairs adjacent samples of ONE
// lane, so the pool is a max of two accumulator columns in the thread that owns the lane.
//
//   in (3000 level, 8 ch; one TMA box: rows = lanes, 8 planes = the lane's 8 samples)
//        --convs.1 folded over the 8 samples: A from shared memory, row taps R-1 / R / R+1, banded weights, N = 8 x 16-->
//   D1[r][t * 16 + c] --epilogue A: ReLU(max(t, t+1) + b) --> A2 (TMEM: 10 slots = samples 4R-3 .. 4R+6 of the 1500 level, 16 ch)
//        --convs.2: 4 blocks (one conv sample each, N = 16, 7 taps = 7 slots), A from TMEM-->
//   D2[r][t * 16 + c] --epilogue B--> A3 (TMEM: 8 slots = samples 2R-3 .. 2R+4 of the 750 level; halos come from lanes r +- 1, r +- 2)
//        --convs.3: 2 blocks (N = 32), A from TMEM-->
//   D3[r][t * 32 + c] --epilogue B: ReLU(max + b) --> y [split][window][375][32] channel-last 16-bit in HBM
//
// Lanes 3 .. 96 carry valid outputs (94 rows per item, four items per window).  Rows outside the window are zeros at every level
// (the convs' zero padding); the TMA box delivers the input's zeros.  TMEM columns: [0, 128) D1; [128, 288) A2; [288, 416) A3;
// [416, 480) D2 / D3.  The issuer interleaves convs.1 of item n + 1 with the TMEM half of item n.  25 warps: tcgen05 issuer, 16
// epilogue warps A (D1 -> A2, one pooled sample each; they also issue the TMA loads), 8 epilogue warps B (D2 -> A3, D3 -> HBM).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fused.cuh"
#include "fused_ts.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

constexpr int EA_THREADS = 32 * 25;
constexpr int EA_LANE_LO = 3, EA_USE = 94;
constexpr int EA_IN_ROWS = 130;  // input rows R0 - 1 .. R0 + 128
constexpr uint32_t EA_COL_D1 = 0, EA_COL_A2 = 128, EA_A2_LO = 80, EA_COL_A3 = 288, EA_A3_LO = 64, EA_COL_D23 = 416;
constexpr int EA_B1 = 0, EA_B2 = 16, EA_B3 = 32;  // bias_c offsets

#ifdef VP_EA_PROF
__device__ long long ea_prof_out[8];
#define EA_LAP(i) do { const long long _n = clock64(); ea_acc[i] += _n - ea_t; ea_t = _n; } while (0)
#else
#define EA_LAP(i) do { } while (0)
#endif

struct FzEncAK {
    alignas(64) CUtensorMap x_map;  // (8 channels, row of the 375 level, sample in the row, window, split) over the 3000-sample level input
    FzEncA p;
};

// convs.1 folded over the eight samples of a row: K steps = (row tap, plane pair); planes 4-7 of row R - 1, all of row R, 0-3 of R + 1
template <int SPLIT>
__device__ __forceinline__ void ea_conv1_tile(uint32_t d_tmem, uint32_t a16, uint32_t w16, uint32_t idesc) {
    constexpr int NTERM = SPLIT == 2 ? 3 : 1;
    constexpr uint32_t ROWS = EA_IN_ROWS;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
    const uint32_t a_base = a16 | (ROWS << 16);          // LBO = plane pitch
    const uint32_t b_base = w16 | (128u << 16);          // LBO = 128 rows of 16 bytes between the K halves
    constexpr int JR[8] = {0, 0, 1, 1, 1, 1, 2, 2}, P0[8] = {4, 6, 0, 2, 4, 6, 0, 2};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int t = 0; t < NTERM; ++t) {
            const int sa = (t == 2) ? 1 : 0, sb = (t == 1) ? 1 : 0;
            const uint32_t a_off = (uint32_t)((sa * 8 + P0[ks]) * ROWS + JR[ks]);
            const uint32_t b_off = (uint32_t)((ks * SPLIT + sb) * 2 * 128);
            umma_f16(d_tmem, desc_hi | (uint64_t)(a_base + a_off), desc_hi | (uint64_t)(b_base + b_off), idesc, (ks == 0 && t == 0) ? 0u : 1u);
        }
}

// 16 channels of two adjacent conv samples (accumulator columns colA.., colB..) -> relu(max + bias): Conv + ReLU + MaxPool1d(2)
__device__ __forceinline__ void ea_load_pool16(const float *bias, uint32_t tacc, int colA, int colB, float (&v)[16]) {
    uint32_t ra[16], rb[16];
    tmem_ld16_nowait(tacc + (uint32_t)colA, ra);
    tmem_ld16_nowait(tacc + (uint32_t)colB, rb);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 16; n += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + n);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[n + e] = fmaxf(fmaxf(__uint_as_float(ra[n + e]), __uint_as_float(rb[n + e])) + bb[e], 0.f);
    }
}

// one halo slot from the registers of the warp that produced the sample: slot `dst` of every lane <- (oh, ol) of lane (lane + delta)
template <int SPLIT>
__device__ __forceinline__ void ea_halo_regs(uint32_t tl, uint32_t col0, uint32_t lo_off, const uint32_t (&oh)[8], const uint32_t (&ol)[8], int dst,
                                             int delta, int lane, bool have, const uint32_t *post_a, const uint32_t *post_b) {
    uint32_t hh[8], ll[8];
    const int from = (lane + delta) & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        hh[i] = __shfl_sync(0xffffffffu, oh[i], from);
        ll[i] = SPLIT == 2 ? __shfl_sync(0xffffffffu, ol[i], from) : 0u;
    }
    // lanes whose source lane is in the neighbour quarter take that quarter's posted copy: delta > 0: lanes 32 - delta .. 31 (the
    // outermost takes post_b when two lanes are outside), delta < 0: lanes 0 .. -delta - 1
    const int out = delta > 0 ? lane - (32 - delta) : (-delta - 1) - lane;  // >= 0: outside; 0 = nearest the inside
    if (out >= 0) ts_fetch16((out == 0 || post_b == nullptr) ? post_a : post_b, have, hh, ll);
    ts_st_slot<SPLIT>(tl, col0 + 8 * dst, lo_off, hh, ll);
}

// ------------------------------------------------------------------------------------------ the kernel
template <int SPLIT>
__global__ void __launch_bounds__(EA_THREADS, 1) enca_kernel(const __grid_constant__ FzEncAK K) {
    const FzEncA &p = K.p;
    extern __shared__ __align__(128) uint8_t ea_smem[];
    __shared__ __align__(8) uint64_t in_full, d1_full, a2_full, d2_full, a3_full, d3_full, d23_free;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(ea_smem);
    const int n_items = p.B * p.tiles_per_seq;
    const int n_my = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(&in_full, 1);
        mbar_init(&d1_full, 1);
        mbar_init(&a2_full, 16);
        mbar_init(&d2_full, 1);
        mbar_init(&a3_full, 8);
        mbar_init(&d3_full, 1);
        mbar_init(&d23_free, 8);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    {   // resident weights
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.blob);
        for (int idx = tid; idx < p.blob_bytes / 16; idx += EA_THREADS) cp_async16(sbase + p.blob_off + idx * 16, wg + idx, 16u);
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    auto item_R0 = [&](int k, int &b) {
        const int it = blockIdx.x + k * gridDim.x;
        b = it / p.tiles_per_seq;
        return EA_USE * (it - b * p.tiles_per_seq) - EA_LANE_LO;
    };
    auto load_item = [&](int k) {  // the 130 input rows of item k as one TMA box [split][8 samples][130 rows][8 channels]
        int b;
        const int R0 = item_R0(k, b);
        mbar_arrive_expect_tx(&in_full, (uint32_t)(EA_IN_ROWS * 8 * 16 * SPLIT));
        tma_load_5d(sbase + p.in_off, &K.x_map, &in_full, 0, R0 - 1, 0, b, 0);
    };

    if (warp == 0) {
        // ================= tcgen05 issuer
        const uint32_t fmt = SPLIT == 2 ? 0u : 1u;
        const uint32_t id16 = umma_idesc(16, fmt), id32 = umma_idesc(32, fmt), id128 = umma_idesc(128, fmt);
        const uint32_t in16 = (sbase + p.in_off) >> 4;
        const uint32_t w1 = (sbase + p.w1_off) >> 4, w2 = (sbase + p.w2_off) >> 4, w3 = (sbase + p.w3_off) >> 4;
#ifdef VP_EA_PROF
        long long ea_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ea_t = clock64();
#endif
        for (int n = -1; n < n_my; ++n) {
            if (n >= 0) {  // convs.2: A2 (TMEM) -> D2
                mbar_wait(&a2_full, n & 1);
                EA_LAP(0);
                if (n > 0) mbar_wait(&d23_free, (n - 1) & 1);
                EA_LAP(1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll 1
                    for (int t = 0; t < 4; ++t)
                        umma_ts_block<16, SPLIT, 7>(tmem_base + EA_COL_D23 + 16 * t, tmem_base + EA_COL_A2 + 8 * t, EA_A2_LO, w2, id16);
                    umma_commit(&d2_full);
                }
                __syncwarp();
                EA_LAP(4);
            }
            if (n + 1 < n_my) {  // convs.1 of the next item: input box (shared memory) -> D1
                mbar_wait(&in_full, (n + 1) & 1);
                EA_LAP(2);
                tc_fence_after();
                if (elect_one()) {
                    ea_conv1_tile<SPLIT>(tmem_base + EA_COL_D1, in16, w1, id128);
                    umma_commit(&d1_full);
                }
                __syncwarp();
                EA_LAP(5);
            }
            if (n >= 0) {  // convs.3: A3 (TMEM) -> D3
                mbar_wait(&a3_full, n & 1);
                EA_LAP(3);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll 1
                    for (int t = 0; t < 2; ++t)
                        umma_ts_block<32, SPLIT, 7>(tmem_base + EA_COL_D23 + 32 * t, tmem_base + EA_COL_A3 + 8 * t, EA_A3_LO, w3, id32);
                    umma_commit(&d3_full);
                }
                __syncwarp();
                EA_LAP(6);
            }
        }
#ifdef VP_EA_PROF
        if (blockIdx.x == 0 && lane == 0)
            for (int i = 0; i < 8; ++i) ea_prof_out[i] = ea_acc[i];
#endif
    } else if (warp < 17) {
        // ================= epilogue warps A (16: four per lane quarter): D1 -> A2.  Warp (q, s) converts the conv samples 2 s, 2 s + 1 of
        // its lanes = pooled sample s, stores slot 3 + s and, from the same registers, the halo slots its neighbours need.
        const int q = warp & 3, s = (warp - 1) >> 2, r = q * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t *xa = reinterpret_cast<uint32_t *>(ea_smem + p.xch_off);  // [parity][quarter][first | last][3 samples][16]
        if (warp == 1 && lane == 0 && n_my > 0) load_item(0);
        for (int m = 0; m < n_my; ++m) {
            int b;
            const int R0 = item_R0(m, b);
            const bool valid = (unsigned)(R0 + r) < (unsigned)p.T0;
            mbar_wait(&d1_full, m & 1);
            if (warp == 1 && lane == 0 && m + 1 < n_my) load_item(m + 1);  // convs.1 has retired: the input box is free
            if (m > 0) mbar_wait(&d2_full, (m - 1) & 1);                    // convs.2 of the previous item has read A2
            tc_fence_after();
            uint32_t *xw = xa + (size_t)(((m & 1) * 4 + q) * 2) * 48;
            float v[16];
            ea_load_pool16(&p.bias_c[EA_B1], tl + EA_COL_D1, 32 * s, 32 * s + 16, v);
            uint32_t oh[8], ol[8];
            ts_pack16<SPLIT>(v, valid, oh, ol);
            // edge lanes post what the neighbour quarter needs: lane 0 its samples 0-2, lane 31 its samples 1-3
            if (lane == 0 && s < 3) ts_post16(xw + s * 16, oh, ol);
            if (lane == 31 && s >= 1) ts_post16(xw + 48 + (s - 1) * 16, oh, ol);
            ts_st_slot<SPLIT>(tl, EA_COL_A2 + 8 * (3 + s), EA_A2_LO, oh, ol);
            named_bar_sync(1, 512);  // the posts are visible
            if (s < 3) {  // right halo of lane r - 1 ... seen from the receiver: slot 7 + s = sample s of lane r + 1
                uint32_t hh[8], ll[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    hh[i] = __shfl_down_sync(0xffffffffu, oh[i], 1);
                    ll[i] = __shfl_down_sync(0xffffffffu, ol[i], 1);
                }
                if (lane == 31) ts_fetch16(xa + (size_t)(((m & 1) * 4 + (q + 1) % 4) * 2) * 48 + s * 16, q < 3, hh, ll);
                ts_st_slot<SPLIT>(tl, EA_COL_A2 + 8 * (7 + s), EA_A2_LO, hh, ll);
            }
            if (s >= 1) {  // left halo: slot s - 1 = sample s of lane r - 1
                uint32_t hh[8], ll[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    hh[i] = __shfl_up_sync(0xffffffffu, oh[i], 1);
                    ll[i] = __shfl_up_sync(0xffffffffu, ol[i], 1);
                }
                if (lane == 0) ts_fetch16(xa + (size_t)(((m & 1) * 4 + (q + 3) % 4) * 2 + 1) * 48 + (s - 1) * 16, q > 0, hh, ll);
                ts_st_slot<SPLIT>(tl, EA_COL_A2 + 8 * (s - 1), EA_A2_LO, hh, ll);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a2_full);
        }
    } else {
        // ================= epilogue warps B: D2 -> A3 (h = pooled sample of the lane's two), D3 -> y (h = channel half)
        const int q = warp & 3, h = (warp - 17) >> 2, r = q * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t *xb = reinterpret_cast<uint32_t *>(ea_smem + p.xch_off + 3072);  // [parity][quarter][first | last][3][16]
        for (int n = 0; n < n_my; ++n) {
            int b;
            const int R0 = item_R0(n, b);
            const bool valid = (unsigned)(R0 + r) < (unsigned)p.T0;
            // ---- E2
            mbar_wait(&d2_full, n & 1);
            tc_fence_after();
            uint32_t *xw = xb + (size_t)(((n & 1) * 4 + q) * 2) * 48;
            uint32_t oh[8], ol[8];  // this warp's pooled sample h of every lane, kept for the halo slots it supplies
            {
                float v[16];
                ea_load_pool16(&p.bias_c[EA_B2], tl + EA_COL_D23, 32 * h, 32 * h + 16, v);
                ts_pack16<SPLIT>(v, valid, oh, ol);
                ts_st_slot<SPLIT>(tl, EA_COL_A3 + 8 * (3 + h), EA_A3_LO, oh, ol);
                // posts: first = [lane 0 sample 0, lane 0 sample 1, lane 1 sample 0]; last = [lane 31 sample 0, lane 31 sample 1, lane 30 sample 1]
                if (lane == 0) ts_post16(xw + h * 16, oh, ol);
                if (lane == 1 && h == 0) ts_post16(xw + 2 * 16, oh, ol);
                if (lane == 31) ts_post16(xw + 48 + h * 16, oh, ol);
                if (lane == 30 && h == 1) ts_post16(xw + 48 + 2 * 16, oh, ol);
            }
            named_bar_sync(2, 256);  // the posts are visible
            {
                // The six halo slots by OWNER of the data (shuffled from registers, no TMEM re-read): sample 0 (warp h = 0) feeds slot 5
                // of lane r - 1, slot 7 of lane r - 2 and slot 1 of lane r + 1; sample 1 (warp h = 1) feeds slot 6 of lane r - 1, slot 2
                // of lane r + 1 and slot 0 of lane r + 2.  Seen from the receiving lane: (destination slot, lane distance of the source).
                const uint32_t *xn = xb + (size_t)(((n & 1) * 4 + (q + 1) % 4) * 2) * 48;      // next quarter: first = [l0 s0, l0 s1, l1 s0]
                const uint32_t *xp = xb + (size_t)(((n & 1) * 4 + (q + 3) % 4) * 2 + 1) * 48;  // previous quarter: last = [l31 s0, l31 s1, l30 s1]
                const bool hn = q < 3, hp = q > 0;
                if (h == 0) {
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 5, 1, lane, hn, xn, nullptr);
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 7, 2, lane, hn, xn, xn + 32);
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 1, -1, lane, hp, xp, nullptr);
                } else {
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 6, 1, lane, hn, xn + 16, nullptr);
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 2, -1, lane, hp, xp + 16, nullptr);
                    ea_halo_regs<SPLIT>(tl, EA_COL_A3, EA_A3_LO, oh, ol, 0, -2, lane, hp, xp + 16, xp + 32);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a3_full);
            // ---- E3: channels 16 h .. 16 h + 15 of the row
            mbar_wait(&d3_full, n & 1);
            tc_fence_after();
            {
                float v[16];
                ea_load_pool16(&p.bias_c[EA_B3 + 16 * h], tl + EA_COL_D23, 16 * h, 32 + 16 * h, v);
                uint32_t hh[8], ll[8];
                ts_pack16<SPLIT>(v, true, hh, ll);
                const int row = R0 + r;
                if (r >= EA_LANE_LO && r < EA_LANE_LO + EA_USE && (unsigned)row < (unsigned)p.T0) {
                    uint16_t *yb = p.y + ((size_t)b * p.T0 + row) * 32 + 16 * h;
                    *reinterpret_cast<uint4 *>(yb) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<uint4 *>(yb + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                    if (SPLIT == 2) {
                        *reinterpret_cast<uint4 *>(yb + p.y_split) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        *reinterpret_cast<uint4 *>(yb + p.y_split + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d23_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ host
static void ea_to16(float w, int split, uint16_t &hi, uint16_t &lo) {
    if (split == 2) {
        const __half h = __float2half_rn(w);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(w - __half2float(h)));
    } else {
        hi = __bfloat16_as_ushort(__float2bfloat16_rn(w));
        lo = 0;
    }
}

int enca_build(EncAPlan &plan, const TcLayer &enc2, const TcLayer &enc3, const float *w1 /*(16, 8, 9)*/, const float *b1, int split) {
    FzEncA &p = plan.p;
    std::memset(&p, 0, sizeof(p));
    plan.split = split;
    plan.ready = false;
    VP_REQUIRE(enc2.ph == 1 && enc2.cin == 16 && enc2.nout == 16 && enc2.sched_taps == 7 && enc2.sched_nq == 1 && enc2.row0 == -3 && enc2.groups == 1 &&
                   enc2.n_blocks == 7,
               VP_ERR_UNSUPPORTED, "enca: encoder.convs.2 differs from the compiled schedule");
    VP_REQUIRE(enc3.ph == 1 && enc3.cin == 16 && enc3.nout == 32 && enc3.sched_taps == 7 && enc3.sched_nq == 1 && enc3.row0 == -3 && enc3.groups == 1 &&
                   enc3.n_blocks == 7,
               VP_ERR_UNSUPPORTED, "enca: encoder.convs.3 differs from the compiled schedule");
    VP_REQUIRE(!enc2.blocks.empty() && !enc3.blocks.empty(), VP_ERR_ARG, "enca: host weight blocks are gone");
    p.T0 = 375;
    p.tiles_per_seq = (p.T0 + EA_USE - 1) / EA_USE;
    auto up128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    size_t off = 0;
    p.in_off = (int)off;
    off += up128((size_t)EA_IN_ROWS * 8 * 16 * split);
    p.xch_off = (int)off;
    off += 2 * 3072;
    p.blob_off = (int)off;
    const size_t blk1 = (size_t)split * 2 * 128 * 8, blk2 = (size_t)split * 2 * 16 * 8, blk3 = (size_t)split * 2 * 32 * 8;  // 16-bit elements
    size_t rel = 0;
    const size_t r1 = rel;
    rel += up128(8 * blk1 * 2);
    const size_t r2 = rel;
    rel += up128(7 * blk2 * 2);
    const size_t r3 = rel;
    rel += up128(7 * blk3 * 2);
    p.blob_bytes = (int)rel;
    p.w1_off = p.blob_off + (int)r1;
    p.w2_off = p.blob_off + (int)r2;
    p.w3_off = p.blob_off + (int)r3;
    p.smem_bytes = p.blob_off + p.blob_bytes;
    VP_REQUIRE(p.smem_bytes <= 226 * 1024, VP_ERR_UNSUPPORTED, "enca: %d bytes of shared memory", p.smem_bytes);
    plan.blob.assign(rel / 2, 0);
    uint16_t *dst = plan.blob.data();
    std::memcpy(dst + r2 / 2, enc2.blocks.data(), 7 * blk2 * 2);
    std::memcpy(dst + r3 / 2, enc3.blocks.data(), 7 * blk3 * 2);
    // encoder.convs.1 (16, 8, 9), 'same', folded over the 8 samples of a 375-level row: conv sample t (0..7) of row R reads the
    // 3000-level samples 8 R + d, d = t - 4 + k.  K step ks = (row tap jr, plane pair p0): d = (jr - 1) * 8 + p0 + k-half; K index
    // inside the step = k-half * 8 + ci; column n = t * 16 + co.  Blocks [ks][split][k-half][128][8].
    static const int JR[8] = {0, 0, 1, 1, 1, 1, 2, 2}, P0[8] = {4, 6, 0, 2, 4, 6, 0, 2};
    uint16_t *wb = dst + r1 / 2;
    for (int ks = 0; ks < 8; ++ks)
        for (int kh = 0; kh < 2; ++kh) {
            const int d = (JR[ks] - 1) * 8 + P0[ks] + kh;
            for (int t = 0; t < 8; ++t) {
                const int k = d - t + 4;
                if (k < 0 || k > 8) continue;
                for (int co = 0; co < 16; ++co)
                    for (int ci = 0; ci < 8; ++ci) {
                        uint16_t hi, lo;
                        ea_to16(w1[((size_t)co * 8 + ci) * 9 + k], split, hi, lo);
                        const int n = t * 16 + co;
                        wb[(size_t)ks * blk1 + ((size_t)(0 * 2 + kh) * 128 + n) * 8 + ci] = hi;
                        if (split == 2) wb[(size_t)ks * blk1 + ((size_t)(1 * 2 + kh) * 128 + n) * 8 + ci] = lo;
                    }
            }
        }
    for (int n = 0; n < 16; ++n) p.bias_c[EA_B1 + n] = b1 ? b1[n] : 0.f;
    for (int n = 0; n < 16; ++n) p.bias_c[EA_B2 + n] = enc2.bias[n];
    for (int n = 0; n < 32; ++n) p.bias_c[EA_B3 + n] = enc3.bias[n];
    return VP_OK;
}

int enca_upload(EncAPlan &plan) {
    VP_CUDA_CHECK(cudaMalloc(&plan.d_blob, plan.blob.size() * sizeof(uint16_t)));
    VP_CUDA_CHECK(cudaMemcpy(plan.d_blob, plan.blob.data(), plan.blob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    plan.blob.clear();
    plan.blob.shrink_to_fit();
    plan.ready = true;
    return VP_OK;
}

void enca_free(EncAPlan &plan) {
    if (plan.d_blob) cudaFree(plan.d_blob);
    plan.d_blob = nullptr;
    plan.ready = false;
}

template <int SPLIT>
static int enca_launch_t(const FzEncA &p, dim3 grid, cudaStream_t s) {
    auto kern = enca_kernel<SPLIT>;
    FzEncAK K;
    K.p = p;
    {
        VP_REQUIRE((p.x_split * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(p.x) % 16 == 0, VP_ERR_ARG, "enca: the input must be 16-byte aligned");
        const uint64_t dims[5] = {8, (uint64_t)p.T0, 8, (uint64_t)p.B, (uint64_t)SPLIT};
        const uint64_t strides[4] = {128, 16, (uint64_t)8 * p.T0 * 16, (uint64_t)p.x_split * 2};
        const uint32_t box[5] = {8, (uint32_t)EA_IN_ROWS, 8, 1, (uint32_t)SPLIT};
        if (int rc = tma_encode_u16(&K.x_map, p.x, 5, dims, strides, box)) return rc;
    }
    if (int rc = ensure_dyn_smem((const void *)kern, (size_t)p.smem_bytes)) return rc;
    KTimer kt(KC_TCCONV, s);
    kern<<<grid, EA_THREADS, p.smem_bytes, s>>>(K);
    VP_LAUNCH_CHECK();
#ifdef VP_EA_PROF
    {
        cudaStreamSynchronize(s);
        long long h[8];
        cudaMemcpyFromSymbol(h, ea_prof_out, sizeof(h));
        const int n_my = (p.B * p.tiles_per_seq - 1) / (int)grid.x + 1;
        fprintf(stderr, "[enca prof] %d items per CTA; issuer cycles per item: waits a2_full %.0f d23_free %.0f in_full %.0f a3_full %.0f | issue convs.2 %.0f convs.1 %.0f convs.3 %.0f\n",
                n_my, (double)h[0] / n_my, (double)h[1] / n_my, (double)h[2] / n_my, (double)h[3] / n_my, (double)h[4] / n_my, (double)h[5] / n_my, (double)h[6] / n_my);
    }
#endif
    return VP_OK;
}

int enca_launch(const EncAPlan &plan, const uint16_t *x, long long x_split, int B, uint16_t *y, long long y_split, cudaStream_t s) {
    VP_REQUIRE(plan.ready, VP_ERR_UNSUPPORTED, "enca: plan not uploaded");
    if (B == 0) return VP_OK;
    FzEncA p = plan.p;
    p.x = x;
    p.x_split = x_split;
    p.B = B;
    p.blob = plan.d_blob;
    p.y = y;
    p.y_split = y_split;
    const int n_items = B * p.tiles_per_seq;
    dim3 grid((unsigned)std::min(std::max(device_sm_count(), 1), n_items));
    return plan.split == 2 ? enca_launch_t<2>(p, grid, s) : enca_launch_t<1>(p, grid, s);
}

}  // namespace vp
