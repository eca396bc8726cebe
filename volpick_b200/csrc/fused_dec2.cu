// Fused decoder tail, second generation (sm_100a): decoder.convs.3 -> 4 -> 5 -> 6 -> sigmoid(conv k11) with the activations of
// the two widest levels in TENSOR MEMORY, read by tcgen05.mma as its A operand.
//
// Replaces the same reference ops as fused_dec.cu (SeisBench eqtransformer.py Decoder stages 3-6 + conv_d / pick_convs of the
// three decoders; restated at oracle/nets.py Decoder).  Why a second kernel: fused_dec.cu reads every A operand from shared
// memory, and a tcgen05.mma of N <= 64 then costs 39 - 48 cycles whatever its math (4 KB of A at 128 B / cycle; measured with
// tools/ts_probe.cu) -- the 3000- and 6000-sample levels (N = 32 / 16) ran at a quarter of the tensor pipe's rate and the
// kernel saturated the shared-memory pipe (profiles/r02_decb_experiments.md).  With A in tensor memory the same MMAs cost
// N / 2 cycles (16 / 9.9 measured) and touch shared memory only for their 0.5 - 1 KB of weights.
//
// Layout.  A work item = 128 consecutive rows R0 .. R0 + 127 of the 375-sample level of one (window, decoder); TMEM lane r
// holds everything that descends from row R0 + r: 2 samples of the 750 level, 4 of the 1500 level, 8 of the 3000 level, 16
// outputs.  A conv tap cannot shift TMEM lanes, so each level stores, next to a lane's own samples, the halo samples of its
// two neighbour lanes (copied by warp shuffles in the epilogue that produces the level); a tap is then a COLUMN offset of
// the A operand:
//
//   in (375 level, 32 ch, shared memory, one TMA box)                 --dec3: A from shared memory, taps = row offsets-->
//   D0[r][phase * 32 + c]  --epilogue A--> S1 (shared memory, rows = lanes, K = 2 samples x 32 ch)
//                                                     --dec4 folded over 2 samples: 3 row taps, banded weights, N = 4 x 16-->
//   D1[r][sample * 16 + c] --epilogue A--> A2 (TMEM: 8 slots = samples 4R-2 .. 4R+5 of the 1500 level, 16 ch, fp16 hi | lo)
//                                                     --dec5: 4 blocks (one input sample -> 2 outputs, N = 32), A from TMEM-->
//   D2[r][block * 32 + phase * 16 + c] --epilogue B--> A3 (TMEM: 14 slots = samples 8R-3 .. 8R+10 of the 3000 level)
//                                                     --dec6: 8 blocks (N = 16), A from TMEM-->
//   D3[r][block * 16 + phase * 8 + c]  --epilogue B--> fp32 planar [c][16 r + i] in shared memory --head warps--> y
//
// Every layer loses one lane on each side of the item (its halo comes from the neighbour lane): lanes 4 .. 123 carry valid
// outputs; an item uses lanes 4 .. 115 = 112 rows = 1792 output samples, three items cover the 5000 kept samples of a blinded window.  Rows
// outside the sequence are written as zeros at every level (they are the convs' zero padding).
//
// TMEM columns (512): [0, 128) A2, which also hosts the accumulators D0 / D1 of the NEXT item while A2 is dead; [128, 352) A3;
// [352, 480) D2 / D3.  The issuer interleaves the shared-memory half of item n + 1 with the TMEM half of item n
// (L2(n), L0(n+1), L1(n+1), L3(n)), so each epilogue runs under the other item's MMAs.
// Warps (24): tcgen05 issuer; 8 epilogue warps A (D0 -> S1, D1 -> A2; they also issue the TMA load of the next item); 8 epilogue warps
// B (D2 -> A3, D3 -> head buffer); 7 head warps (sigmoid(conv k11), 8 outputs per thread).  Two epilogue warps share a TMEM lane
// quarter and split its columns.  Hand-over by mbarriers; the two halves of a group meet at one named barrier per item (halo exchange).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fused.cuh"
#include "fused_ts.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

constexpr int D2_THREADS = 32 * 24;
constexpr int D2_LANE_LO = 4, D2_USE = 112;  // lanes [4, 124) of an item are valid at the output; [4, 116) are used: three items still cover
                                             // the 313 kept rows of a blinded window, and 2 x 112 head threads are seven warps -> 24 warps, 80 registers
constexpr int D2_IN_ROWS = 132;              // input rows R0 - 2 .. R0 + 129
constexpr int D2_S1_ROWS = 130;              // lanes -1 .. 128 of the 750 level (rows 0 and 129 stay zero)
constexpr int D2_RP = 2048;                  // floats per channel of the head buffer
constexpr uint32_t D2_COL_A2 = 0, D2_A2_LO = 64, D2_COL_A3 = 128, D2_A3_LO = 112, D2_COL_D23 = 352;
// offsets into FzDecB2::bias_c (floats): dec3 [64] | dec4 [16] | dec5 [32] | dec6 [16]
constexpr int D2_B0 = 0, D2_B1 = 64, D2_B2 = 80, D2_B3 = 112;

// -DVP_D2_PROF build (tools/build_variant.sh): every warp of CTA (0, 0) accumulates the cycles of its sections (waits by barrier,
// work by epilogue) and decb2_launch prints them.  Empty in the product build.
#ifdef VP_D2_PROF
__device__ long long d2_prof_out[25 * 8];
struct D2Prof {
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t = 0;
    __device__ __forceinline__ void start() { t = clock64(); }
    __device__ __forceinline__ void lap(int s) {
        const long long n = clock64();
        acc[s] += n - t;
        t = n;
    }
    __device__ __forceinline__ void flush(int warp, int lane) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
            for (int i = 0; i < 8; ++i) d2_prof_out[warp * 8 + i] = acc[i];
    }
};
#else
struct D2Prof {
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void lap(int) {}
    __device__ __forceinline__ void flush(int, int) {}
};
#endif

struct FzDecB2K {
    alignas(64) CUtensorMap x_map;  // (8 channels, t, group * B + window, 8-channel plane, split) over the 375-sample level input
    FzDecB2 p;
};

// Head buffer: fp32 planar [c][sample], addressed in 16-byte units of 4 samples.  Both its writers (lane r stores unit 4 r + bp)
// and its readers (head thread e loads units 14 + 4 e + q) have a lane stride of four units; XOR-ing the low two unit bits with
// bits 3-4 spreads the eight lanes of a quarter warp over the eight 16-byte bank groups.
__device__ __forceinline__ int d2_unit(int u) { return u ^ ((u >> 3) & 3); }

// sigmoid(conv k11, 8 -> 1): head thread (rr, half) = the outputs 8 half .. 8 half + 7 of lane 4 + rr.  Consecutive lanes of a warp are
// consecutive TMEM lanes (16 samples = four 16-byte units apart, the stride d2_unit() is conflict-free for); the two halves of a lane
// run in different warps.  Weights come through the constant bank.  Rolled over the channels on purpose: fully unrolled, the head alone
// was 40 KB of straight-line code executed once per item next to the other warp roles with bodies of the same size, and the
// instruction caches (6 KB L0, 32 KB L1.5) served none of them (9.24 -> 7.38 ms per station-day from rolling this kernel's loops).
__device__ __forceinline__ void d2_head(const FzDecB2 &p, int g, int b, int R0, int rr, int half, const float *buf) {
    const int row = R0 + D2_LANE_LO + rr;
    if (rr >= D2_USE || row >= p.row_hi || (unsigned)row >= (unsigned)p.T0) return;  // rows past the kept range: blinded, never read
    const int u0 = 4 * (D2_LANE_LO + rr) + 2 * half - 2;  // unit of sample (first output) - 8
    const float *hw = p.head_c[g];
    // Packed FFMA2 (two fp32 FMAs per instruction, scalar weight broadcast): taps with an even first input index pair the
    // outputs (0,1) .. (6,7) in accA, the others pair (1,2) .. (5,6) in accB and leave outputs 0 and 7 to scalar FMAs: 50 instead of 88
    // FMA-pipe instructions per channel.
    float2 accA[4], accB[3];
    float s0 = 0.f, s7 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) accA[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) accB[i] = make_float2(0.f, 0.f);
    int uo[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) uo[q] = 4 * d2_unit(u0 + q);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        float xv[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const float4 x4 = *reinterpret_cast<const float4 *>(buf + (size_t)c * D2_RP + uo[q]);
            xv[4 * q] = x4.x;
            xv[4 * q + 1] = x4.y;
            xv[4 * q + 2] = x4.z;
            xv[4 * q + 3] = x4.w;
        }
        const float4 w0 = *reinterpret_cast<const float4 *>(hw + c * 12), w1 = *reinterpret_cast<const float4 *>(hw + c * 12 + 4),
                     w2 = *reinterpret_cast<const float4 *>(hw + c * 12 + 8);
        const float wk[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int k = 0; k < 11; ++k) {  // output o reads xv[o + k + 3] (xv[0] = sample (first output) - 8)
            const float w = wk[k];
            if (k & 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ffma2(accA[i], xv[2 * i + k + 3], xv[2 * i + k + 4], w);
            } else {
                s0 = fmaf(w, xv[k + 3], s0);
#pragma unroll
                for (int i = 0; i < 3; ++i) ffma2(accB[i], xv[2 * i + 1 + k + 3], xv[2 * i + 1 + k + 4], w);
                s7 = fmaf(w, xv[7 + k + 3], s7);
            }
        }
    }
    const float hb = hw[96];
    const float acc[8] = {accA[0].x + s0 + hb,        accA[0].y + accB[0].x + hb, accA[1].x + accB[0].y + hb, accA[1].y + accB[1].x + hb,
                          accA[2].x + accB[1].y + hb, accA[2].y + accB[2].x + hb, accA[3].x + accB[2].y + hb, accA[3].y + s7 + hb};
    // (the SFU form of the sigmoid -- ex2.approx + rcp.approx, 27 % fewer head instructions -- was measured neutral: 5.93 vs 5.86 ms)
    float *yb = p.y + ((size_t)b * 3 + g) * p.L_out + (size_t)16 * row + 8 * half;
#pragma unroll
    for (int o = 0; o < 8; o += 4) {
        float4 o4;
        o4.x = 1.f / (1.f + expf(-acc[o]));
        o4.y = 1.f / (1.f + expf(-acc[o + 1]));
        o4.z = 1.f / (1.f + expf(-acc[o + 2]));
        o4.w = 1.f / (1.f + expf(-acc[o + 3]));
        *reinterpret_cast<float4 *>(yb + o) = o4;
    }
}

// ------------------------------------------------------------------------------------------ the kernel
// Warps: 0 issuer; 1-8 epilogue A (two per TMEM lane quarter: quarter q = warp % 4, half h); 9-16 epilogue B; 17-24 head.
template <int SPLIT>
__global__ void __launch_bounds__(D2_THREADS, 1) decb2_kernel(const __grid_constant__ FzDecB2K K) {
    const FzDecB2 &p = K.p;
    extern __shared__ __align__(128) uint8_t d2_smem[];
    __shared__ __align__(8) uint64_t in_full, d0_full, s1_full, d1_full, a2_full, d2_full[4], a3_full, d3_full[4], d23_free, head_go, head_done;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const uint32_t sbase = smem_u32(d2_smem);
    const int n_items = p.B * p.tiles_per_seq;
    const int n_my = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(&in_full, 1);
        mbar_init(&d0_full, 1);
        mbar_init(&s1_full, 8);
        mbar_init(&d1_full, 1);
        mbar_init(&a2_full, 8);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d3_full[i], 1);
        }
        mbar_init(&a3_full, 8);
        mbar_init(&d23_free, 8);
        mbar_init(&head_go, 8);
        mbar_init(&head_done, 7);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    {   // resident weights of this decoder; the two halo rows of S1 that no lane writes
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.blob + (long long)g * (p.blob_bytes / 2));
        for (int idx = tid; idx < p.blob_bytes / 16; idx += D2_THREADS) cp_async16(sbase + p.blob_off + idx * 16, wg + idx, 16u);
        if (tid < SPLIT * 8 * 2) {
            const int pl = tid >> 1, row = (tid & 1) ? D2_S1_ROWS - 1 : 0;
            *reinterpret_cast<uint4 *>(d2_smem + p.s1_off + ((size_t)pl * D2_S1_ROWS + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // the 132 input rows of item k as one TMA box (rows outside the sequence arrive as zeros); issued by one thread of epilogue A as
    // soon as the previous item's first layer has retired (its accumulator is full), so there is no loader warp and no "empty" barrier
    auto load_item = [&](int k) {
        const int it = blockIdx.x + k * gridDim.x;
        const int b = it / p.tiles_per_seq, j = it - b * p.tiles_per_seq;
        const int R0 = p.row_off0 + D2_USE * j - D2_LANE_LO;
        mbar_arrive_expect_tx(&in_full, (uint32_t)(D2_IN_ROWS * 4 * 16 * SPLIT));
        tma_load_5d(sbase + p.in_off, &K.x_map, &in_full, 0, R0 - 2, g * p.B + b, 0, 0);
    };

    if (warp == 0) {
        // ================= tcgen05 issuer
        const uint32_t fmt = SPLIT == 2 ? 0u : 1u;
        const uint32_t id16 = umma_idesc(16, fmt), id32 = umma_idesc(32, fmt), id64 = umma_idesc(64, fmt), id128 = umma_idesc(128, fmt);
        const uint32_t in16 = (sbase + p.in_off) >> 4, s116 = (sbase + p.s1_off) >> 4;
        const uint32_t w0 = (sbase + p.w0_off) >> 4, w1 = (sbase + p.w1_off) >> 4, w2 = (sbase + p.w2_off) >> 4, w3 = (sbase + p.w3_off) >> 4;
        (void)id128;
        D2Prof prof;  // 0 in_full, 1 s1_full, 2 a2_full, 3 d23_free, 4 a3_full, 5 issue
        prof.start();
        auto issue_L0 = [&](int m) {  // dec3: in (shared) -> D0 = columns [0, 64) (+ [64, 128): the stacked W_lo half)
            mbar_wait(&in_full, m & 1);
            prof.lap(0);
            tc_fence_after();
            if (elect_one()) {
                if constexpr (SPLIT == 2)
                    ts_conv_tile_stacked<64, 5, 2>(tmem_base + D2_COL_A2, in16, D2_IN_ROWS, w0, id128, id64);
                else
                    ts_conv_tile<64, 1, 5, 2>(tmem_base + D2_COL_A2, in16, D2_IN_ROWS, w0, id64);
                umma_commit(&d0_full);
            }
            __syncwarp();
        };
        auto issue_L1 = [&](int m) {  // dec4 folded: S1 (shared) -> D1 = columns [0, 64) (+ [64, 128): the stacked W_lo half)
            mbar_wait(&s1_full, m & 1);
            prof.lap(1);
            tc_fence_after();
            if (elect_one()) {
                if constexpr (SPLIT == 2)
                    ts_conv_tile_stacked<64, 3, 4>(tmem_base + D2_COL_A2, s116, D2_S1_ROWS, w1, id128, id64);
                else
                    ts_conv_tile<64, 1, 3, 4>(tmem_base + D2_COL_A2, s116, D2_S1_ROWS, w1, id64);
                umma_commit(&d1_full);
            }
            __syncwarp();
        };
        auto issue_L2 = [&](int n) {  // dec5: A2 (TMEM) -> D2 blocks
            mbar_wait(&a2_full, n & 1);
            prof.lap(2);
            // (releasing the D2 / D3 columns pair by pair, so that dec5 of the next item starts under the last head-buffer stores,
            // was measured SLOWER twice: 6.99 vs 6.57 ms per station-day, and 6.17 vs 5.86 under the final issue order)
            if (n > 0) mbar_wait(&d23_free, (n - 1) & 1);
            prof.lap(3);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll 1
                for (int b = 0; b < 4; ++b) {
                    umma_ts_block<32, SPLIT, 5>(tmem_base + D2_COL_D23 + 32 * b, tmem_base + D2_COL_A2 + 8 * b, D2_A2_LO, w2, id32);
                    umma_commit(&d2_full[b]);
                }
            }
            __syncwarp();
        };
        auto issue_L3 = [&](int n, int half) {  // dec6: A3 (TMEM) -> D3 blocks
            if (half == 0) {
                mbar_wait(&a3_full, n & 1);
                prof.lap(4);
                tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll 1
                for (int bb = 0; bb < 4; ++bb) {
                    const int b = 4 * half + bb;
                    umma_ts_block<16, SPLIT, 7>(tmem_base + D2_COL_D23 + 16 * b, tmem_base + D2_COL_A3 + 8 * b, D2_A3_LO, w3, id16);
                    if (b & 1) umma_commit(&d3_full[b >> 1]);
                }
            }
            __syncwarp();
        };
        // one call site per layer (the lambdas are inlined; this kernel's speed follows the size of its hot code): iteration n = -1
        // is the prologue, the first two layers of item 0
        for (int n = -1; n < n_my; ++n) {
            if (n >= 0) {
                issue_L2(n);
                prof.lap(5);
            }
            if (n + 1 < n_my) {
                if (n >= 0 && (p.dbg & 16)) {  // debug: drain dec5 before the next item's first accumulator lands in the A2 columns
                    mbar_wait(&d2_full[3], n & 1);
                    tc_fence_after();
                }
                issue_L0(n + 1);
                prof.lap(5);
            }
            // The second shared-memory layer of the next item goes BEFORE dec6 of the current one (5.77 ms per station-day; between
            // the halves of dec6, VP_DECB_DBG bit 5, it is 6.28): its epilogue (D1 -> A2) then runs under all of dec6, and the
            // level-3000 epilogue, the longest of the chain, under both shared-memory layers.
            const int l1_at = (p.dbg & 32) ? 0 : -1;
            if (l1_at < 0 && n + 1 < n_my) {
                issue_L1(n + 1);
                prof.lap(5);
            }
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (n >= 0) {
                    issue_L3(n, half);
                    prof.lap(5);
                }
                if (half == l1_at && n + 1 < n_my) {
                    issue_L1(n + 1);
                    prof.lap(5);
                }
            }
        }
        prof.flush(warp, lane);
    } else if (warp < 9) {
        // ================= epilogue warps A: D0 -> S1 (shared memory), D1 -> A2 (tensor memory).  Half h = phase (E0) / sample pair (E1).
        const int q = warp & 3, h = (warp - 1) >> 2, r = q * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t *xa = reinterpret_cast<uint32_t *>(d2_smem + p.xch_off);  // [parity][quarter][side][sample 2][16]
        uint8_t *s1 = d2_smem + p.s1_off;
        float b1r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) b1r[i] = p.bias_c[g][D2_B1 + i];
        D2Prof prof;  // 0 d0_full, 1 d1_full, 2 d2_full[3], 3 E0, 4 E1
        prof.start();
        if (warp == 1 && lane == 0 && n_my > 0) load_item(0);
        for (int m = 0; m < n_my; ++m) {
            const int it = blockIdx.x + m * gridDim.x;
            const int b = it / p.tiles_per_seq, j = it - b * p.tiles_per_seq;
            const int R0 = p.row_off0 + D2_USE * j - D2_LANE_LO;
            const bool valid = (unsigned)(R0 + r) < (unsigned)p.T0;
            // ---- E0: this warp converts phase h (accumulator columns 32 h .. 32 h + 31)
            mbar_wait(&d0_full, m & 1);
            prof.lap(0);
            if (warp == 1 && lane == 0 && m + 1 < n_my) load_item(m + 1);  // dec3 has retired: the input box is free
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = 32 * h + 16 * cc;
                float v[16];
                ts_load16<SPLIT == 2 ? 64 : 0>(&p.bias_c[g][D2_B0 + c0], tl + D2_COL_A2, c0, v);
                uint32_t hh[8], ll[8];
                ts_pack16<SPLIT>(v, valid, hh, ll);
                const int pl = h * 4 + 2 * cc;  // plane' = phase * 4 + channel plane
                uint8_t *d = s1 + ((size_t)pl * D2_S1_ROWS + r + 1) * 16;
                *reinterpret_cast<uint4 *>(d) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                *reinterpret_cast<uint4 *>(d + (size_t)D2_S1_ROWS * 16) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                if (SPLIT == 2) {
                    *reinterpret_cast<uint4 *>(d + (size_t)8 * D2_S1_ROWS * 16) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                    *reinterpret_cast<uint4 *>(d + (size_t)9 * D2_S1_ROWS * 16) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s1_full);
            prof.lap(3);
            // ---- E1: this warp converts samples 2 h, 2 h + 1.  D1 sits in the A2 columns: the own slots (columns 16 .. 47 / 80 .. 111)
            // may be stored right away -- warp h = 0 reads columns 0 .. 31 / 64 .. 95 and stores 16 .. 31 / 80 .. 95, warp h = 1 reads
            // 32 .. 63 / 96 .. 127 and stores 32 .. 47 / 96 .. 111 -- the halo slots only after the group barrier (everyone has read).
            mbar_wait(&d1_full, m & 1);
            prof.lap(1);
            if (m > 0) mbar_wait(&d2_full[3], (m - 1) & 1);  // dec5 of the previous item has read A2
            prof.lap(2);
            tc_fence_after();
            uint32_t oh[2][8], ol[2][8];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                float v[16];
                ts_load16r<SPLIT == 2 ? 64 : 0>(b1r, tl + D2_COL_A2, 16 * (2 * h + s), v);
                ts_pack16<SPLIT>(v, valid, oh[s], ol[s]);
            }
            // edge lanes post what the neighbour quarter needs: lane 0 of h = 0 its samples 0, 1; lane 31 of h = 1 its samples 2, 3
            uint32_t *xw = xa + (size_t)(((m & 1) * 4 + q) * 2 + h) * 32;
            if (lane == (h ? 31 : 0)) {
#pragma unroll
                for (int s = 0; s < 2; ++s) ts_post16(xw + s * 16, oh[s], ol[s]);
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) ts_st_slot<SPLIT>(tl, D2_COL_A2 + 8 * (2 + 2 * h + s), D2_A2_LO, oh[s], ol[s]);
            tc_fence_before();
            named_bar_sync(1, 256);
            tc_fence_after();
            {   // h = 1: left halo, slots 0, 1 = samples 2, 3 of lane r - 1; h = 0: right halo, slots 6, 7 = samples 0, 1 of lane r + 1
                const uint32_t *xe = h ? xa + (size_t)(((m & 1) * 4 + (q + 3) % 4) * 2 + 1) * 32 : xa + (size_t)(((m & 1) * 4 + (q + 1) % 4) * 2) * 32;
                const int from = (lane + (h ? 31 : 1)) & 31, edge = h ? 0 : 31;
                const bool have = h ? q > 0 : q < 3;
                const uint32_t dst0 = D2_COL_A2 + 8 * (h ? 0 : 6);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    uint32_t hh[8], ll[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        hh[i] = __shfl_sync(0xffffffffu, oh[s][i], from);
                        ll[i] = __shfl_sync(0xffffffffu, ol[s][i], from);
                    }
                    if (lane == edge) ts_fetch16(xe + s * 16, have, hh, ll);
                    ts_st_slot<SPLIT>(tl, dst0 + 8 * s, D2_A2_LO, hh, ll);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a2_full);
            prof.lap(4);
        }
        prof.flush(warp, lane);
    } else if (warp < 17) {
        // ================= epilogue warps B: D2 -> A3 (tensor memory), D3 -> head buffer (shared memory).  Half h = phase (E2) / block
        // pairs 2 h, 2 h + 1 (E3).
        const int q = warp & 3, h = (warp - 9) >> 2, r = q * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t *xb = reinterpret_cast<uint32_t *>(d2_smem + p.xch_off + 2048);  // [parity][quarter][side][sample 3][16]
        float *hbuf = reinterpret_cast<float *>(d2_smem + p.head_off);
        float b2r[16], b3r[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) b2r[i] = p.bias_c[g][D2_B2 + 16 * h + i];
#pragma unroll
        for (int i = 0; i < 8; ++i) b3r[i] = p.bias_c[g][D2_B3 + i];
        D2Prof prof;  // 0 d2_full, 1 head_done, 2 d3_full, 3 E2 blocks, 4 halo exchange, 5 E3
        prof.start();
        for (int n = 0; n < n_my; ++n) {
            const int it = blockIdx.x + n * gridDim.x;
            const int b = it / p.tiles_per_seq, j = it - b * p.tiles_per_seq;
            const int R0 = p.row_off0 + D2_USE * j - D2_LANE_LO;
            const bool valid = (unsigned)(R0 + r) < (unsigned)p.T0;
            // ---- E2: block bb, phase h = sample s = 2 bb + h of the lane's eight 3000-level samples -> slot 3 + s
            uint32_t *xw = xb + (size_t)(((n & 1) * 4 + q) * 2) * 48;
#pragma unroll 1
            for (int bb = 0; bb < 4; ++bb) {
                mbar_wait(&d2_full[bb], n & 1);
                prof.lap(0);
                tc_fence_after();
                float v[16];
                ts_load16r<0>(b2r, tl + D2_COL_D23, 32 * bb + 16 * h, v);
                uint32_t hh[8], ll[8];
                ts_pack16<SPLIT>(v, valid, hh, ll);
                const int s = 2 * bb + h;
                ts_st_slot<SPLIT>(tl, D2_COL_A3 + 8 * (3 + s), D2_A3_LO, hh, ll);
                // the quarter's edge lanes post what their neighbour quarter needs: lane 0 samples 0-2, lane 31 samples 5-7
                if (s < 3 && lane == 0) ts_post16(xw + s * 16, hh, ll);
                if (s >= 5 && lane == 31) ts_post16(xw + 48 + (s - 5) * 16, hh, ll);
                prof.lap(3);
            }
            tmem_st_wait();
            tc_fence_before();
            named_bar_sync(2, 256);
            tc_fence_after();
            {
                // h = 0: right halo, slot 11 + i = sample i of lane r + 1 (its slot 3 + i); h = 1: left halo, slot i = sample 5 + i of lane
                // r - 1 (its slot 8 + i).  One code path for both directions and a rolled loop: this kernel's speed follows the size of
                // its hot code (25 warps in four roles share a 32 KB instruction cache).
                const uint32_t *xe = h == 0 ? xb + (size_t)(((n & 1) * 4 + (q + 1) % 4) * 2) * 48      // next quarter's lane 0: samples 0-2
                                            : xb + (size_t)(((n & 1) * 4 + (q + 3) % 4) * 2 + 1) * 48;  // previous quarter's lane 31: samples 5-7
                const uint32_t src0 = D2_COL_A3 + 8 * (h == 0 ? 3 : 8), dst0 = D2_COL_A3 + 8 * (h == 0 ? 11 : 0);
                const int from = (lane + (h == 0 ? 1 : 31)) & 31, edge = h == 0 ? 31 : 0;
                const bool have = h == 0 ? q < 3 : q > 0;
#pragma unroll 1
                for (int i3 = 0; i3 < 3; ++i3) {
                    uint32_t hh[8], ll[8];
                    tmem_ld8_nowait(tl + src0 + 8 * i3, hh);
                    if (SPLIT == 2) tmem_ld8_nowait(tl + src0 + D2_A3_LO + 8 * i3, ll);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        hh[i] = __shfl_sync(0xffffffffu, hh[i], from);
                        if (SPLIT == 2) ll[i] = __shfl_sync(0xffffffffu, ll[i], from);
                    }
                    if (lane == edge) ts_fetch16(xe + i3 * 16, have, hh, ll);
                    ts_st_slot<SPLIT>(tl, dst0 + 8 * i3, D2_A3_LO, hh, ll);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a3_full);
            prof.lap(4);
            // ---- E3: blocks 2 bp, 2 bp + 1 = samples 4 bp .. 4 bp + 3 of the lane's 16 outputs -> one 16-byte unit per channel
            if (n > 0) {
                mbar_wait(&head_done, (n - 1) & 1);
#ifdef VP_RACECHECK_BARRIERS
                named_bar_sync(4, 480);
#endif
            }
            prof.lap(1);
#pragma unroll 1
            for (int bq = 0; bq < 2; ++bq) {
                const int bp = 2 * h + bq;
                mbar_wait(&d3_full[bp], n & 1);
                prof.lap(2);
                tc_fence_after();
                uint32_t ra[16], rb[16];
                tmem_ld16_nowait(tl + D2_COL_D23 + 32 * bp, ra);       // samples 4 bp, 4 bp + 1: [phase][8]
                tmem_ld16_nowait(tl + D2_COL_D23 + 32 * bp + 16, rb);  // samples 4 bp + 2, 4 bp + 3
                tmem_ld_wait();
                float *d = hbuf + 4 * d2_unit(4 * r + bp);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 o;
                    o.x = valid ? fmaxf(__uint_as_float(ra[c]) + b3r[c], 0.f) : 0.f;
                    o.y = valid ? fmaxf(__uint_as_float(ra[8 + c]) + b3r[c], 0.f) : 0.f;
                    o.z = valid ? fmaxf(__uint_as_float(rb[c]) + b3r[c], 0.f) : 0.f;
                    o.w = valid ? fmaxf(__uint_as_float(rb[8 + c]) + b3r[c], 0.f) : 0.f;
                    *reinterpret_cast<float4 *>(d + (size_t)c * D2_RP) = o;
                }
                prof.lap(5);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&head_go);
                mbar_arrive(&d23_free);
            }
#ifdef VP_RACECHECK_BARRIERS  // debug build for compute-sanitizer racecheck (it models bar.sync, not mbarriers): same program points
            named_bar_sync(3, 480);
#endif
        }
        prof.flush(warp, lane);
    } else {
        // ================= head warps: sigmoid(conv k11) of item n while the pipeline runs item n + 1
        const int e = (warp - 17) * 32 + lane;  // 224 head threads: outputs 8 half .. 8 half + 7 of lane 4 + rr
        const int half = e >= D2_USE ? 1 : 0, rr = e - half * D2_USE;
        const float *hbuf = reinterpret_cast<const float *>(d2_smem + p.head_off);
        D2Prof prof;  // 0 head_go, 1 head
        prof.start();
        for (int n = 0; n < n_my; ++n) {
            const int it = blockIdx.x + n * gridDim.x;
            const int b = it / p.tiles_per_seq, j = it - b * p.tiles_per_seq;
            const int R0 = p.row_off0 + D2_USE * j - D2_LANE_LO;
            mbar_wait(&head_go, n & 1);
#ifdef VP_RACECHECK_BARRIERS
            named_bar_sync(3, 480);
#endif
            prof.lap(0);
            if (!(p.dbg & 4)) d2_head(p, g, b, R0, rr, half, hbuf);
            __syncwarp();
            if (lane == 0) mbar_arrive(&head_done);
#ifdef VP_RACECHECK_BARRIERS
            if (n + 1 < n_my) named_bar_sync(4, 480);  // pairs with the next item's head_done wait of epilogue B
#endif
            prof.lap(1);
        }
        prof.flush(warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ host: plan
static inline int d2_floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

static void d2_to16(float w, int split, uint16_t &hi, uint16_t &lo) {
    if (split == 2) {
        const __half h = __float2half_rn(w);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(w - __half2float(h)));
    } else {
        hi = __bfloat16_as_ushort(__float2bfloat16_rn(w));
        lo = 0;
    }
}

int decb2_build(DecB2Plan &plan, const TcLayer *dec, const float *const *w4, const float *const *b4, int split, const float (*head_w)[88],
                const float *head_b) {
    FzDecB2 &p = plan.p;
    std::memset(&p, 0, sizeof(p));
    plan.split = split;
    plan.ready = false;
    const int G = 3;
    const TcLayer &L0 = dec[3], &L1 = dec[4], &L2 = dec[5], &L3 = dec[6];
    VP_REQUIRE(L0.ph == 2 && L0.cin == 32 && L0.nout == 64 && L0.sched_taps == 5 && L0.sched_nq == 2 && L0.row0 == -2 && L0.groups == G,
               VP_ERR_UNSUPPORTED, "decb2: decoder.convs.3 differs from the compiled schedule");
    VP_REQUIRE(L1.cin == 32 && L1.cout == 16 && L1.k == 7, VP_ERR_UNSUPPORTED, "decb2: decoder.convs.4 must be 32 -> 16, k = 7");
    VP_REQUIRE(L2.ph == 2 && L2.cin == 16 && L2.nout == 32 && L2.sched_taps == 5 && L2.sched_nq == 1 && L2.row0 == -2 && L2.groups == G,
               VP_ERR_UNSUPPORTED, "decb2: decoder.convs.5 differs from the compiled schedule");
    VP_REQUIRE(L3.ph == 2 && L3.cin == 16 && L3.nout == 16 && L3.sched_taps == 7 && L3.sched_nq == 1 && L3.row0 == -3 && L3.groups == G,
               VP_ERR_UNSUPPORTED, "decb2: decoder.convs.6 differs from the compiled schedule");
    VP_REQUIRE(!L0.blocks.empty() && !L2.blocks.empty() && !L3.blocks.empty(), VP_ERR_ARG, "decb2: host weight blocks are gone");
    p.T0 = 375;
    p.L_out = 6000;
    auto up128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    // shared memory: input box | S1 | head buffer | shuffle exchange | weight blob
    size_t off = 0;
    p.in_off = (int)off;
    off += up128((size_t)D2_IN_ROWS * 4 * 16 * split);
    p.s1_off = (int)off;
    off += up128((size_t)D2_S1_ROWS * 8 * 16 * split);
    p.head_off = (int)off;
    off += (size_t)8 * D2_RP * sizeof(float);
    p.xch_off = (int)off;
    off += 2048 + 3072;
    p.blob_off = (int)off;
    // blob per group
    const size_t blk0 = (size_t)split * 2 * 64 * 8, blk2 = (size_t)split * 2 * 32 * 8, blk3 = (size_t)split * 2 * 16 * 8;  // 16-bit elements
    const size_t blk1 = (size_t)split * 2 * 64 * 8;
    size_t rel = 0;
    const size_t r0 = rel;
    rel += up128(10 * blk0 * 2);
    const size_t r1 = rel;
    rel += up128(12 * blk1 * 2);
    const size_t r2 = rel;
    rel += up128(5 * blk2 * 2);
    const size_t r3 = rel;
    rel += up128(7 * blk3 * 2);
    p.blob_bytes = (int)rel;
    p.w0_off = p.blob_off + (int)r0;
    p.w1_off = p.blob_off + (int)r1;
    p.w2_off = p.blob_off + (int)r2;
    p.w3_off = p.blob_off + (int)r3;
    p.smem_bytes = p.blob_off + p.blob_bytes;
    VP_REQUIRE(p.smem_bytes <= 227 * 1024 - 256, VP_ERR_UNSUPPORTED, "decb2: %d bytes of shared memory", p.smem_bytes);
    VP_REQUIRE(L0.n_blocks == 10 && L2.n_blocks == 5 && L3.n_blocks == 7, VP_ERR_UNSUPPORTED, "decb2: weight block counts");
    plan.blob.assign((size_t)G * rel / 2, 0);
    for (int g = 0; g < G; ++g) {
        uint16_t *dst = plan.blob.data() + (size_t)g * rel / 2;
        if (split == 2) {  // [block][split][k-half][64][8] -> [block][k-half][W_hi rows 0..63 | W_lo rows 64..127][8]
            const uint16_t *src = L0.blocks.data() + (size_t)g * 10 * blk0;
            for (int bk = 0; bk < 10; ++bk)
                for (int sp = 0; sp < 2; ++sp)
                    for (int kh = 0; kh < 2; ++kh)
                        std::memcpy(dst + r0 / 2 + (size_t)bk * blk0 + ((size_t)kh * 128 + (size_t)sp * 64) * 8,
                                    src + (size_t)bk * blk0 + ((size_t)sp * 2 + kh) * 64 * 8, (size_t)64 * 8 * sizeof(uint16_t));
        } else {
            std::memcpy(dst + r0 / 2, L0.blocks.data() + (size_t)g * 10 * blk0, 10 * blk0 * 2);
        }
        std::memcpy(dst + r2 / 2, L2.blocks.data() + (size_t)g * 5 * blk2, 5 * blk2 * 2);
        std::memcpy(dst + r3 / 2, L3.blocks.data() + (size_t)g * 7 * blk3, 7 * blk3 * 2);
        // decoder.convs.4 (16, 32, 7) after x2 up-sampling, folded over the two 750-level samples of a 375-level row:
        //   out1500[4 R + s'] = sum_kk W[kk] xup[4 R + s' + kk - 3],  xup[u] = x750[u div 2]  ->  x750[2 R + d], d in [-2, 3]
        //   row tap jr = d div 2 + 1 (rows R - 1, R, R + 1), K index = (d mod 2) * 32 + ci, column n = s' * 16 + co.
        // Block = jr * 4 + K step; f16x3: [k-half][W_hi rows 0..63 | W_lo rows 64..127][8] (stacked along N), bf16: [k-half][64][8].
        std::vector<float> wf((size_t)3 * 64 * 64, 0.f);  // [jr][kidx][n]
        const float *W = w4[g];
        for (int sp = 0; sp < 4; ++sp)
            for (int kk = 0; kk < 7; ++kk) {
                const int u = sp + kk - 3;              // up-sampled position relative to 4 R
                const int d = d2_floordiv2(u);          // 750-level position relative to 2 R
                const int jr = d2_floordiv2(d) + 1, ph = d - 2 * d2_floordiv2(d);
                VP_REQUIRE(jr >= 0 && jr < 3, VP_ERR_UNSUPPORTED, "decb2: row tap %d", jr);
                for (int co = 0; co < 16; ++co)
                    for (int ci = 0; ci < 32; ++ci)
                        wf[((size_t)jr * 64 + ph * 32 + ci) * 64 + sp * 16 + co] += W[((size_t)co * 32 + ci) * 7 + kk];
            }
        uint16_t *w1 = dst + r1 / 2;
        for (int jr = 0; jr < 3; ++jr)
            for (int kidx = 0; kidx < 64; ++kidx)
                for (int n = 0; n < 64; ++n) {
                    const int blk = jr * 4 + kidx / 16, kh = (kidx % 16) / 8, e = kidx % 8;
                    uint16_t hi, lo;
                    d2_to16(wf[((size_t)jr * 64 + kidx) * 64 + n], split, hi, lo);
                    if (split == 2) {
                        w1[(size_t)blk * blk1 + ((size_t)kh * 128 + n) * 8 + e] = hi;
                        w1[(size_t)blk * blk1 + ((size_t)kh * 128 + 64 + n) * 8 + e] = lo;
                    } else {
                        w1[(size_t)blk * blk1 + ((size_t)kh * 64 + n) * 8 + e] = hi;
                    }
                }
        for (int n = 0; n < 64; ++n) p.bias_c[g][D2_B0 + n] = L0.bias[(size_t)g * 64 + n];
        for (int n = 0; n < 16; ++n) p.bias_c[g][D2_B1 + n] = b4[g] ? b4[g][n] : 0.f;
        for (int n = 0; n < 32; ++n) p.bias_c[g][D2_B2 + n] = L2.bias[(size_t)g * 32 + n];
        for (int n = 0; n < 16; ++n) p.bias_c[g][D2_B3 + n] = L3.bias[(size_t)g * 16 + n];
        for (int c = 0; c < 8; ++c)
            for (int k = 0; k < 12; ++k) p.head_c[g][c * 12 + k] = k < 11 ? head_w[g][c * 11 + k] : 0.f;
        p.head_c[g][96] = head_b[g];
    }
    return VP_OK;
}

int decb2_upload(DecB2Plan &plan) {
    VP_CUDA_CHECK(cudaMalloc(&plan.d_blob, plan.blob.size() * sizeof(uint16_t)));
    VP_CUDA_CHECK(cudaMemcpy(plan.d_blob, plan.blob.data(), plan.blob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    plan.blob.clear();
    plan.blob.shrink_to_fit();
    plan.ready = true;
    return VP_OK;
}

void decb2_free(DecB2Plan &plan) {
    if (plan.d_blob) cudaFree(plan.d_blob);
    plan.d_blob = nullptr;
    plan.ready = false;
}

template <int SPLIT>
static int decb2_launch_t(const FzDecB2 &p, dim3 grid, cudaStream_t s) {
    auto kern = decb2_kernel<SPLIT>;
    FzDecB2K K;
    K.p = p;
    {
        VP_REQUIRE(p.x_gs == (long long)p.B * p.T0 * 32 && (p.x_split * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(p.x) % 16 == 0, VP_ERR_ARG,
                   "decb2: the input groups must be contiguous ([split][group][B][375][32]) and 16-byte aligned");
        const uint64_t dims[5] = {8, (uint64_t)p.T0, (uint64_t)3 * p.B, 4, (uint64_t)SPLIT};
        const uint64_t strides[4] = {64, (uint64_t)p.T0 * 64, 16, (uint64_t)p.x_split * 2};
        const uint32_t box[5] = {8, (uint32_t)D2_IN_ROWS, 1, 4, (uint32_t)SPLIT};
        if (int rc = tma_encode_u16(&K.x_map, p.x, 5, dims, strides, box)) return rc;
    }
    if (int rc = ensure_dyn_smem((const void *)kern, (size_t)p.smem_bytes)) return rc;
    KTimer kt(KC_DECB, s);
    kern<<<grid, D2_THREADS, p.smem_bytes, s>>>(K);
    VP_LAUNCH_CHECK();
#ifdef VP_D2_PROF
    {
        cudaStreamSynchronize(s);
        long long h[25 * 8];
        cudaMemcpyFromSymbol(h, d2_prof_out, sizeof(h));
        const int n_items = p.B * p.tiles_per_seq, n_my = (n_items - 1) / (int)grid.x + 1;
        auto role = [](int w) { return w == 0 ? "issuer" : w < 9 ? "epiA" : w < 17 ? "epiB" : "head"; };
        fprintf(stderr, "[decb2 prof] split %d, %d items per CTA; cycles per item by section\n", SPLIT, n_my);
        for (int w = 0; w < 25; w += (w == 0 ? 1 : 4)) {
            fprintf(stderr, "  warp %2d %-6s", w, role(w));
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %8.0f", (double)h[w * 8 + i] / n_my);
            fprintf(stderr, "\n");
        }
    }
#endif
    return VP_OK;
}

int decb2_launch(const DecB2Plan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, float *y, int keep_lo, int keep_hi,
                 cudaStream_t s) {
    VP_REQUIRE(plan.ready, VP_ERR_UNSUPPORTED, "decb2: plan not uploaded");
    FzDecB2 p = plan.p;
    keep_lo = std::max(0, std::min(keep_lo, p.L_out));
    keep_hi = std::max(keep_lo, std::min(keep_hi, p.L_out));
    if (keep_hi == keep_lo) return VP_OK;
    p.row_off0 = keep_lo / 16;
    p.row_hi = (keep_hi + 15) / 16;
    p.tiles_per_seq = (p.row_hi - p.row_off0 + D2_USE - 1) / D2_USE;
    p.x = x;
    p.x_split = x_split;
    p.x_gs = x_gs;
    p.B = B;
    p.blob = plan.d_blob;
    p.y = y;
    {
        const char *e = getenv("VP_DECB_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    const int n_items = B * p.tiles_per_seq;
    if (n_items == 0) return VP_OK;
    dim3 grid((unsigned)std::min(std::max(device_sm_count() / 3, 1), n_items), 3);
    return plan.split == 2 ? decb2_launch_t<2>(p, grid, s) : decb2_launch_t<1>(p, grid, s);
}

}  // namespace vp
