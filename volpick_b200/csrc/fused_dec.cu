// Fused decoder tail on tcgen05 (sm_100a): decoder.convs.3 -> 4 -> 5 -> 6 -> sigmoid(conv k11) in ONE kernel.
//
// Replaces, for the three EQTransformer decoders (decoder_d, pick_decoders.0/1; SURVEY.md Appendix A,
// seisbench/models/eqtransformer.py Decoder + conv_d / pick_convs), four
// nn.Upsample(x2) + F.conv1d + ReLU stages and the sigmoid head that the layer-by-layer path ran as five
// kernels with every activation round-tripping through HBM.  Here one work item = (window, decoder, time
// tile of m rows of the 375-sample level = 16 m output samples); its activations live in shared memory:
//
//   in[2] (375-level, 32 ch)  --dec3-->  Y (750, 32 ch)  --dec4-->  X (1500, 16 ch)  --dec5-->  Y (3000, 16 ch)
//         --dec6-->  X (6000, 8 ch, fp32 planar)  --head (CUDA cores, fp32)-->  y (B, 3, 6000) in HBM
//
// Tiles overlap by the receptive-field halo (5-6 rows per level), recomputed per tile; rows outside the
// sequence are written as zeros (they are the convs' zero padding).  Every layer is the polyphase implicit
// GEMM of tcconv.cu: a tap is a row offset of the SAME shared-memory activation buffer (no-swizzle K-major
// UMMA descriptors), accumulators in TMEM, f16x3 (fp16 hi/lo split, fp32-equivalent) or bf16 operands.  The
// kernel is bound by the shared-memory reads of the MMA operands (4 KB of A per MMA for N = 16 .. 64), so in
// f16x3 mode the layers with N <= 32 stack W_hi and W_lo along N (two A-tile reads per K step instead of three)
// and the epilogue adds the two column halves, the bias and the ReLU before the 16-bit split + store.
//
// Ordering between the generic-proxy actors (layer-epilogue warps, head warps) goes through mbarriers (acc_full / done_bar carry
// the tensor core's completions, head_go / head_done the hand-over of buffer X).  compute-sanitizer's racecheck only models
// bar.sync, so -DVP_RACECHECK_BARRIERS adds a named barrier at each of those program points (same positions, same participants:
// the 4 epilogue warps of a pipeline per step; epilogue + head warps at the two hand-overs): that build is racecheck-clean,
// which shows that synchronisation AT THESE POINTS orders every shared-memory access pair (profiles/r02_sanitizer.md).
//
// A CTA runs FZ_NPIPE independent pipelines over alternating work items (the layer chain of one item is a
// dependency chain: while one pipeline's epilogue converts a tile, the other pipeline's MMAs own the tensor
// pipe).  Per pipeline: one cp.async loader warp (next item's input rows, double buffered), one tcgen05
// issuer warp walking a host-built step list (layer, tile) with per-step producer dependencies, four epilogue
// warps (TMEM -> ReLU -> fp16 hi/lo planes of the next layer's A operand) that also run the head.  The
// decoder's weights are resident in shared memory and shared by the pipelines.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fused.cuh"
#include "tc_ptx.cuh"
#include "tma.cuh"

namespace vp {

// ------------------------------------------------------------------------------------------ epilogues
template <int SPLIT>
__device__ __forceinline__ void fz_pack8(const float *v, bool valid, uint4 &hi, uint4 &lo) {
    if (valid) {
        pack8_split16<SPLIT>(v, hi, lo);
    } else {
        hi = make_uint4(0u, 0u, 0u, 0u);
        lo = hi;
    }
}

// CH accumulator columns starting at col0 (+ the stacked half at col0 + NST when NST > 0) -> relu(acc + bias).
// The biases are read from shared memory as broadcast 16-byte loads (6 % of the kernel's shared-memory wavefronts).  Reading
// them through the constant bank instead (kernel parameters, one LDC with a register offset per value) was measured
// SLOWER: 10.38 vs 10.03 ms per station-day.
template <int CH, int NST>
__device__ __forceinline__ void fz_load_cols(const float *bias /*shared memory, column col0*/, uint32_t tacc, int col0, float (&v)[CH]) {
    uint32_t r[CH], r2[NST > 0 ? CH : 1];
    if constexpr (CH == 16) {
        tmem_ld16_nowait(tacc + (uint32_t)col0, r);
        if constexpr (NST > 0) tmem_ld16_nowait(tacc + (uint32_t)(col0 + NST), r2);
    } else {
        static_assert(CH == 8, "column slice");
        tmem_ld8_nowait(tacc + (uint32_t)col0, r);
        if constexpr (NST > 0) tmem_ld8_nowait(tacc + (uint32_t)(col0 + NST), r2);
    }
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < CH; n += 4) {
        const float4 b4 = *reinterpret_cast<const float4 *>(bias + n);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = __uint_as_float(r[n + e]);
            if constexpr (NST > 0) a += __uint_as_float(r2[n + e]);
            v[n + e] = fmaxf(a + bb[e], 0.f);
        }
    }
}

// Polyphase layer -> 16-bit planes of the next layer.  Thread = one input row s; it owns output rows 2s, 2s+1.
// The two rows are 32 bytes apart: a plain "phase 0, then phase 1" store has consecutive lanes 32 bytes apart
// (2-way bank conflict).  Instead lanes with bit 2 set store phase 1 first: every quarter warp then covers eight
// distinct 16-byte bank groups.  16 channels (two planes) of both phases are in registers at a time.
template <int COUTP, int SPLIT, bool STACK>
__device__ __forceinline__ void fz_epi16(const FzLayer &L, const float *bias, uint32_t tacc, int t, int r, int R0, uint8_t *arena) {
    constexpr int P = COUTP / 8;
    constexpr int NST = STACK ? 2 * COUTP : 0;
    const int s_rel = L.s_lo + 128 * t + r;
    const int lrow0 = 2 * s_rel - L.out_lo;
    const int grow0 = 2 * ((R0 << L.lvl) + s_rel);  // R0: 375-level row of the work item's origin
    const uint32_t plane = (uint32_t)L.out_rows * 16u;
    const int sel = (r >> 2) & 1;  // phase stored first by this lane
    const int rowA = lrow0 + sel, rowB = lrow0 + 1 - sel;
    const bool inA = (unsigned)rowA < (unsigned)L.out_rows, inB = (unsigned)rowB < (unsigned)L.out_rows;
    const bool valid0 = (unsigned)grow0 < (unsigned)L.T_out, valid1 = (unsigned)(grow0 + 1) < (unsigned)L.T_out;
    uint8_t *dstA = arena + L.out_off + (size_t)rowA * 16, *dstB = arena + L.out_off + (size_t)rowB * 16;
#pragma unroll
    for (int c0 = 0; c0 < COUTP; c0 += 16) {
        float v0[16], v1[16];
        fz_load_cols<16, NST>(bias + c0, tacc, c0, v0);
        fz_load_cols<16, NST>(bias + COUTP + c0, tacc, COUTP + c0, v1);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int pl = c0 / 8 + k;
            uint4 h0, l0, h1, l1;
            fz_pack8<SPLIT>(&v0[k * 8], valid0, h0, l0);
            fz_pack8<SPLIT>(&v1[k * 8], valid1, h1, l1);
            const uint4 hA = sel ? h1 : h0, hB = sel ? h0 : h1;
            if (inA) *reinterpret_cast<uint4 *>(dstA + pl * plane) = hA;
            if (inB) *reinterpret_cast<uint4 *>(dstB + pl * plane) = hB;
            if (SPLIT == 2) {
                const uint4 lA = sel ? l1 : l0, lB = sel ? l0 : l1;
                if (inA) *reinterpret_cast<uint4 *>(dstA + (P + pl) * plane) = lA;
                if (inB) *reinterpret_cast<uint4 *>(dstB + (P + pl) * plane) = lB;
            }
        }
    }
}

// fp32 planar head input [c][row]: the head reads it as 16-byte units (4 rows).  With 8 outputs per head thread the lane
// stride is two units, so a quarter warp would hit only four of the eight 16-byte bank groups (2-way conflict on every
// load); flipping the low unit bit in every other group of eight units makes the eight accesses distinct.  12 outputs per
// thread (lane stride three units) are conflict-free as they are.
__device__ __forceinline__ int fz_head_unit(int u, int swz) { return swz ? (u ^ ((u >> 3) & 1)) : u; }

// Last polyphase layer (8 channels per phase) -> fp32 planar [c][row] for the CUDA-core head.
template <bool STACK>
__device__ __forceinline__ void fz_epi32(const FzLayer &L, const float *bias, uint32_t tacc, int t, int r, int R0, uint8_t *arena, int swz) {
    const int s_rel = L.s_lo + 128 * t + r;
    const int lrow0 = 2 * s_rel - L.out_lo;  // even
    const int grow0 = 2 * ((R0 << L.lvl) + s_rel);
    float v[16];
    fz_load_cols<16, STACK ? 16 : 0>(bias, tacc, 0, v);
    if ((unsigned)lrow0 < (unsigned)L.out_rows) {  // out_rows is even: both phases are in range together
        const bool valid = (unsigned)grow0 < (unsigned)L.T_out;  // T_out even: both phases valid together
        float *d = reinterpret_cast<float *>(arena + L.out_off) + 4 * fz_head_unit(lrow0 >> 2, swz) + (lrow0 & 3);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float2 o;
            o.x = valid ? v[c] : 0.f;
            o.y = valid ? v[8 + c] : 0.f;
            *reinterpret_cast<float2 *>(d + (size_t)c * L.out_rp) = o;
        }
    }
}

// sigmoid(conv k11, 8 -> 1) on the CUDA cores: thread e = OPT consecutive output samples (OPT a multiple of 4).
template <int OPT>
__device__ __forceinline__ void fz_head(const FzDecB &p, const float *hw /*shared: [8][12] weights, [96] bias*/, int g, int b, int R0, int e,
                                        const uint8_t *arena) {
    static_assert(OPT % 4 == 0, "the head reads 16-byte units");
    const int tl0 = OPT * e;
    if (tl0 >= p.W) return;
    const float *d6 = reinterpret_cast<const float *>(arena + p.head_in_off);
    const int RP = p.head_rp, swz = p.head_swz;
    float acc[OPT];
#pragma unroll
    for (int o = 0; o < OPT; ++o) acc[o] = hw[96];
    // buffer rows tl0 .. tl0 + OPT + 11 (row 0 = output - 6): NU units of 4 rows; the loads of channel c + 1 are issued before
    // the FMAs of channel c (with two warps per scheduler nothing else hides the shared-memory latency)
    constexpr int NU = (OPT + 12) / 4;
    int uo[NU];
#pragma unroll
    for (int q = 0; q < NU; ++q) uo[q] = 4 * fz_head_unit((tl0 >> 2) + q, swz);
    float4 cur[NU], nxt[NU];
#pragma unroll
    for (int q = 0; q < NU; ++q) cur[q] = *reinterpret_cast<const float4 *>(d6 + uo[q]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c + 1 < 8) {
#pragma unroll
            for (int q = 0; q < NU; ++q) nxt[q] = *reinterpret_cast<const float4 *>(d6 + (size_t)(c + 1) * RP + uo[q]);
        }
        float xv[4 * NU];
#pragma unroll
        for (int q = 0; q < NU; ++q) {
            xv[4 * q] = cur[q].x;
            xv[4 * q + 1] = cur[q].y;
            xv[4 * q + 2] = cur[q].z;
            xv[4 * q + 3] = cur[q].w;
        }
        // weights of this channel: three broadcast 16-byte reads (a kernel-parameter array indexed by the run-time
        // group costs one constant load with a register offset per FMA)
        const float4 w0 = *reinterpret_cast<const float4 *>(hw + c * 12), w1 = *reinterpret_cast<const float4 *>(hw + c * 12 + 4),
                     w2 = *reinterpret_cast<const float4 *>(hw + c * 12 + 8);
        const float wk[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = wk[k];
#pragma unroll
            for (int o = 0; o < OPT; ++o) acc[o] = fmaf(w, xv[o + k + 1], acc[o]);  // buffer row 0 = output - 6
        }
#pragma unroll
        for (int q = 0; q < NU; ++q) cur[q] = nxt[q];
    }
    const int t_out = 16 * R0 + tl0;
    float *yb = p.y + ((size_t)b * 3 + g) * p.L_out + t_out;
#pragma unroll
    for (int o = 0; o < OPT; o += 4) {
        if (tl0 + o < p.W && t_out + o < p.L_out) {  // W, L_out and t_out are multiples of 4 (16-byte store)
            float4 o4;
            o4.x = 1.f / (1.f + expf(-acc[o]));
            o4.y = 1.f / (1.f + expf(-acc[o + 1]));
            o4.z = 1.f / (1.f + expf(-acc[o + 2]));
            o4.w = 1.f / (1.f + expf(-acc[o + 3]));
            *reinterpret_cast<float4 *>(yb + o) = o4;
        }
    }
}

// One layer of one work item on the issuer warp: per M tile wait for the accumulator buffer and the producer
// tiles, then issue the layer's compile-time MMA schedule (the first MMA overwrites the accumulator).
template <int LYR, int SPLIT>
__device__ __forceinline__ void fz_issue_layer(const FzDecB &p, int pp, int slot, int n, uint32_t &i, uint32_t tmem_base,
                                               uint32_t arena16, uint32_t sB16, uint64_t (*in_full)[FZ_NSLOT], uint64_t (*in_empty)[FZ_NSLOT],
                                               uint64_t (*acc_full)[FZ_NBUF], uint64_t (*done_bar)[FZ_NBUF]) {
    constexpr int l = LYR;
    constexpr int NOUT = FZ_DEC_NOUT[l];
    constexpr bool STACK = SPLIT == 2 && FZ_DEC_STACK[l];
    const FzLayer &L = p.L[l];
    const uint32_t idesc = umma_idesc(NOUT, SPLIT == 2 ? 0 : 1);
    const uint32_t in16 = arena16 + ((uint32_t)(L.in_off + (l == 0 ? slot * p.in_slot_bytes : 0)) >> 4) + (uint32_t)L.a_row0;
    const uint32_t w16 = sB16 + ((uint32_t)L.w_off >> 4);
    const uint32_t in_rows = (uint32_t)L.in_pitch;
    for (int t = 0; t < L.n_tiles; ++t, ++i) {
        const uint32_t buf = i & (FZ_NBUF - 1);
        mbar_wait(&done_bar[pp][buf], ((i / FZ_NBUF) & 1) ^ 1);  // accumulator free: step i - NBUF retired
#pragma unroll
        for (int dd = 0; dd < 2; ++dd) {
            const uint32_t rel = L.dep[t][dd];
            if (rel != 0 && rel < FZ_NBUF) {
                const uint32_t d = i - rel;
                mbar_wait(&done_bar[pp][d & (FZ_NBUF - 1)], (d / FZ_NBUF) & 1);
            }
        }
        if (l == 0 && t == 0) mbar_wait(&in_full[pp][slot], (n / FZ_NSLOT) & 1);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t a16 = in16 + (uint32_t)t * 128u;
        const uint32_t d_tmem = tmem_base + (pp * FZ_NBUF + buf) * FZ_NCOLS;
        if (elect_one()) {
            if (p.dbg & 1) {
            } else if constexpr (STACK)
                umma_conv_tile_stacked<NOUT, FZ_DEC_NTAPS[l], FZ_DEC_NQ[l]>(d_tmem, a16, in_rows, w16, umma_idesc(2 * NOUT, 0), idesc, 0u);
            else
                umma_conv_tile<NOUT, SPLIT, FZ_DEC_NTAPS[l], FZ_DEC_NQ[l]>(d_tmem, a16, in_rows, w16, idesc, 0u);
            umma_commit(&acc_full[pp][buf]);
            if (l == 0 && t == L.n_tiles - 1) umma_commit(&in_empty[pp][slot]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ the kernel
struct FzDecBK {
    alignas(64) CUtensorMap x_map;  // (8 channels, t, group * B + window, 8-channel plane, split) over the 375-sample level input
    FzDecB p;
};

template <int SPLIT, int OPT>
__global__ void __launch_bounds__(FZ_THREADS, 1) decb_kernel(const __grid_constant__ FzDecBK K) {
    const FzDecB &p = K.p;
    extern __shared__ __align__(128) uint8_t fz_smem[];
    __shared__ __align__(8) uint64_t in_full[FZ_NPIPE][FZ_NSLOT], in_empty[FZ_NPIPE][FZ_NSLOT], acc_full[FZ_NPIPE][FZ_NBUF],
        done_bar[FZ_NPIPE][FZ_NBUF], head_go[FZ_NPIPE], head_done[FZ_NPIPE];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const uint32_t sbase = smem_u32(fz_smem);
    const int n_items = p.B * p.tiles_per_seq;

    if (tid == 0) {
        for (int pp = 0; pp < FZ_NPIPE; ++pp) {
            for (int i = 0; i < FZ_NSLOT; ++i) {
                mbar_init(&in_full[pp][i], 1);
                mbar_init(&in_empty[pp][i], 1);
            }
            for (int i = 0; i < FZ_NBUF; ++i) {
                mbar_init(&acc_full[pp][i], 1);
                mbar_init(&done_bar[pp][i], 4);
            }
            mbar_init(&head_go[pp], 4);
            mbar_init(&head_done[pp], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, FZ_NPIPE * FZ_NBUF * FZ_NCOLS);
    {   // resident weights of this decoder
        const uint4 *wg = reinterpret_cast<const uint4 *>(p.blob + (long long)g * (p.blob_bytes / 2));
        for (int idx = tid; idx < p.blob_bytes / 16; idx += FZ_THREADS) cp_async16(sbase + p.blob_off + idx * 16, wg + idx, 16u);
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < FZ_NPIPE) {
        // ================= loaders: the input rows of the pipeline's next work item as ONE TMA box
        // [8 channels][in_rows][1 sequence][cin / 8 planes][split] -> [split][plane][row][16 B]; rows outside the sequence lie outside
        // the tensor map and arrive as zeros (the conv's padding).  Round 1 copied them as 16-byte cp.async pieces.
        const int pp = warp;
        if (lane == 0 && !(p.dbg & 8)) {
            const uint32_t bytes = (uint32_t)(p.L[0].in_rows * p.L[0].cin8 * 16 * SPLIT);
            int n = 0;
            for (int item = blockIdx.x + pp * gridDim.x; item < n_items; item += FZ_NPIPE * gridDim.x, ++n) {
                const int slot = n % FZ_NSLOT;
                mbar_wait(&in_empty[pp][slot], ((n / FZ_NSLOT) & 1) ^ 1);
                const int b = item / p.tiles_per_seq, j = item - b * p.tiles_per_seq;
                const int row_base = p.c0 * j + p.row_off0 + p.in_lo0;
                const uint32_t dst0 = sbase + pp * p.pipe_stride + p.L[0].in_off + slot * p.in_slot_bytes;
                mbar_arrive_expect_tx(&in_full[pp][slot], bytes);
                tma_load_5d(dst0, &K.x_map, &in_full[pp][slot], 0, row_base, g * p.B + b, 0, 0);
            }
        }
    } else if (warp < 2 * FZ_NPIPE) {
        // ================= tcgen05 issuers =================
        const int pp = warp - FZ_NPIPE;
        const uint32_t sB16 = sbase >> 4;
        const uint32_t arena16 = (sbase + pp * p.pipe_stride) >> 4;
        uint32_t i = 0;
        int n = 0;
        for (int item = blockIdx.x + pp * gridDim.x; item < n_items; item += FZ_NPIPE * gridDim.x, ++n) {
            const int slot = n % FZ_NSLOT;
            fz_issue_layer<0, SPLIT>(p, pp, slot, n, i, tmem_base, arena16, sB16, in_full, in_empty, acc_full, done_bar);
            fz_issue_layer<1, SPLIT>(p, pp, slot, n, i, tmem_base, arena16, sB16, in_full, in_empty, acc_full, done_bar);
            fz_issue_layer<2, SPLIT>(p, pp, slot, n, i, tmem_base, arena16, sB16, in_full, in_empty, acc_full, done_bar);
            fz_issue_layer<3, SPLIT>(p, pp, slot, n, i, tmem_base, arena16, sB16, in_full, in_empty, acc_full, done_bar);
        }
    } else if (warp < 6 * FZ_NPIPE) {
        // ================= layer epilogue warps =================
        const int pp = (warp - 2 * FZ_NPIPE) >> 2, q = warp & 3;
        const int r = q * 32 + lane;
        uint8_t *arena = fz_smem + pp * p.pipe_stride;
        uint32_t i = 0;
        int n = 0;
        for (int item = blockIdx.x + pp * gridDim.x; item < n_items; item += FZ_NPIPE * gridDim.x, ++n) {
            const int b = item / p.tiles_per_seq, j = item - b * p.tiles_per_seq;
            const int R0 = p.c0 * j + p.row_off0;
            (void)b;
            for (int l = 0; l < p.n_layers; ++l) {
                const FzLayer &L = p.L[l];
                // layer 1 is the first writer of buffer X, which the head warps may still be reading (previous item)
                if (l == p.head_wait_layer && n > 0) {
                    mbar_wait(&head_done[pp], (n - 1) & 1);
#ifdef VP_RACECHECK_BARRIERS
                    named_bar_sync(4 + pp, 256);
#endif
                }
                for (int t = 0; t < L.n_tiles; ++t, ++i) {
                    const uint32_t buf = i & (FZ_NBUF - 1);
                    mbar_wait(&acc_full[pp][buf], (i / FZ_NBUF) & 1);
                    tc_fence_after();
#ifdef VP_RACECHECK_BARRIERS  // debug build for compute-sanitizer racecheck (it models bar.sync, not mbarriers): see the kernel comment
                    named_bar_sync(6 + pp, 128);
#endif
                    const uint32_t tacc = tmem_base + (pp * FZ_NBUF + buf) * FZ_NCOLS + ((uint32_t)(q * 32) << 16);
                    constexpr bool ST = SPLIT == 2;  // stacked layers: FZ_DEC_STACK (layers 1-3)
                    const float *bias = reinterpret_cast<const float *>(fz_smem + p.bias_off) + l * FZ_NCOLS;
                    // rows of this warp: outputs [2 (s_lo + 128 t + 32 q) - out_lo, + 64); a warp whose rows all fall
                    // outside the output buffer (partially filled last tile of a layer) has nothing to convert
                    const int wrow0 = 2 * (L.s_lo + 128 * t + 32 * q) - L.out_lo;
                    if (wrow0 < L.out_rows && wrow0 + 64 > 0 && !(p.dbg & 2)) {
                        if (l == 0) fz_epi16<32, SPLIT, ST && FZ_DEC_STACK[0]>(L, bias, tacc, t, r, R0, arena);
                        else if (l == 3) fz_epi32<ST && FZ_DEC_STACK[3]>(L, bias, tacc, t, r, R0, arena, p.head_swz);
                        else fz_epi16<16, SPLIT, ST && FZ_DEC_STACK[1]>(L, bias, tacc, t, r, R0, arena);
                    }
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&done_bar[pp][buf]);
                }
            }
            if (lane == 0) mbar_arrive(&head_go[pp]);  // the last layer's rows of this warp are in shared memory
#ifdef VP_RACECHECK_BARRIERS
            named_bar_sync(2 + pp, 256);
#endif
        }
    } else {
        // ================= head warps: sigmoid(conv k11) of the previous item while the pipeline runs the next one
        const int pp = (warp - 6 * FZ_NPIPE) >> 2;
        const int e = ((warp - 6 * FZ_NPIPE) & 3) * 32 + lane;
        const uint8_t *arena = fz_smem + pp * p.pipe_stride;
        const float *hw = reinterpret_cast<const float *>(fz_smem + p.bias_off) + FZ_MAX_LAYERS * FZ_NCOLS;
        int n = 0;
        for (int item = blockIdx.x + pp * gridDim.x; item < n_items; item += FZ_NPIPE * gridDim.x, ++n) {
            const int b = item / p.tiles_per_seq, j = item - b * p.tiles_per_seq;
            const int R0 = p.c0 * j + p.row_off0;
            mbar_wait(&head_go[pp], n & 1);
#ifdef VP_RACECHECK_BARRIERS
            named_bar_sync(2 + pp, 256);
#endif
            if (!(p.dbg & 4)) fz_head<OPT>(p, hw, g, b, R0, e, arena);
            __syncwarp();
            if (lane == 0) mbar_arrive(&head_done[pp]);
#ifdef VP_RACECHECK_BARRIERS
            if (item + FZ_NPIPE * (int)gridDim.x < n_items) named_bar_sync(4 + pp, 256);  // pairs with the next item's layer-1 wait
#endif
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, FZ_NPIPE * FZ_NBUF * FZ_NCOLS);
}

// ------------------------------------------------------------------------------------------ host: plan
static inline int fdiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }
static inline int cdiv2(int a) { return -fdiv2(-a); }

int decb_build(DecBPlan &plan, const TcLayer *dec, int split, int m, const float (*head_w)[88], const float *head_b) {
    FzDecB &p = plan.p;
    std::memset(&p, 0, sizeof(p));
    plan.split = split;
    plan.ready = false;
    const int NL = 4, G = 3;
    p.n_layers = NL;
    p.T0 = 375;
    p.cin0 = dec[3].cin;
    p.c0 = m;
    p.tiles_per_seq = (p.T0 + m - 1) / m;
    p.W = 16 * m;
    p.L_out = 6000;
    plan.opt = ((p.W + 127) / 128 + 3) & ~3;  // head: 128 threads per pipeline, a multiple of 4 outputs each (16-byte units)
    VP_REQUIRE(plan.opt == 8 || plan.opt == 12, VP_ERR_UNSUPPORTED, "decb: no head instance for %d outputs per thread", plan.opt);
    p.head_swz = plan.opt == 8 ? 1 : 0;
    int lo[5], hi[5], c[5];
    for (int k = 0; k <= NL; ++k) c[k] = m << k;
    lo[NL] = -6;  // the head (k11) needs -5; even so that both phases of a row land on an aligned float2
    hi[NL] = c[NL] + 6;
    int s_lo[4], n_tiles[4];
    for (int l = NL - 1; l >= 0; --l) {
        const TcLayer &TL = dec[3 + l];
        VP_REQUIRE(TL.ph == 2 && TL.cin % 16 == 0 && TL.groups == G, VP_ERR_UNSUPPORTED, "decb: layer %d is not a 3-group polyphase layer", 3 + l);
        VP_REQUIRE(!TL.blocks.empty(), VP_ERR_ARG, "decb: host weight blocks of layer %d are gone", 3 + l);
        const int ntaps = TL.halo + 1, o_min = TL.row0;
        s_lo[l] = fdiv2(lo[l + 1]);
        const int s_hi = cdiv2(hi[l + 1]);
        n_tiles[l] = (s_hi - s_lo[l] + 127) / 128;
        VP_REQUIRE(n_tiles[l] <= FZ_MAX_TILES, VP_ERR_UNSUPPORTED, "decb: %d tiles per item", n_tiles[l]);
        lo[l] = s_lo[l] + o_min;
        hi[l] = s_hi + o_min + ntaps - 1;
        if (l > 0 && (lo[l] & 1)) --lo[l];
    }
    static_assert(FZ_DEC_STACK[1] == FZ_DEC_STACK[2] && !FZ_DEC_STACK[0], "epilogue dispatch in decb_kernel");
    // shared memory: per pipeline { in[2] | X | Y }, then the weight blob
    const int esz = 16 * split;  // bytes per (row, plane)
    // input slot: filled by one TMA box, plane pitch = rows of the box
    const int pitch0 = hi[0] - lo[0];
    auto lvl_bytes = [&](int k) {
        if (k == NL) return (size_t)8 * (size_t)((((hi[k] - lo[k]) + 12 + 3) & ~3) + 4) * 4;  // + one unit: the swizzle swaps unit pairs
        const int ch = (k == 0) ? dec[3].cin : dec[3 + k - 1].cout;
        return (size_t)(k == 0 ? pitch0 : hi[k] - lo[k]) * (ch / 8) * esz;
    };
    auto up128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    p.in_slot_bytes = (int)up128(lvl_bytes(0));
    const size_t off_x = FZ_NSLOT * (size_t)p.in_slot_bytes;
    // Buffer X holds level 2 and then the head's input (level 4): the next item's layer 1 stores its output only after the
    // head of the previous item is done.  (A separate level-2 buffer was measured: 10.61 vs 10.63 ms -- the head is not on
    // the critical path.)
    const size_t off_y = off_x + up128(std::max(lvl_bytes(2), lvl_bytes(4)));
    p.pipe_stride = (int)(off_y + up128(std::max(lvl_bytes(1), lvl_bytes(3))));
    p.blob_off = FZ_NPIPE * p.pipe_stride;
    p.head_wait_layer = 1;
    const size_t lvl_off[5] = {0, off_y, off_x, off_y, off_x};
    // blob per group: the weight blocks of the four layers
    size_t blob = 0;
    size_t w_rel[4];
    for (int l = 0; l < NL; ++l) {
        const TcLayer &TL = dec[3 + l];
        w_rel[l] = blob;
        blob += up128((size_t)TL.n_blocks * split * 2 * TL.nout * 16);
    }
    const size_t bias_rel = blob;  // fp32 [layer][FZ_NCOLS], then the head: weights [8][12] (11 taps + pad), bias at [96]
    blob += up128(((size_t)FZ_MAX_LAYERS * FZ_NCOLS + 128) * sizeof(float));
    p.blob_bytes = (int)blob;
    p.bias_off = p.blob_off + (int)bias_rel;
    plan.blob.assign((size_t)G * blob / 2, 0);
    int step = 0;
    int step0[4];
    for (int l = 0; l < NL; ++l) {
        const TcLayer &TL = dec[3 + l];
        FzLayer &L = p.L[l];
        L.cin8 = TL.cin / 8;
        L.nout = TL.nout;
        L.coutp = (TL.cout + 7) / 8 * 8;
        const bool stack = split == 2 && FZ_DEC_STACK[l];
        VP_REQUIRE(L.nout == 2 * L.coutp && (stack ? 2 : 1) * L.nout <= FZ_NCOLS, VP_ERR_UNSUPPORTED, "decb: layer %d N=%d coutp=%d", 3 + l,
                   L.nout, L.coutp);
        L.n_tiles = n_tiles[l];
        L.in_off = (int)lvl_off[l];
        L.in_rows = hi[l] - lo[l];
        L.in_pitch = (l == 0) ? pitch0 : L.in_rows;
        L.out_off = (int)lvl_off[l + 1];
        L.out_rows = hi[l + 1] - lo[l + 1];
        L.out_rp = (l == NL - 1) ? (((L.out_rows + 12 + 3) & ~3) + 4) : 0;
        L.s_lo = s_lo[l];
        L.lvl = l;
        L.c_in = c[l];
        L.out_lo = lo[l + 1];
        L.T_out = p.T0 << (l + 1);
        L.out_kind = (l == NL - 1) ? 1 : 0;
        VP_REQUIRE(L.out_kind == 0 || (L.coutp == 8 && L.out_rows % 2 == 0), VP_ERR_UNSUPPORTED, "decb: last layer must have 8 channels");
        VP_REQUIRE(L.out_kind == 1 || L.coutp == (l == 0 ? 32 : 16), VP_ERR_UNSUPPORTED, "decb: coutp %d of layer %d", L.coutp, 3 + l);
        L.w_off = p.blob_off + (int)w_rel[l];
        // TcLayer blocks: [group][block][split][k-half][nout][8]; stacked layers want [block][k-half][split * nout + n][8]
        const size_t blk_elems = (size_t)split * 2 * TL.nout * 8;
        for (int g = 0; g < G; ++g) {
            uint16_t *dst = plan.blob.data() + (size_t)g * blob / 2 + w_rel[l] / 2;
            const uint16_t *src = TL.blocks.data() + (size_t)g * TL.n_blocks * blk_elems;
            if (!stack) {
                std::memcpy(dst, src, (size_t)TL.n_blocks * blk_elems * sizeof(uint16_t));
            } else {
                for (int b = 0; b < TL.n_blocks; ++b)
                    for (int sp = 0; sp < 2; ++sp)
                        for (int kh = 0; kh < 2; ++kh)
                            std::memcpy(dst + (size_t)b * blk_elems + ((size_t)kh * 2 * TL.nout + (size_t)sp * TL.nout) * 8,
                                        src + (size_t)b * blk_elems + ((size_t)sp * 2 + kh) * TL.nout * 8, (size_t)TL.nout * 8 * sizeof(uint16_t));
            }
        }
        for (int g = 0; g < G; ++g) {  // fp32 biases of this decoder, read by the epilogue from shared memory
            float *bd = reinterpret_cast<float *>(plan.blob.data() + (size_t)g * blob / 2 + bias_rel / 2) + (size_t)l * FZ_NCOLS;
            for (int n = 0; n < TL.nout; ++n) bd[n] = TL.bias[(size_t)g * TL.nout + n];
        }
        const int a_row0 = (s_lo[l] + TL.row0) - lo[l];
        L.a_row0 = a_row0;
        VP_REQUIRE(TL.nout == FZ_DEC_NOUT[l] && TL.sched_taps == FZ_DEC_NTAPS[l] && TL.sched_nq == FZ_DEC_NQ[l] && a_row0 >= 0,
                   VP_ERR_UNSUPPORTED, "decb: layer %d (N=%d, %d taps, %d pairs) differs from the compiled schedule", 3 + l,
                   TL.nout, TL.sched_taps, TL.sched_nq);
        VP_REQUIRE((int)TL.mma.size() == FZ_DEC_NTAPS[l] * FZ_DEC_NQ[l], VP_ERR_UNSUPPORTED, "decb: layer %d has %d K steps", 3 + l,
                   (int)TL.mma.size());
        for (int j = 0, kk = 0; j < FZ_DEC_NTAPS[l]; ++j)  // the compiled schedule walks (tap, pair) with block j * NQ + q
            for (int q = 0; q < FZ_DEC_NQ[l]; ++q, ++kk) {
                const TcMma &e = TL.mma[kk];
                VP_REQUIRE(e.a_row == j && e.a_plane == 2 * q && e.a_rowk == 0 && e.b_block == kk, VP_ERR_UNSUPPORTED,
                           "decb: compiled MMA schedule of layer %d differs from the host-built one", 3 + l);
            }
        step0[l] = step;
        const int ntaps_l = TL.halo + 1;
        // producer dependencies (previous layer's tiles that write the rows this tile reads)
        for (int t = 0; t < L.n_tiles; ++t) {
            L.dep[t][0] = L.dep[t][1] = 0;
            if (l == 0) continue;
            const int ntaps = ntaps_l;
            const int first = a_row0 + 128 * t;
            const int last = std::min(first + 127 + ntaps - 1, L.in_rows - 1);
            VP_REQUIRE(2 * s_lo[l - 1] - lo[l] == 0 && first >= 0, VP_ERR_UNSUPPORTED, "decb: producer rows are not tile aligned");
            int ta = first / 256, tb = last / 256;  // producer tile t' writes local rows [256 t', 256 t' + 256)
            ta = std::min(ta, n_tiles[l - 1] - 1);
            tb = std::min(tb, n_tiles[l - 1] - 1);
            VP_REQUIRE(tb - ta <= 1, VP_ERR_UNSUPPORTED, "decb: tile depends on more than two producer tiles");
            L.dep[t][0] = (uint8_t)((step + t) - (step0[l - 1] + ta));
            if (tb != ta) L.dep[t][1] = (uint8_t)((step + t) - (step0[l - 1] + tb));
        }
        step += L.n_tiles;
    }
    p.steps_per_item = step;
    p.in_lo0 = lo[0];
    p.head_in_off = (int)off_x;
    p.head_rp = p.L[NL - 1].out_rp;
    VP_REQUIRE(-5 - lo[NL] == 1, VP_ERR_UNSUPPORTED, "decb: head row offset");
    for (int g = 0; g < G; ++g) {
        std::memcpy(p.head_w[g], head_w[g], 88 * sizeof(float));
        p.head_b[g] = head_b[g];
        float *hd = reinterpret_cast<float *>(plan.blob.data() + (size_t)g * blob / 2 + bias_rel / 2) + (size_t)FZ_MAX_LAYERS * FZ_NCOLS;
        for (int c = 0; c < 8; ++c)
            for (int k = 0; k < 11; ++k) hd[c * 12 + k] = head_w[g][c * 11 + k];
        hd[96] = head_b[g];
    }
    p.smem_bytes = p.blob_off + p.blob_bytes;
    VP_REQUIRE(p.smem_bytes <= 226 * 1024, VP_ERR_UNSUPPORTED, "decb: %d bytes of shared memory", p.smem_bytes);
    return VP_OK;
}

int decb_upload(DecBPlan &plan) {
    VP_CUDA_CHECK(cudaMalloc(&plan.d_blob, plan.blob.size() * sizeof(uint16_t)));
    VP_CUDA_CHECK(cudaMemcpy(plan.d_blob, plan.blob.data(), plan.blob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    plan.blob.clear();
    plan.blob.shrink_to_fit();
    plan.ready = true;
    return VP_OK;
}

void decb_free(DecBPlan &plan) {
    if (plan.d_blob) cudaFree(plan.d_blob);
    plan.d_blob = nullptr;
    plan.ready = false;
}

template <int SPLIT, int OPT>
static int decb_launch_t(const FzDecB &p, dim3 grid, cudaStream_t s) {
    auto kern = decb_kernel<SPLIT, OPT>;
    FzDecBK K;
    K.p = p;
    {
        const int c8 = p.L[0].cin8;
        VP_REQUIRE(p.x_gs == (long long)p.B * p.T0 * p.cin0 && (p.x_split * 2) % 16 == 0 && reinterpret_cast<uintptr_t>(p.x) % 16 == 0 &&
                       p.L[0].in_rows <= 256 && p.L[0].in_pitch == p.L[0].in_rows,
                   VP_ERR_ARG, "decb: the input groups must be contiguous ([split][group][B][375][32]) and 16-byte aligned");
        const uint64_t dims[5] = {8, (uint64_t)p.T0, (uint64_t)3 * p.B, (uint64_t)c8, (uint64_t)SPLIT};
        const uint64_t strides[4] = {(uint64_t)p.cin0 * 2, (uint64_t)p.T0 * p.cin0 * 2, 16, (uint64_t)p.x_split * 2};
        const uint32_t box[5] = {8, (uint32_t)p.L[0].in_rows, 1, (uint32_t)c8, (uint32_t)SPLIT};
        if (int rc = tma_encode_u16(&K.x_map, p.x, 5, dims, strides, box)) return rc;
    }
    if (int rc = ensure_dyn_smem((const void *)kern, (size_t)p.smem_bytes)) return rc;
    KTimer kt(KC_DECB, s);
    kern<<<grid, FZ_THREADS, p.smem_bytes, s>>>(K);
    VP_LAUNCH_CHECK();
    return VP_OK;
}

int decb_launch(const DecBPlan &plan, const uint16_t *x, long long x_split, long long x_gs, int B, float *y, int keep_lo,
                int keep_hi, cudaStream_t s) {
    VP_REQUIRE(plan.ready, VP_ERR_UNSUPPORTED, "decb: plan not uploaded");
    FzDecB p = plan.p;
    {   // tiles that produce at least one kept output sample: tile t covers samples [16 (row_off0 + m t), 16 (row_off0 + m (t + 1)))
        keep_lo = std::max(0, std::min(keep_lo, p.L_out));
        keep_hi = std::max(keep_lo, std::min(keep_hi, p.L_out));
        if (keep_hi == keep_lo) return VP_OK;
        p.row_off0 = keep_lo / 16;
        p.tiles_per_seq = (keep_hi - 16 * p.row_off0 + p.W - 1) / p.W;
    }
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
    dim3 grid((unsigned)std::min(std::max(device_sm_count() / 3, 1), (n_items + FZ_NPIPE - 1) / FZ_NPIPE), 3);
    if (plan.split == 2 && plan.opt == 8) return decb_launch_t<2, 8>(p, grid, s);
    if (plan.split == 2 && plan.opt == 12) return decb_launch_t<2, 12>(p, grid, s);
    if (plan.split == 1 && plan.opt == 8) return decb_launch_t<1, 8>(p, grid, s);
    if (plan.split == 1 && plan.opt == 12) return decb_launch_t<1, 12>(p, grid, s);
    set_error("decb: no kernel instance for split %d, %d outputs per thread", plan.split, plan.opt);
    return VP_ERR_UNSUPPORTED;
}

}  // namespace vp
