// tcgen05 / TMEM / mbarrier / cp.async PTX helpers shared by the tensor-core kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace vp {

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef VP_MBAR_SLEEP  // A/B build: back off with nanosleep between polls (the waiting warps of a long hand-over chain otherwise
                      // take issue slots from the working warps of their scheduler)
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)VP_MBAR_SLEEP)
        : "memory");
    return;
#endif
#ifdef VP_MBAR_NOHINT  // A/B build: default time limit of try_wait (the warp re-polls every ~30 cycles)
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
    return;
#endif
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)  // suspend-time hint (as CUTLASS passes it).  Measured neutral on sm_100a: the polls stay
                                     // 27 - 67 % of the executed instructions of these kernels, with or without the hint
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// .ca: the 32-byte sector of a 16-byte piece is kept in L1, so the neighbouring piece (the next 8-channel plane of the same
// row, copied by the same thread right after) does not go to L2 again.  Only for data written by EARLIER kernels.
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 / bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = byte distance between the two K halves,
// SBO = byte distance between 8-row groups.  (cute::UMMA::SmemDescriptor, version 1 = Blackwell.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor for kind::f16: D = f32, A/B = fp16 (0) or bf16 (1), both K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(int n, int fmt16) {
    return (1u << 4) | ((uint32_t)fmt16 << 7) | ((uint32_t)fmt16 << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void split16(float v, int fmt16, int split, uint16_t &hi, uint16_t &lo) {
    if (fmt16 == 1) {
        hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        lo = 0;
    } else {
        const __half h = __float2half_rn(v);
        hi = __half_as_ushort(h);
        lo = (split == 2) ? __half_as_ushort(__float2half_rn(v - __half2float(h))) : (uint16_t)0;
    }
}

// 8 fp32 values -> 8 x 16-bit (one 16-byte row of an 8-channel plane): fp16 hi + fp16 lo (SPLIT == 2) or bf16.
template <int SPLIT>
__device__ __forceinline__ void pack8_split16(const float *v, uint4 &hi, uint4 &lo) {
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

// d += (a0, a1) * w: one FFMA2 (sm_100: two fp32 FMAs per instruction; the scalar multiplier is broadcast to both halves).  Each
// half is an ordinary fp32 fused multiply-add (round to nearest): results equal fmaf() per element.
__device__ __forceinline__ void ffma2(float2 &d, float a0, float a1, float w) {
    unsigned long long dd, aa, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(ww));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Wider TMEM loads: one tcgen05.ld of NC consecutive 32-bit columns for the 32 lanes of this warp's quarter.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Named barrier among `nthreads` threads (ids 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// All tcgen05.mma of one 128-row tile of an implicit-GEMM conv layer, fully unrolled at compile time so that
// every descriptor is (base + immediate): the issuing thread spends ~3 uniform instructions per MMA instead of
// a dependent constant-bank load per descriptor (measured: ~80 cycles per MMA with run-time schedule tables).
//   NQ > 0: input of 16*NQ channels, tap j of channel pair q reads rows [j, j+128) of planes 2q, 2q+1
//           (LBO = plane pitch `rows`);  NQ == 0: 8-channel input, K step = taps 2j, 2j+1 (LBO = one row).
//   a16 / w16: shared-memory addresses (16-byte units) of staged row 0 (hi split) / weight block 0.
//   Weight blocks: [block][split][k-half][NOUT][8]; activation planes: [split][plane][rows][8].
// Order = (tap, pair, split term), the order of the host-built schedule (TcLayer::mma).
template <int NOUT, int SPLIT, int NTAPS, int NQ>
__device__ __forceinline__ void umma_conv_tile(uint32_t d_tmem, uint32_t a16, uint32_t rows, uint32_t w16, uint32_t idesc,
                                               uint32_t accumulate_first) {
    constexpr int NTERM = SPLIT == 2 ? 3 : 1;  // hi*hi, hi*lo, lo*hi
    constexpr int CIN8 = NQ == 0 ? 1 : 2 * NQ;
    constexpr int QN = NQ == 0 ? 1 : NQ;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1 (Blackwell), SBO = 128 B
    const uint32_t a_base = a16 | ((NQ == 0 ? 1u : rows) << 16);
    const uint32_t b_base = w16 | ((uint32_t)NOUT << 16);
#pragma unroll
    for (int j = 0; j < NTAPS; ++j)
#pragma unroll
        for (int q = 0; q < QN; ++q)
#pragma unroll
            for (int t = 0; t < NTERM; ++t) {
                const int sa = (t == 2) ? 1 : 0, sb = (t == 1) ? 1 : 0;
                const uint32_t a_off = (uint32_t)(sa * CIN8 + (NQ == 0 ? 0 : 2 * q)) * rows + (uint32_t)(NQ == 0 ? 2 * j : j);
                const uint32_t b_off = (uint32_t)(((NQ == 0 ? j : j * NQ + q) * SPLIT + sb) * 2 * NOUT);
                const bool first = (j == 0 && q == 0 && t == 0);
                umma_f16(d_tmem, desc_hi | (uint64_t)(a_base + a_off), desc_hi | (uint64_t)(b_base + b_off), idesc,
                         first ? accumulate_first : 1u);
            }
}

// f16x3 variant with the weight splits stacked along N: weight blocks [block][k-half][2 NOUT][8] (rows 0..NOUT-1 =
// W_hi, NOUT..2 NOUT-1 = W_lo).  Per K step:  D[:, 0:2N] += A_hi [W_hi ; W_lo]   and   D[:, 0:N] += A_lo W_hi,
// so the epilogue adds columns n and NOUT + n.  The first MMA (accumulate_first == 0) initialises all 2 NOUT columns.
template <int NOUT, int NTAPS, int NQ>
__device__ __forceinline__ void umma_conv_tile_stacked(uint32_t d_tmem, uint32_t a16, uint32_t rows, uint32_t w16, uint32_t idesc_2n,
                                                       uint32_t idesc_n, uint32_t accumulate_first) {
    static_assert(NQ > 0, "stacked schedule is for >= 16-channel inputs");
    constexpr int CIN8 = 2 * NQ;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1 (Blackwell), SBO = 128 B
    const uint32_t a_base = a16 | (rows << 16);
    const uint32_t b_base = w16 | ((uint32_t)(2 * NOUT) << 16);  // k-half stride = 2 NOUT rows
#pragma unroll
    for (int j = 0; j < NTAPS; ++j)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const uint32_t a_hi = (uint32_t)(2 * q) * rows + (uint32_t)j;
            const uint32_t a_lo = (uint32_t)(CIN8 + 2 * q) * rows + (uint32_t)j;
            const uint32_t b_off = (uint32_t)((j * NQ + q) * 4 * NOUT);
            const bool first = (j == 0 && q == 0);
            umma_f16(d_tmem, desc_hi | (uint64_t)(a_base + a_hi), desc_hi | (uint64_t)(b_base + b_off), idesc_2n,
                     first ? accumulate_first : 1u);
            umma_f16(d_tmem, desc_hi | (uint64_t)(a_base + a_lo), desc_hi | (uint64_t)(b_base + b_off), idesc_n, 1u);
        }
}

// 256-bit global accesses (sm_100: LDG / STG .256).  The epilogues' traffic is 16-byte pieces per lane at a 128 / 256-byte
// lane stride: in the res-CNN stack the L1 data pipe (58 %) and the L2 tag stage (54 %, 1.2 sectors per request) were the
// busiest units, i.e. such a kernel is bound by the NUMBER of memory transactions.  A thread now moves whole 32-byte sectors.
__device__ __forceinline__ void ld_global_256(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_global_256(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st_global_256(uint16_t *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                 "r"(b.z), "r"(b.w)
                 : "memory");
}

// Two 8-channel groups of one 16-bit row: one 32-byte store when they are adjacent and 32-byte aligned, else two 16-byte stores.
__device__ __forceinline__ void st_pair16(uint16_t *p0, const uint4 &a, uint16_t *p1, const uint4 &b) {
    if (p1 == p0 + 8 && (reinterpret_cast<uintptr_t>(p0) & 31u) == 0) {
        st_global_256(p0, a, b);
    } else {
        *reinterpret_cast<uint4 *>(p0) = a;
        *reinterpret_cast<uint4 *>(p1) = b;
    }
}

}  // namespace vp
