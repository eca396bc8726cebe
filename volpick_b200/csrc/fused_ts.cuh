// Device helpers shared by the fused kernels whose activations live in tensor memory (fused_dec2.cu, fused_enc.cu): tcgen05.mma
// with the A operand in TMEM, tcgen05.st, the rolled shared-memory-operand tiles, and the epilogue pieces (accumulator columns ->
// bias + ReLU -> fp16 hi / lo (or bf16) packed registers -> a TMEM slot; edge exchange between the lane quarters).
#pragma once

#include "tc_ptx.cuh"

namespace vp {

// ------------------------------------------------------------------------------------------ PTX: A operand in tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One block of a polyphase layer whose input lives in TMEM: input sample slot s0 (8 columns per slot: 16 channels as fp16
// pairs; the lo split LO_OFF columns further) -> NOUT accumulator columns.  Tap j reads slot s0 + j.  Weight blocks as in
// tcconv.cu: [tap][split][k-half][NOUT][8].  f16x3: A_hi W_hi + A_hi W_lo + A_lo W_hi.
template <int NOUT, int SPLIT, int NTAPS>
__device__ __forceinline__ void umma_ts_block(uint32_t d_tmem, uint32_t a_hi, uint32_t lo_off, uint32_t w16, uint32_t idesc) {
    constexpr int NTERM = SPLIT == 2 ? 3 : 1;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;  // version 1 (Blackwell), SBO = 128 B
    const uint32_t b_base = w16 | ((uint32_t)NOUT << 16);
#pragma unroll
    for (int j = 0; j < NTAPS; ++j)
#pragma unroll
        for (int t = 0; t < NTERM; ++t) {
            const int sa = (t == 2) ? 1 : 0, sb = (t == 1) ? 1 : 0;
            const uint32_t a = a_hi + (uint32_t)(8 * j) + (sa ? lo_off : 0u);
            const uint32_t b_off = (uint32_t)((j * SPLIT + sb) * 2 * NOUT);
            umma_f16_ts(d_tmem, a, desc_hi | (uint64_t)(b_base + b_off), idesc, (j == 0 && t == 0) ? 0u : 1u);
        }
}

// The shared-memory layers (tc_ptx.cuh umma_conv_tile / umma_conv_tile_stacked) rolled over the taps: one loop iteration = the MMAs
// of one tap with (base + immediate) descriptors, the tap adds a register offset.  Same MMA order as the unrolled forms.
template <int NOUT, int SPLIT, int NTAPS, int NQ>
__device__ __forceinline__ void ts_conv_tile(uint32_t d_tmem, uint32_t a16, uint32_t rows, uint32_t w16, uint32_t idesc) {
    constexpr int NTERM = SPLIT == 2 ? 3 : 1;
    constexpr int CIN8 = 2 * NQ;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
    const uint32_t a_base = a16 | (rows << 16);
    const uint32_t b_base = w16 | ((uint32_t)NOUT << 16);
#pragma unroll 1
    for (int j = 0; j < NTAPS; ++j) {
        const uint32_t aj = a_base + (uint32_t)j, bj = b_base + (uint32_t)(j * NQ * SPLIT * 2 * NOUT);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int t = 0; t < NTERM; ++t) {
                const int sa = (t == 2) ? 1 : 0, sb = (t == 1) ? 1 : 0;
                const uint32_t a_off = (uint32_t)(sa * CIN8 + 2 * q) * rows;
                const uint32_t b_off = (uint32_t)((q * SPLIT + sb) * 2 * NOUT);
                umma_f16(d_tmem, desc_hi | (uint64_t)(aj + a_off), desc_hi | (uint64_t)(bj + b_off), idesc, (q == 0 && t == 0) ? (j > 0 ? 1u : 0u) : 1u);
            }
    }
}
template <int NOUT, int NTAPS, int NQ>
__device__ __forceinline__ void ts_conv_tile_stacked(uint32_t d_tmem, uint32_t a16, uint32_t rows, uint32_t w16, uint32_t idesc_2n,
                                                     uint32_t idesc_n) {
    constexpr int CIN8 = 2 * NQ;
    const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
    const uint32_t a_base = a16 | (rows << 16);
    const uint32_t b_base = w16 | ((uint32_t)(2 * NOUT) << 16);
#pragma unroll 1
    for (int j = 0; j < NTAPS; ++j) {
        const uint32_t aj = a_base + (uint32_t)j, bj = b_base + (uint32_t)(j * NQ * 4 * NOUT);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const uint32_t a_hi = (uint32_t)(2 * q) * rows, a_lo = (uint32_t)(CIN8 + 2 * q) * rows;
            const uint32_t b_off = (uint32_t)(q * 4 * NOUT);
            umma_f16(d_tmem, desc_hi | (uint64_t)(aj + a_hi), desc_hi | (uint64_t)(bj + b_off), idesc_2n, q == 0 ? (j > 0 ? 1u : 0u) : 1u);
            umma_f16(d_tmem, desc_hi | (uint64_t)(aj + a_lo), desc_hi | (uint64_t)(bj + b_off), idesc_n, 1u);
        }
    }
}

// ------------------------------------------------------------------------------------------ epilogue pieces
// 16 accumulator columns from col0 (+ the stacked half NST columns further) -> relu(acc + bias[0..15])
template <int NST>
__device__ __forceinline__ void ts_load16(const float *bias, uint32_t tacc, int col0, float (&v)[16]) {
    uint32_t r[16], r2[NST > 0 ? 16 : 1];
    tmem_ld16_nowait(tacc + (uint32_t)col0, r);
    if constexpr (NST > 0) tmem_ld16_nowait(tacc + (uint32_t)(col0 + NST), r2);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 16; n += 4) {
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

// the same with the 16 biases in registers (epilogues on the hand-over chain: no shared-memory access between tcgen05.ld and tcgen05.st)
template <int NST>
__device__ __forceinline__ void ts_load16r(const float (&bias)[16], uint32_t tacc, int col0, float (&v)[16]) {
    uint32_t r[16], r2[NST > 0 ? 16 : 1];
    tmem_ld16_nowait(tacc + (uint32_t)col0, r);
    if constexpr (NST > 0) tmem_ld16_nowait(tacc + (uint32_t)(col0 + NST), r2);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        float a = __uint_as_float(r[n]);
        if constexpr (NST > 0) a += __uint_as_float(r2[n]);
        v[n] = fmaxf(a + bias[n], 0.f);
    }
}

// edge exchange between the lane quarters: 8 + 8 packed registers of one sample as four 16-byte shared-memory accesses
__device__ __forceinline__ void ts_post16(uint32_t *dst, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    d[0] = make_uint4(h[0], h[1], h[2], h[3]);
    d[1] = make_uint4(h[4], h[5], h[6], h[7]);
    d[2] = make_uint4(l[0], l[1], l[2], l[3]);
    d[3] = make_uint4(l[4], l[5], l[6], l[7]);
}
__device__ __forceinline__ void ts_fetch16(const uint32_t *src, bool have, uint32_t (&h)[8], uint32_t (&l)[8]) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    const uint4 a = have ? s4[0] : z, b = have ? s4[1] : z, c = have ? s4[2] : z, d = have ? s4[3] : z;
    h[0] = a.x, h[1] = a.y, h[2] = a.z, h[3] = a.w, h[4] = b.x, h[5] = b.y, h[6] = b.z, h[7] = b.w;
    l[0] = c.x, l[1] = c.y, l[2] = c.z, l[3] = c.w, l[4] = d.x, l[5] = d.y, l[6] = d.z, l[7] = d.w;
}

// 16 channels of one sample -> the 8 + 8 packed registers of its TMEM slot (zeros when the row lies outside the sequence)
template <int SPLIT>
__device__ __forceinline__ void ts_pack16(const float (&v)[16], bool valid, uint32_t (&h)[8], uint32_t (&l)[8]) {
    uint4 h0, l0, h1, l1;
    if (valid) {
        pack8_split16<SPLIT>(&v[0], h0, l0);
        pack8_split16<SPLIT>(&v[8], h1, l1);
    } else {
        h0 = l0 = h1 = l1 = make_uint4(0u, 0u, 0u, 0u);
    }
    h[0] = h0.x, h[1] = h0.y, h[2] = h0.z, h[3] = h0.w, h[4] = h1.x, h[5] = h1.y, h[6] = h1.z, h[7] = h1.w;
    l[0] = l0.x, l[1] = l0.y, l[2] = l0.z, l[3] = l0.w, l[4] = l1.x, l[5] = l1.y, l[6] = l1.z, l[7] = l1.w;
}

template <int SPLIT>
__device__ __forceinline__ void ts_st_slot(uint32_t tl, uint32_t col_hi, uint32_t lo_off, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
    tmem_st8(tl + col_hi, h);
    if (SPLIT == 2) tmem_st8(tl + col_hi + lo_off, l);
}


}  // namespace vp
