// Probe (GPU box): tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st) -- layout check against a host
// reference and cycles per MMA for N = 16 .. 256, next to the shared-memory-A form the kernels use today.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ts_probe tools/ts_probe.cu && gpurun_out/ts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../volpick_b200/csrc/tc_ptx.cuh"

using namespace vp;

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

__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 7 + k * 3) % 17 - 8) / 8.f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) / 4.f; }

// ---------------------------------------------------------------- correctness: D[128 x N] = A[128 x K] B[N x K]^T, A in TMEM
template <int N, int K>
__global__ void __launch_bounds__(128, 1) ts_check_kernel(float *out, int a_col0) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    // B: per K step [k-half][N][8]
    __half *sb = reinterpret_cast<__half *>(smem);
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        const int ks = k / 16, kh = (k % 16) / 8, ke = k % 8;
        sb[((ks * 2 + kh) * N + n) * 8 + ke] = __float2half(b_val(n, k));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    // A: lane = row, column a_col0 + k / 2 holds (k even: low half, k odd: high half)
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) {
            const __half2 h = __floats2half2_rn(a_val(tid, k0 + 2 * j), a_val(tid, k0 + 2 * j + 1));
            r[j] = *reinterpret_cast<const uint32_t *>(&h);
        }
        tmem_st8(lane_addr + (uint32_t)(a_col0 + k0 / 2), r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t d_col = 256;
    if (warp == 0) {
        if (elect_one()) {
            const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
            const uint32_t b16 = (smem_u32(smem) >> 4) | ((uint32_t)N << 16);
            for (int ks = 0; ks < K / 16; ++ks)
                umma_f16_ts(tmem_base + d_col, tmem_base + (uint32_t)(a_col0 + ks * 8), desc_hi | (uint64_t)(b16 + ks * 2 * N), umma_idesc(N, 0),
                            ks ? 1u : 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld8(lane_addr + d_col + c0, v);
        for (int j = 0; j < 8; ++j) out[tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------- throughput: NM back-to-back MMAs of one shape
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long *cycles, int nm) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    for (int idx = tid; idx < 16384; idx += 128) reinterpret_cast<uint32_t *>(smem)[idx] = 0u;  // 64 KB of zeros: A tiles and B blocks
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        if (elect_one()) {
            const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
            const uint32_t s16 = smem_u32(smem) >> 4;
            const uint32_t a16 = s16 | (136u << 16);            // planes of 136 rows
            const uint32_t b16 = (s16 + 1024) | ((uint32_t)N << 16);
            const uint32_t idesc = umma_idesc(N, 0);
            const long long t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < nm; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (TS)
                        umma_f16_ts(tmem_base + 256, tmem_base + (uint32_t)(8 * j), desc_hi | (uint64_t)(b16 + j * 2 * N), idesc, 1u);
                    else
                        umma_f16(tmem_base + 256, desc_hi | (uint64_t)(a16 + j), desc_hi | (uint64_t)(b16 + j * 2 * N), idesc, 1u);
                }
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            cycles[0] = t1 - t0;
            cycles[1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------- TMEM load / store rate with NW warps (4 or 8), optionally under MMAs
template <bool ST>
__global__ void __launch_bounds__(256, 1) ldst_rate_kernel(long long *cycles, int iters, float *sink) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t lane_addr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (ST) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) r[j] = acc + j + i;
#pragma unroll
            for (int c = 0; c < 128; c += 8) tmem_st8(lane_addr + c, r);
            tmem_st_wait();
        } else {
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 128; c += 16) {
                tmem_ld16_nowait(lane_addr + c, r);
                tmem_ld_wait();
                acc += r[0] ^ r[15];
            }
        }
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[warp] = t1 - t0;
    if (acc == 0x12345678u) sink[tid] = 1.f;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)


// ---------------------------------------------------------------- tcgen05.ld latency while another warp keeps the tensor pipe busy
template <int N, bool TS, int RD>
__global__ void __launch_bounds__(256, 1) ld_under_mma_kernel(long long *out, int nm) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int go, done;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        go = 0;
        done = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    for (int idx = tid; idx < 16384; idx += 256) reinterpret_cast<uint32_t *>(smem)[idx] = 0u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        if (elect_one()) {
            const uint64_t desc_hi = (uint64_t)(0x4000u | (128u >> 4)) << 32;
            const uint32_t s16 = smem_u32(smem) >> 4;
            const uint32_t a16 = s16 | (136u << 16);
            const uint32_t b16 = (s16 + 1024) | ((uint32_t)N << 16);
            const uint32_t idesc = umma_idesc(N, 0);
            go = 1;
            const long long t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < nm; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (TS)
                        umma_f16_ts(tmem_base + 256, tmem_base + (uint32_t)(8 * j), desc_hi | (uint64_t)(b16 + j * 2 * N), idesc, 1u);
                    else
                        umma_f16(tmem_base + 256, desc_hi | (uint64_t)(a16 + j), desc_hi | (uint64_t)(b16 + j * 2 * N), idesc, 1u);
                }
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            done = 1;
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        __syncwarp();
    } else if (warp >= 4) {
        // readers: warps 4-7 (lane quarters 0-3) read 16 columns of an untouched TMEM region in a loop
        while (!go) {
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 448;
        long long tmax = 0, tsum = 0;
        int cnt = 0;
        uint32_t acc = 0;
        while (!done && cnt < 100000) {
            uint32_t r[16];
            const long long a = clock64();
            if (RD == 0) {
                tmem_ld16_nowait(lane_addr, r);
                tmem_ld_wait();
                acc += r[0] ^ r[7];
            } else if (RD == 1) {  // broadcast 16-byte shared-memory load (the epilogues' bias reads)
                uint4 v4;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w) : "r"(smem_u32(smem) + 60000 + (cnt & 3) * 16) : "memory");
                acc += v4.x ^ v4.w;
            } else {  // 16-byte store per lane (conflict-free), then a load to wait for it
                const uint32_t ad = smem_u32(smem) + 61440 + (tid & 127) * 16;
                uint4 v4;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(acc), "r"(1u), "r"(2u), "r"(3u) : "memory");
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w) : "r"(ad) : "memory");
                acc += v4.x ^ v4.w;
            }
            asm volatile("" ::"r"(acc) : "memory");
            const long long b = clock64();
            tsum += b - a;
            tmax = (b - a) > tmax ? (b - a) : tmax;
            ++cnt;
        }
        if ((tid & 31) == 0) {
            out[2 + (warp - 4) * 3] = cnt;
            out[3 + (warp - 4) * 3] = cnt ? tsum / cnt : 0;
            out[4 + (warp - 4) * 3] = tmax;
        }
        if (acc == 0x1234567u) out[20] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, bool TS, int RD>
static int ld_under_mma() {
    long long *d;
    CK(cudaMalloc(&d, 32 * sizeof(long long)));
    CK(cudaMemset(d, 0, 32 * sizeof(long long)));
    auto kern = ld_under_mma_kernel<N, TS, RD>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int nm = 1024;
    for (int rep = 0; rep < 2; ++rep) kern<<<1, 256, 65536>>>(d, nm);
    CK(cudaDeviceSynchronize());
    long long h[32];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%s under %s MMAs N=%3d: MMAs issue %.1f / complete %.1f cycles each; reader warp 4: %lld loads, mean %lld, max %lld cycles per access\n",
           RD == 0 ? "tcgen05.ld.x16" : RD == 1 ? "LDS.128 broadcast" : "STS.128 + LDS.128", TS ? "TS" : "SS", N, (double)h[0] / nm, (double)h[1] / nm, h[2], h[3], h[4]);
    cudaFree(d);
    return 0;
}

template <int N, int K>
static int check(int a_col0) {
    float *d;
    CK(cudaMalloc(&d, 128 * N * sizeof(float)));
    auto kern = ts_check_kernel<N, K>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    kern<<<1, 128, 65536>>>(d, a_col0);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(128 * N);
    CK(cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)a_val(r, k) * (double)b_val(n, k);
            const double e = fabs(ref - h[r * N + n]);
            if (e > maxerr) maxerr = e;
        }
    printf("TS check N=%d K=%d a_col0=%d: max |err| = %g  (D[0][0]=%g D[5][3]=%g)\n", N, K, a_col0, maxerr, h[0], h[5 * N + 3]);
    cudaFree(d);
    return 0;
}

template <int N, bool TS>
static int rate() {
    long long *d;
    CK(cudaMalloc(&d, 2 * sizeof(long long)));
    auto kern = mma_rate_kernel<N, TS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int nm = 4096;
    for (int rep = 0; rep < 2; ++rep) kern<<<1, 128, 65536>>>(d, nm);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%s N=%3d: issue %.1f cycles / MMA, complete %.1f cycles / MMA\n", TS ? "TS (A in TMEM)" : "SS (A in smem)", N, (double)h[0] / nm,
           (double)h[1] / nm);
    cudaFree(d);
    return 0;
}

template <bool ST>
static int ldst(int nwarps) {
    long long *d;
    float *sink;
    CK(cudaMalloc(&d, 8 * sizeof(long long)));
    CK(cudaMalloc(&sink, 256 * sizeof(float)));
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) ldst_rate_kernel<ST><<<1, 32 * nwarps>>>(d, iters, sink);
    CK(cudaDeviceSynchronize());
    long long h[8];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < nwarps; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)nwarps * 32 * 128 * 4 * iters;
    printf("tcgen05.%s %d warps: %.1f cycles per 128-column pass, %.1f B / cycle / SM\n", ST ? "st" : "ld", nwarps, (double)mx / iters, bytes / mx);
    cudaFree(d);
    cudaFree(sink);
    return 0;
}

int main() {
    if (check<32, 64>(0)) return 1;
    if (check<32, 64>(40)) return 1;
    if (check<16, 112>(8)) return 1;
    if (check<64, 32>(128)) return 1;
    rate<16, true>();
    rate<32, true>();
    rate<64, true>();
    rate<128, true>();
    rate<16, false>();
    rate<32, false>();
    rate<64, false>();
    rate<128, false>();
    ld_under_mma<64, false, 0>();
    ld_under_mma<16, true, 0>();
    ld_under_mma<64, false, 1>();
    ld_under_mma<16, true, 1>();
    ld_under_mma<64, false, 2>();
    ld_under_mma<16, true, 2>();
    ld_under_mma<128, false, 1>();
    ldst<false>(4);
    ldst<false>(8);
    ldst<true>(4);
    ldst<true>(8);
    return 0;
}
