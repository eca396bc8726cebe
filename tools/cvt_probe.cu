// Probe (GPU box): issue cost of the fp32 -> fp16 hi / lo split used by the tensor-core epilogues, per warp instruction and SM
// sub-partition, for 1 / 2 / 4 warps per sub-partition:  cvt.rn.f16x2.f32 (F2FP), f16 -> f32 (HADD2.F32), FADD, FMNMX, FFMA, and
// the complete pack8_split16 sequence against a Veltkamp split on the FMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/cvt_probe tools/cvt_probe.cu
#include <cstdio>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../volpick_b200/csrc/tc_ptx.cuh"

using namespace vp;

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long *cycles, float *sink, int iters, float seed) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = seed * (float)(threadIdx.x + 1) + (float)i * 0.37f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 8 F2FP
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                acc ^= *reinterpret_cast<const uint32_t *>(&h);
                v[2 * i] += 1.f;
            }
        } else if (MODE == 1) {  // 16 half -> float
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t b = acc + i + it;
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&b));
                v[2 * i] += f.x;
                v[2 * i + 1] += f.y;
            }
        } else if (MODE == 2) {  // 16 FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += seed;
        } else if (MODE == 3) {  // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], seed + (float)it);
        } else if (MODE == 4) {  // the epilogue's split of 16 values (pack8_split16 x 2)
            uint4 h0, l0, h1, l1;
            pack8_split16<2>(&v[0], h0, l0);
            pack8_split16<2>(&v[8], h1, l1);
            acc ^= h0.x ^ h0.y ^ h0.z ^ h0.w ^ l0.x ^ l0.y ^ l0.z ^ l0.w ^ h1.x ^ h1.y ^ h1.z ^ h1.w ^ l1.x ^ l1.y ^ l1.z ^ l1.w;
            v[0] += 1.f;
        } else if (MODE == 5) {  // Veltkamp split (FMA pipe) + 2 F2FP per pair
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = v[2 * i], b = v[2 * i + 1];
                const float ca = a * 8193.f, cb = b * 8193.f;
                const float ha = ca - (ca - a), hb = cb - (cb - b);
                const __half2 hh = __floats2half2_rn(ha, hb);
                const __half2 ll = __floats2half2_rn(a - ha, b - hb);
                acc ^= *reinterpret_cast<const uint32_t *>(&hh) ^ *reinterpret_cast<const uint32_t *>(&ll);
            }
            v[0] += 1.f;
        } else if (MODE == 6) {  // 16 FFMA (3 registers)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], seed, v[(i + 1) & 15]);
        } else if (MODE == 7) {  // integer split: hi = fp16(v) by F2FP, float(hi) rebuilt on the integer pipe (shift / mask), lo by F2FP
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = v[2 * i], b = v[2 * i + 1];
                const __half2 hh = __floats2half2_rn(a, b);
                const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hh);
                // fp16 -> fp32 for normal, positive values: exponent rebias (+112) and mantissa shift
                const float fa = __uint_as_float(((hb & 0x7fffu) << 13) + 0x38000000u);
                const float fb = __uint_as_float(((hb >> 16) << 13) + 0x38000000u);
                const __half2 ll = __floats2half2_rn(a - fa, b - fb);
                acc ^= hb ^ *reinterpret_cast<const uint32_t *>(&ll);
            }
            v[0] += 1.f;
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 1234.5f || acc == 0x12345u) sink[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int ninstr) {
    long long *d;
    float *sink;
    cudaMalloc(&d, 8);
    cudaMalloc(&sink, 4096);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        probe<MODE><<<1, 32 * warps>>>(d, sink, iters, 1.0001f);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %2d warps / SM: %7.1f cycles per iteration", name, warps, (double)h / iters);
        if (ninstr) printf(" = %.2f cycles per warp instruction and sub-partition", (double)h / iters / ninstr / (warps / 4));
        printf("\n");
    }
    cudaFree(d);
    cudaFree(sink);
}

int main() {
    run<0>("8 x cvt.rn.f16x2.f32 (F2FP)", 8);
    run<1>("16 x f16 -> f32 (HADD2.F32)", 16);
    run<2>("16 x FADD", 16);
    run<3>("16 x FMNMX", 16);
    run<6>("16 x FFMA", 16);
    run<4>("hi / lo split of 16 values (current)", 0);
    run<5>("Veltkamp split of 16 values", 0);
    run<7>("split with integer fp16 -> fp32", 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
