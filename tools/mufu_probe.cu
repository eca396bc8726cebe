// MUFU throughput probe (B200): lanes per clock per SM of rcp.approx / ex2.approx, alone and mixed with FFMA, at full occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, long long *cycles) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
    float acc = 0.f, acc2 = 0.f, q = 1.0001f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        q += 1e-6f;  // loop-carried: nothing below is invariant
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {  // FADD + MUFU.RCP, dependent chain per element
                float x = a[i] + q;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(x));
            }
            if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 2) {  // the attention inner step: FFMA, MUFU.RCP, FFMA (two accumulator chains)
                float x = fmaf(a[i], q, 1.f), r;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
                if (i & 1) acc2 = fmaf(a[i], r, acc2);
                else acc = fmaf(a[i], r, acc);
            }
            if (MODE == 3) {  // two reciprocals from one MUFU: 1/x = y rcp(xy), 1/y = x rcp(xy)
                if (i & 1) continue;
                float x = fmaf(a[i], q, 1.f), y = fmaf(a[i + 1], q, 1.f), r;
                float xy = x * y;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xy));
                acc = fmaf(a[i], y * r, acc);
                acc2 = fmaf(a[i + 1], x * r, acc2);
            }
        }
    }
    const long long t1 = clock64();
    float s = acc + acc2 + q;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, iters = 65536;
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaMalloc(&cyc, 8);
    const char *names[4] = {"FADD + rcp.approx (dependent)", "ex2.approx alone", "FFMA + rcp + FFMA per element", "pair: 2 FFMA + FMUL + rcp + 2 FMUL + 2 FFMA per 2 elements"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 4; ++mode)
        for (int cps = 1; cps <= 8; cps *= 2) {  // CTAs (256 threads) per SM
            float ms = 0.f;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<sms * cps, 256>>>(out, iters, cyc);
                if (mode == 1) probe<1><<<sms * cps, 256>>>(out, iters, cyc);
                if (mode == 2) probe<2><<<sms * cps, 256>>>(out, iters, cyc);
                if (mode == 3) probe<3><<<sms * cps, 256>>>(out, iters, cyc);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                cudaEventElapsedTime(&ms, e0, e1);
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double elems = (double)iters * 8 * 256 * cps;  // per SM
            printf("%-62s %2d warps/SM: %8.3f ms, block 0: %lld clk -> %.2f elements / clk / SM by events at %d MHz, %.2f by clock64\n", names[mode],
                   cps * 8, ms, h, elems / (ms * 1e-3 * khz * 1e3), khz / 1000, elems / (double)h);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
