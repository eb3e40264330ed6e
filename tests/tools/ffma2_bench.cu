// Dev tool (GPU box): FP32 issue/throughput micro-benchmarks for sm_100a — scalar FFMA vs packed FFMA2
// (fma.rn.f32x2), alone and mixed with ALU-pipe and shared-memory instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(seed, seed + 1, seed + 2, seed + 3);
    __syncthreads();
    float a[8];
    u64 p[8];
    unsigned q[8];
    for (int i = 0; i < 8; i++) {
        a[i] = seed + threadIdx.x + i;
        float2 t = make_float2(a[i], a[i] + 0.5f);
        p[i] = *reinterpret_cast<u64*>(&t);
        q[i] = threadIdx.x * 7 + i;
    }
    const float m = 0.999f + seed * 1e-6f, c = 1e-3f + seed;
    float2 mm = make_float2(m, m), cc = make_float2(c, c);
    const u64 pm = *reinterpret_cast<u64*>(&mm), pc = *reinterpret_cast<u64*>(&cc);
    float4 acc4 = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0 || MODE == 2 || MODE == 4) a[i] = fmaf(a[i], m, c);
                if (MODE == 1 || MODE == 3 || MODE == 5) p[i] = ffma2(p[i], pm, pc);
                if (MODE == 2 || MODE == 3) q[i] = (q[i] ^ (q[i] >> 3)) + 0x9e37u;  // 2 ALU instr per FMA
                if ((MODE == 4 || MODE == 5) && (i & 3) == 0) {  // one LDS.128 per 4 FMA
                    float4 v = sh[(threadIdx.x + u * 8 + i + it) & 255];
                    acc4.x += v.x;
                }
            }
        }
    }
    float r = acc4.x;
    for (int i = 0; i < 8; i++) {
        float2 t = *reinterpret_cast<float2*>(&p[i]);
        r += a[i] + t.x + t.y + (float)q[i];
    }
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double fma_per_iter_thread, float* d) {
    int dev_sms;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000, grid = dev_sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, iters, 0.0f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 0.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = fma_per_iter_thread * iters * (double)grid * 256;
    printf("%-34s %8.3f ms  %7.2f TFLOP/s (FMA = 2 flops)\n", name, ms, 2.0 * fma / ms / 1e9);
}

int main() {
    float* d;
    cudaMalloc(&d, 1 << 24);
    run<0>("FFMA  x64/iter", 64, d);
    run<1>("FFMA2 x64/iter (128 fma)", 128, d);
    run<2>("FFMA  + 2 ALU each", 64, d);
    run<3>("FFMA2 + 2 ALU each", 128, d);
    run<4>("FFMA  + LDS.128 per 4", 64, d);
    run<5>("FFMA2 + LDS.128 per 4", 128, d);
    return 0;
}
