// Micro-benchmark (dev tool, not part of the library): issue rate of the legacy warp-level tensor path
// (mma.sync -> SASS HMMA) on B200 for TF32 m16n8k8 and FP16 m16n8k16, next to packed FFMA2, as a function
// of resident warps per SM.  Decides whether mma.sync tiles can pay inside the fused log-mel kernel.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tf32(float* out, int iters) {
    float d[8][4];
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) d[j][q] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f16(float* out, int iters) {
    float d[8][4];
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) d[j][q] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
    unsigned long long d[16];
    for (int j = 0; j < 16; ++j) d[j] = threadIdx.x + j;
    unsigned long long a = 0x3f8000013f800001ull, b = 0x3a0000003a000000ull;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[j]) : "l"(a), "l"(b));
    }
    unsigned long long s = 0;
    for (int j = 0; j < 16; ++j) s ^= d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}
__global__ void k_ffma(float* out, int iters) {
    float d[16];
    for (int j = 0; j < 16; ++j) d[j] = threadIdx.x + j;
    float a = 1.0000001f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d[j]) : "f"(a), "f"(b));
    }
    float s = 0;
    for (int j = 0; j < 16; ++j) s += d[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static double run(K kern, int grid, int block, float* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, block>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    float* out;
    cudaMalloc(&out, sizeof(float) * 148 * 4 * 1024);
    const int iters = 20000;
    printf("{\"sms\": %d, \"clock_khz\": %d, \"rows\": [\n", sms, clk);
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (one CTA per SM)
        const int block = wps * 32;
        const double t1 = run(k_tf32, sms, block, out, iters);
        const double t2 = run(k_f16, sms, block, out, iters);
        const double t3 = run(k_ffma2, sms, block, out, iters);
        const double t4 = run(k_ffma, sms, block, out, iters);
        // per SM: warp-instructions per ns
        const double n_mma = (double)iters * 8 * wps, n_f = (double)iters * 16 * wps;
        printf("  {\"warps_per_sm\": %d, \"tf32_m16n8k8_per_sm_per_us\": %.1f, \"tf32_tflops\": %.1f, \"f16_m16n8k16_per_sm_per_us\": %.1f, \"f16_tflops\": %.1f, "
               "\"ffma2_per_sm_per_us\": %.1f, \"ffma_per_sm_per_us\": %.1f}%s\n", wps,
               n_mma / (t1 * 1e3), n_mma * sms * 2048.0 / (t1 * 1e-3) / 1e12, n_mma / (t2 * 1e3), n_mma * sms * 4096.0 / (t2 * 1e-3) / 1e12,
               n_f / (t3 * 1e3), n_f / (t4 * 1e3), wps < 32 ? "," : "");
    }
    printf("]}\n");
    return 0;
}
