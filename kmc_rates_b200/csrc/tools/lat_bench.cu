// Single-warp latency / issue-rate probes for the instructions on the panel kernel's dependency chain (needs a GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 lat_bench.cu -o lat_bench && ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(double *out, long long *clk, double a, double b) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    sm[lane] = a + lane; sm[lane + 32] = b;
    __syncwarp();
    double x = a + lane * 1e-9;
    long long t0, t1;
    int slot = 0;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fma(x, b, a);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    // dependent DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = x + b;
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    // 8 independent DFMA chains (throughput for one warp)
    double y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = x + k;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = fma(y[k], b, a);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
#pragma unroll
    for (int k = 0; k < 8; ++k) x += y[k];
    // 32 independent DFMA (one pass like the row update)
    double z[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) z[k] = x + k;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 32; ++k) z[k] = fma(z[k], b, a);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
#pragma unroll
    for (int k = 0; k < 32; ++k) x += z[k];
    // dependent reciprocal chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) x = 1.0 / (x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    // dependent shuffle chain (64-bit)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    // dependent shared-memory round trip: store, syncwarp, load
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        sm[lane] = x;
        __syncwarp();
        x = sm[(lane + 1) & 31];
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    // dependent FFMA chain for reference
    float fx = (float)x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) fx = fmaf(fx, (float)b, (float)a);
    t1 = clock64();
    if (threadIdx.x == 0) clk[slot] = t1 - t0; slot++;
    out[threadIdx.x] = x + fx;
}

int main() {
    double *out; long long *clk;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&clk, 16 * 8);
    for (int threads : {32, 128, 512}) {
        probe<<<1, threads>>>(out, clk, 0.25, 0.999);
        probe<<<1, threads>>>(out, clk, 0.25, 0.999);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        std::printf("threads/CTA %d (cycles per op, warp 0)\n", threads);
        std::printf("  dependent DFMA            %.1f\n", h[0] / 64.0);
        std::printf("  dependent DADD            %.1f\n", h[1] / 64.0);
        std::printf("  8 independent DFMA chains %.1f per DFMA\n", h[2] / 128.0);
        std::printf("  32 independent DFMA       %.1f per DFMA\n", h[3] / 128.0);
        std::printf("  dependent 1.0/x (+DADD)   %.1f\n", h[4] / 16.0);
        std::printf("  dependent 64-bit shuffle  %.1f\n", h[5] / 32.0);
        std::printf("  STS + syncwarp + LDS      %.1f\n", h[6] / 32.0);
        std::printf("  dependent FFMA            %.1f\n", h[7] / 64.0);
    }
    std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
