// tools/ubench_mix.cu -- do the ALU (LOP3) and XU (FLO/POPC/BREV) pipes overlap on B200?  Independent streams, mixed ratios.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int NLOP, int NXU, int NSTS>
__global__ void k(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[1024];
    uint32_t a[8], x[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 31 + i * 7;
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = seed * 3 + threadIdx.x + i;
    uint32_t m = seed | 0x0F0F0F0F, c = seed ^ 0x33333333;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NLOP; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(a[i & 7]) : "r"(c), "r"(m));
#pragma unroll
        for (int i = 0; i < NXU; i++) asm volatile("bfind.u32 %0, %0;" : "+r"(x[i & 3]));
#pragma unroll
        for (int i = 0; i < NSTS; i++) asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x * 7 + i * 33) & 1023])), "r"(a[i & 7]) : "memory");
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
#pragma unroll
    for (int i = 0; i < 4; i++) r ^= x[i];
    if (r == 0x12345) out[0] = r + sm[5];
}
template <int NLOP, int NXU, int NSTS>
void run() {
    uint32_t *d; cudaMalloc(&d, 4);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(sms * 4), block(256);
    k<NLOP, NXU, NSTS><<<grid, block>>>(d, 12345);
    cudaEventRecord(e0);
    k<NLOP, NXU, NSTS><<<grid, block>>>(d, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double clk = ms * 1e-3 * 1.965e9;
    double iters_per_sm = (double)grid.x * 8 * ITERS / sms;   // warp-iterations per SM
    printf("LOP3 x%2d + FLO x%d + STS x%d : %.2f clk per warp-iteration per SM  (LOP3 alone would be %.2f, FLO alone %.2f)\n", NLOP, NXU, NSTS,
           clk / iters_per_sm, NLOP / 2.0, NXU / 0.5);
}
int main() {
    run<16, 0, 0>(); run<0, 2, 0>(); run<16, 2, 0>(); run<8, 1, 0>(); run<16, 1, 0>(); run<16, 0, 2>(); run<16, 2, 2>(); run<0, 0, 2>();
    return 0;
}
