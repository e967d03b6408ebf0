// tools/ubench_pipes.cu -- instruction throughput per SM for the integer ops the stage-1 kernel is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu ; run on the B200 box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed) {
    uint32_t a[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++) a[i] = seed + threadIdx.x * 31 + i * 7;
    uint32_t m = seed | 0x0F0F0F0F, c = seed ^ 0x33333333;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            uint32_t x = a[i];
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(x) : "r"(c), "r"(m));
            if (OP == 1) asm volatile("shr.u32 %0, %0, 1; xor.b32 %0, %0, %1;" : "+r"(x) : "r"(c));   // SHF + LOP3
            if (OP == 2) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x) : "r"(c));
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(c));
            if (OP == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(m));
            if (OP == 5) asm volatile("bfind.u32 %0, %0; add.u32 %0, %0, %1;" : "+r"(x) : "r"(c));   // FLO + IADD
            if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(c));
            if (OP == 7) asm volatile("popc.b32 %0, %0; add.u32 %0, %0, %1;" : "+r"(x) : "r"(c));
            if (OP == 8) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(x) : "r"(c), "r"(m)); // ISETP + SEL
            if (OP == 9) asm volatile("shl.b32 %0, %0, 3; xor.b32 %0, %0, %1;" : "+r"(x) : "r"(c));  // shl (IMAD.SHL or SHF) + LOP3
            if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4; mad.lo.u32 %0, %0, %2, %1;" : "+r"(x) : "r"(c), "r"(m)); // LOP3 + IMAD interleaved
            a[i] = x;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++) r ^= a[i];
    if (r == 0x12345) out[0] = r;
}

template <int OP>
void run(const char *name, int per_iter) {
    uint32_t *d;
    cudaMalloc(&d, 4);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps_per_sm : {16, 32}) {
        dim3 grid(sms * (warps_per_sm / 8)), block(256);
        k<OP><<<grid, block>>>(d, 12345);
        cudaEventRecord(e0);
        k<OP><<<grid, block>>>(d, 12345);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        int clk_khz;
        cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
        double inst = (double)grid.x * 8 * ITERS * UNROLL * per_iter;  // warp instructions
        double per_sm_per_clk = inst / sms / (ms * 1e-3 * clk_khz * 1e3);
        printf("%-28s warps/SM=%2d  %.3f ms  warp-instr/clk/SM=%.3f (at %d MHz nominal)\n", name, warps_per_sm, ms, per_sm_per_clk, clk_khz / 1000);
    }
    cudaFree(d);
}

int main() {
    run<0>("LOP3", 1);
    run<1>("SHF.R+LOP3", 2);
    run<2>("PRMT", 1);
    run<3>("IMAD", 1);
    run<4>("IMAD.HI", 1);
    run<5>("FLO+IADD", 2);
    run<6>("IADD", 1);
    run<7>("POPC+IADD", 2);
    run<8>("ISETP+SEL", 2);
    run<9>("SHL+LOP3", 2);
    run<10>("LOP3+IMAD", 2);
    return 0;
}
