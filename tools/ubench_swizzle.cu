// Where does a 128-byte-swizzled tensor-map copy put the 16-byte pieces of a 16 x 128-byte box?  (sm_100a)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ubench_swizzle tools/ubench_swizzle.cu ; run on a B200
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__global__ void k(const __grid_constant__ CUtensorMap tmap, uint32_t row0, uint8_t *out, uint32_t *info) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem), b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        info[0] = dst;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2048u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0u), "r"(row0), "r"(b)
                     : "memory");
    }
    __syncthreads();
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b)
        : "memory");
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult st;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess || !f) { printf("no encoder\n"); return 1; }
    const size_t rows = 64, len = rows * 128;
    uint8_t *h = (uint8_t *)malloc(len), *d, *dout, hout[2048];
    for (size_t i = 0; i < len; i++) h[i] = (uint8_t)(i / 16);   // every 16-byte piece carries its own number (row * 8 + piece)
    cudaMalloc(&d, len); cudaMalloc(&dout, 2048);
    uint32_t *dinfo, hinfo[4];
    cudaMalloc(&dinfo, 16);
    cudaMemcpy(d, h, len, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {128, rows}, gstr[1] = {128};
    const cuuint32_t box[2] = {128, 16}, estr[2] = {1, 1};
    CUresult r = ((Enc)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    for (uint32_t row0 : {0u, 16u, 56u}) {   // the last one runs 8 rows past the end of the tensor
        k<<<1, 128, 2048>>>(tm, row0, dout, dinfo);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hout, dout, 2048, cudaMemcpyDeviceToHost);
        cudaMemcpy(hinfo, dinfo, 16, cudaMemcpyDeviceToHost);
        printf("row0=%u sync=%s smem base=0x%x (mod 1024 = %u)\n", row0, cudaGetErrorString(e), hinfo[0], hinfo[0] & 1023u);
        int ok = 1;
        for (int r16 = 0; r16 < 16; r16++) {
            printf("  smem row %2d:", r16);
            for (int p = 0; p < 8; p++) {
                const int v = hout[r16 * 128 + p * 16];        // piece number (mod 256) found at physical piece p of row r16
                const int want = (int)(((row0 + r16) * 8 + (p ^ (r16 & 7))) & 255);
                printf(" %3d%s", v, v == want ? "" : "!");
                if (v != want && row0 + r16 < rows) ok = 0;
            }
            printf("\n");
        }
        printf("  physical piece p of row r holds logical piece p ^ (r & 7): %s\n", ok ? "yes" : "NO");
    }
    return 0;
}
