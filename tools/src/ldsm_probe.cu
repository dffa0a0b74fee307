// ldsm_probe.cu -- prints (a) the fragment layout of ldmatrix.m16n16.trans.b8 and (b) the
// shared-memory image of a 3-D TMA box with the 32 B swizzle, as used by xengine_tma.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/ldsm_probe tools/src/ldsm_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void k_ldsm(uint32_t *out)
{
    __shared__ __align__(1024) uint8_t s[512];
    for (int i = threadIdx.x; i < 512; i += 32) s[i] = (uint8_t)i;        // matrix 0: row*16+col, matrix 1: same
    __syncwarp();
    uint32_t a = (uint32_t)__cvta_generic_to_shared(s) + (threadIdx.x & 15) * 16 + (threadIdx.x >> 4) * 256;
    uint32_t r0, r1, r2, r3;
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
    out[threadIdx.x * 4 + 0] = r0;
    out[threadIdx.x * 4 + 1] = r1;
    out[threadIdx.x * 4 + 2] = r2;
    out[threadIdx.x * 4 + 3] = r3;
}

__global__ void k_tma(const __grid_constant__ CUtensorMap tm, uint8_t *out, int c0, int c1, int c2, int bytes)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    uint32_t base = ((uint32_t)__cvta_generic_to_shared(sm) + 1023u) & ~1023u;
    uint8_t *p = sm + (base - (uint32_t)__cvta_generic_to_shared(sm));
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = p[i];
}

int main()
{
    uint32_t *d;
    cudaMalloc(&d, 512);
    k_ldsm<<<1, 32>>>(d);
    uint32_t h[128];
    if (cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("ldsm failed\n"); return 1; }
    printf("ldmatrix.m16n16.x2.trans.b8: smem byte (matrix m, row r, col c) = r*16+c; per lane, per reg, 4 bytes as m:(r,c)\n");
    for (int t = 0; t < 32; t++) {
        printf("lane %2d:", t);
        for (int q = 0; q < 4; q++) {
            printf("  r%d=", q);
            for (int b = 0; b < 4; b++) {
                int v = (h[t * 4 + q] >> (8 * b)) & 255;
                printf("(%d,%d)", v >> 4, v & 15);
            }
        }
        printf("\n");
    }
    // TMA: tensor [t=8][station=4][row bytes=64], box 32 B x 8 t x 4 stations, 32 B swizzle
    const int T = 8, A = 4, RB = 64;
    std::vector<uint8_t> hin(T * A * RB);
    for (int t = 0; t < T; t++)
        for (int s = 0; s < A; s++)
            for (int b = 0; b < RB; b++) hin[(t * A + s) * RB + b] = (uint8_t)((t << 5) | (s << 3) | (b >> 4 & 1) << 2 | (b & 3));
    uint8_t *din, *dout;
    cudaMalloc(&din, hin.size());
    cudaMalloc(&dout, 4096);
    cudaMemcpy(din, hin.data(), hin.size(), cudaMemcpyHostToDevice);
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    cuuint64_t dims[3] = {RB, T, A}, strides[2] = {(cuuint64_t)RB * A, RB};
    cuuint32_t box[3] = {32, 8, 4}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    k_tma<<<1, 128, 4096>>>(tm, dout, 32, 0, 0, 1024);
    std::vector<uint8_t> ho(1024);
    cudaError_t e = cudaMemcpy(ho.data(), dout, 1024, cudaMemcpyDeviceToHost);
    printf("tma rc=%s; smem image, 16 B chunks as t.s.half (expect [station][t][32 B], chunk ^= (addr>>7)&1):\n", cudaGetErrorString(e));
    for (int c = 0; c < 64; c++) {
        int v = ho[c * 16];
        printf("%d.%d.%d%s", v >> 5, (v >> 3) & 3, (v >> 2) & 1, (c % 8 == 7) ? "\n" : "  ");
    }
    return 0;
}
