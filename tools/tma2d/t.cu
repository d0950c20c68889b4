#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, int rows, int sep, float* out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(rows * sep * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                         (unsigned)__cvta_generic_to_shared(sm)), "l"(&m), "r"(x), "r"(y), "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
    unsigned ok = 0;
    for (int i = 0; i < 1000000 && !ok; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0) : "memory");
    __syncwarp();
    for (int i = threadIdx.x; i < rows * sep; i += 32) out[i] = sm[i];
    if (threadIdx.x == 0) out[rows * sep] = (float)ok;
}
int main() {
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    printf("entry %d\n", (int)cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    const int Npad = 1024, nv = 11;
    float* G; cudaMalloc(&G, Npad * nv * 4);
    float* h = new float[Npad * nv];
    for (int r = 0; r < nv; ++r) for (int i = 0; i < Npad; ++i) h[r * Npad + i] = r * 10000 + i;
    cudaMemcpy(G, h, Npad * nv * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 64 * 1024);
    for (int sep : {64, 68}) for (int rows : {11, 4}) for (int x : {0, 37}) {
        CUtensorMap m;
        cuuint64_t gd[2] = {(cuuint64_t)Npad, (cuuint64_t)nv}; cuuint64_t gs[1] = {(cuuint64_t)Npad * 4};
        cuuint32_t box[2] = {(cuuint32_t)sep, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        CUresult cr = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, G, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        k<<<1, 32, rows * sep * 4 + 128>>>(m, x, 0, rows, sep, out);
        cudaError_t e = cudaDeviceSynchronize();
        float r[3] = {0, 0, 0};
        if (e == cudaSuccess) { cudaMemcpy(&r[0], out + 0, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&r[1], out + sep + 1, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&r[2], out + rows * sep, 4, cudaMemcpyDeviceToHost); }
        printf("sep %d rows %d x %d: encode %d kernel %s first %.0f row1[1] %.0f (want %d, %d) done %.0f\n", sep, rows, x, (int)cr, cudaGetErrorString(e), r[0], r[1], x, 10000 + x + 1, r[2]);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
