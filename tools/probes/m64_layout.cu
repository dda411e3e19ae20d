// Probe: where do the rows of an M = 64 tcgen05.mma (cta_group::1, kind::f16) land in TMEM?  A[i][0] = i + 1, B[n][0] = 1,
// so D[i][n] = i + 1; every TMEM lane is read back with 32x32b loads and printed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/m64 tools/probes/m64_layout.cu && /tmp/m64
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../cadm_b200/csrc/ptx.cuh"
using namespace cadm;

__global__ void __launch_bounds__(128, 1) probe(int M, float* out) {
    __shared__ __align__(1024) unsigned char sA[128 * 32];     // [2 k-chunks][128 rows][16 B]: K-major no-swizzle core matrices
    __shared__ __align__(1024) unsigned char sB[16 * 32];      // [2 k-chunks][16 rows][16 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32 / 2; i += 128) reinterpret_cast<__half*>(sA)[i] = __float2half(0.f);
    for (int i = tid; i < 16 * 32 / 2; i += 128) reinterpret_cast<__half*>(sB)[i] = __float2half(0.f);
    __syncthreads();
    if (tid < 128) reinterpret_cast<__half*>(sA + tid * 16)[0] = __float2half((float)(tid + 1));   // chunk 0, row tid, k = 0
    if (tid < 16) reinterpret_cast<__half*>(sB + tid * 16)[0] = __float2half(1.f);
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) { __syncwarp(); tc::tmem_alloc(&tmem_slot, 32); tc::tmem_relinquish(); }
    ptx::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    // clear the accumulator region first (garbage otherwise in lanes the MMA does not write)
    {
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl), "r"(0xff800000u) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl + 8), "r"(0xff800000u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0 && ptx::elect_one()) {
        // K-major A: LBO = bytes between the two k-chunks (M rows x 16 B ... we stored 128 rows per chunk), SBO = 128 B between 8-row groups
        const uint64_t a = tc::smem_desc(ptx::smem_u32(sA), 128 * 16, 128);
        const uint64_t b = tc::smem_desc(ptx::smem_u32(sB), 16 * 16, 128);
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | (((uint32_t)M >> 4) << 24);
        tc::mma_f16_ss(tmem, a, b, idesc, 0u);
        tc::mma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    uint32_t v[8];
    tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_wait_ld();
    out[tid * 2] = __uint_as_float(v[0]);
    out[tid * 2 + 1] = __uint_as_float(v[7]);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tc::tmem_dealloc(tmem, 32); }
}

int main() {
    float* d;
    cudaMalloc(&d, 256 * 4);
    for (int M : {128, 64}) {
        probe<<<1, 128>>>(M, d);
        cudaError_t e = cudaDeviceSynchronize();
        float h[256];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M = %d (%s): TMEM lane -> D row (column 0 | column 7)\n", M, cudaGetErrorString(e));
        for (int l = 0; l < 128; ++l) printf("%s%3d:%4.0f|%4.0f", l % 8 ? "  " : "\n  ", l, h[2 * l], h[2 * l + 1]);
        printf("\n");
    }
    return 0;
}
