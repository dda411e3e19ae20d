// Probe: how fast can a CTA push bytes into its cluster peer's shared memory?  (a) st.shared::cluster.v4 from 512 threads + release
// arrive, (b) cp.async.bulk.shared::cluster.shared::cta (TMA engine, smem -> remote smem) with complete_tx on the peer's mbarrier.
// Both CTAs of the pair send at the same time (the kernel's traffic pattern).  Prints cycles per 16 KB and bytes per cycle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/dsmem_bw tools/probes/dsmem_bw_probe.cu && /tmp/dsmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cadm_b200/csrc/ptx.cuh"
using namespace cadm;

constexpr int kBytes = 16384;
constexpr int kIters = 64;

__global__ void __launch_bounds__(512, 1) probe(int mode, int chunk, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* src = smem;                       // [16 KB]
    unsigned char* dst = smem + kBytes;              // [16 KB] written by the peer
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kBytes);
    const uint32_t rank = ptx::cluster_ctarank(), peer = rank ^ 1u;
    const int tid = threadIdx.x;
    for (int i = tid; i < kBytes / 4; i += 512) reinterpret_cast<uint32_t*>(src)[i] = i + rank;
    if (tid == 0) { ptx::mbar_init(bar, mode == 0 ? 16 : 1); ptx::fence_mbar_init(); }
    __syncthreads();
    ptx::cluster_sync();
    const uint32_t smem0 = ptx::smem_u32(smem);
    const uint32_t rdelta = ptx::mapa(smem0, peer) - smem0;
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < kIters; ++it) {
        if (mode == 0) {
            // 512 threads x 2 x 16 B = 16 KB, then fence + release arrive per warp (what rollout_tcp.cu did first)
            const uint32_t a = ptx::smem_u32(dst) + rdelta + tid * 16;
            ptx::sts128_cluster(a, tid, it, 2, 3);
            ptx::sts128_cluster(a + 8192, tid, it, 4, 5);
            __syncwarp();
            if ((tid & 31) == 0) ptx::mbar_arrive_cluster(ptx::smem_u32(bar) + rdelta);
        } else {
            // one thread: expect the peer's bytes on my barrier, push mine in `chunk`-byte bulk copies
            if (tid == 0) {
                ptx::mbar_arrive_expect_tx(bar, kBytes);
                ptx::fence_proxy_async();
                for (int o = 0; o < kBytes; o += chunk)
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     ptx::smem_u32(dst) + rdelta + o),
                                 "r"(ptx::smem_u32(src) + o), "r"(chunk), "r"(ptx::smem_u32(bar) + rdelta)
                                 : "memory");
            }
        }
        ptx::mbar_wait_cluster(bar, ph);             // the peer's 16 KB have landed here
        ph ^= 1u;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    __syncthreads();
    ptx::cluster_sync();
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    const int smem_bytes = 2 * kBytes + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    struct { int mode, chunk; const char* name; } cases[] = {{0, 0, "st.shared::cluster.v4 x 512 threads + release arrive"},
                                                             {1, 16384, "cp.async.bulk smem->peer smem, 1 x 16 KB"},
                                                             {1, 2048, "cp.async.bulk smem->peer smem, 8 x 2 KB"},
                                                             {1, 1024, "cp.async.bulk smem->peer smem, 16 x 1 KB"}};
    for (auto& c : cases) {
        cudaMemset(d, 0, 64);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe, c.mode, c.chunk, d);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-58s %8.0f cycles per 16 KB exchange (both directions at once)  = %5.1f B/cycle per direction   [%s %s]\n", c.name, (double)h / kIters,
               kBytes / ((double)h / kIters), cudaGetErrorString(e), cudaGetErrorString(e2));
    }
    return 0;
}
