// Probe: distributed shared memory addressing in a cluster of 2: mapa results, remote stores by address arithmetic
// (local shared::cta address + (mapa(base, peer) - base)), remote mbarrier arrive, multicast tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/dsmem_probe tools/probes/dsmem_probe.cu && /tmp/dsmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cadm_b200/csrc/ptx.cuh"
using namespace cadm;

__global__ void __launch_bounds__(128, 1) probe(unsigned* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem);                 // [256]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2048);          // remote-arrive barrier, count 128 (peer's threads)
    uint64_t* cbar = bar + 1;                                          // multicast commit barrier, count 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const uint32_t rank = ptx::cluster_ctarank(), peer = rank ^ 1u;
    const int tid = threadIdx.x;
    buf[tid] = 0xdead0000u + rank;
    buf[128 + tid] = 0xbeef0000u + rank;
    if (tid == 0) { ptx::mbar_init(bar, 128); ptx::mbar_init(cbar, 2); ptx::fence_mbar_init(); }
    if (tid < 32) { tc::tmem_alloc(tmem_slot, 32); tc::tmem_relinquish(); }
    tc::fence_before_sync();
    __syncthreads();
    ptx::cluster_sync();
    tc::fence_after_sync();
    const uint32_t smem0 = ptx::smem_u32(smem);
    const uint32_t m_own = ptx::mapa(smem0, rank), m_peer = ptx::mapa(smem0, peer);
    const uint32_t rdelta = m_peer - smem0;
    // remote store: element tid of the peer's first half <- 0x1000 * (rank + 1) + tid
    ptx::sts32_cluster(ptx::smem_u32(&buf[tid]) + rdelta, 0x1000u * (rank + 1) + tid);
    ptx::mbar_arrive_cluster(ptx::smem_u32(bar) + rdelta);
    ptx::mbar_wait_cluster(bar, 0);                                    // all 128 threads of the peer have stored + arrived
    const uint32_t got = buf[tid];
    // multicast commit (no MMAs outstanding: arrives immediately) from one thread of each CTA to both CTAs
    if (tid == 0) tc::mma_commit_multicast(cbar, (uint16_t)3);
    ptx::mbar_wait_cluster(cbar, 0);
    if (tid < 4) {
        unsigned* o = out + (blockIdx.x * 4 + tid) * 8;
        o[0] = smem0; o[1] = m_own; o[2] = m_peer; o[3] = got; o[4] = buf[128 + tid]; o[5] = rank; o[6] = 1;
    }
    __syncthreads();
    ptx::cluster_sync();
    if (tid < 32) tc::tmem_dealloc(*tmem_slot, 32);
}

int main() {
    unsigned* d;
    cudaMalloc(&d, 4096);
    cudaMemset(d, 0, 4096);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 4096;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    unsigned h[128];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("launch %s, sync %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    for (int b = 0; b < 4; ++b)
        for (int t = 0; t < 2; ++t) {
            unsigned* o = h + (b * 4 + t) * 8;
            printf("block %d tid %d: rank %u smem0 %08x mapa(own) %08x mapa(peer) %08x  received %08x (expect %08x)  untouched %08x  done %u\n", b, t, o[5],
                   o[0], o[1], o[2], o[3], 0x1000u * ((o[5] ^ 1u) + 1) + t, o[4], o[6]);
        }
    return 0;
}
