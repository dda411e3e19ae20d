// Probe: what slows the swapped kernel's MMA stream when the epilogue warps run beside it?  One CTA: an elected thread issues
// the kernel's MMA pattern (per K16 block: M=128 x N=64 "W_hi [X_hi;X_lo]" then M=128 x N=32 "W_lo X_hi", K-major A, MN-major B,
// resident shared memory) while the 16 epilogue warps run ONE kind of work in a loop.  Prints cycles per K-block pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_interference tools/probes/mma_interference.cu && /tmp/mma_interference
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cadm_b200/csrc/ptx.cuh"

using namespace cadm;

constexpr int kThreads = 576;
constexpr int kRows = 32;
constexpr int kXsbo = (208 / 8) * 128;
constexpr int kXbytes = (kRows / 8) * kXsbo;           // one operand half
constexpr int kSlot = 4 * 8192;                        // 4 K blocks of [hi | lo] of a 128-row tile

__host__ __device__ constexpr uint32_t idesc_sw(uint32_t rows, uint32_t M = 128u) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 16) | ((rows >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1) probe(int mode, int npairs, const unsigned char* src, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* xb = smem;                                      // [X_hi | X_lo] + a second buffer the epilogue writes
    unsigned char* ring = smem + 4 * kXbytes;                      // 4 slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 4 * kSlot);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(bars + 7);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (4 * kXbytes + 4 * kSlot) / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) ptx::mbar_init(&bars[i], 1);
        *stop = 0u;
        ptx::fence_mbar_init();
    }
    if (warp == 1) { __syncwarp(); tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
    ptx::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t x16 = ptx::smem_u32(xb) >> 4, w16 = ptx::smem_u32(ring) >> 4;
            const uint32_t hi32 = (1u << 14);
            const uint64_t a_top = (uint64_t)(hi32 | (128u >> 4)) << 32, b_top = (uint64_t)(hi32 | ((uint32_t)kXsbo >> 4)) << 32;
            const uint32_t b_lbo = (128u >> 4) << 16, R = (mode & 0x200) ? 64 : 128, MM = R;      // 0x200: M = 64 tiles (64 weight rows)
            const long long t0 = clock64();
            for (int p = 0; p < npairs; p += 4) {
                const uint32_t slot = w16 + (uint32_t)((p >> 2) & 3) * (kSlot >> 4);
                uint64_t a_hi = a_top | (slot | (R << 16)), a_lo = a_top | ((slot + 4u * 2u * R) | (R << 16));
                uint64_t b = b_top | ((x16 + (uint32_t)((p % 12)) * 16u) | b_lbo);
                const uint32_t d = tmem_base + ((p >> 2) & 1) * 256u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tc::mma_f16_ss(d, a_hi, b, idesc_sw(2 * kRows, MM), j ? 1u : 0u);
                    tc::mma_f16_ss(d, a_lo, b, idesc_sw(kRows, MM), 1u);
                    a_hi += 2u * R; a_lo += 2u * R; b += 16u;
                }
                tc::mma_commit(&bars[1]);
            }
            tc::mma_commit(&bars[0]);
            ptx::mbar_wait(&bars[0], 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
            *stop = 1u;
        }
        __syncwarp();
    } else if (warp == 0) {
        if (tid == 0 && (mode & 0x100)) {                      // weight stream: bulk copies into the ring while it is read
            uint32_t ph[4] = {0u, 0u, 0u, 0u};
            long long n = 0;
            for (int i = 0; i < 4; ++i) {
                ptx::mbar_arrive_expect_tx(&bars[2 + i], kSlot);
                ptx::bulk_g2s(ring + i * kSlot, src + (size_t)(blockIdx.x / 25) * 622592 + ((n++ * kSlot) % 589824), kSlot, &bars[2 + i]);
            }
            while (!*stop) {
                for (int i = 0; i < 4; ++i) {
                    ptx::mbar_wait(&bars[2 + i], ph[i]);
                    ph[i] ^= 1u;
                    ptx::mbar_arrive_expect_tx(&bars[2 + i], kSlot);
                    ptx::bulk_g2s(ring + i * kSlot, src + (size_t)(blockIdx.x / 25) * 622592 + ((n++ * kSlot) % 589824), kSlot, &bars[2 + i]);
                }
            }
            for (int i = 0; i < 4; ++i) ptx::mbar_wait(&bars[2 + i], ph[i]);
            if (blockIdx.x == 0) out[1] = n * kSlot;
        }
    } else {
        const int m = mode & 0xff;
        const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128u;      // columns the MMAs do not write
        unsigned char* dst = xb + 2 * kXbytes + ((warp - 2) >> 2) * kXsbo + lane * 16;
        float x[8];
        for (int j = 0; j < 8; ++j) x[j] = 0.5f + 0.01f * (lane + j);
        long long iters = 0;
        while (!*stop) {
            ++iters;
            if (m == 0) { __nanosleep(500); continue; }
            if (m == 1 || m == 5) {                          // the epilogue's arithmetic: 16 MUFU + packed FMA per 8 values
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 y = tc::swish_pair_u(tc::ffma2(make_float2(x[2 * j], x[2 * j + 1]), make_float2(0.7f, 0.7f), make_float2(0.1f, 0.1f)));
                    x[2 * j] = y.x; x[2 * j + 1] = y.y;
                }
            }
            if (m == 2 || m == 5) {                          // TMEM loads (two x8 per chunk, as the merged accumulator needs)
                uint32_t v[8], v1[8];
                tc::tmem_ld8(tl, v);
                tc::tmem_ld8(tl + 32, v1);
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] += 1e-30f * __uint_as_float(v[j] ^ v1[j]);
            }
            if (m == 3 || m == 5) {                          // the two 16-byte operand stores of a chunk
                uint32_t h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) tc::split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
                *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(dst + kXbytes) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            if (m == 4) {                                    // shared-memory loads
                uint4 q;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(ptx::smem_u32(dst)));
                x[0] += 1e-30f * __uint_as_float(q.x ^ q.y ^ q.z ^ q.w);
            }
            if (m == 6 || m == 5) ptx::fence_proxy_async();
            if (m == 7) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], 0.999f, 0.001f);
            }
            if (m == 8) {                                    // stores at the epilogue's duty cycle (~1 chunk per 300 cycles per warp)
                uint32_t h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) tc::split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
                *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(dst + kXbytes) = make_uint4(l[0], l[1], l[2], l[3]);
                __nanosleep(200);
            }
        }
        float acc = 0.f;
        for (int j = 0; j < 8; ++j) acc += x[j];
        if (acc == 123.456f) out[3] = 1;
        if (tid == 64 && blockIdx.x == 0) out[2] = iters;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tc::tmem_dealloc(tmem_base, 512); }
}

int main() {
    long long* d;
    unsigned char* src;
    cudaMalloc(&d, 64);
    cudaMalloc(&src, 8 << 20);
    cudaMemset(src, 0x2c, 8 << 20);
    const int smem_bytes = 4 * kXbytes + 4 * kSlot + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const char* names[] = {"idle warps", "MUFU + packed FMA", "TMEM loads", "split + 2 STS.128", "LDS.128", "full epilogue chunk loop", "fence.proxy.async",
                           "FFMA", "split + 2 STS.128 + sleep"};
    const int npairs = 13 * 2 * 40;
    for (int m64 = 0; m64 < 2; ++m64) {
        cudaMemset(d, 0, 64);
        probe<<<1, kThreads, smem_bytes>>>(m64 ? 0x200 : 0, npairs, src, d);
        cudaDeviceSynchronize();
        long long h[4] = {0, 0, 0, 0};
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M = %3d tiles, N = 64 + 32 per K block: %7.1f cycles per K-block pair\n", m64 ? 64 : 128, (double)h[0] / npairs);
    }
    for (int grid : {1})
    for (int stream = 0; stream < 2; ++stream)
        for (int m = 0; m <= 8; ++m) {
            if (grid > 1 && !(m == 0 || m == 5)) continue;
            cudaMemset(d, 0, 64);
            printf("grid %3d | ", grid);
            probe<<<grid, kThreads, smem_bytes>>>(m | (stream ? 0x100 : 0), npairs, src, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[4] = {0, 0, 0, 0};
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("weight stream %s | epilogue warps: %-28s  %7.1f cycles per K-block pair   (loop iterations per warp %lld, copied %.1f B/clk)  %s\n",
                   stream ? "on " : "off", names[m], (double)h[0] / npairs, h[2], h[0] ? (double)h[1] / h[0] : 0.0, cudaGetErrorString(e));
        }
    return 0;
}
