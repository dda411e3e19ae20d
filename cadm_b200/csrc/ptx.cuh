// Thin inline-PTX wrappers used by the sm_100a kernels: mbarrier, 1-D bulk async copy (TMA engine,
// SASS UBLKCP), named barriers, tcgen05 / TMEM.  No library dependencies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cadm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- 1-D bulk copy global -> shared, completion on an mbarrier (TMA engine) -----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- named barriers -------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 16-byte shared-memory store to a 32-bit shared address (a generic pointer costs an address conversion per use)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}

// ---- thread-block clusters: rank, barrier, distributed shared memory -------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of THIS CTA) in the CTA of rank `cta`
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void sts128_cluster(uint32_t caddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32_cluster(uint32_t caddr, uint32_t v) {
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16_cluster(uint32_t caddr, uint32_t v) {
    asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(caddr), "h"((unsigned short)v) : "memory");
}
// arrive on an mbarrier of another CTA of the cluster (release at cluster scope: the stores above are visible to its waiters)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// wait with acquire at cluster scope (the arrivals may come from the other CTA of the pair)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}
// generic-proxy writes (of any CTA of the cluster) -> async-proxy reads (tcgen05.mma operands), all state spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace ptx
}  // namespace cadm

// =================================================================================================
// tcgen05 / TMEM (5th-gen tensor cores).  Spellings follow the PTX ISA as used by CUTLASS' cute/arch/*sm100*.hpp.
// =================================================================================================
namespace cadm {
namespace tc {

// ---- TMEM allocation (one full warp; power-of-two column count >= 32) ---------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- shared-memory matrix descriptor: K-major, no swizzle ---------------------------------------
// canonical layout (16-byte units): ((8, n), 2) : ((1, SBO), LBO)  -- cute/atom/mma_traits_sm100.hpp make_umma_desc
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// ---- instruction descriptor, kind::f16, fp32 accumulate, K-major A and B, M = 128 ----------------
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t a_fmt, uint32_t b_fmt, uint32_t n) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mbarrier arrive when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}

// the same arrive on the mbarrier at this shared-memory offset in EVERY CTA of `cta_mask` (a CTA pair hands "my half of the
// layer is done" to both halves with one instruction)
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, 16 / 8 consecutive columns --------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- fp32 -> (fp16 hi, fp16 lo) split, two values at a time ----------------------------------------
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (both saturating).  Error <= max(2^-22 |x|, 2^-25): the residual
// may be an fp16 subnormal (absolute spacing 2^-24), so operands are pre-scaled by powers of two (kXScale, kWScale)
// to keep typical residuals in the normal range.  A and B of one tcgen05.mma must share a format (mixing f16 with
// bf16 traps as an illegal instruction on sm_100a), hence fp16 for both halves.
constexpr float kXScale = 8.0f;        // activations are stored as 8 x
constexpr float kWScale = 64.0f;       // weights are stored as 64 w     -> accumulators hold 512 x.w
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    uint32_t h;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));   // low half <- x0
#if defined(CADM_SPLIT_FHADD) && CADM_SPLIT_FHADD
    // Experimental (off by default, not yet measured on the device): mixed-precision add of PTX ISA 8.6 -- neg.f16 + add.f32.f16
    // fold into one FHADD per value, the same x - float(h) bit for bit (tools/probes/split_fhadd.cu).
    float r0, r1;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tneg.f16 a, a;\n\tneg.f16 b, b;\n\tadd.f32.f16 %0, a, %3;\n\tadd.f32.f16 %1, b, %4;\n\t}"
        : "=f"(r0), "=f"(r1) : "r"(h), "f"(x0), "f"(x1));
#else
    float h0, h1;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(h0), "=f"(h1) : "r"(h));
    const float r0 = x0 - h0, r1 = x1 - h1;
#endif
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
    hi = h;
}

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two IEEE-rounded results) ----------
// The epilogues are issue-slot / MUFU bound, not FMA-pipe bound, so halving the instruction count of the plain
// arithmetic around the two MUFU ops per activation is what shortens them.  Each lane is rounded exactly like the
// scalar instruction (rn), so results do not depend on how values are paired.
__device__ __forceinline__ unsigned long long pk(float2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ float2 upk(unsigned long long a) { return *reinterpret_cast<float2*>(&a); }
#if defined(CADM_NO_F32X2) && CADM_NO_F32X2      // A/B switch: the same arithmetic with scalar instructions
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)));
    return upk(d);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return upk(d);
}
#endif
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Two activations of a hidden layer from their pre-activations in the exponent's units:
//   u = -log2(e) * pre      ->      y = 8 u / (1 + 2^u) = -8 log2(e) * swish(pre)
// i.e. the stored activation carries the scale kActScale = -8 log2(e) instead of kXScale; the next layer's epilogue
// removes it together with the weight scale (accumulator * 1/512 = -log2(e) * (W . swish): the exponent's units again,
// with an exact power-of-two factor).  Per pair: 1 FFMA2 + 2 EX2 + 2 RCP + 1 FMUL2 (the scalar form needed 2 FMUL more).
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kActScale = -8.0f * kLog2e;
__device__ __forceinline__ float2 swish_pair_u(float2 u) {
    const float2 t = make_float2(ex2_approx(u.x), ex2_approx(u.y));
    const float2 d = ffma2(t, make_float2(0.125f, 0.125f), make_float2(0.125f, 0.125f));     // (1 + t) / 8, exact scaling
    const float2 r = make_float2(rcp_approx(d.x), rcp_approx(d.y));
    return fmul2(u, r);
}

// 8 * swish(x) given x8 = 8 x:  x8 * sigmoid(x8 / 8), with ex2.approx / rcp.approx (each <= 2 ulp)
__device__ __forceinline__ float swish8_fast(float x8) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x8 * (-1.4426950408889634f / 8.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return x8 * r;
}

__device__ __forceinline__ float swish_fast(float x) {
    // x * sigmoid(x) with ex2.approx / rcp.approx (each <= 2 ulp); exact limits: x -> -inf gives -0, +inf gives x
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return x * r;
}

}  // namespace tc
}  // namespace cadm
