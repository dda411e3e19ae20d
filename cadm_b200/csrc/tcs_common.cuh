// Pieces shared by the swapped-operand tensor-core kernels (rollout_tcs.cu: one CTA per row tile; rollout_tcp.cu: a CTA pair per
// two row tiles): instruction descriptor, MMA issue of one weight stage, MN-major operand stores, fast sin / cos, cp.async.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cadm {
namespace tcs {

// instruction descriptor: kind::f16, fp32 accumulate, A K-major, B MN-major (bit 16), M = 128 (or 64), N = rows.
// M = 64 (the heads, when they have at most 64 outputs): the 4 KB weight fetch that paces a small-N MMA halves (measured 60.5
// cycles per K block instead of 88.7, tools/probes/m64_layout.cu); accumulator row j then lives in TMEM lane 32 (j / 16) + j % 16.
__host__ __device__ constexpr uint32_t idesc(uint32_t rows, uint32_t M = 128u) {
    return (1u << 4) | (tc::kFmtF16 << 7) | (tc::kFmtF16 << 10) | (1u << 16) | ((rows >> 3) << 17) | ((M >> 4) << 24);
}

struct Ring {
    int stage;
    uint32_t phase;
    int n;
    __device__ __forceinline__ void advance() {
        if (++stage == n) { stage = 0; phase ^= 1u; }
    }
};

// 8 consecutive rows of one feature / hidden unit k -> one 16-byte store per operand half (MN-major core matrix row)
__device__ __forceinline__ void store_rows8(unsigned char* xhi, unsigned char* xlo, int xsbo, int rgroup, int k, const float (&y)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc::split2(y[2 * j], y[2 * j + 1], h[j], l[j]);
    const int o = rgroup * xsbo + (k >> 3) * 128 + (k & 7) * 16;
    *reinterpret_cast<uint4*>(xhi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(xlo + o) = make_uint4(l[0], l[1], l[2], l[3]);
}

// sin / cos with a two-constant Cody-Waite reduction to [-pi, pi] and the MUFU approximations (|error| ~ 1e-6 for |x| up to a few
// hundred; the accurate sincosf costs ~350 cycles on the per-step critical path)
__device__ __forceinline__ void sincos_reduced(float x, float& s, float& c) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, x);          // 2 pi rounded to fp32 ...
    r = fmaf(k, 1.7484556000744883e-07f, r);             // ... and its remainder
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(r));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(r));
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The MMAs of one weight stage (KBS K16 blocks of one M tile), fully unrolled, descriptors advanced by constant adds.
// TERMS == 3: per block  D[0, 2 rows) (+)= W_hi [X_hi ; X_lo]  (one MMA with N = 2 rows)  and then  D[0, rows) += W_lo X_hi;
// TERMS == 1: D[0, rows) (+)= W_hi X_hi.   a_hi / a_lo / b are complete 64-bit shared-memory descriptors of the first block.
template <int TERMS, int KBS>
__device__ __forceinline__ void issue_blocks(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b, uint32_t a_step, uint32_t rows,
                                             uint32_t idesc1, uint32_t idesc2, uint32_t accum) {
#pragma unroll
    for (int j = 0; j < KBS; ++j) {
        const uint32_t acc = j == 0 ? accum : 1u;
        if (TERMS == 3) {
            tc::mma_f16_ss(d_tmem, a_hi, b, idesc2, acc);                   // D[0, 2 rows) (+)= W_hi [X_hi ; X_lo]
            tc::mma_f16_ss(d_tmem, a_lo, b, idesc1, 1u);                    // D[0, rows)    += W_lo X_hi
        } else {
            tc::mma_f16_ss(d_tmem, a_hi, b, idesc1, acc);
        }
        a_hi += a_step; a_lo += a_step;      // two k-chunks of R rows x 16 B, in 16-byte units
        b += 16u;                            // 256 B per K block
    }
}

//   slot16    (shared-memory address >> 4) of the stage: [hi: 2 kbs k-chunks x (R rows x 16 B)] [lo: same]
//   xhi16     (address >> 4) of the stage's first K block in the X_hi buffer; X_lo starts `rows / 8` row groups behind X_hi
template <int TERMS>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t slot16, uint32_t R, int kbs, uint32_t accum, uint32_t xhi16,
                                            uint32_t rows, uint32_t xsbo) {
    const uint32_t hi32 = (1u << 14);                                   // descriptor version 1 (bit 46)
    const uint64_t a_top = (uint64_t)(hi32 | (128u >> 4)) << 32;        // A: SBO = 128 B between 8-row groups
    const uint64_t b_top = (uint64_t)(hi32 | (xsbo >> 4)) << 32;        // B: SBO = stride between 8-row (N) groups
    const uint32_t a_lbo = R << 16;                                     // A: LBO = 16 R bytes between the two k-chunks
    const uint32_t b_lbo = (128u >> 4) << 16;                           // B: LBO = 128 B between the two k-groups
    const uint64_t a_hi = a_top | (slot16 | a_lbo);
    const uint64_t a_lo = a_top | ((slot16 + (uint32_t)kbs * 2u * R) | a_lbo);
    const uint64_t b = b_top | (xhi16 | b_lbo);
    const uint32_t idesc1 = idesc(rows), idesc2 = idesc(2u * rows);
    switch (kbs) {
        case 1: issue_blocks<TERMS, 1>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum); break;
        case 2: issue_blocks<TERMS, 2>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum); break;
        case 3: issue_blocks<TERMS, 3>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum); break;
        default: issue_blocks<TERMS, 4>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum); break;
    }
}

// Ring position of the MMA thread (slot index + phase parity), advanced with two instructions per stage.
struct RingPos {
    uint32_t slot, phase;
};

}  // namespace tcs
}  // namespace cadm
