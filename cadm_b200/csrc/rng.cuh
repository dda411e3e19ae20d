// Counter-based noise of the engine: Philox4x32-10 with the counter layout specified in oracle/philox.py
// (the oracle is the written spec; this file is the device implementation the product uses).
// Stands in for tf.random.truncated_normal (cadm/dynamics/core/utils.py:135), tf.random.normal (:90) and
// tf.random.uniform (:195,198) -- TensorFlow's own streams are unseeded in the reference and cannot be reproduced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cadm {

constexpr uint32_t kStreamZ = 1, kStreamEps = 2, kStreamU = 3, kStreamUD = 4;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0;
        k1 += W1;
    }
    return c;
}

__device__ __forceinline__ uint4 philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    return philox4x32_10(make_uint4(c0, c1, c2, c3), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 23-bit uniform strictly inside (0, 1); exact in fp32
__device__ __forceinline__ float u01(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 1.1920928955078125e-07f; }

// N(0,1) truncated to [-2, 2] by inverse CDF
__device__ __forceinline__ float trunc_normal(uint32_t w) {
    float v = 2.0f * u01(w) - 1.0f;
    float z = 1.41421356237309505f * erfinvf(0.9544997361036416f * v);
    return fminf(fmaxf(z, -2.0f), 2.0f);
}

__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& n0, float& n1) {
    float r = sqrtf(-2.0f * logf(u01(wa)));
    float s, c;
    sincospif(2.0f * u01(wb), &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// four N(0,1) values of block j of row `rid` at (it, t)
__device__ __forceinline__ void normal4(uint64_t seed, uint32_t j, uint32_t rid, uint32_t t, uint32_t it, float out[4]) {
    uint4 w = philox(seed, j, rid, t, (it << 8) | kStreamEps);
    box_muller(w.x, w.y, out[0], out[1]);
    box_muller(w.z, w.w, out[2], out[3]);
}

// same stream with MUFU transcendentals (lg2 / sqrt / sin / cos .approx): |error| ~ 1e-6 absolute per normal
__device__ __forceinline__ void box_muller_fast(uint32_t wa, uint32_t wb, float& n0, float& n1) {
    float l, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u01(wa)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * (-2.0f * 0.6931471805599453f)));
    const float th = 6.283185307179586f * (u01(wb) - 0.5f);        // theta - pi, in [-pi, pi)
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    n0 = -r * c;                                                    // cos(theta) = -cos(theta - pi)
    n1 = -r * s;
}
__device__ __forceinline__ void normal4_fast(uint64_t seed, uint32_t j, uint32_t rid, uint32_t t, uint32_t it, float out[4]) {
    uint4 w = philox(seed, j, rid, t, (it << 8) | kStreamEps);
    box_muller_fast(w.x, w.y, out[0], out[1]);
    box_muller_fast(w.z, w.w, out[2], out[3]);
}

__device__ __forceinline__ uint32_t word_of(const uint4& w, int lane) {
    return lane == 0 ? w.x : lane == 1 ? w.y : lane == 2 ? w.z : w.w;
}

}  // namespace cadm
