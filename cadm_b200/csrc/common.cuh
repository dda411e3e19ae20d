// Shared definitions of the cadm_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cadm_b200.h"

namespace cadm {

constexpr int kMaxObs = 48;      // D  <= 48 (HalfCheetah 18, Ant 28, SlimHumanoid 45)
constexpr int kMaxAct = 20;      // A  <= 20
constexpr int kMaxCtx = 16;      // C  <= 16
constexpr int kMaxHidden = 208;  // padded hidden width of the fp32 path
constexpr int kChunkK = 8;       // k-rows per streamed weight chunk

enum RowMode : int {
    kRowsPlanner = 0,   // rows = (mi, nl, pi) of a CEM / RS rollout
    kRowsPredict = 1,   // rows = (e, b) of the training-graph layout [E, B, .]
};

// Everything the rollout kernels need; passed by value (__grid_constant__).
struct RolloutParams {
    // problem dims
    int env_id, D, P, A, C, In;
    int H, n_hidden;
    int E, p, q;              // q = p / E
    int n_local, n_global, n_offset;
    int m, h;
    int env_offset;           // index of environment 0 in an environment-sharded decision (enters the Philox counters only)
    int deterministic, discrete;
    int it;                   // CEM iteration (context pairing + RNG counter)
    int ctx_mode;             // 0 none, 1 reference pairing, 2 matched pairing, 3 per-row (predict)
    int row_mode;
    int rows_per_member;      // planner: q*m*n_local ; predict: B
    int rows_per_cta;
    int row_lo, row_hi;       // the rows of every member THIS launch covers, [row_lo, row_hi); row_hi == 0: all of them.  The engine
                              // may split a batch between two kernels (full waves of 128-row tiles + a remainder of small tiles)
    float max_torque;
    // padded dims of the packed fp32 weights
    int Kp0;                  // In rounded up to kChunkK
    int Hp;                   // H rounded up to 8
    int NHp;                  // 2*D rounded up to 8
    long long member_stride;  // floats per member in wpack
    long long bias_stride;    // floats per member in bpack
    // inputs
    const float* wpack;       // [E][ layer0 Kp0 x Hp | (n_hidden-1) x H x Hp | head H x NHp ]  (k-major rows)
    const float* bpack;       // [E][ n_hidden x Hp | NHp ]
    const float* obs0;        // planner [m, D] ; predict [E, B, D]
    const float* actions;     // planner [m, n_local, h, A] ; predict [E, B, A]
    const int*   actions_int; // discrete planner [m, n_local, h]
    const float* ctx;         // planner: ctx_raw [E, m, C] ; predict: [E, B, C]
    const float* eps;         // nullable; planner [h, E, R_global, D] ; predict [E, B, D]
    const float* obs_mean; const float* obs_std;       // [P]
    const float* act_mean; const float* act_std;       // [A]
    const float* delta_mean; const float* delta_std;   // [D]
    const float* max_lv; const float* min_lv;          // [D]
    unsigned long long seed;
    // outputs
    float* ret_p;             // planner [m, n_local, p]
    float* states;            // nullable; planner [h, m, n_local, p, D]
    float* next_obs;          // predict [E, B, D] (nullable)
    float* mu_out;            // predict (nullable)
    float* lv_out;            // predict (nullable)
};

__host__ __device__ inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The three kernels of a CEM iteration (sample_actions -> rollout -> refit) are short
// next to their launch latency and strictly ordered; launched with cudaLaunchAttributeProgrammaticStreamSerialization the next
// kernel's CTAs become resident -- and run whatever does not depend on its predecessor: barrier / TMEM set-up, bias and
// feature tables, the first weight stages -- while the predecessor is still running, and block in griddep_wait() until the
// predecessor's grid has completed and its writes are visible.  Rule for every kernel of the chain: nothing the predecessor
// writes is read, and nothing it reads is written, before griddep_wait().  Without the launch attribute both calls are no-ops.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern bool g_pdl;      // engine.cu: CADM_PDL=0 or cadm_set_option("pdl", 0) turns the attribute off (A/B, debugging)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// environment closures (see include/cadm_b200.h CADM_ENV_* for the reference lines)
// ---------------------------------------------------------------------------------------------

// processed-observation element k of state s (obs_preproc)
__device__ __forceinline__ float env_preproc(int env_id, const float* s, int k) {
    switch (env_id) {
        case CADM_ENV_HALFCHEETAH:
            if (k == 0) return s[1];
            if (k == 1) return sinf(s[2]);
            if (k == 2) return cosf(s[2]);
            return s[k];
        case CADM_ENV_ANT:
            return s[k + 1];
        default:
            return s[k];
    }
}

// next-state element d from state s and predicted delta (obs_postproc)
__device__ __forceinline__ float env_postproc(int env_id, float s_d, float delta_d, int d) {
    if (env_id == CADM_ENV_HALFCHEETAH || env_id == CADM_ENV_ANT) return d == 0 ? delta_d : s_d + delta_d;
    return s_d + delta_d;
}

__device__ __forceinline__ bool env_reward_reads_next(int env_id) { return env_id == CADM_ENV_CARTPOLE; }

// reward terms that read the CURRENT observation and the action (tf_reward_fn; quirk Q4)
__device__ __forceinline__ float env_reward_current(int env_id, const float* s, const float* a, int A, float max_torque) {
    float e = 0.f;
    switch (env_id) {
        case CADM_ENV_HALFCHEETAH:
            for (int i = 0; i < A; ++i) e += a[i] * a[i];
            return s[0] - 0.1f * e;
        case CADM_ENV_ANT:
            for (int i = 0; i < A; ++i) e += a[i] * a[i];
            return ((s[0] + (-0.005f) * e) + 0.0f) + 0.05f;
        case CADM_ENV_SLIM_HUMANOID: {
            for (int i = 0; i < A; ++i) e += a[i] * a[i];
            float alive = (s[1] > 1.0f && s[1] < 2.0f) ? 5.0f : 0.0f;
            return (((0.25f / 0.015f) * s[22] - 0.1f * e) - 0.0f) + alive;
        }
        case CADM_ENV_PENDULUM: {
            const float pi = 3.14159265358979323846f;
            float theta = atan2f(s[1], s[0]);
            float x = theta + pi;
            float y = 2.0f * pi;
            float r = fmodf(x, y);
            if (r < 0.f) r += y;                          // floored modulo (python / tf `%`)
            float tn = r - pi;
            float tq = fminf(fmaxf(a[0], -max_torque), max_torque);
            return -((tn * tn + 0.1f * (s[2] * s[2])) + 0.001f * (tq * tq));
        }
        default:
            return 0.f;
    }
}

// reward terms that read the NEXT observation (CartPole)
__device__ __forceinline__ float env_reward_next(int env_id, const float* sn) {
    if (env_id == CADM_ENV_CARTPOLE) {
        const float x_thr = 2.4f;
        const float th_thr = (float)(12.0 * 2.0 * 3.14159265358979323846 / 360.0);
        float cond = (sn[0] > x_thr ? 1.f : 0.f) + (sn[0] < -x_thr ? 1.f : 0.f) + (sn[2] > th_thr ? 1.f : 0.f) +
                     (sn[2] < -th_thr ? 1.f : 0.f);
        return 1.0f - cond * 1.0f;
    }
    return 0.f;
}

// ---------------------------------------------------------------------------------------------
// scalar math of forward() (cadm/dynamics/core/utils.py:73-92)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float swishf(float x) { return x * (1.0f / (1.0f + expf(-x))); }

__device__ __forceinline__ float softplusf(float x) {   // log(1 + e^x), overflow-safe
    return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float bounded_logvar(float lv, float max_lv, float min_lv) {
    lv = max_lv - softplusf(max_lv - lv);                // utils.py:84
    lv = min_lv + softplusf(lv - min_lv);                // utils.py:85
    return lv;
}

// MUFU-based variants for the tensor-core path (ex2.approx / lg2.approx: ~2^-22 relative; the 1e-4 bar has 2 orders of
// headroom and the fp32 path keeps the libdevice versions)
__device__ __forceinline__ float fast_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float fast_log(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.6931471805599453f;
}
__device__ __forceinline__ float fast_softplus(float x) { return fmaxf(x, 0.f) + fast_log(1.0f + fast_exp(-fabsf(x))); }
__device__ __forceinline__ float fast_bounded_logvar(float lv, float max_lv, float min_lv) {
    lv = max_lv - fast_softplus(max_lv - lv);
    lv = min_lv + fast_softplus(lv - min_lv);
    return lv;
}

// planner row -> (mi, nl, pi); rl indexes the rows of member e in the order j*m*n_local + mi*n_local + nl
__device__ __forceinline__ void planner_row(const RolloutParams& P, int e, int rl, int& mi, int& nl, int& pi) {
    int mn = P.m * P.n_local;
    int j = rl / mn;
    int rem = rl - j * mn;
    mi = rem / P.n_local;
    nl = rem - mi * P.n_local;
    pi = e * P.q + j;
}

// which [e', m'] entry of ctx_raw a planner row reads (quirks Q2 / Q3 of cadm/dynamics/core/utils.py:433-439)
__device__ __forceinline__ int planner_ctx_index(const RolloutParams& P, int e, int mi, int pi) {
    if (P.ctx_mode == 2) return e * P.m + mi;                       // matched
    int ce = pi % P.E;
    if ((P.it & 1) == 0) return ce * P.m + mi;                      // transposed view [m, E] of [E, m]
    return mi * P.E + ce;                                           // memory reinterpretation on odd iterations
}

}  // namespace cadm
