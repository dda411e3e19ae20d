// The small kernels around the rollout: CEM candidate sampling, particle mean, elite selection + refit,
// context encoder, weight packing.  All are latency-bound helpers; the rollout kernel holds the FLOPs.
#include "common.cuh"
#include "kernels.cuh"
#include "rng.cuh"

namespace cadm {

// ------------------------------------------------------------------------------------------------
// candidate action sequences
// ------------------------------------------------------------------------------------------------
// One value of a CEM candidate: mean + sqrt(constrained var) * z      (core/utils.py:131-135)
__device__ __forceinline__ float cem_action_value(float mean, float var, float z) {
    const float lb = mean - (-1.0f), ub = 1.0f - mean;
    const float cvar = fminf(fminf((lb / 2.0f) * (lb / 2.0f), (ub / 2.0f) * (ub / 2.0f)), var);
    return fmaf(sqrtf(cvar), z, mean);
}


__global__ void sample_actions_kernel(const SampleParams S) {
    griddep_launch();
    griddep_wait();                 // mean / var come from the previous refit, which also still reads the actions written here
    const int nb = (S.hA + 3) / 4;
    const long long total = (long long)S.m * S.n_local * nb;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % nb);
        const long long c = i / nb;
        const int nl = (int)(c % S.n_local);
        const int mi = (int)(c / S.n_local);
        const int ng = S.n_offset + nl;
        uint4 w = make_uint4(0, 0, 0, 0);
        const bool need_rng = (S.mode == 2) ? (S.u_int == nullptr) : (S.z == nullptr);
        if (need_rng) {
            const uint32_t stream = S.mode == 0 ? kStreamZ : (S.mode == 1 ? kStreamU : kStreamUD);
            w = philox(S.seed, (uint32_t)j, (uint32_t)ng, (uint32_t)(mi + S.env_offset), ((uint32_t)S.it << 8) | stream);
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int k = 4 * j + l;
            if (k >= S.hA) break;
            const size_t o = ((size_t)mi * S.n_local + nl) * S.hA + k;
            if (S.mode == 0) {
                const float z = S.z ? S.z[((size_t)mi * S.n_global + ng) * S.hA + k] : trunc_normal(word_of(w, l));
                S.actions[o] = cem_action_value(S.mean[(size_t)mi * S.hA + k], S.var[(size_t)mi * S.hA + k], z);
            } else if (S.mode == 1) {
                S.actions[o] = S.z ? S.z[o] : 2.0f * u01(word_of(w, l)) - 1.0f;
            } else {
                S.actions_int[o] = S.u_int ? S.u_int[o] : (int)(word_of(w, l) % (uint32_t)S.A);
            }
        }
    }
}

cudaError_t launch_sample_actions(const SampleParams& S, cudaStream_t stream) {
    const long long total = (long long)S.m * S.n_local * ((S.hA + 3) / 4);
    const int threads = 256;
    const int blocks = (int)min((total + threads - 1) / threads, (long long)148 * 8);
    return launch_chain(sample_actions_kernel, dim3(max(blocks, 1)), dim3(threads), 0, stream, S);
}

// ------------------------------------------------------------------------------------------------
// mean over particles (core/utils.py:170), fixed summation order -> independent of the sharding
// ------------------------------------------------------------------------------------------------
__global__ void particle_mean_kernel(const float* __restrict__ ret_p, float* __restrict__ out, int count, int p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float* r = ret_p + (size_t)i * p;
    float s = 0.f;
    for (int k = 0; k < p; ++k) s += r[k];
    out[i] = s / (float)p;
}

cudaError_t launch_particle_mean(const float* ret_p, float* out, int count, int p, cudaStream_t stream) {
    particle_mean_kernel<<<(count + 255) / 256, 256, 0, stream>>>(ret_p, out, count, p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// elite selection + refit (core/utils.py:171-182); one CTA per environment
// ------------------------------------------------------------------------------------------------

// ascending-sortable key: larger return first, ties -> lower index first (tf.nn.top_k); NaN sorts last
__device__ __forceinline__ unsigned long long elite_key(float v, int idx) {
    if (v != v) v = -INFINITY;
    if (v == 0.f) v = 0.f;          // -0 == +0
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending order of value
    return ((unsigned long long)(~u) << 32) | (uint32_t)idx;
}

__device__ __forceinline__ float elite_action(const RefitParams& R, int mi, int ni, int k, int hA, float mu0, float var0) {
    if (ni >= R.n_offset && ni < R.n_offset + R.n_local)
        return R.actions[((size_t)mi * R.n_local + (ni - R.n_offset)) * hA + k];
    // not ours: regenerate from the counter-based stream (identical arithmetic on every rank)
    float z;
    if (R.z) z = R.z[((size_t)mi * R.n_global + ni) * hA + k];
    else {
        uint4 w = philox(R.seed, (uint32_t)(k >> 2), (uint32_t)ni, (uint32_t)(mi + R.env_offset), ((uint32_t)R.it << 8) | kStreamZ);
        z = trunc_normal(word_of(w, k & 3));
    }
    return cem_action_value(mu0, var0, z);
}

__global__ void __launch_bounds__(1024, 1) refit_kernel(const RefitParams R) {
    extern __shared__ unsigned long long keys[];      // [npad] then (optionally) the staged elite rows [K][hA] fp32
    __shared__ int s_el[256];
    __shared__ unsigned long long s_sel[256];
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_need, s_count;
    const int mi = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = R.n_global;
    const int K = R.k_elites;
    const int hA = R.h * R.A;
    float* el = reinterpret_cast<float*>(keys + R.npad);
    const float* returns_buf = R.returns_buf;
    griddep_launch();
    griddep_wait();                 // the rollout's particle returns
    if (R.peers != nullptr) {
        // Fused all-gather over peer memory (multi-GPU): this CTA averages the particle returns of ITS environment's local
        // candidates (the fixed summation order of core/utils.py:170, so the value does not depend on the sharding), stores
        // the slice into the returns buffer of EVERY rank (NVLink / NVSwitch stores), publishes the iteration's epoch in flag
        // (rank, environment) of every rank with system-scope release ordering, and then waits until every rank's slice of
        // this environment has landed in OUR buffer.  No separate scatter kernel, no collective call, no host round trip.
        const long long slice = ((long long)R.rank * R.m + mi) * R.n_local;
        for (int i = tid; i < R.n_local; i += blockDim.x) {
            const float* r = R.ret_p_local + ((size_t)mi * R.n_local + i) * R.p;
            float sum = 0.f;
            for (int q = 0; q < R.p; ++q) sum += r[q];
            const float v = sum / (float)R.p;
            for (int g = 0; g < R.world; ++g) reinterpret_cast<float*>(R.peers[g] + R.slice_off)[slice + i] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (tid < R.world) {
            int* flag = reinterpret_cast<int*>(R.peers[tid] + R.flag_off) + R.rank * R.m_max + mi;
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(R.peer_epoch) : "memory");
            // bounded wait (a rank that died must not leave the others spinning for ever): gives up after R.timeout_cycles and
            // reports through the host-mapped word; the host raises on the call that ends this decision
            const int* mine = reinterpret_cast<const int*>(R.peers[R.rank] + R.flag_off) + tid * R.m_max + mi;
            int seen;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
                if (seen - R.peer_epoch < 0 && clock64() - t0 > R.timeout_cycles) {
                    if (R.peer_timeout) *reinterpret_cast<volatile int*>(R.peer_timeout) = 1 + tid;
                    break;
                }
            } while (seen - R.peer_epoch < 0);
        }
        __syncthreads();
        returns_buf = reinterpret_cast<const float*>(R.peers[R.rank] + R.slice_off);
    }
    for (int i = tid; i < R.npad; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < n) {
            float v;
            if (R.ret_p) {                                  // single rank: fold the particle mean (core/utils.py:170) in
                const float* r = R.ret_p + ((size_t)mi * n + i) * R.p;
                float sum = 0.f;
                for (int q = 0; q < R.p; ++q) sum += r[q];
                v = sum / (float)R.p;
            } else {
                const int g = i / R.n_local, nl = i - g * R.n_local;
                v = returns_buf[((size_t)g * R.m + mi) * R.n_local + nl];
            }
            if (R.returns_log) R.returns_log[(size_t)mi * n + i] = v;
            key = elite_key(v, i);
        }
        keys[i] = key;
    }
    __syncthreads();
    if (n <= 256 && !R.mode_rs) {
        // rank by counting: keys are distinct (the index is part of the key), so rank = #smaller keys.  O(n^2 / 32) broadcast
        // loads: 1.4k for n = 200
        if (tid < n) {
            const unsigned long long mine = keys[tid];
            int rank = 0;
#pragma unroll 4
            for (int jx = 0; jx < n; ++jx) rank += keys[jx] < mine ? 1 : 0;
            if (rank < K) s_el[rank] = tid;
        }
    } else {
        // Radix select of the K smallest keys (tf.nn.top_k order: larger return first, ties -> lower index first; keys are
        // distinct, so "the K smallest" is exact): the key's 32 value bits and 16 index bits, most significant byte first, one
        // 256-bin shared-memory histogram per byte over the keys that match the prefix found so far.  6 passes over n keys
        // instead of the log^2(n) compare-exchange rounds of a full sort (n = 1600: 66 rounds with a barrier each).
        if (tid == 0) { s_prefix = 0ull; s_need = (unsigned)K; }
        __syncthreads();
        for (int pass = 0; pass < 6; ++pass) {
            const int shift = pass < 4 ? 56 - 8 * pass : (pass == 4 ? 8 : 0);
            if (tid < 256) s_hist[tid] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = shift + 8 >= 64 ? 0ull : (~0ull << (shift + 8));
            for (int i = tid; i < n; i += blockDim.x) {
                const unsigned long long k = keys[i];
                if ((k & himask) == (prefix & himask)) atomicAdd(&s_hist[(unsigned)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                // warp 0: the digit at which the running count reaches `need` (8 bins per lane, then a warp scan)
                unsigned loc[8], sum = 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) { loc[q] = s_hist[tid * 8 + q]; sum += loc[q]; }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned up = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += up; }
                const unsigned excl = incl - sum, need = s_need;
                if (excl < need && need <= incl) {            // exactly one lane
                    unsigned run = excl;
                    int d = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { if (run < need && need <= run + loc[q]) { d = q; break; } run += loc[q]; }
                    s_prefix = prefix | ((unsigned long long)(tid * 8 + d) << shift);
                    s_need = need - run;
                }
            }
            __syncthreads();
        }
        // s_prefix is now the K-th smallest key (bits 16..31 of a key are zero: indices are below 2^14): gather what is <= it
        if (tid == 0) s_count = 0u;
        __syncthreads();
        const unsigned long long kth = s_prefix;
        for (int i = tid; i < n; i += blockDim.x) {
            const unsigned long long k = keys[i];
            if (k <= kth) { const unsigned pos = atomicAdd(&s_count, 1u); if (pos < 256u) s_sel[pos] = k; }
        }
        __syncthreads();
        if (tid < K) {                                        // order the K selected keys by counting
            const unsigned long long mine = s_sel[tid];
            int rank = 0;
            for (int jx = 0; jx < K; ++jx) rank += s_sel[jx] < mine ? 1 : 0;
            s_el[rank] = (int)(mine & 0xffffffffu);
        }
        if (R.mode_rs) {
            __syncthreads();
            if (tid == 0) R.best[mi] = s_el[0];
            return;
        }
    }
    __syncthreads();
    for (int i = tid; i < K; i += blockDim.x)
        if (R.elites_log) R.elites_log[(size_t)mi * K + i] = s_el[i];
    if (R.stage_elites) {
        // gather the elite sequences, one work item = (elite, block of 4 coordinates): a local elite is copied from this rank's
        // actions; an elite owned by another rank is regenerated from the counter-based stream with ONE Philox call per block
        // (the arithmetic of sample_actions_kernel, so every rank holds the same bits), then reduce from shared memory
        const int nb = (hA + 3) >> 2;
        for (int i = tid; i < K * nb; i += blockDim.x) {
            const int jx = i / nb, b = i - jx * nb;
            const int ni = s_el[jx];
            const bool local = ni >= R.n_offset && ni < R.n_offset + R.n_local;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (!local && !R.z) w = philox(R.seed, (uint32_t)b, (uint32_t)ni, (uint32_t)(mi + R.env_offset), ((uint32_t)R.it << 8) | kStreamZ);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const int k = 4 * b + l;
                if (k >= hA) break;
                float v;
                if (local) {
                    v = R.actions[((size_t)mi * R.n_local + (ni - R.n_offset)) * hA + k];
                } else {
                    const float z = R.z ? R.z[((size_t)mi * R.n_global + ni) * hA + k] : trunc_normal(word_of(w, l));
                    v = cem_action_value(R.mean[(size_t)mi * hA + k], R.var[(size_t)mi * hA + k], z);
                }
                el[jx * hA + k] = v;
            }
        }
        __syncthreads();
    }
    for (int k = tid; k < hA; k += blockDim.x) {
        const float mu0 = R.mean[(size_t)mi * hA + k];
        const float var0 = R.var[(size_t)mi * hA + k];
        float sum = 0.f;
        for (int jx = 0; jx < K; ++jx) sum += R.stage_elites ? el[jx * hA + k] : elite_action(R, mi, s_el[jx], k, hA, mu0, var0);
        const float new_mean = sum / (float)K;
        float sq = 0.f;
        for (int jx = 0; jx < K; ++jx) {
            const float d = (R.stage_elites ? el[jx * hA + k] : elite_action(R, mi, s_el[jx], k, hA, mu0, var0)) - new_mean;
            sq += d * d;
        }
        const float new_var = sq / (float)K;
        R.mean[(size_t)mi * hA + k] = mu0 * R.alpha + (1.0f - R.alpha) * new_mean;
        R.var[(size_t)mi * hA + k] = var0 * R.alpha + (1.0f - R.alpha) * new_var;
    }
}

static int g_refit_smem = 0;
cudaError_t launch_refit(RefitParams R, cudaStream_t stream) {
    size_t smem = (size_t)R.npad * 8;
    const size_t stage = (size_t)R.k_elites * R.h * R.A * 4;
    R.stage_elites = (!R.mode_rs && smem + stage <= 200 * 1024) ? 1 : 0;
    if (R.stage_elites) smem += stage;
    if (smem > 48 * 1024 && (int)smem > g_refit_smem) {
        cudaError_t e = cudaFuncSetAttribute(refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        g_refit_smem = (int)smem;
    }
    const int threads = 1024;      // the elite gather and the per-coordinate reductions want every thread the CTA can have
    return launch_chain(refit_kernel, dim3(R.m), dim3(threads), smem, stream, R);
}

// random shooting: gather the first action of the best candidate (core/utils.py:239-245)
__global__ void rs_gather_kernel(const float* actions, const int* actions_int, const int* best, int n_local, int h, int A,
                                 float* action, int* action_int) {
    const int mi = blockIdx.x;
    const int b = best[mi];
    if (actions_int) {
        if (threadIdx.x == 0 && action_int) action_int[mi] = actions_int[((size_t)mi * n_local + b) * h];
    } else {
        for (int a = threadIdx.x; a < A; a += blockDim.x)
            if (action) action[(size_t)mi * A + a] = actions[(((size_t)mi * n_local + b) * h) * A + a];
    }
}

cudaError_t launch_rs_gather(const float* actions, const int* actions_int, const int* best, int m, int n_local, int h,
                             int A, float* action, int* action_int, cudaStream_t stream) {
    rs_gather_kernel<<<m, 32, 0, stream>>>(actions, actions_int, best, n_local, h, A, action, action_int);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// context encoder (core/utils.py:400-407, 591-617): relu hidden layers, linear output; one CTA per (mi, e)
// ------------------------------------------------------------------------------------------------

// One CTA per (environment, member).  A layer is a mat-vec whose 10^4..10^5 weights are read exactly once, usually from HBM
// (the planner's working set between two decisions does not keep them in L2), so the kernel is a latency problem: what counts
// is how many independent loads are in flight.  Thread t owns FOUR consecutive outputs (one 16-byte load per weight row,
// coalesced: a warp reads 512 contiguous bytes of a row) and every (512 / (out / 4))-th input row; eight rows are loaded
// before the first FMA, i.e. 64 KB in flight per CTA.  The partial sums of the row groups are combined through shared memory
// in a fixed order, so the result does not depend on the launch.  (The first version gave eight lanes one output each and
// walked the rows four at a time: 30 dependent round trips to HBM per layer, 152 us per decision under a cold L2; this one
// takes about a tenth.)  Layers whose width is not a multiple of four (the 10-wide output) take the scalar path.
__global__ void __launch_bounds__(512) encoder_kernel(const EncoderParams Q) {
    extern __shared__ float ebuf[];      // two ping-pong vectors of max width, then the row-group partial sums [nrg][out]
    const int mi = blockIdx.x, e = blockIdx.y;
    int wmax = 0;
    for (int l = 0; l <= Q.n_layers; ++l) wmax = max(wmax, Q.dims[l]);
    float* x = ebuf;
    float* y = ebuf + wmax;
    float* part = ebuf + 2 * wmax;       // 2048 floats
    const int no = Q.D * Q.K, na = Q.A * Q.K;
    for (int i = threadIdx.x; i < no + na; i += blockDim.x) {
        float v;
        if (i < no) v = __fdiv_rn(Q.cp_obs[(size_t)mi * no + i] - Q.cpo_mean[i], Q.cpo_std[i] + 1e-10f);
        else { const int a = i - no; v = __fdiv_rn(Q.cp_act[(size_t)mi * na + a] - Q.cpa_mean[a], Q.cpa_std[a] + 1e-10f); }
        x[i] = v;
    }
    __syncthreads();
    for (int l = 0; l < Q.n_layers; ++l) {
        const int in = Q.dims[l], out = Q.dims[l + 1];
        const float* W = Q.W[l] + (size_t)e * in * out;
        const float* b = Q.b[l] + (size_t)e * out;
        const int ncol4 = out >> 2;
        if ((out & 3) == 0 && ncol4 <= 512 && (512 / ncol4) * out <= 2048 && ((size_t)W & 15) == 0) {
            const int nrg = 512 / ncol4;                         // row groups
            const int cj = threadIdx.x % ncol4, rg = threadIdx.x / ncol4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rg < nrg) {
                const float4* Wv = reinterpret_cast<const float4*>(W) + cj;
                for (int k0 = rg; k0 < in; k0 += 8 * nrg) {
                    float4 w[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = k0 + u * nrg;
                        w[u] = k < in ? __ldg(Wv + (size_t)k * ncol4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = k0 + u * nrg;
                        const float xk = k < in ? x[k] : 0.f;
                        acc.x = fmaf(xk, w[u].x, acc.x); acc.y = fmaf(xk, w[u].y, acc.y);
                        acc.z = fmaf(xk, w[u].z, acc.z); acc.w = fmaf(xk, w[u].w, acc.w);
                    }
                }
                *reinterpret_cast<float4*>(part + rg * out + 4 * cj) = acc;
            }
            __syncthreads();
            for (int j = threadIdx.x; j < out; j += blockDim.x) {
                float sum = 0.f;
                for (int r = 0; r < nrg; ++r) sum += part[r * out + j];
                sum += b[j];
                if (l < Q.n_layers - 1) sum = fmaxf(sum, 0.f);
                y[j] = sum;
            }
        } else {
            constexpr int kSplit = 8;            // lanes per output unit
            const int sub = threadIdx.x % kSplit, grp = threadIdx.x / kSplit, ngrp = blockDim.x / kSplit;
            for (int j0 = 0; j0 < out; j0 += ngrp) {
                const int j = j0 + grp;
                float acc = 0.f;
                if (j < out) {
#pragma unroll 8
                    for (int i = sub; i < in; i += kSplit) acc = fmaf(x[i], __ldg(W + (size_t)i * out + j), acc);
                }
#pragma unroll
                for (int o = kSplit / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (j < out && sub == 0) {
                    acc += b[j];
                    if (l < Q.n_layers - 1) acc = fmaxf(acc, 0.f);
                    y[j] = acc;
                }
            }
        }
        __syncthreads();
        float* t = x; x = y; y = t;
    }
    for (int j = threadIdx.x; j < Q.C; j += blockDim.x) Q.ctx[((size_t)e * Q.m + mi) * Q.C + j] = x[j];
}

cudaError_t launch_encoder(const EncoderParams& Q, cudaStream_t stream) {
    int wmax = 0;
    for (int l = 0; l <= Q.n_layers; ++l) wmax = max(wmax, Q.dims[l]);
    encoder_kernel<<<dim3(Q.m, Q.E), 512, (2 * wmax + 2048) * sizeof(float), stream>>>(Q);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// sampler-side state on the device: what cadm/samplers/sampler.py keeps in NumPy between two
// get_actions() calls -- the warm-start plan prev_sol, the K-step history buffers and their fill counters
// ------------------------------------------------------------------------------------------------
// after a decision (sampler.py:118-120): actions = clip(plan)[:, 0]; prev_sol[:, :-1] = clip(plan)[:, 1:]; prev_sol[:, -1] = 0
// (the clip is get_action's np.clip, mlp_ensemble_cem_dynamics.py:205-206)
__global__ void session_shift_kernel(const float* __restrict__ mean, float* __restrict__ prev_sol, float* __restrict__ action, int h, int A) {
    const int mi = blockIdx.x, hA = h * A;
    for (int k = threadIdx.x; k < hA; k += blockDim.x) {
        const float v = fminf(fmaxf(mean[(size_t)mi * hA + k], -1.0f), 1.0f);
        if (k < A) action[(size_t)mi * A + k] = v;
        else prev_sol[(size_t)mi * hA + k - A] = v;
    }
    __syncthreads();        // every read of row mi precedes the zeroing of its tail (mean and prev_sol may alias)
    for (int k = threadIdx.x; k < A; k += blockDim.x) prev_sol[(size_t)mi * hA + hA - A + k] = 0.f;
}

cudaError_t launch_session_shift(const float* mean, float* prev_sol, float* action, int m, int h, int A, cudaStream_t stream) {
    session_shift_kernel<<<m, 256, 0, stream>>>(mean, prev_sol, action, h, A);
    return cudaGetLastError();
}

// after the environment step (sampler.py:164-195): append (next_obs - obs | obs, action) to the history -- left to right
// while it fills, sliding afterwards -- then, for finished episodes, clear the history, the counter and the warm start
__global__ void session_observe_kernel(const float* __restrict__ obs, const float* __restrict__ next_obs, const float* __restrict__ action,
                                       const unsigned char* __restrict__ done, float* hist_obs, float* hist_act, int* counts,
                                       float* prev_sol, int D, int A, int K, int hA, int state_diff) {
    const int mi = blockIdx.x, tid = threadIdx.x;
    float* ho = hist_obs + (size_t)mi * D * K;
    float* ha = hist_act + (size_t)mi * A * K;
    const int cnt = counts[mi];
    const bool fin = done != nullptr && done[mi] != 0;
    if (!fin) {
        if (cnt >= K) {
            // slide by one entry; each thread moves its own elements from the back of a register copy
            for (int base = 0; base < D * (K - 1); base += blockDim.x) {
                const int i = base + tid;
                const float v = i < D * (K - 1) ? ho[i + D] : 0.f;
                __syncthreads();
                if (i < D * (K - 1)) ho[i] = v;
                __syncthreads();
            }
            for (int base = 0; base < A * (K - 1); base += blockDim.x) {
                const int i = base + tid;
                const float v = i < A * (K - 1) ? ha[i + A] : 0.f;
                __syncthreads();
                if (i < A * (K - 1)) ha[i] = v;
                __syncthreads();
            }
        }
        const int slot = cnt < K ? cnt : K - 1;
        for (int d = tid; d < D; d += blockDim.x) {
            const float o = obs[(size_t)mi * D + d];
            const float nx = next_obs[(size_t)mi * D + d];
            ho[slot * D + d] = state_diff == 2 ? nx : (state_diff ? nx - o : o);
        }
        for (int a = tid; a < A; a += blockDim.x) ha[slot * A + a] = action[(size_t)mi * A + a];
        if (tid == 0) counts[mi] = cnt + 1;
    } else {
        for (int i = tid; i < D * K; i += blockDim.x) ho[i] = 0.f;
        for (int i = tid; i < A * K; i += blockDim.x) ha[i] = 0.f;
        for (int i = tid; i < hA; i += blockDim.x) prev_sol[(size_t)mi * hA + i] = 0.f;
        if (tid == 0) counts[mi] = 0;
    }
}

cudaError_t launch_session_observe(const float* obs, const float* next_obs, const float* action, const unsigned char* done,
                                   float* hist_obs, float* hist_act, int* counts, float* prev_sol, int m, int D, int A, int K, int hA,
                                   int state_diff, cudaStream_t stream) {
    session_observe_kernel<<<m, 256, 0, stream>>>(obs, next_obs, action, done, hist_obs, hist_act, counts, prev_sol, D, A, K, hA, state_diff);
    return cudaGetLastError();
}

__global__ void session_reset_kernel(const unsigned char* __restrict__ mask, float* hist_obs, float* hist_act, int* counts,
                                     float* prev_sol, int DK, int AK, int hA) {
    const int mi = blockIdx.x, tid = threadIdx.x;
    if (mask != nullptr && mask[mi] == 0) return;
    for (int i = tid; i < DK; i += blockDim.x) hist_obs[(size_t)mi * DK + i] = 0.f;
    for (int i = tid; i < AK; i += blockDim.x) hist_act[(size_t)mi * AK + i] = 0.f;
    for (int i = tid; i < hA; i += blockDim.x) prev_sol[(size_t)mi * hA + i] = 0.f;
    if (tid == 0) counts[mi] = 0;
}

cudaError_t launch_session_reset(const unsigned char* mask, float* hist_obs, float* hist_act, int* counts, float* prev_sol, int m,
                                 int DK, int AK, int hA, cudaStream_t stream) {
    session_reset_kernel<<<m, 256, 0, stream>>>(mask, hist_obs, hist_act, counts, prev_sol, DK, AK, hA);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// weight packing for the fp32 path: [E, in, out] -> zero-padded k-major image
// ------------------------------------------------------------------------------------------------
__global__ void pack_f32_kernel(float* dst, const float* src, int E, int in, int out, int Kp, int Np, int col0,
                                long long member_stride, long long layer_off, int clear) {
    const long long total = (long long)E * Kp * Np;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Np);
        const int k = (int)((i / Np) % Kp);
        const int e = (int)(i / ((long long)Np * Kp));
        float* d = dst + e * member_stride + layer_off + (long long)k * Np + c;
        const int cs = c - col0;
        if (cs >= 0 && cs < out && k < in) *d = src[((size_t)e * in + k) * out + cs];
        else if (clear) *d = 0.f;
    }
}

cudaError_t launch_pack_f32(float* dst, const float* src, int E, int in, int out, int Kp, int Np, int col0,
                            long long member_stride, long long layer_off, int clear, cudaStream_t stream) {
    const long long total = (long long)E * Kp * Np;
    pack_f32_kernel<<<(int)min((total + 255) / 256, (long long)1184), 256, 0, stream>>>(dst, src, E, in, out, Kp, Np, col0,
                                                                                        member_stride, layer_off, clear);
    return cudaGetLastError();
}

__global__ void pack_bias_kernel(float* dst, const float* src, int E, int out, int col0, long long bias_stride, long long off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * out) return;
    const int e = i / out, c = i - e * out;
    dst[e * bias_stride + off + col0 + c] = src[(size_t)e * out + c];
}

cudaError_t launch_pack_bias(float* dst, const float* src, int E, int out, int col0, long long bias_stride, long long off,
                             cudaStream_t stream) {
    pack_bias_kernel<<<(E * out + 255) / 256, 256, 0, stream>>>(dst, src, E, out, col0, bias_stride, off);
    return cudaGetLastError();
}

}  // namespace cadm
