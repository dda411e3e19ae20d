// Persistent fp32 rollout kernel (the exact-fp32 path, CADM_PREC_FP32).
//
// One CTA owns a tile of TR rows of ONE ensemble member and carries them through all h horizon steps of
// one CEM iteration: state, return accumulator and the activations never leave the SM.  Per step it fuses
//   input assembly (obs_preproc + normalize, action normalize, context concat)   core/utils.py:141-156 / 442-460
//   n_hidden swish layers + the mu / logvar heads                                core/utils.py:73-77, 635-647
//   denormalize, bounded logvar, Gaussian sample                                 core/utils.py:79-90
//   obs_postproc, reward, return accumulation                                    core/utils.py:158-168
// The tile/transpose/reshape/concat ops of the TF graph are index arithmetic here (never materialised).
//
// Weights are streamed every step from L2 through a 4-stage shared-memory ring by a producer warp using 1-D
// bulk async copies (TMA engine, mbarrier completion); the packed weight image of a member is contiguous, so
// the stream is a linear walk.  8 compute warps do register-tiled FFMA: thread tile (TR/8) rows x 8 columns,
// activations stored [k][row] so one LDS.128 feeds 4 rows.
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace cadm {

constexpr int kStages = 4;
constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;
constexpr int kStageFloats = kChunkK * kMaxHidden;   // 8 x 208

struct SmemLayoutF32 {
    int x_floats, s_floats, bias_floats;
    size_t off_xa, off_xb, off_w, off_s, off_bias, off_vec, off_rowi, off_bar, total;
};

__host__ __device__ inline SmemLayoutF32 smem_layout_f32(int TR, int D, int n_hidden, int Hp, int NHp) {
    SmemLayoutF32 L;
    int kmax = kMaxHidden;
    L.x_floats = kmax * TR;
    L.s_floats = TR * (D + 1);
    L.bias_floats = n_hidden * Hp + NHp;
    size_t o = 0;
    L.off_xa = o; o += (size_t)L.x_floats * 4;
    L.off_xb = o; o += (size_t)L.x_floats * 4;
    L.off_w = o; o += (size_t)kStages * kStageFloats * 4;
    L.off_s = o; o += (size_t)round_up(L.s_floats, 4) * 4;
    L.off_bias = o; o += (size_t)round_up(L.bias_floats, 4) * 4;
    L.off_vec = o; o += (size_t)(2 * kMaxObs + 2 * kMaxAct + 5 * kMaxObs) * 4;   // norm vectors
    L.off_rowi = o; o += (size_t)TR * 6 * 4;
    o = (o + 7) / 8 * 8;
    L.off_bar = o; o += (size_t)2 * kStages * 8;
    L.total = o;
    return L;
}

struct Pipe {
    int stage;
    uint32_t phase;
    __device__ __forceinline__ void advance() {
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
};

// acc[RT][CT] += X[k][rows] * W[k][cols] over `nchunks` streamed chunks of kChunkK k-rows.
template <int TR, int RT, int CT>
__device__ __forceinline__ void gemm_stream(float (&acc)[RT][CT], const float* __restrict__ X, const float* Wst,
                                            uint64_t* full, uint64_t* empty, Pipe& pipe, int nchunks, int Np,
                                            int row0, int col0, bool active, int lane) {
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        ptx::mbar_wait(&full[pipe.stage], pipe.phase);
        if (active) {
            const float* Ws = Wst + pipe.stage * kStageFloats + col0;
            const float* Xk = X + (size_t)c * kChunkK * TR + row0;
#pragma unroll
            for (int kk = 0; kk < kChunkK; ++kk) {
                float a[RT], w[CT];
#pragma unroll
                for (int r = 0; r < RT; r += 4) *reinterpret_cast<float4*>(&a[r]) = *reinterpret_cast<const float4*>(Xk + kk * TR + r);
#pragma unroll
                for (int j = 0; j < CT; j += 4) *reinterpret_cast<float4*>(&w[j]) = *reinterpret_cast<const float4*>(Ws + kk * Np + j);
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int j = 0; j < CT; ++j) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[pipe.stage]);
        pipe.advance();
    }
}

template <int TR>
__global__ void __launch_bounds__(kThreads, 1) rollout_f32_kernel(const __grid_constant__ RolloutParams P) {
    griddep_launch();
    griddep_wait();            // programmatic dependent launch (common.cuh): the candidates' actions come from the kernel before
    constexpr int RT = TR / 8;   // rows per thread
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayoutF32 L = smem_layout_f32(TR, P.D, P.n_hidden, P.Hp, P.NHp);
    float* Xa = reinterpret_cast<float*>(smem + L.off_xa);
    float* Xb = reinterpret_cast<float*>(smem + L.off_xb);
    float* Wst = reinterpret_cast<float*>(smem + L.off_w);
    float* S = reinterpret_cast<float*>(smem + L.off_s);
    float* bias = reinterpret_cast<float*>(smem + L.off_bias);
    float* vec = reinterpret_cast<float*>(smem + L.off_vec);
    int* rowi = reinterpret_cast<int*>(smem + L.off_rowi);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* empty = full + kStages;

    float* v_obs_mean = vec;                       // [P]
    float* v_obs_den = vec + kMaxObs;              // [P]   std + 1e-10
    float* v_act_mean = vec + 2 * kMaxObs;         // [A]
    float* v_act_den = v_act_mean + kMaxAct;       // [A]
    float* v_dmean = v_act_den + kMaxAct;          // [D]
    float* v_dscale = v_dmean + kMaxObs;           // [D]   std + 1e-10
    float* v_2logstd = v_dscale + kMaxObs;         // [D]   2 log(std)
    float* v_maxlv = v_2logstd + kMaxObs;
    float* v_minlv = v_maxlv + kMaxObs;
    int* r_mi = rowi;            // env index (predict: unused)
    int* r_src = rowi + TR;      // planner: mi*n_local+nl ; predict: e*B+b   (index of the action / obs row)
    int* r_pi = rowi + 2 * TR;   // particle
    int* r_ctx = rowi + 3 * TR;  // entry index into ctx (x C)
    int* r_rid = rowi + 4 * TR;  // RNG row id
    int* r_eps = rowi + 5 * TR;  // row index into the injected eps array (x D), without the t term

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int e = blockIdx.y;
    const int tile_row0 = blockIdx.x * P.rows_per_cta;
    const int nrows = min(P.rows_per_cta, P.rows_per_member - tile_row0);
    if (nrows <= 0) return;

    const int chunks0 = P.Kp0 / kChunkK;
    const int chunksH = P.Hp / kChunkK;
    const int chunks_per_step = chunks0 + (P.n_hidden - 1) * chunksH + chunksH;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], kComputeThreads / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    // =========================== producer warp: stream the member's weight image ===================
    if (warp == kComputeThreads / 32) {
        if (lane == 0) {
            const float* wsrc = P.wpack + (size_t)e * P.member_stride;
            Pipe pp{0, 0};
            for (int t = 0; t < P.h; ++t) {
                size_t off = 0;
                for (int c = 0; c < chunks_per_step; ++c) {
                    const bool is_head = c >= chunks_per_step - chunksH;
                    const uint32_t nfl = kChunkK * (is_head ? P.NHp : P.Hp);
                    ptx::mbar_wait(&empty[pp.stage], pp.phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(&full[pp.stage], nfl * 4u);
                    ptx::bulk_g2s(Wst + pp.stage * kStageFloats, wsrc + off, nfl * 4u, &full[pp.stage]);
                    off += nfl;
                    pp.advance();
                }
            }
        }
        return;
    }

    // =========================== compute warps =====================================================
    // ---- one-time setup: norm vectors, biases, row descriptors, initial state
    for (int i = tid; i < P.P; i += kComputeThreads) {
        v_obs_mean[i] = P.obs_mean[i];
        v_obs_den[i] = P.obs_std[i] + 1e-10f;
    }
    for (int i = tid; i < P.A; i += kComputeThreads) {
        v_act_mean[i] = P.act_mean[i];
        v_act_den[i] = P.act_std[i] + 1e-10f;
    }
    for (int i = tid; i < P.D; i += kComputeThreads) {
        v_dmean[i] = P.delta_mean[i];
        v_dscale[i] = P.delta_std[i] + 1e-10f;
        v_2logstd[i] = 2.0f * logf(P.delta_std[i]);
        v_maxlv[i] = P.max_lv[i];
        v_minlv[i] = P.min_lv[i];
    }
    for (int i = tid; i < L.bias_floats; i += kComputeThreads) bias[i] = P.bpack[(size_t)e * P.bias_stride + i];
    for (int r = tid; r < TR; r += kComputeThreads) {
        int mi = 0, src = 0, pi = 0, cidx = 0, rid = 0, er = 0;
        if (r < nrows) {
            const int rl = tile_row0 + r;
            if (P.row_mode == kRowsPlanner) {
                int nl;
                planner_row(P, e, rl, mi, nl, pi);
                src = mi * P.n_local + nl;
                const int ng = P.n_offset + nl;
                rid = ((mi + P.env_offset) * P.n_global + ng) * P.p + pi;
                cidx = P.ctx_mode ? planner_ctx_index(P, e, mi, pi) : 0;
                const int jq = pi - e * P.q;
                er = e * (P.q * P.m * P.n_global) + (jq * P.m + mi) * P.n_global + ng;
            } else {
                src = e * P.rows_per_member + rl;
                rid = src;
                cidx = src;
                er = src;
            }
        }
        r_mi[r] = mi; r_src[r] = src; r_pi[r] = pi; r_ctx[r] = cidx; r_rid[r] = rid; r_eps[r] = er;
    }
    ptx::bar_sync(1, kComputeThreads);
    for (int i = tid; i < TR * P.D; i += kComputeThreads) {
        const int r = i / P.D, d = i - r * P.D;
        float v = 0.f;
        if (r < nrows) v = (P.row_mode == kRowsPlanner) ? P.obs0[r_mi[r] * P.D + d] : P.obs0[(size_t)r_src[r] * P.D + d];
        S[r * (P.D + 1) + d] = v;
    }
    ptx::bar_sync(1, kComputeThreads);

    // thread -> tile coordinates
    const int rg = tid & 7;                   // row group: rows rg*RT ..
    const int cg = tid >> 3;                  // hidden: 8-column group ; head: 4-column group
    const int row0 = rg * RT;
    const bool act_hidden = cg < P.Hp / 8;
    const bool act_head = cg < P.NHp / 4;
    const size_t eps_step_stride = (size_t)P.E * P.q * P.m * P.n_global * P.D;   // planner: floats per horizon step

    Pipe pipe{0, 0};
    float ret = 0.f;                          // return accumulator of row `tid` (tid < TR)
    const int A = P.A;

#pragma unroll 1
    for (int t = 0; t < P.h; ++t) {
        // ---------------- phase A: reward of the current state, then the MLP input -------------
        if (tid < TR && tid < nrows) {
            const float* s = S + tid * (P.D + 1);
            if (env_reward_reads_next(P.env_id)) {
                if (t > 0) ret += env_reward_next(P.env_id, s);
            } else {
                float a[kMaxAct];
                const float* ap = P.actions + ((size_t)r_src[tid] * P.h + t) * A;
#pragma unroll 1
                for (int i = 0; i < A; ++i) a[i] = __ldg(ap + i);
                ret += env_reward_current(P.env_id, s, a, A, P.max_torque);
            }
        }
        for (int i = tid; i < TR * P.Kp0; i += kComputeThreads) {
            const int k = i / TR, r = i - k * TR;
            float v = 0.f;
            if (r < nrows && k < P.In) {
                if (k < P.P) {
                    v = __fdiv_rn(env_preproc(P.env_id, S + r * (P.D + 1), k) - v_obs_mean[k], v_obs_den[k]);
                } else if (k < P.P + A) {
                    const int ai = k - P.P;
                    if (P.discrete) {
                        const int ida = P.row_mode == kRowsPlanner ? __ldg(P.actions_int + (size_t)r_src[r] * P.h + t) : -1;
                        v = P.row_mode == kRowsPlanner ? (ida == ai ? 1.f : 0.f)
                                                       : __ldg(P.actions + (size_t)r_src[r] * A + ai);
                    } else {
                        const float av = __ldg(P.actions + ((size_t)r_src[r] * P.h + t) * A + ai);
                        v = __fdiv_rn(av - v_act_mean[ai], v_act_den[ai]);
                    }
                } else {
                    v = __ldg(P.ctx + (size_t)r_ctx[r] * P.C + (k - P.P - A));
                }
            }
            Xa[k * TR + r] = v;
        }
        ptx::bar_sync(1, kComputeThreads);

        // ---------------- hidden layers ---------------------------------------------------------
        float* cur = Xa;
        float* oth = Xb;
#pragma unroll 1
        for (int l = 0; l < P.n_hidden; ++l) {
            float acc[RT][8];
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
            gemm_stream<TR, RT, 8>(acc, cur, Wst, full, empty, pipe, l == 0 ? chunks0 : chunksH, P.Hp, row0, cg * 8,
                                   act_hidden, lane);
            if (act_hidden) {
                const float* bl = bias + l * P.Hp + cg * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float o[RT];
#pragma unroll
                    for (int r = 0; r < RT; ++r) o[r] = swishf(acc[r][j] + bl[j]);
#pragma unroll
                    for (int r = 0; r < RT; r += 4)
                        *reinterpret_cast<float4*>(oth + (cg * 8 + j) * TR + row0 + r) = *reinterpret_cast<float4*>(&o[r]);
                }
            }
            ptx::bar_sync(1, kComputeThreads);
            float* tmp = cur; cur = oth; oth = tmp;
        }

        // ---------------- heads: [mu | logvar] -> oth as Hd[c][row] -----------------------------
        {
            float acc[RT][4];
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
            gemm_stream<TR, RT, 4>(acc, cur, Wst, full, empty, pipe, chunksH, P.NHp, row0, cg * 4, act_head, lane);
            if (act_head) {
                const float* bl = bias + P.n_hidden * P.Hp + cg * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float o[RT];
#pragma unroll
                    for (int r = 0; r < RT; ++r) o[r] = acc[r][j] + bl[j];
#pragma unroll
                    for (int r = 0; r < RT; r += 4)
                        *reinterpret_cast<float4*>(oth + (cg * 4 + j) * TR + row0 + r) = *reinterpret_cast<float4*>(&o[r]);
                }
            }
        }
        ptx::bar_sync(1, kComputeThreads);

        // ---------------- phase Z: sample, next state ------------------------------------------
        {
            const float* Hd = oth;
            const int D = P.D;
            const int nblk = (D + 3) / 4;
            for (int item = tid; item < TR * nblk; item += kComputeThreads) {
                const int j = item / TR, r = item - j * TR;
                if (r >= nrows) continue;
                float nz[4] = {0.f, 0.f, 0.f, 0.f};
                if (!P.deterministic) {
                    if (P.eps != nullptr) {
                        const float* ep = P.eps + (P.row_mode == kRowsPlanner ? (size_t)t * eps_step_stride : 0) +
                                          (size_t)r_eps[r] * D + 4 * j;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (4 * j + i < D) nz[i] = __ldg(ep + i);
                    } else {
                        normal4(P.seed, (uint32_t)j, (uint32_t)r_rid[r], (uint32_t)t, (uint32_t)P.it, nz);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int d = 4 * j + i;
                    if (d >= D) break;
                    const float mu = Hd[d * TR + r];
                    float lv = Hd[(D + d) * TR + r];
                    const float dmu = mu * v_dscale[d] + v_dmean[d];
                    float delta = dmu;
                    if (!P.deterministic) {
                        lv = bounded_logvar(lv, v_maxlv[d], v_minlv[d]);
                        const float sd = expf((lv + v_2logstd[d]) / 2.0f);
                        delta = dmu + nz[i] * sd;
                    }
                    float* sp = S + r * (D + 1) + d;
                    const float sn = env_postproc(P.env_id, *sp, delta, d);
                    *sp = sn;
                    if (P.row_mode == kRowsPlanner) {
                        if (P.states != nullptr)
                            P.states[(((size_t)t * P.m * P.n_local + r_src[r]) * P.p + r_pi[r]) * D + d] = sn;
                    } else {
                        const size_t o = (size_t)r_src[r] * D + d;
                        if (P.next_obs) P.next_obs[o] = sn;
                        if (P.mu_out) P.mu_out[o] = mu;
                        if (P.lv_out) P.lv_out[o] = lv;
                    }
                }
            }
        }
        ptx::bar_sync(1, kComputeThreads);
    }

    if (P.row_mode == kRowsPlanner && tid < TR && tid < nrows) {
        if (env_reward_reads_next(P.env_id)) ret += env_reward_next(P.env_id, S + tid * (P.D + 1));
        P.ret_p[(size_t)r_src[tid] * P.p + r_pi[tid]] = ret;
    }
}

// ------------------------------------------------------------------------------------------------
static int g_smem_set[2] = {0, 0};

static int pick_tile_rows(const RolloutParams& P, int num_sms) {
    // 64-row tiles halve the weight traffic per row; use them once they still fill the machine
    const long long tiles64 = (long long)P.E * ((P.rows_per_member + 63) / 64);
    return tiles64 >= num_sms ? 64 : 32;
}

cudaError_t launch_rollout_f32(RolloutParams P, int num_sms, cudaStream_t stream, const char** name) {
    const int TR = pick_tile_rows(P, num_sms);
    // balance the rows over the tiles of one member
    int tiles = (P.rows_per_member + TR - 1) / TR;
    if (TR == 32) {
        // try to use every SM once: E * tiles <= num_sms while keeping <= 32 rows per tile
        int want = num_sms / P.E;
        if (want > tiles) {
            int cand = want;
            if ((P.rows_per_member + cand - 1) / cand >= 8) tiles = cand;   // do not shrink tiles below 8 rows
        }
    }
    P.rows_per_cta = (P.rows_per_member + tiles - 1) / tiles;
    tiles = (P.rows_per_member + P.rows_per_cta - 1) / P.rows_per_cta;
    const SmemLayoutF32 L = smem_layout_f32(TR, P.D, P.n_hidden, P.Hp, P.NHp);
    dim3 grid(tiles, P.E), block(kThreads);
    cudaError_t err;
    if (TR == 64) {
        if (!g_smem_set[1]) {
            err = cudaFuncSetAttribute(rollout_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (err != cudaSuccess) return err;
            g_smem_set[1] = 1;
        }
        if (name) *name = "rollout_f32_kernel<64>";
        err = launch_chain(rollout_f32_kernel<64>, dim3(grid), dim3(block), L.total, stream, P);
    } else {
        if (!g_smem_set[0]) {
            err = cudaFuncSetAttribute(rollout_f32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (err != cudaSuccess) return err;
            g_smem_set[0] = 1;
        }
        if (name) *name = "rollout_f32_kernel<32>";
        err = launch_chain(rollout_f32_kernel<32>, dim3(grid), dim3(block), L.total, stream, P);
    }
    return err != cudaSuccess ? err : cudaGetLastError();
}

}  // namespace cadm
