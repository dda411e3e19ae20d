// C ABI of the engine (include/cadm_b200.h): handle, device buffers, weight packing, launch sequencing.
// No torch types, no exceptions across the boundary, no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

using namespace cadm;

namespace cadm {
// programmatic dependent launch of the sample -> rollout -> refit chain (common.cuh); CADM_PDL=0 or option "pdl" turn it off
bool g_pdl = [] { const char* v = getenv("CADM_PDL"); return !(v && v[0] == '0'); }();
}  // namespace cadm

namespace {

thread_local std::string g_create_error;

struct Engine {
    CadmConfig cfg{};
    int device = 0;
    int num_sms = 148;
    std::string err;
    // derived dims
    int In = 0, Kp0 = 0, Hp = 0, NHp = 0, n_local = 0, n_offset = 0, q = 0, hA = 0;
    long long member_stride = 0, bias_stride = 0;
    // parameters
    float* wpack = nullptr;
    float* bpack = nullptr;
    // tensor-core image (hi/lo split, UMMA layout) and its 16-padded biases
    unsigned char* wimg = nullptr;
    unsigned char* wimg_s = nullptr;   // same weights packed for the swapped-operand kernel (rollout_tcs.cu)
    int tc_variant = 0;                // 0 auto, 1 row tiles (rollout_tc.cu), 2 swapped operands (rollout_tcs.cu), 3 CTA pairs (rollout_tcp.cu)
    int tcs_rows = 0;                  // rows per tile of the swapped kernel (0 = pick)
    bool trace = false;                // clock64 phase trace of CTA 0 (perturbs that CTA: off unless asked for)
    int tcs_skew = 0;                  // start-delay step (cycles) that de-phases the CTAs of the swapped kernel
    int env_offset = 0;                // first environment of this engine in an environment-sharded decision (Philox counters)
    int tcs_kps = 4;                   // K16 blocks per weight stage the swapped image is packed with
    float* bpack_tc = nullptr;
    int Np16 = 0, NHp16 = 0, nkb0 = 0, nkbH = 0;
    long long wimg_member_stride = 0, bias_stride_tc = 0;
    int precision = CADM_PREC_FP32;
    long long* dbg = nullptr;        // clock64 trace of the tensor-core kernel (diagnostic)
    float* max_lv = nullptr;
    float* min_lv = nullptr;
    bool have_weights = false, have_encoder = false, have_norm = false;
    int enc_layers = 0;
    int enc_dims[6] = {0, 0, 0, 0, 0, 0};
    float* encW[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float* encB[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *obs_mean = nullptr, *obs_std = nullptr, *act_mean = nullptr, *act_std = nullptr, *d_mean = nullptr, *d_std = nullptr;
    float *cpo_mean = nullptr, *cpo_std = nullptr, *cpa_mean = nullptr, *cpa_std = nullptr;
    // decision state
    int m = 0;
    bool in_flight = false;
    uint64_t cur_seed = 0;          // seed / injected z of the iteration being rolled out (the refit needs them to
    const float* cur_z = nullptr;   // regenerate elite sequences owned by other ranks)
    float *obs = nullptr, *cp_obs = nullptr, *cp_act = nullptr, *mean = nullptr, *var = nullptr, *ctx = nullptr;
    float* actions = nullptr;
    int* actions_int = nullptr;
    float* ret_p = nullptr;
    float* returns_buf = nullptr;
    float* returns_log = nullptr;
    // exchange block (one allocation, exported over CUDA IPC): [2 parities][world, m_max, n_local] returns | [2][world, m_max] flags
    unsigned char* xchg = nullptr;
    size_t xchg_ret_bytes = 0, xchg_flag_bytes = 0, xchg_bytes = 0;
    long long peer_timeout_cycles = 60000000000ll;   // bound of the device-side wait for a peer's slice (~30 s at 2 GHz); cadm_set_option "peer_timeout_ms"
    unsigned char** peer_tab = nullptr;      // device array [world]: every rank's exchange block as seen from this device
    std::vector<void*> peer_open;            // mappings to close
    bool peers_on = false;
    int epoch = 0;                           // exchange epoch: one per rollout since the peers were attached
    int* peer_timeout_host = nullptr;        // host-mapped word: a refit kernel gave up waiting for a peer's slice
    int* peer_timeout_dev = nullptr;
    int* elites_log = nullptr;
    int* best = nullptr;
    // sampler-side state kept on the device (cadm_session_*)
    float *s_prev = nullptr, *s_var = nullptr, *s_obs = nullptr, *s_next = nullptr, *s_act = nullptr, *s_hobs = nullptr, *s_hact = nullptr;
    int* s_counts = nullptr;
    unsigned char* s_mask = nullptr;
    bool s_var_set = false;
    // instrumentation
    long long launches = 0;
    const char* kernel_name = "rollout_f32_kernel";
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    int ev_used = 0;
    std::vector<void*> allocs;
};

int fail(Engine* E, int code, const std::string& msg) {
    if (E) E->err = msg;
    else g_create_error = msg;
    return code;
}

#define CU(E, call)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(E, CADM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));        \
    } while (0)

template <typename T>
cudaError_t dalloc(Engine* E, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    E->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T));
}

int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

Engine* H(void* h) { return reinterpret_cast<Engine*>(h); }

RolloutParams base_params(Engine* E) {
    const CadmConfig& c = E->cfg;
    RolloutParams P{};
    P.env_id = c.env_id; P.D = c.obs_dim; P.P = c.proc_obs_dim; P.A = c.act_dim; P.C = c.ctx_dim; P.In = E->In;
    P.H = c.hidden; P.n_hidden = c.n_hidden;
    P.E = c.ensemble; P.p = c.particles; P.q = E->q;
    P.n_local = E->n_local; P.n_global = c.candidates; P.n_offset = E->n_offset;
    P.h = c.horizon;
    P.env_offset = E->env_offset;
    P.deterministic = c.deterministic; P.discrete = c.discrete;
    P.max_torque = c.max_torque > 0.f ? c.max_torque : 2.0f;
    P.Kp0 = E->Kp0; P.Hp = E->Hp; P.NHp = E->NHp;
    P.member_stride = E->member_stride; P.bias_stride = E->bias_stride;
    P.wpack = E->wpack; P.bpack = E->bpack;
    P.obs_mean = E->obs_mean; P.obs_std = E->obs_std; P.act_mean = E->act_mean; P.act_std = E->act_std;
    P.delta_mean = E->d_mean; P.delta_std = E->d_std; P.max_lv = E->max_lv; P.min_lv = E->min_lv;
    return P;
}

int check_ready(Engine* E, bool need_encoder) {
    if (!E->have_weights) return fail(E, CADM_ERR_STATE, "cadm_plan_set_weights has not been called");
    if (!E->have_norm) return fail(E, CADM_ERR_STATE, "cadm_plan_set_norm has not been called");
    if (need_encoder && E->cfg.ctx_dim > 0 && !E->have_encoder)
        return fail(E, CADM_ERR_STATE, "cadm_plan_set_encoder has not been called");
    return CADM_OK;
}

// The refit kernel's wait for a peer's slice is bounded; when it gave up it says so in a host-mapped word.  The launches are
// asynchronous, so the report surfaces on the first engine call after the kernel has run (begin / rollout / finish).
int check_peers(Engine* E) {
    if (E->peer_timeout_host && *reinterpret_cast<volatile int*>(E->peer_timeout_host) != 0) {
        const int who = *E->peer_timeout_host - 1;
        return fail(E, CADM_ERR_CUDA, "fused all-gather: gave up waiting for the returns slice of rank " + std::to_string(who) +
                                          " (peer process gone or out of lockstep); results since then are invalid");
    }
    return CADM_OK;
}

int run_rollout(Engine* E, RolloutParams& P, cudaStream_t s) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (E->timing) {
        if ((int)E->ev.size() < E->ev_used + 2) {
            cudaEvent_t a, b;
            CU(E, cudaEventCreate(&a));
            CU(E, cudaEventCreate(&b));
            E->ev.push_back(a);
            E->ev.push_back(b);
        }
        e0 = E->ev[E->ev_used];
        e1 = E->ev[E->ev_used + 1];
        E->ev_used += 2;
        CU(E, cudaEventRecord(e0, s));
    }
    switch (E->precision) {
        case CADM_PREC_FP32:
            CU(E, launch_rollout_f32(P, E->num_sms, s, &E->kernel_name));
            break;
        case CADM_PREC_TC_3X:
        case CADM_PREC_TC_1X:
            P.bpack = E->bpack_tc;
            P.bias_stride = E->bias_stride_tc;
            {
                // Which tensor-core kernel(s), and for the swapped-operand one how many rows per tile: the launch time is
                // (waves of tiles over the SMs) x (time of one wave), and one wave costs about the same for every workload of the
                // reference architecture -- measured on B200 (profiles/r2_variant_sweep.log, us per launch at h = 30):
                //   128-row tiles (rollout_tc.cu) 570,   swapped operands (rollout_tcs.cu) 238 / 293 / 375 for 32 / 48 / 64 rows.
                // Per row the 128-row kernel is the cheapest (4.5 us against 6.1 - 7.4) but only in FULL waves, so a large batch is
                // split: k full waves of 128-row tiles, and the rows that would start a mostly empty wave go to one wave of the
                // swapped kernel with small tiles (two launches, disjoint rows of every member).  The model reproduces the measured
                // launch times of C2 (m = 1, 4, 10), C3 (m = 2) and C4 to 3 %.
                const bool swapped_ok = E->Np16 <= 256 && E->NHp16 <= 128;
                const int rows_pm = P.rows_per_member, sms = E->num_sms;
                int rows_tc = E->tc_variant == 1 ? rows_pm : 0;           // rows of every member given to the 128-row kernel
                int rows_pick = E->tcs_rows;
                if (E->tc_variant == 2) rows_tc = 0;
                if (E->tc_variant == 0) {
                    rows_tc = rows_pm;
                    if (swapped_ok) {
                        auto waves = [&](int span, int rows) { return (P.E * ((span + rows - 1) / rows) + sms - 1) / sms; };
                        const int cost[3] = {238, 293, 375};
                        auto best_tcs = [&](int span, int& pick) {       // cheapest swapped launch over `span` rows per member
                            int best = 1 << 30;
                            for (int i = 0; i < 3; ++i) {
                                if (E->tcs_rows != 0 && E->tcs_rows != 32 + 16 * i) continue;
                                const int c = waves(span, 32 + 16 * i) * cost[i];
                                if (c < best) { best = c; pick = 32 + 16 * i; }
                            }
                            return best;
                        };
                        int best = waves(rows_pm, 128) * 570;            // everything on 128-row tiles
                        int pick = 32;
                        const int all_tcs = best_tcs(rows_pm, pick);
                        if (all_tcs < best) { best = all_tcs; rows_tc = 0; rows_pick = pick; }
                        for (int k = 1; k <= 64; ++k) {                  // k full waves of 128-row tiles + the rest swapped
                            const int tiles_pm = k * sms / P.E;
                            const int r1 = tiles_pm * 128;
                            if (tiles_pm < 1) continue;
                            if (r1 >= rows_pm) break;
                            int pk = 32;
                            const int c = k * 570 + best_tcs(rows_pm - r1, pk);
                            if (c < best) { best = c; rows_tc = r1; rows_pick = pk; }
                        }
                    }
                }
                const int terms = E->precision == CADM_PREC_TC_3X ? 3 : 1;
                if (E->tc_variant == 3) {
                    if (!tcp_supported(P, E->tcs_kps)) return fail(E, CADM_ERR_UNSUPPORTED, "tc_variant 3 (CTA pairs) covers the reference architecture only");
                    CU(E, launch_rollout_tcp(P, E->wimg_s, E->wimg_member_stride, terms, E->tcs_rows, E->num_sms, s, &E->kernel_name, E->trace ? E->dbg : nullptr));
                } else {
                    const char* n1 = nullptr;
                    const char* n2 = nullptr;
                    if (rows_tc > 0) {
                        P.row_lo = 0; P.row_hi = rows_tc;
                        CU(E, launch_rollout_tc(P, E->wimg, E->wimg_member_stride, terms, E->num_sms, s, &n1, E->trace ? E->dbg : nullptr));
                    }
                    if (rows_tc < rows_pm) {
                        P.row_lo = rows_tc; P.row_hi = rows_pm;
                        CU(E, launch_rollout_tcs(P, E->wimg_s, E->wimg_member_stride, terms, E->tcs_kps, rows_pick, E->tcs_skew, E->num_sms, s, &n2,
                                                 (E->trace && rows_tc == 0) ? E->dbg : nullptr));
                        if (rows_tc > 0) E->launches++;
                    }
                    P.row_lo = 0; P.row_hi = 0;
                    E->kernel_name = (n1 && n2) ? (terms == 3 ? "rollout_tc_kernel(128-row tiles, full waves) + rollout_tcs_kernel(swapped operands, the rest), fp16 hi/lo x3"
                                                              : "rollout_tc_kernel(128-row tiles, full waves) + rollout_tcs_kernel(swapped operands, the rest), f16 x1")
                                                : (n1 ? n1 : n2);
                }
            }
            break;
        default:
            return fail(E, CADM_ERR_UNSUPPORTED, "unknown precision mode");
    }
    E->launches++;
    if (E->timing) CU(E, cudaEventRecord(e1, s));
    return CADM_OK;
}

int run_encoder(Engine* E, int m, const float* cp_obs, const float* cp_act, float* ctx, cudaStream_t s) {
    const CadmConfig& c = E->cfg;
    EncoderParams Q{};
    Q.m = m; Q.E = c.ensemble; Q.D = c.obs_dim; Q.A = c.act_dim; Q.K = c.hist_len; Q.C = c.ctx_dim;
    Q.n_layers = E->enc_layers;
    for (int i = 0; i <= E->enc_layers; ++i) Q.dims[i] = E->enc_dims[i];
    for (int i = 0; i < E->enc_layers; ++i) { Q.W[i] = E->encW[i]; Q.b[i] = E->encB[i]; }
    Q.cp_obs = cp_obs; Q.cp_act = cp_act;
    Q.cpo_mean = E->cpo_mean; Q.cpo_std = E->cpo_std; Q.cpa_mean = E->cpa_mean; Q.cpa_std = E->cpa_std;
    Q.ctx = ctx;
    CU(E, launch_encoder(Q, s));
    E->launches++;
    return CADM_OK;
}

}  // namespace

extern "C" {

int cadm_abi_version(void) { return CADM_ABI_VERSION; }

const char* cadm_last_error(const void* handle) {
    if (!handle) return g_create_error.c_str();
    return reinterpret_cast<const Engine*>(handle)->err.c_str();
}

int cadm_plan_create(const CadmConfig* cfg, void** handle) {
    if (!cfg || !handle) return fail(nullptr, CADM_ERR_ARG, "null argument");
    if (cfg->struct_size != (int)sizeof(CadmConfig)) return fail(nullptr, CADM_ERR_ARG, "CadmConfig size mismatch (ABI)");
    const CadmConfig& c = *cfg;
    auto bad = [&](const char* m) { return fail(nullptr, CADM_ERR_ARG, m); };
    if (c.env_id < 0 || c.env_id > CADM_ENV_PENDULUM) return bad("unknown env_id");
    if (c.obs_dim < 1 || c.obs_dim > kMaxObs) return bad("obs_dim out of range (1..48)");
    if (c.proc_obs_dim < 1 || c.proc_obs_dim > kMaxObs) return bad("proc_obs_dim out of range");
    if (c.act_dim < 1 || c.act_dim > kMaxAct) return bad("act_dim out of range (1..20)");
    if (c.ctx_dim < 0 || c.ctx_dim > kMaxCtx) return bad("ctx_dim out of range (0..16)");
    if (c.hidden < 8 || c.hidden > kMaxHidden) return bad("hidden width out of range (8..208)");
    if (c.n_hidden < 1 || c.n_hidden > 8) return bad("n_hidden out of range (1..8)");
    if (c.ensemble < 1 || c.particles < 1 || c.particles % c.ensemble) return bad("particles must be a positive multiple of ensemble");
    if (c.world < 1 || c.rank < 0 || c.rank >= c.world) return bad("bad rank/world");
    if (c.candidates < 1 || c.candidates % c.world) return bad("candidates must be a positive multiple of world");
    if (c.horizon < 1 || c.m_max < 1) return bad("horizon and m_max must be positive");
    if (c.num_elites < 1 || c.num_elites > 256 || c.num_elites > c.candidates) return bad("num_elites must be in 1..min(256, candidates)");
    if (c.cem_iters < 1) return bad("cem_iters must be positive");
    if (next_pow2(c.candidates) > 16384) return bad("candidates > 16384 not supported by the elite selection kernel");
    if (c.env_id == CADM_ENV_SLIM_HUMANOID && c.obs_dim < 23) return bad("slim humanoid needs obs_dim >= 23");
    if (c.env_id == CADM_ENV_ANT && c.proc_obs_dim != c.obs_dim - 1) return bad("ant: proc_obs_dim must be obs_dim - 1");
    if (c.env_id != CADM_ENV_ANT && c.proc_obs_dim != c.obs_dim) return bad("proc_obs_dim must equal obs_dim for this env");
    if (c.ctx_dim > 0 && c.hist_len < 1) return bad("hist_len must be positive with a context encoder");
    if (c.precision < CADM_PREC_FP32 || c.precision > CADM_PREC_TC_1X) return bad("unknown precision mode");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, CADM_ERR_CUDA, "no CUDA device: cadm_b200 has no CPU fallback");
    Engine* E = new (std::nothrow) Engine();
    if (!E) return fail(nullptr, CADM_ERR_ARG, "out of host memory");
    E->cfg = c;
    cudaGetDevice(&E->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, E->device) == cudaSuccess) {
        E->num_sms = prop.multiProcessorCount;
        if (prop.major != 10) {
            delete E;
            return fail(nullptr, CADM_ERR_UNSUPPORTED, "cadm_b200 is built for sm_100a (B200) only");
        }
    }
    E->In = c.proc_obs_dim + c.act_dim + c.ctx_dim;
    E->Kp0 = round_up(E->In, kChunkK);
    E->Hp = round_up(c.hidden, 8);
    E->NHp = round_up(2 * c.obs_dim, 8);
    E->n_local = c.candidates / c.world;
    E->n_offset = c.rank * E->n_local;
    E->q = c.particles / c.ensemble;
    E->hA = c.horizon * c.act_dim;
    E->precision = c.precision;
    E->Np16 = round_up(c.hidden, 16);
    E->NHp16 = round_up(2 * c.obs_dim, 16);
    E->nkb0 = round_up(E->In, 16) / 16;
    E->nkbH = E->Np16 / 16;
    E->wimg_member_stride = ((long long)(E->nkb0 + (c.n_hidden - 1) * E->nkbH) * E->Np16 + (long long)E->nkbH * E->NHp16) * 64;
    E->bias_stride_tc = (long long)c.n_hidden * E->Np16 + E->NHp16;
    E->member_stride = (long long)E->Kp0 * E->Hp + (long long)(c.n_hidden - 1) * E->Hp * E->Hp + (long long)E->Hp * E->NHp;
    E->bias_stride = (long long)c.n_hidden * E->Hp + E->NHp;

    const size_t mm = c.m_max;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(dalloc(E, &E->wpack, (size_t)c.ensemble * E->member_stride));
    A(dalloc(E, &E->bpack, (size_t)c.ensemble * E->bias_stride));
    A(dalloc(E, &E->wimg, (size_t)c.ensemble * E->wimg_member_stride));
    A(dalloc(E, &E->wimg_s, (size_t)c.ensemble * E->wimg_member_stride));
    A(dalloc(E, &E->bpack_tc, (size_t)c.ensemble * E->bias_stride_tc));
    A(dalloc(E, &E->max_lv, c.obs_dim));
    A(dalloc(E, &E->min_lv, c.obs_dim));
    A(dalloc(E, &E->obs_mean, c.proc_obs_dim)); A(dalloc(E, &E->obs_std, c.proc_obs_dim));
    A(dalloc(E, &E->act_mean, c.act_dim)); A(dalloc(E, &E->act_std, c.act_dim));
    A(dalloc(E, &E->d_mean, c.obs_dim)); A(dalloc(E, &E->d_std, c.obs_dim));
    const int K = c.ctx_dim > 0 ? c.hist_len : 1;
    A(dalloc(E, &E->cpo_mean, (size_t)c.obs_dim * K)); A(dalloc(E, &E->cpo_std, (size_t)c.obs_dim * K));
    A(dalloc(E, &E->cpa_mean, (size_t)c.act_dim * K)); A(dalloc(E, &E->cpa_std, (size_t)c.act_dim * K));
    A(dalloc(E, &E->obs, mm * c.obs_dim));
    A(dalloc(E, &E->cp_obs, mm * c.obs_dim * K)); A(dalloc(E, &E->cp_act, mm * c.act_dim * K));
    A(dalloc(E, &E->mean, mm * E->hA)); A(dalloc(E, &E->var, mm * E->hA));
    A(dalloc(E, &E->ctx, (size_t)c.ensemble * mm * std::max(c.ctx_dim, 1)));
    A(dalloc(E, &E->actions, mm * E->n_local * E->hA));
    A(dalloc(E, &E->actions_int, mm * E->n_local * c.horizon));
    A(dalloc(E, &E->ret_p, mm * E->n_local * c.particles));
    E->xchg_ret_bytes = (mm * c.candidates * sizeof(float) + 255) / 256 * 256;
    E->xchg_flag_bytes = ((size_t)c.world * mm * sizeof(int) + 255) / 256 * 256;
    E->xchg_bytes = 2 * E->xchg_ret_bytes + 2 * E->xchg_flag_bytes;
    A(dalloc(E, &E->xchg, E->xchg_bytes));
    E->returns_buf = reinterpret_cast<float*>(E->xchg);          // parity 0 doubles as the buffer of the NCCL path
    A(dalloc(E, &E->peer_tab, (size_t)c.world));
    A(dalloc(E, &E->returns_log, (size_t)c.cem_iters * mm * c.candidates));
    A(dalloc(E, &E->elites_log, (size_t)c.cem_iters * mm * c.num_elites));
    A(dalloc(E, &E->best, mm));
    A(dalloc(E, &E->s_prev, mm * E->hA)); A(dalloc(E, &E->s_var, mm * E->hA));
    A(dalloc(E, &E->s_obs, mm * c.obs_dim)); A(dalloc(E, &E->s_next, mm * c.obs_dim)); A(dalloc(E, &E->s_act, mm * c.act_dim));
    A(dalloc(E, &E->s_hobs, mm * c.obs_dim * K)); A(dalloc(E, &E->s_hact, mm * c.act_dim * K));
    A(dalloc(E, &E->s_counts, mm)); A(dalloc(E, &E->s_mask, mm));
    A(dalloc(E, &E->dbg, 64 * 64));
    if (e != cudaSuccess) {
        std::string msg = std::string("device allocation failed: ") + cudaGetErrorString(e);
        for (void* p : E->allocs) cudaFree(p);
        delete E;
        return fail(nullptr, CADM_ERR_CUDA, msg);
    }
    *handle = E;
    return CADM_OK;
}

int cadm_plan_destroy(void* handle) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    cudaDeviceSynchronize();
    for (void* p : E->peer_open) cudaIpcCloseMemHandle(p);
    if (E->peer_timeout_host) cudaFreeHost(E->peer_timeout_host);
    for (void* p : E->allocs) cudaFree(p);
    for (cudaEvent_t ev : E->ev) cudaEventDestroy(ev);
    delete E;
    return CADM_OK;
}

int cadm_plan_set_weights(void* handle, const float* const* W, const float* const* b, int32_t n_layers,
                          const float* max_logvar, const float* min_logvar, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (!W || !b || n_layers != c.n_hidden + 2) return fail(E, CADM_ERR_ARG, "expected n_hidden + 2 layers (hidden..., mu, logvar)");
    if (!max_logvar || !min_logvar) return fail(E, CADM_ERR_ARG, "max_logvar/min_logvar are required");
    for (int l = 0; l < n_layers; ++l)
        if (!W[l] || !b[l]) return fail(E, CADM_ERR_ARG, "null weight pointer");
    cudaStream_t s = (cudaStream_t)stream;
    long long off = 0, boff = 0;
    for (int l = 0; l < c.n_hidden; ++l) {
        const int in = l == 0 ? E->In : c.hidden;
        const int Kp = l == 0 ? E->Kp0 : E->Hp;
        CU(E, launch_pack_f32(E->wpack, W[l], c.ensemble, in, c.hidden, Kp, E->Hp, 0, E->member_stride, off, 1, s));
        CU(E, launch_pack_bias(E->bpack, b[l], c.ensemble, c.hidden, 0, E->bias_stride, boff, s));
        off += (long long)Kp * E->Hp;
        boff += E->Hp;
        E->launches += 2;
    }
    // heads: mu -> columns [0, D), logvar -> columns [D, 2D)
    CU(E, launch_pack_f32(E->wpack, W[c.n_hidden], c.ensemble, c.hidden, c.obs_dim, E->Hp, E->NHp, 0, E->member_stride, off, 1, s));
    CU(E, launch_pack_f32(E->wpack, W[c.n_hidden + 1], c.ensemble, c.hidden, c.obs_dim, E->Hp, E->NHp, c.obs_dim, E->member_stride, off, 0, s));
    CU(E, launch_pack_bias(E->bpack, b[c.n_hidden], c.ensemble, c.obs_dim, 0, E->bias_stride, boff, s));
    CU(E, launch_pack_bias(E->bpack, b[c.n_hidden + 1], c.ensemble, c.obs_dim, c.obs_dim, E->bias_stride, boff, s));
    E->launches += 4;
    // tensor-core image: per layer nkb K16 blocks of [hi | lo], N padded to 16
    {
        long long toff = 0, tboff = 0;
        for (int l = 0; l < c.n_hidden; ++l) {
            const int in = l == 0 ? E->In : c.hidden;
            const int nkb = l == 0 ? E->nkb0 : E->nkbH;
            CU(E, launch_pack_tc(E->wimg, W[l], c.ensemble, in, c.hidden, 0, nkb, E->Np16, E->wimg_member_stride, toff, 1, s));
            CU(E, launch_pack_bias(E->bpack_tc, b[l], c.ensemble, c.hidden, 0, E->bias_stride_tc, tboff, s));
            toff += (long long)nkb * E->Np16 * 64;
            tboff += E->Np16;
            E->launches += 2;
        }
        CU(E, launch_pack_tc(E->wimg, W[c.n_hidden], c.ensemble, c.hidden, c.obs_dim, 0, E->nkbH, E->NHp16, E->wimg_member_stride, toff, 1, s));
        CU(E, launch_pack_tc(E->wimg, W[c.n_hidden + 1], c.ensemble, c.hidden, c.obs_dim, c.obs_dim, E->nkbH, E->NHp16, E->wimg_member_stride, toff, 0, s));
        CU(E, launch_pack_bias(E->bpack_tc, b[c.n_hidden], c.ensemble, c.obs_dim, 0, E->bias_stride_tc, tboff, s));
        CU(E, launch_pack_bias(E->bpack_tc, b[c.n_hidden + 1], c.ensemble, c.obs_dim, c.obs_dim, E->bias_stride_tc, tboff, s));
        E->launches += 4;
    }
    // the same image in the stage order / tiling of the swapped-operand kernel
    {
        long long toff = 0;
        const int kps = E->tcs_kps;
        for (int l = 0; l < c.n_hidden; ++l) {
            const int in = l == 0 ? E->In : c.hidden;
            const int nkb = l == 0 ? E->nkb0 : E->nkbH;
            CU(E, launch_pack_tcs(E->wimg_s, W[l], c.ensemble, in, c.hidden, 0, nkb, E->Np16, kps, E->wimg_member_stride, toff, 1, s));
            toff += (long long)nkb * E->Np16 * 64;
            E->launches++;
        }
        CU(E, launch_pack_tcs(E->wimg_s, W[c.n_hidden], c.ensemble, c.hidden, c.obs_dim, 0, E->nkbH, E->NHp16, kps, E->wimg_member_stride, toff, 1, s));
        CU(E, launch_pack_tcs(E->wimg_s, W[c.n_hidden + 1], c.ensemble, c.hidden, c.obs_dim, c.obs_dim, E->nkbH, E->NHp16, kps, E->wimg_member_stride, toff, 0, s));
        E->launches += 2;
    }
    CU(E, cudaMemcpyAsync(E->max_lv, max_logvar, c.obs_dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(E, cudaMemcpyAsync(E->min_lv, min_logvar, c.obs_dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
    E->have_weights = true;
    return CADM_OK;
}

int cadm_plan_set_encoder(void* handle, const float* const* W, const float* const* b, int32_t n_layers, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (c.ctx_dim <= 0) return fail(E, CADM_ERR_STATE, "engine was created without a context encoder (ctx_dim == 0)");
    if (!W || !b || n_layers < 1 || n_layers > 4) return fail(E, CADM_ERR_ARG, "encoder must have 1..4 layers");
    cudaStream_t s = (cudaStream_t)stream;
    int dims[6];
    dims[0] = (c.obs_dim + c.act_dim) * c.hist_len;
    for (int l = 0; l < n_layers - 1; ++l) {
        if (c.enc_hidden[l] < 1 || c.enc_hidden[l] > 1024) return fail(E, CADM_ERR_ARG, "enc_hidden out of range");
        dims[l + 1] = c.enc_hidden[l];
    }
    dims[n_layers] = c.ctx_dim;
    for (int l = 0; l < n_layers; ++l) {
        if (!W[l] || !b[l]) return fail(E, CADM_ERR_ARG, "null encoder pointer");
        const size_t nw = (size_t)c.ensemble * dims[l] * dims[l + 1], nb = (size_t)c.ensemble * dims[l + 1];
        if (!E->encW[l] || E->enc_dims[l] != dims[l] || E->enc_dims[l + 1] != dims[l + 1]) {
            CU(E, dalloc(E, &E->encW[l], nw));
            CU(E, dalloc(E, &E->encB[l], nb));
        }
        CU(E, cudaMemcpyAsync(E->encW[l], W[l], nw * sizeof(float), cudaMemcpyDeviceToDevice, s));
        CU(E, cudaMemcpyAsync(E->encB[l], b[l], nb * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    for (int l = 0; l <= n_layers; ++l) E->enc_dims[l] = dims[l];
    E->enc_layers = n_layers;
    E->have_encoder = true;
    return CADM_OK;
}

int cadm_plan_set_norm(void* handle, const float* obs_mean, const float* obs_std, const float* act_mean,
                       const float* act_std, const float* delta_mean, const float* delta_std, const float* cp_obs_mean,
                       const float* cp_obs_std, const float* cp_act_mean, const float* cp_act_std, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (!obs_mean || !obs_std || !act_mean || !act_std || !delta_mean || !delta_std)
        return fail(E, CADM_ERR_ARG, "obs/act/delta statistics are required");
    cudaStream_t s = (cudaStream_t)stream;
    auto cp = [&](float* d, const float* src, size_t n) { return cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s); };
    CU(E, cp(E->obs_mean, obs_mean, c.proc_obs_dim)); CU(E, cp(E->obs_std, obs_std, c.proc_obs_dim));
    CU(E, cp(E->act_mean, act_mean, c.act_dim)); CU(E, cp(E->act_std, act_std, c.act_dim));
    CU(E, cp(E->d_mean, delta_mean, c.obs_dim)); CU(E, cp(E->d_std, delta_std, c.obs_dim));
    if (c.ctx_dim > 0) {
        if (!cp_obs_mean || !cp_obs_std || !cp_act_mean || !cp_act_std)
            return fail(E, CADM_ERR_ARG, "cp_obs/cp_act statistics are required with a context encoder");
        const size_t K = c.hist_len;
        CU(E, cp(E->cpo_mean, cp_obs_mean, c.obs_dim * K)); CU(E, cp(E->cpo_std, cp_obs_std, c.obs_dim * K));
        CU(E, cp(E->cpa_mean, cp_act_mean, c.act_dim * K)); CU(E, cp(E->cpa_std, cp_act_std, c.act_dim * K));
    }
    E->have_norm = true;
    return CADM_OK;
}

int cadm_encode_context(void* handle, int32_t m, const float* cp_obs, const float* cp_act, float* ctx, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (E->cfg.ctx_dim <= 0) return fail(E, CADM_ERR_STATE, "no context encoder in this engine");
    if (!E->have_encoder || !E->have_norm) return fail(E, CADM_ERR_STATE, "encoder weights / norm stats not set");
    if (m < 1 || !cp_obs || !cp_act || !ctx) return fail(E, CADM_ERR_ARG, "bad arguments");
    return run_encoder(E, m, cp_obs, cp_act, ctx, (cudaStream_t)stream);
}

int cadm_predict(void* handle, int32_t B, const float* obs, const float* act, const float* ctx, const float* eps,
                 uint64_t seed, float* next_obs, float* mu, float* logvar, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (int r = check_ready(E, false)) return r;
    if (B < 1 || !obs || !act) return fail(E, CADM_ERR_ARG, "bad arguments");
    if (E->cfg.ctx_dim > 0 && !ctx) return fail(E, CADM_ERR_ARG, "ctx is required for a context model");
    RolloutParams P = base_params(E);
    P.row_mode = kRowsPredict;
    P.rows_per_member = B;
    P.m = 1; P.h = 1; P.it = 0;
    P.n_local = B; P.n_global = B; P.n_offset = 0;
    P.ctx_mode = E->cfg.ctx_dim > 0 ? 3 : 0;
    P.obs0 = obs; P.actions = act; P.ctx = ctx; P.eps = eps; P.seed = seed;
    P.next_obs = next_obs; P.mu_out = mu; P.lv_out = logvar;
    return run_rollout(E, P, (cudaStream_t)stream);
}

int cadm_rollout(void* handle, int32_t m, int32_t it, const float* obs, const float* actions, const float* ctx_raw,
                 const float* eps, uint64_t seed, float* particle_returns, float* states, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (int r = check_ready(E, false)) return r;
    const CadmConfig& c = E->cfg;
    if (m < 1 || m > c.m_max || !obs || !actions || !particle_returns) return fail(E, CADM_ERR_ARG, "bad arguments");
    if (c.discrete) return fail(E, CADM_ERR_UNSUPPORTED, "cadm_rollout takes continuous actions");
    if (c.ctx_dim > 0 && !ctx_raw) return fail(E, CADM_ERR_ARG, "ctx_raw is required for a context model");
    RolloutParams P = base_params(E);
    P.row_mode = kRowsPlanner;
    P.m = m; P.it = it;
    P.rows_per_member = E->q * m * E->n_local;
    P.ctx_mode = c.ctx_dim > 0 ? (c.context_layout == CADM_CTX_MATCHED ? 2 : 1) : 0;
    P.obs0 = obs; P.actions = actions; P.ctx = ctx_raw; P.eps = eps; P.seed = seed;
    P.ret_p = particle_returns; P.states = states;
    return run_rollout(E, P, (cudaStream_t)stream);
}

int cadm_cem_begin(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act,
                   const float* init_mean, const float* init_var, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (int r = check_ready(E, true)) return r;
    if (int r = check_peers(E)) return r;
    const CadmConfig& c = E->cfg;
    if (c.discrete) return fail(E, CADM_ERR_UNSUPPORTED, "CEM needs continuous actions (the reference builds RS for discrete envs)");
    if (m < 1 || m > c.m_max) return fail(E, CADM_ERR_ARG, "m out of range (1..m_max)");
    if (!obs || !init_mean || !init_var) return fail(E, CADM_ERR_ARG, "obs/init_mean/init_var are required");
    if (c.ctx_dim > 0 && (!cp_obs || !cp_act)) return fail(E, CADM_ERR_ARG, "cp_obs/cp_act are required for a context model");
    cudaStream_t s = (cudaStream_t)stream;
    auto cp = [&](float* d, const float* src, size_t n) {
        return d == src ? cudaSuccess : cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
    };
    CU(E, cp(E->obs, obs, (size_t)m * c.obs_dim));
    CU(E, cp(E->mean, init_mean, (size_t)m * E->hA));
    CU(E, cp(E->var, init_var, (size_t)m * E->hA));
    E->m = m;
    E->ev_used = 0;
    if (c.ctx_dim > 0) {
        if (int r = run_encoder(E, m, cp_obs, cp_act, E->ctx, s)) return r;
    }
    E->in_flight = true;
    return CADM_OK;
}

int cadm_cem_rollout(void* handle, int32_t it, uint64_t seed, const float* z, const float* eps, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (!E->in_flight) return fail(E, CADM_ERR_STATE, "cadm_cem_begin has not been called");
    const CadmConfig& c = E->cfg;
    if (it < 0 || it >= c.cem_iters) return fail(E, CADM_ERR_ARG, "iteration out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const int m = E->m;
    E->cur_seed = seed;
    E->cur_z = z;
    SampleParams S{};
    S.m = m; S.n_local = E->n_local; S.n_global = c.candidates; S.n_offset = E->n_offset; S.hA = E->hA; S.A = c.act_dim;
    S.it = it; S.mode = 0; S.seed = seed; S.env_offset = E->env_offset;
    S.mean = E->mean; S.var = E->var;
    S.z = z ? z + (size_t)it * m * c.candidates * E->hA : nullptr;
    S.actions = E->actions;
    CU(E, launch_sample_actions(S, s));
    E->launches++;

    RolloutParams P = base_params(E);
    P.row_mode = kRowsPlanner;
    P.m = m; P.it = it;
    P.rows_per_member = E->q * m * E->n_local;
    P.ctx_mode = c.ctx_dim > 0 ? (c.context_layout == CADM_CTX_MATCHED ? 2 : 1) : 0;
    P.obs0 = E->obs; P.actions = E->actions; P.ctx = E->ctx; P.seed = seed;
    P.eps = eps ? eps + (size_t)it * c.horizon * c.ensemble * ((size_t)E->q * m * c.candidates) * c.obs_dim : nullptr;
    P.ret_p = E->ret_p; P.states = nullptr;
    if (int r = run_rollout(E, P, s)) return r;

    if (c.world > 1) {      // single rank: the refit kernel folds the particle mean in
        if (E->peers_on) {
            // fused: the refit kernel averages the particles, stores the slice into every rank's exchange block and waits for the
            // others (cem_kernels.cu); parity = epoch & 1 so that a rank one iteration ahead cannot overwrite a slice a slower
            // rank is still reading
            ++E->epoch;
        } else {
            CU(E, launch_particle_mean(E->ret_p, E->returns_buf + (size_t)c.rank * m * E->n_local, m * E->n_local, c.particles, s));
            E->launches++;
        }
    }
    return CADM_OK;
}

float* cadm_cem_returns_buffer(void* handle) { return handle ? H(handle)->returns_buf : nullptr; }

int64_t cadm_cem_returns_slice_elems(void* handle) {
    Engine* E = H(handle);
    return E ? (int64_t)E->m * E->n_local : 0;
}

static int refit_common(Engine* E, int it, uint64_t seed, const float* z, cudaStream_t s) {
    const CadmConfig& c = E->cfg;
    const int m = E->m;
    RefitParams R{};
    R.m = m; R.n_local = E->n_local; R.n_global = c.candidates; R.n_offset = E->n_offset; R.world = c.world;
    R.h = c.horizon; R.A = c.act_dim; R.k_elites = c.num_elites; R.it = it; R.env_offset = E->env_offset;
    R.npad = next_pow2(c.candidates);
    R.alpha = c.alpha; R.seed = seed;
    R.returns_buf = E->returns_buf; R.actions = E->actions;
    if (c.world > 1 && E->peers_on) {
        const long long par = E->epoch & 1;
        R.peers = E->peer_tab;
        R.slice_off = par * (long long)E->xchg_ret_bytes;
        R.flag_off = 2 * (long long)E->xchg_ret_bytes + par * (long long)E->xchg_flag_bytes;
        R.ret_p_local = E->ret_p;
        R.rank = c.rank; R.m_max = c.m_max;
        R.peer_epoch = E->epoch;
        R.timeout_cycles = E->peer_timeout_cycles;
        R.peer_timeout = E->peer_timeout_dev;
    }
    R.z = z ? z + (size_t)it * m * c.candidates * E->hA : nullptr;
    R.mean = E->mean; R.var = E->var;
    R.returns_log = E->returns_log + (size_t)it * m * c.candidates;
    R.elites_log = E->elites_log + (size_t)it * m * c.num_elites;
    R.mode_rs = 0; R.best = E->best;
    R.ret_p = c.world == 1 ? E->ret_p : nullptr; R.p = c.particles;
    CU(E, launch_refit(R, s));
    E->launches++;
    return CADM_OK;
}

int cadm_cem_refit(void* handle, int32_t it, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (!E->in_flight) return fail(E, CADM_ERR_STATE, "cadm_cem_begin has not been called");
    if (it < 0 || it >= E->cfg.cem_iters) return fail(E, CADM_ERR_ARG, "iteration out of range");
    return refit_common(E, it, E->cur_seed, E->cur_z, (cudaStream_t)stream);
}

int cadm_cem_finish(void* handle, float* mean, float* var, float* returns, int32_t* elites, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (!E->in_flight) return fail(E, CADM_ERR_STATE, "cadm_cem_begin has not been called");
    const CadmConfig& c = E->cfg;
    cudaStream_t s = (cudaStream_t)stream;
    if (E->peers_on) {
        // the refit kernels report a peer that never delivered through a host-mapped word; they are asynchronous, so the word is
        // only meaningful once they have run: wait for them here and fail THIS decision instead of the next one
        CU(E, cudaStreamSynchronize(s));
        if (int r = check_peers(E)) { E->in_flight = false; return r; }
    }
    const size_t m = E->m;
    if (mean) CU(E, cudaMemcpyAsync(mean, E->mean, m * E->hA * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (var) CU(E, cudaMemcpyAsync(var, E->var, m * E->hA * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (returns) CU(E, cudaMemcpyAsync(returns, E->returns_log, (size_t)c.cem_iters * m * c.candidates * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (elites) CU(E, cudaMemcpyAsync(elites, E->elites_log, (size_t)c.cem_iters * m * c.num_elites * sizeof(int), cudaMemcpyDeviceToDevice, s));
    E->in_flight = false;
    return CADM_OK;
}

}  // extern "C"

extern "C" {

int cadm_plan_cem(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act,
                  const float* init_mean, const float* init_var, uint64_t seed, const float* z, const float* eps,
                  float* mean, float* var, float* returns, int32_t* elites, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (E->cfg.world != 1 && !E->peers_on)
        return fail(E, CADM_ERR_STATE, "cadm_plan_cem at world > 1 needs the fused exchange (cadm_peer_attach); without it use the phase calls "
                                       "with your own all-gather between cadm_cem_rollout and cadm_cem_refit");
    if (int r = cadm_cem_begin(handle, m, obs, cp_obs, cp_act, init_mean, init_var, stream)) return r;
    for (int it = 0; it < E->cfg.cem_iters; ++it) {
        if (int r = cadm_cem_rollout(handle, it, seed, z, eps, stream)) return r;
        if (int r = cadm_cem_refit(handle, it, stream)) return r;
    }
    return cadm_cem_finish(handle, mean, var, returns, elites, stream);
}

int cadm_plan_cem_host(void* handle, int32_t m, const float* obs_host, const float* cp_obs_host, const float* cp_act_host,
                       const float* init_mean_host, const float* init_var_host, uint64_t seed, float* action_host,
                       void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (c.world != 1 && !E->peers_on)
        return fail(E, CADM_ERR_STATE, "cadm_plan_cem_host at world > 1 needs the fused exchange (cadm_peer_attach)");
    if (m < 1 || m > c.m_max) return fail(E, CADM_ERR_ARG, "m out of range (1..m_max)");
    if (!obs_host || !init_mean_host || !init_var_host || !action_host) return fail(E, CADM_ERR_ARG, "null host pointer");
    if (c.ctx_dim > 0 && (!cp_obs_host || !cp_act_host)) return fail(E, CADM_ERR_ARG, "cp_obs/cp_act are required for a context model");
    cudaStream_t s = (cudaStream_t)stream;
    auto up = [&](float* d, const float* src, size_t n) { return cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyHostToDevice, s); };
    CU(E, up(E->obs, obs_host, (size_t)m * c.obs_dim));
    CU(E, up(E->mean, init_mean_host, (size_t)m * E->hA));
    CU(E, up(E->var, init_var_host, (size_t)m * E->hA));
    if (c.ctx_dim > 0) {
        CU(E, up(E->cp_obs, cp_obs_host, (size_t)m * c.obs_dim * c.hist_len));
        CU(E, up(E->cp_act, cp_act_host, (size_t)m * c.act_dim * c.hist_len));
    }
    if (int r = cadm_plan_cem(handle, m, E->obs, E->cp_obs, E->cp_act, E->mean, E->var, seed, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, stream))
        return r;
    const size_t n = (size_t)m * E->hA;
    CU(E, cudaMemcpyAsync(action_host, E->mean, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(E, cudaStreamSynchronize(s));
    for (size_t i = 0; i < n; ++i) action_host[i] = fminf(fmaxf(action_host[i], -1.0f), 1.0f);   // get_action clip
    return CADM_OK;
}

// ---- sampler-side state on the device ----------------------------------------------------------
int cadm_session_reset(void* handle, int32_t m, const uint8_t* mask_host, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (m < 1 || m > c.m_max) return fail(E, CADM_ERR_ARG, "m out of range (1..m_max)");
    cudaStream_t s = (cudaStream_t)stream;
    const int K = c.ctx_dim > 0 ? c.hist_len : 1;
    if (mask_host) CU(E, cudaMemcpyAsync(E->s_mask, mask_host, m, cudaMemcpyHostToDevice, s));
    CU(E, launch_session_reset(mask_host ? E->s_mask : nullptr, E->s_hobs, E->s_hact, E->s_counts, E->s_prev, m, c.obs_dim * K,
                               c.act_dim * K, E->hA, s));
    E->launches++;
    if (!E->s_var_set) {        // init_var = 2^2 / 16 for every coordinate, never updated (sampler.py:53, quirk Q7)
        std::vector<float> v((size_t)c.m_max * E->hA, 0.25f);
        CU(E, cudaMemcpyAsync(E->s_var, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        CU(E, cudaStreamSynchronize(s));
        E->s_var_set = true;
    }
    return CADM_OK;
}

int cadm_session_act(void* handle, int32_t m, const float* obs_host, uint64_t seed, float* action_host, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (c.world != 1 && !E->peers_on) return fail(E, CADM_ERR_STATE, "sessions at world > 1 need the fused exchange (cadm_peer_attach)");
    if (c.discrete) return fail(E, CADM_ERR_UNSUPPORTED, "sessions plan with CEM (continuous actions)");
    if (m < 1 || m > c.m_max || !obs_host || !action_host) return fail(E, CADM_ERR_ARG, "bad arguments");
    if (!E->s_var_set) return fail(E, CADM_ERR_STATE, "cadm_session_reset has not been called");
    cudaStream_t s = (cudaStream_t)stream;
    CU(E, cudaMemcpyAsync(E->s_obs, obs_host, (size_t)m * c.obs_dim * sizeof(float), cudaMemcpyHostToDevice, s));
    // plan from the warm start with the constant variance; the history buffers feed the context encoder
    if (int r = cadm_plan_cem(handle, m, E->s_obs, E->s_hobs, E->s_hact, E->s_prev, E->s_var, seed, nullptr, nullptr, nullptr, nullptr,
                              nullptr, nullptr, stream))
        return r;
    CU(E, launch_session_shift(E->mean, E->s_prev, E->s_act, m, c.horizon, c.act_dim, s));
    E->launches++;
    CU(E, cudaMemcpyAsync(action_host, E->s_act, (size_t)m * c.act_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(E, cudaStreamSynchronize(s));
    return CADM_OK;
}

int cadm_session_observe(void* handle, int32_t m, const float* next_obs_host, const uint8_t* done_host, int32_t state_diff,
                         void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (m < 1 || m > c.m_max || !next_obs_host || state_diff < 0 || state_diff > 2) return fail(E, CADM_ERR_ARG, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int K = c.ctx_dim > 0 ? c.hist_len : 1;
    CU(E, cudaMemcpyAsync(E->s_next, next_obs_host, (size_t)m * c.obs_dim * sizeof(float), cudaMemcpyHostToDevice, s));
    if (done_host) CU(E, cudaMemcpyAsync(E->s_mask, done_host, m, cudaMemcpyHostToDevice, s));
    CU(E, launch_session_observe(E->s_obs, E->s_next, E->s_act, done_host ? E->s_mask : nullptr, E->s_hobs, E->s_hact, E->s_counts,
                                 E->s_prev, m, c.obs_dim, c.act_dim, K, E->hA, state_diff, s));
    E->launches++;
    return CADM_OK;
}

int cadm_session_state(void* handle, int32_t m, float* prev_sol_host, float* hist_obs_host, float* hist_act_host,
                       int32_t* counts_host, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (m < 1 || m > c.m_max) return fail(E, CADM_ERR_ARG, "m out of range (1..m_max)");
    cudaStream_t s = (cudaStream_t)stream;
    const int K = c.ctx_dim > 0 ? c.hist_len : 1;
    if (prev_sol_host) CU(E, cudaMemcpyAsync(prev_sol_host, E->s_prev, (size_t)m * E->hA * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (hist_obs_host) CU(E, cudaMemcpyAsync(hist_obs_host, E->s_hobs, (size_t)m * c.obs_dim * K * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (hist_act_host) CU(E, cudaMemcpyAsync(hist_act_host, E->s_hact, (size_t)m * c.act_dim * K * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (counts_host) CU(E, cudaMemcpyAsync(counts_host, E->s_counts, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(E, cudaStreamSynchronize(s));
    return CADM_OK;
}

int cadm_plan_rs(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act, uint64_t seed,
                 const float* u, const int32_t* u_int, const float* eps, float* action, int32_t* action_int,
                 float* returns, int32_t* best, void* stream) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (int r = check_ready(E, true)) return r;
    const CadmConfig& c = E->cfg;
    if (c.world != 1) return fail(E, CADM_ERR_UNSUPPORTED, "random shooting is single-rank");
    if (m < 1 || m > c.m_max || !obs) return fail(E, CADM_ERR_ARG, "bad arguments");
    if (c.ctx_dim > 0 && (!cp_obs || !cp_act)) return fail(E, CADM_ERR_ARG, "cp_obs/cp_act are required for a context model");
    cudaStream_t s = (cudaStream_t)stream;
    if (obs != E->obs) CU(E, cudaMemcpyAsync(E->obs, obs, (size_t)m * c.obs_dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
    E->m = m;
    if (c.ctx_dim > 0)
        if (int r = run_encoder(E, m, cp_obs, cp_act, E->ctx, s)) return r;
    SampleParams S{};
    S.m = m; S.n_local = E->n_local; S.n_global = c.candidates; S.n_offset = 0; S.A = c.act_dim; S.it = 0; S.seed = seed; S.env_offset = E->env_offset;
    if (c.discrete) { S.mode = 2; S.hA = c.horizon; S.u_int = u_int; S.actions_int = E->actions_int; }
    else { S.mode = 1; S.hA = E->hA; S.z = u; S.actions = E->actions; }
    CU(E, launch_sample_actions(S, s));
    E->launches++;
    RolloutParams P = base_params(E);
    P.row_mode = kRowsPlanner;
    P.m = m; P.it = 0;                       // RS transposes the context once (core/utils.py:512-513)
    P.rows_per_member = E->q * m * E->n_local;
    P.ctx_mode = c.ctx_dim > 0 ? (c.context_layout == CADM_CTX_MATCHED ? 2 : 1) : 0;
    P.obs0 = E->obs; P.actions = E->actions; P.actions_int = E->actions_int; P.ctx = E->ctx; P.seed = seed; P.eps = eps;
    P.ret_p = E->ret_p;
    if (int r = run_rollout(E, P, s)) return r;
    CU(E, launch_particle_mean(E->ret_p, E->returns_buf, m * E->n_local, c.particles, s));
    RefitParams R{};
    R.m = m; R.n_local = E->n_local; R.n_global = c.candidates; R.n_offset = 0; R.world = 1;
    R.h = c.horizon; R.A = c.act_dim; R.k_elites = 1; R.npad = next_pow2(c.candidates);
    R.returns_buf = E->returns_buf; R.returns_log = E->returns_log; R.mode_rs = 1; R.best = E->best;
    CU(E, launch_refit(R, s));
    CU(E, launch_rs_gather(E->actions, c.discrete ? E->actions_int : nullptr, E->best, m, E->n_local, c.horizon, c.act_dim,
                           action, action_int, s));
    E->launches += 3;
    if (returns) CU(E, cudaMemcpyAsync(returns, E->returns_log, (size_t)m * c.candidates * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (best) CU(E, cudaMemcpyAsync(best, E->best, (size_t)m * sizeof(int), cudaMemcpyDeviceToDevice, s));
    return CADM_OK;
}

int cadm_set_precision(void* handle, int32_t precision) {
    Engine* E = H(handle);
    if (!E) return CADM_ERR_ARG;
    if (precision < CADM_PREC_FP32 || precision > CADM_PREC_TC_1X) return fail(E, CADM_ERR_ARG, "unknown precision mode");
    E->precision = precision;
    return CADM_OK;
}

int cadm_selftest_tc_gemm(const float* X, const float* W, int32_t K, int32_t N, int32_t terms, float* out, void* stream) {
    if (!X || !W || !out || K < 1 || K > 208 || N < 16 || N > 208 || N % 16 || (terms != 1 && terms != 3))
        return fail(nullptr, CADM_ERR_ARG, "selftest: need 1 <= K <= 208, N a multiple of 16 in 16..208, terms 1 or 3");
    cudaStream_t s = (cudaStream_t)stream;
    const int nkb = (K + 15) / 16;
    unsigned char* img = nullptr;
    cudaError_t e = cudaMalloc(&img, (size_t)nkb * N * 64);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    e = launch_pack_tc(img, W, 1, K, N, 0, nkb, N, (long long)nkb * N * 64, 0, 1, s);
    if (e == cudaSuccess) e = launch_tc_gemm_selftest(X, img, K, N, terms, out, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(img);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    return CADM_OK;
}

int cadm_peer_export(void* handle, void* ipc_handle_out) {
    Engine* E = H(handle);
    if (!E || !ipc_handle_out) return CADM_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == CADM_IPC_HANDLE_BYTES, "CADM_IPC_HANDLE_BYTES must match cudaIpcMemHandle_t");
    cudaIpcMemHandle_t h;
    CU(E, cudaIpcGetMemHandle(&h, E->xchg));
    memcpy(ipc_handle_out, &h, sizeof(h));
    return CADM_OK;
}

int cadm_peer_attach(void* handle, const void* ipc_handles, int32_t count) {
    Engine* E = H(handle);
    if (!E || !ipc_handles) return CADM_ERR_ARG;
    const CadmConfig& c = E->cfg;
    if (count != c.world) return fail(E, CADM_ERR_ARG, "expected one IPC handle per rank");
    if (E->in_flight) return fail(E, CADM_ERR_STATE, "cannot attach peers in the middle of a decision");
    std::vector<unsigned char*> tab(c.world, nullptr);
    for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) { tab[r] = E->xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const unsigned char*>(ipc_handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (void* q : E->peer_open) cudaIpcCloseMemHandle(q);
            E->peer_open.clear();
            return fail(E, CADM_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        E->peer_open.push_back(p);
        tab[r] = reinterpret_cast<unsigned char*>(p);
    }
    if (!E->peer_timeout_host) {
        CU(E, cudaHostAlloc(reinterpret_cast<void**>(&E->peer_timeout_host), sizeof(int), cudaHostAllocMapped));
        *E->peer_timeout_host = 0;
        CU(E, cudaHostGetDevicePointer(reinterpret_cast<void**>(&E->peer_timeout_dev), E->peer_timeout_host, 0));
    }
    CU(E, cudaMemcpy(E->peer_tab, tab.data(), sizeof(unsigned char*) * c.world, cudaMemcpyHostToDevice));
    CU(E, cudaMemset(E->xchg + 2 * E->xchg_ret_bytes, 0, 2 * E->xchg_flag_bytes));      // flags: epoch 0
    CU(E, cudaDeviceSynchronize());
    E->epoch = 0;
    E->peers_on = true;
    return CADM_OK;
}

int cadm_peer_enabled(void* handle) { return handle && H(handle)->peers_on ? 1 : 0; }

int cadm_set_option(void* handle, const char* name, int32_t value) {
    Engine* E = H(handle);
    if (!E || !name) return CADM_ERR_ARG;
    const std::string k(name);
    if (k == "tc_variant") {
        if (value < 0 || value > 3) return fail(E, CADM_ERR_ARG, "tc_variant must be 0 (auto), 1 (row tiles), 2 (swapped operands) or 3 (CTA pairs)");
        E->tc_variant = value;
    } else if (k == "tcs_rows") {
        if (value < 0 || value > 64 || value % 16) return fail(E, CADM_ERR_ARG, "tcs_rows must be 0 (pick) or a multiple of 16 up to 64");
        E->tcs_rows = value;
    } else if (k == "tcs_kps") {
        if (value < 1 || value > 4) return fail(E, CADM_ERR_ARG, "tcs_kps must be in 1..4");
        if (E->have_weights) return fail(E, CADM_ERR_STATE, "tcs_kps must be set before cadm_plan_set_weights");
        E->tcs_kps = value;
    } else if (k == "env_offset") {
        if (value < 0) return fail(E, CADM_ERR_ARG, "env_offset must be non-negative");
        if (E->in_flight) return fail(E, CADM_ERR_STATE, "env_offset cannot change in the middle of a decision");
        E->env_offset = value;
    } else if (k == "peer_timeout_ms") {
        if (value < 1) return fail(E, CADM_ERR_ARG, "peer_timeout_ms must be positive");
        E->peer_timeout_cycles = (long long)value * 2000000ll;           // clock64 ticks at ~2 GHz
    } else if (k == "peer_clear_timeout") {
        // after a peer was slow rather than dead: forget the report so that the engine can be used again (the caller has
        // re-synchronised the ranks, e.g. with a process-group barrier)
        if (E->peer_timeout_host) *reinterpret_cast<volatile int*>(E->peer_timeout_host) = 0;
    } else if (k == "trace") {
        E->trace = value != 0;
    } else if (k == "pdl") {
        cadm::g_pdl = value != 0;              // process-wide: programmatic dependent launch of the CEM kernel chain
    } else if (k == "tcs_skew") {
        if (value < 0) return fail(E, CADM_ERR_ARG, "tcs_skew must be non-negative (bits 20+ are diagnostic switches)");
        E->tcs_skew = value;
    } else {
        return fail(E, CADM_ERR_ARG, "unknown option: " + k);
    }
    return CADM_OK;
}

int cadm_selftest_tcs_gemm(const float* X, const float* W, int32_t rows, int32_t K, int32_t N, int32_t kps, int32_t terms,
                           float* out, void* stream) {
    if (!X || !W || !out || rows < 16 || rows > 64 || rows % 16 || K < 1 || K > 208 || N < 1 || N > 256 || kps < 1 || kps > 4 ||
        (terms != 1 && terms != 3))
        return fail(nullptr, CADM_ERR_ARG, "selftest: need rows in {16,32,48,64}, 1 <= K <= 208, 1 <= N <= 256, kps 1..4, terms 1 or 3");
    cudaStream_t s = (cudaStream_t)stream;
    const int nkb = (K + 15) / 16;
    const int Npad = round_up(N, 16);
    unsigned char* img = nullptr;
    const size_t bytes = (size_t)nkb * Npad * 64;
    cudaError_t e = cudaMalloc(&img, bytes);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    e = launch_pack_tcs(img, W, 1, K, N, 0, nkb, Npad, kps, (long long)bytes, 0, 1, s);
    if (e == cudaSuccess) e = launch_tcs_gemm_selftest(X, img, rows, K, N, kps, terms, out, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(img);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    return CADM_OK;
}

int cadm_selftest_tc_rate(int32_t N, int32_t n_mma, int32_t a_lbo, int32_t n_acc, int32_t swapped, int32_t background,
                          int64_t* cycles_host) {
    if (N < 16 || N > 208 || N % 16 || n_mma < 1 || !cycles_host || n_acc < 1 || n_acc > 8 || (n_acc & (n_acc - 1)) ||
        (512 / n_acc) < N)
        return fail(nullptr, CADM_ERR_ARG, "bad arguments");
    long long* d = nullptr;
    unsigned char* src = nullptr;
    cudaError_t e = cudaMalloc(&d, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&src, 1 << 20);
    if (e == cudaSuccess) e = cudaMemset(src, 0x3c, 1 << 20);
    if (e == cudaSuccess) e = cudaMemset(d, 0, 4 * sizeof(long long));
    if (e == cudaSuccess) e = launch_tc_mma_rate(N, n_mma, a_lbo, n_acc, swapped, background, src, d, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 3 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(src);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    return CADM_OK;
}

int cadm_selftest_tcs_rate(int32_t rows, int32_t iters, int32_t R, int32_t kps, int32_t mode, int64_t* cycles_host) {
    if (rows < 16 || rows > 64 || rows % 16 || iters < 1 || R < 16 || R > 128 || R % 16 || kps < 1 || kps > 4 || !cycles_host)
        return fail(nullptr, CADM_ERR_ARG, "bad arguments");
    long long* d = nullptr;
    unsigned char* src = nullptr;
    cudaError_t e = cudaMalloc(&d, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&src, (1 << 20) + 32768);
    if (e == cudaSuccess) e = cudaMemset(src, 0x3c, (1 << 20) + 32768);
    if (e == cudaSuccess) e = cudaMemset(d, 0, 4 * sizeof(long long));
    if (e == cudaSuccess) e = launch_tcs_mma_rate(rows, iters, R, kps, mode, src, d, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 3 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(src);
    if (e != cudaSuccess) return fail(nullptr, CADM_ERR_CUDA, cudaGetErrorString(e));
    return CADM_OK;
}

int cadm_debug_trace(void* handle, int64_t* out_host, int32_t count) {
    Engine* E = H(handle);
    if (!E || !out_host || count < 1 || count > 64 * 64) return CADM_ERR_ARG;
    CU(E, cudaDeviceSynchronize());
    CU(E, cudaMemcpy(out_host, E->dbg, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost));
    return CADM_OK;
}

int64_t cadm_launch_count(void* handle) { return handle ? H(handle)->launches : 0; }
const char* cadm_kernel_name(void* handle) { return handle ? H(handle)->kernel_name : ""; }

int cadm_set_timing(void* handle, int32_t on) {
    if (!handle) return CADM_ERR_ARG;
    H(handle)->timing = on != 0;
    return CADM_OK;
}

float cadm_last_rollout_ms(void* handle) {
    Engine* E = H(handle);
    if (!E || E->ev_used == 0) return 0.f;
    float total = 0.f;
    for (int i = 0; i + 1 < E->ev_used; i += 2) {
        if (cudaEventSynchronize(E->ev[i + 1]) != cudaSuccess) return -1.f;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, E->ev[i], E->ev[i + 1]) != cudaSuccess) return -1.f;
        total += ms;
    }
    return total;
}

}  // extern "C"
