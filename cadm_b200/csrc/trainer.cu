// fit() on the device: the training graph of the reference -- the ensemble dynamics MLP with the Gaussian
// NLL of the soft-bounded log-variance (cadm/dynamics/mlp_ensemble_cem_dynamics.py:86-170, core/utils.py:43-97), and for CaDM
// the context encoder trained end to end through the forward and the deterministic backward model
// (mlp_cadm_ensemble_cem_dynamics.py:108-317, core/utils.py:251-372, 569-624) -- as hand-written forward, backward and Adam
// kernels behind the C ABI (cadm_train_*).  No library GEMMs, no autograd.
//
// Data layout: the training set lives in HBM for the whole fit() (cadm_train_set_dataset); a minibatch is an index matrix
// [E, B] (the reference's bootstrap indices) that a gather kernel turns into normalised network inputs and targets.  All
// parameters, gradients and Adam slots are ONE flat fp32 vector each (layout: cadm_train_param_count in the header), so the
// optimiser is a single elementwise launch.  Every reduction has a fixed order: results do not change from run to run.
//
// Kernels: train_gather (obs_preproc + normalisation + concat), train_gemm (batched over the ensemble, generic strides: the
// same kernel is X W (+ bias, activation), dZ W^T (* activation') and X^T dZ (+ weight decay)), train_colsum (bias gradients),
// train_loss (losses and their gradients w.r.t. the head outputs and max / min logvar), train_adam (tf.train.AdamOptimizer).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace cadm {
namespace train {

enum Act : int { kNone = 0, kSwish = 1, kRelu = 2 };
enum Mode : int { kFwd = 0, kDgrad = 1, kWgrad = 2, kPlain = 3 };

struct GemmArgs {
    const float* A; long long sAe; int sAm, sAk;       // A[e](m, k) at A + e sAe + m sAm + k sAk
    const float* B; long long sBe; int sBk, sBn;       // B[e](k, n)
    float* C; long long sCe; int ldc;                   // C[e](m, n) at C + e sCe + m ldc + n
    int M, N, K;
    int mode, act;
    const float* bias; long long sBiasE;                // kFwd: + bias[e][n]
    float* Z; long long sZe; int ldz;                   // kFwd: pre-activation out (nullable); kDgrad: pre-activation of the layer below
    const float* W; float lam;                          // kWgrad: C = acc + lam W (W laid out like C)
    float beta;                                         // kPlain: C = acc + beta C
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float act_fwd(int act, float z) {
    if (act == kSwish) return z * sigmoidf_(z);
    if (act == kRelu) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ float act_bwd(int act, float z) {
    if (act == kSwish) { const float s = sigmoidf_(z); return s * (1.0f + z * (1.0f - s)); }
    if (act == kRelu) return z > 0.f ? 1.0f : 0.f;
    return 1.0f;
}

// 64 x 64 output tile, 256 threads, 4 x 4 outputs per thread, K tiles of 16 through shared memory.  The operand loaders follow
// whichever index is contiguous in memory, so X W, dZ W^T and X^T dZ all read coalesced.  k ascends in a fixed order.
constexpr int BM = 64, BN = 64, BK = 16;
__global__ void __launch_bounds__(256) train_gemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int e = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const float* A = g.A + (long long)e * g.sAe;
    const float* B = g.B + (long long)e * g.sBe;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_k_fast = g.sAk == 1, b_n_fast = g.sBn == 1;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / 256; ++i) {
            const int idx = tid + i * 256;
            const int mm = a_k_fast ? idx / BK : idx % BM, kk = a_k_fast ? idx % BK : idx / BM;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? A[(long long)m * g.sAm + (long long)k * g.sAk] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / 256; ++i) {
            const int idx = tid + i * 256;
            const int nn = b_n_fast ? idx % BN : idx / BK, kk = b_n_fast ? idx / BN : idx % BK;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? B[(long long)k * g.sBk + (long long)n * g.sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (long long)e * g.sCe;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.mode == kFwd) {
                v += g.bias[(long long)e * g.sBiasE + n];
                if (g.Z) g.Z[(long long)e * g.sZe + (long long)m * g.ldz + n] = v;
                v = act_fwd(g.act, v);
            } else if (g.mode == kDgrad) {
                v *= act_bwd(g.act, g.Z[(long long)e * g.sZe + (long long)m * g.ldz + n]);
            } else if (g.mode == kWgrad) {
                if (g.lam != 0.f) v = fmaf(g.lam, g.W[(long long)e * g.sCe + (long long)m * g.ldc + n], v);
            } else if (g.beta != 0.f) {
                v = fmaf(g.beta, C[(long long)m * g.ldc + n], v);
            }
            C[(long long)m * g.ldc + n] = v;
        }
    }
}

// db[e][n] = sum_b dZ[e][b][n]  (fixed order)
__global__ void train_colsum_kernel(const float* __restrict__ dZ, int B, int N, float* __restrict__ out) {
    const int e = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = dZ + (long long)e * B * N + n;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += p[(long long)b * N];
    out[(long long)e * N + n] = s;
}

struct GatherArgs {
    int env_id, D, P, A, C, K, E, B;
    const int* idx;                                       // [E, B] dataset rows
    const float *obs, *act, *delta, *obs_next, *back_delta, *cp_obs, *cp_act;      // dataset (device), nullable where unused
    const float *obs_mean, *obs_std, *act_mean, *act_std, *d_mean, *d_std;         // [P] [P] [A] [A] [D] [D]
    const float *cpo_mean, *cpo_std, *cpa_mean, *cpa_std, *bd_mean, *bd_std;       // [D K] [D K] [A K] [A K] [D] [D]
    float *x_fwd, *t_fwd;                                 // [E, B, P + A + C] (context columns left alone), [E, B, D]
    float *x_back, *t_back;                               // same from obs_next / back_delta (nullable)
    float* x_enc;                                         // [E, B, (D + A) K] (nullable)
};

// One block per (e, b) row: obs_preproc (cadm/envs/*.py), normalisation (core/utils.py:77-81, 355-362) and the concat.
__global__ void train_gather_kernel(const GatherArgs g) {
    const int row = blockIdx.x;                            // e * B + b
    const long long r = g.idx[row];
    const int In = g.P + g.A + g.C;
    const int t = threadIdx.x;
    const float* o = g.obs + r * g.D;
    for (int k = t; k < g.P; k += blockDim.x)
        g.x_fwd[(long long)row * In + k] = (env_preproc(g.env_id, o, k) - g.obs_mean[k]) / (g.obs_std[k] + 1e-10f);
    for (int k = t; k < g.A; k += blockDim.x)
        g.x_fwd[(long long)row * In + g.P + k] = (g.act[r * g.A + k] - g.act_mean[k]) / (g.act_std[k] + 1e-10f);
    for (int d = t; d < g.D; d += blockDim.x)
        g.t_fwd[(long long)row * g.D + d] = (g.delta[r * g.D + d] - g.d_mean[d]) / (g.d_std[d] + 1e-10f);
    if (g.x_back) {
        const float* on = g.obs_next + r * g.D;
        for (int k = t; k < g.P; k += blockDim.x)
            g.x_back[(long long)row * In + k] = (env_preproc(g.env_id, on, k) - g.obs_mean[k]) / (g.obs_std[k] + 1e-10f);
        for (int k = t; k < g.A; k += blockDim.x)
            g.x_back[(long long)row * In + g.P + k] = (g.act[r * g.A + k] - g.act_mean[k]) / (g.act_std[k] + 1e-10f);
        for (int d = t; d < g.D; d += blockDim.x)
            g.t_back[(long long)row * g.D + d] = (g.back_delta[r * g.D + d] - g.bd_mean[d]) / (g.bd_std[d] + 1e-10f);
    }
    if (g.x_enc) {
        const int no = g.D * g.K, na = g.A * g.K;
        for (int k = t; k < no; k += blockDim.x)
            g.x_enc[(long long)row * (no + na) + k] = (g.cp_obs[r * no + k] - g.cpo_mean[k]) / (g.cpo_std[k] + 1e-10f);
        for (int k = t; k < na; k += blockDim.x)
            g.x_enc[(long long)row * (no + na) + no + k] = (g.cp_act[r * na + k] - g.cpa_mean[k]) / (g.cpa_std[k] + 1e-10f);
    }
}

// ctx [rows, C] -> columns [col0, col0 + C) of one or two [rows, ld] inputs
__global__ void train_scatter_ctx_kernel(const float* __restrict__ ctx, long long rows, int C, int ld, int col0, float* x0, float* x1) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float v = ctx[i];
    x0[r * ld + col0 + c] = v;
    if (x1) x1[r * ld + col0 + c] = v;
}

__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float s = 0.f;
    if (w == 0) {
        s = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    return s;                                              // valid in warp 0
}

// The scalars of mlp_ensemble_cem_dynamics.py:150-167 / mlp_cadm_ensemble_cem_dynamics.py:266-314 for one model and their
// gradients.  heads [R, 2 D] = [mu | raw logvar] with R = E B rows; every mean is over D, then over B, summed over E, i.e. a
// sum over all elements times 1 / (B D).  ONE block, fixed summation order.
//   prob:  lv1 = max - softplus(max - lv), lv2 = min + softplus(lv1 - min) (core/utils.py:84-90)
//          mu_loss = sum sq exp(-lv2) w, var_loss = sum lv2 w;  det (and the backward model): mse only
//   gscale multiplies the gradient w.r.t. mu / logvar (back_coeff for the backward model)
//   out[0] = mse, out[1] = mu_loss, out[2] = var_loss;  dmax / dmin (nullable) receive the data term + / - 0.01 (reg_loss)
__global__ void __launch_bounds__(1024) train_loss_kernel(const float* __restrict__ heads, const float* __restrict__ target,
                                                          const float* __restrict__ maxlv, const float* __restrict__ minlv, long long R,
                                                          int B, int D, int deterministic, float gscale, float* __restrict__ dheads,
                                                          float* __restrict__ dmax, float* __restrict__ dmin, float* __restrict__ out) {
    __shared__ float red[32];
    const float w = 1.0f / ((float)B * (float)D);
    float s_mse = 0.f, s_mu = 0.f, s_var = 0.f;
    const long long total = R * D;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        const long long r = i / D;
        const int d = (int)(i - r * D);
        const float mu = heads[r * 2 * D + d];
        const float diff = mu - target[i];
        const float sq = diff * diff;
        s_mse += sq;
        if (deterministic) {
            if (dheads) { dheads[r * 2 * D + d] = gscale * 2.0f * diff * w; dheads[r * 2 * D + D + d] = 0.f; }
        } else {
            const float lv = heads[r * 2 * D + D + d];
            const float a = maxlv[d] - lv;
            const float lv1 = maxlv[d] - softplusf_(a);
            const float c = lv1 - minlv[d];
            const float lv2 = minlv[d] + softplusf_(c);
            const float inv = expf(-lv2);
            s_mu += sq * inv;
            s_var += lv2;
            if (dheads) {
                const float dlv2 = (1.0f - sq * inv) * w;
                dheads[r * 2 * D + d] = gscale * 2.0f * diff * inv * w;
                dheads[r * 2 * D + D + d] = gscale * dlv2 * sigmoidf_(c) * sigmoidf_(a);
            }
        }
    }
    const float t0 = block_sum(s_mse, red);
    const float t1 = block_sum(s_mu, red);
    const float t2 = block_sum(s_var, red);
    if (threadIdx.x == 0) { out[0] = t0 * w; out[1] = t1 * w; out[2] = t2 * w; }
    if (dmax == nullptr) return;
    // d loss / d max_logvar[d], d min_logvar[d]: one warp per state dimension, lanes stride over the rows
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int d = wp; d < D; d += (int)(blockDim.x >> 5)) {
        float gmax = 0.f, gmin = 0.f;
        if (!deterministic) {
            for (long long r = lane; r < R; r += 32) {
                const float mu = heads[r * 2 * D + d], lv = heads[r * 2 * D + D + d];
                const float diff = mu - target[r * D + d];
                const float a = maxlv[d] - lv;
                const float lv1 = maxlv[d] - softplusf_(a);
                const float c = lv1 - minlv[d];
                const float lv2 = minlv[d] + softplusf_(c);
                const float dlv2 = (1.0f - diff * diff * expf(-lv2)) * w;
                const float s1 = sigmoidf_(a), s2 = sigmoidf_(c);
                gmax += dlv2 * s2 * (1.0f - s1);
                gmin += dlv2 * (1.0f - s2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { gmax += __shfl_xor_sync(0xffffffffu, gmax, o); gmin += __shfl_xor_sync(0xffffffffu, gmin, o); }
        }
        if (lane == 0) {
            dmax[d] = deterministic ? 0.f : gmax + 0.01f;
            dmin[d] = deterministic ? 0.f : gmin - 0.01f;
        }
    }
}

// tf.train.AdamOptimizer (TensorFlow 1.x, "epsilon hat" form): m <- b1 m + (1 - b1) g; v <- b2 v + (1 - b2) g^2;
// theta <- theta - lr_t m / (sqrt(v) + eps) with lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) computed on the host in double.
__global__ void train_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                                  long long n, float lr_t, float b1, float b2, float eps) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace train
}  // namespace cadm

// ======================================================================================================================
// host side: handle, buffers, launch sequencing
// ======================================================================================================================
using namespace cadm;
using namespace cadm::train;

namespace {

thread_local std::string g_train_create_error;

struct LayerD {
    int in = 0, out = 0, act = kNone;
    long long w_off = 0, b_off = 0;
    float decay = 0.f;
};

struct NetD {
    std::vector<LayerD> L;
    int in_dim = 0;
    float* X0 = nullptr;                  // [E, B, in_dim]
    std::vector<float*> Z, Aout;          // per layer [E, B, out]
    float* T = nullptr;                   // targets [E, B, D] (dynamics nets)
    float* dHeads = nullptr;              // gradient w.r.t. the last layer's output
};

struct Trainer {
    CadmTrainConfig cfg{};
    std::string err;
    int device = 0;
    int E = 0, D = 0, P = 0, A = 0, C = 0, K = 0;
    bool has_enc = false, has_back = false;
    NetD enc, fwd, back;
    long long n_params = 0, maxlv_off = 0, minlv_off = 0;
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    long long t = 0;
    // normalisation (device)
    float* stats[12] = {nullptr};
    bool have_norm = false;
    // datasets (device): 0 train, 1 valid
    struct DS { long long rows = 0, cap = 0; float *obs = nullptr, *act = nullptr, *delta = nullptr, *obs_next = nullptr, *back_delta = nullptr, *cp_obs = nullptr, *cp_act = nullptr; } ds[2];
    // workspace
    int Bcap = 0;
    int* idx_dev = nullptr;
    int* idx_pin = nullptr;
    float *dz0 = nullptr, *dz1 = nullptr;     // ping-pong gradient buffers [E, B, widest layer]
    float* dctx = nullptr;                    // [E, B, C]
    float* losses_dev = nullptr;              // [8]
    float* losses_pin = nullptr;
    int widest = 0;
    long long launches = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
};

int tfail(Trainer* T, int code, const std::string& msg) {
    if (T) T->err = msg;
    else g_train_create_error = msg;
    return code;
}

#define TCU(T, call)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return tfail(T, CADM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));       \
    } while (0)

Trainer* TH(void* h) { return reinterpret_cast<Trainer*>(h); }

// Every entry point works on the device the handle was created on, whatever the caller's current device is, and leaves the
// caller's device as it found it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

cudaError_t tmalloc(float** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(float));
    if (e != cudaSuccess) return e;
    *p = reinterpret_cast<float*>(q);
    return cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(float));
}

void add_layer(Trainer* T, NetD& net, int in, int out, int act, float decay) {
    LayerD l;
    l.in = in; l.out = out; l.act = act; l.decay = decay;
    l.w_off = T->n_params; T->n_params += (long long)T->E * in * out;
    l.b_off = T->n_params; T->n_params += (long long)T->E * out;
    net.L.push_back(l);
    T->widest = std::max(T->widest, std::max(in, out));
}

void free_net_ws(NetD& n) {
    cudaFree(n.X0); n.X0 = nullptr;
    for (float* p : n.Z) cudaFree(p);
    for (float* p : n.Aout) cudaFree(p);
    n.Z.clear(); n.Aout.clear();
    cudaFree(n.T); n.T = nullptr;
    cudaFree(n.dHeads); n.dHeads = nullptr;
}

cudaError_t alloc_net_ws(Trainer* T, NetD& n, int B, bool dynamics) {
    cudaError_t e;
    const size_t rows = (size_t)T->E * B;
    if ((e = tmalloc(&n.X0, rows * n.in_dim)) != cudaSuccess) return e;
    for (const LayerD& l : n.L) {
        float *z = nullptr, *a = nullptr;
        if ((e = tmalloc(&z, rows * l.out)) != cudaSuccess) return e;
        n.Z.push_back(z);
        if ((e = tmalloc(&a, rows * l.out)) != cudaSuccess) return e;
        n.Aout.push_back(a);
    }
    if (dynamics) {
        if ((e = tmalloc(&n.T, rows * T->D)) != cudaSuccess) return e;
        if ((e = tmalloc(&n.dHeads, rows * 2 * T->D)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int ensure_ws(Trainer* T, int B) {
    if (B <= T->Bcap) return CADM_OK;
    free_net_ws(T->enc); free_net_ws(T->fwd); free_net_ws(T->back);
    cudaFree(T->dz0); cudaFree(T->dz1); cudaFree(T->dctx); cudaFree(T->idx_dev);
    if (T->idx_pin) cudaFreeHost(T->idx_pin);
    T->dz0 = T->dz1 = T->dctx = nullptr; T->idx_dev = nullptr; T->idx_pin = nullptr;
    T->Bcap = 0;
    const size_t rows = (size_t)T->E * B;
    TCU(T, alloc_net_ws(T, T->fwd, B, true));
    if (T->has_back) TCU(T, alloc_net_ws(T, T->back, B, true));
    if (T->has_enc) TCU(T, alloc_net_ws(T, T->enc, B, false));
    TCU(T, tmalloc(&T->dz0, rows * T->widest));
    TCU(T, tmalloc(&T->dz1, rows * T->widest));
    TCU(T, tmalloc(&T->dctx, rows * std::max(T->C, 1)));
    TCU(T, cudaMalloc(reinterpret_cast<void**>(&T->idx_dev), rows * sizeof(int)));
    TCU(T, cudaMallocHost(reinterpret_cast<void**>(&T->idx_pin), rows * sizeof(int)));
    T->Bcap = B;
    return CADM_OK;
}

void launch_gemm(Trainer* T, const GemmArgs& g) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, T->E);
    train_gemm_kernel<<<grid, 256, 0, T->stream>>>(g);
    ++T->launches;
}

// Y = act(X W + b) for every layer of `net` on B rows per member
void net_forward(Trainer* T, NetD& n, int B) {
    const float* x = n.X0;
    int ldx = n.in_dim;
    for (size_t i = 0; i < n.L.size(); ++i) {
        const LayerD& l = n.L[i];
        GemmArgs g{};
        g.A = x; g.sAe = (long long)B * ldx; g.sAm = ldx; g.sAk = 1;
        g.B = T->params + l.w_off; g.sBe = (long long)l.in * l.out; g.sBk = l.out; g.sBn = 1;
        g.C = n.Aout[i]; g.sCe = (long long)B * l.out; g.ldc = l.out;
        g.M = B; g.N = l.out; g.K = l.in;
        g.mode = kFwd; g.act = l.act;
        g.bias = T->params + l.b_off; g.sBiasE = l.out;
        g.Z = l.act != kNone ? n.Z[i] : nullptr; g.sZe = (long long)B * l.out; g.ldz = l.out;
        launch_gemm(T, g);
        x = n.Aout[i];
        ldx = l.out;
    }
}

// Gradients of every layer of `net` given dOut = d loss / d (last layer output) [E, B, out_last].  When dctx is given, the
// gradient w.r.t. the input columns [ctx_col0, ctx_col0 + C) is written (beta = 0) or added (beta = 1) to it.
void net_backward(Trainer* T, NetD& n, int B, const float* dOut, float* dctx, int ctx_col0, float beta) {
    const float* dz = dOut;
    float* bufs[2] = {T->dz0, T->dz1};
    int which = 0;
    for (int i = (int)n.L.size() - 1; i >= 0; --i) {
        const LayerD& l = n.L[i];
        const float* xin = i == 0 ? n.X0 : n.Aout[i - 1];
        const int ldx = i == 0 ? n.in_dim : n.L[i - 1].out;
        {   // dW = X^T dZ + coeff decay W
            GemmArgs g{};
            g.A = xin; g.sAe = (long long)B * ldx; g.sAm = 1; g.sAk = ldx;
            g.B = dz; g.sBe = (long long)B * l.out; g.sBk = l.out; g.sBn = 1;
            g.C = T->grads + l.w_off; g.sCe = (long long)l.in * l.out; g.ldc = l.out;
            g.M = l.in; g.N = l.out; g.K = B;
            g.mode = kWgrad; g.W = T->params + l.w_off; g.lam = T->cfg.weight_decay_coeff * l.decay;
            launch_gemm(T, g);
        }
        train_colsum_kernel<<<dim3((l.out + 127) / 128, T->E), 128, 0, T->stream>>>(dz, B, l.out, T->grads + l.b_off);
        ++T->launches;
        if (i > 0) {   // dZ_below = (dZ W^T) * act'(Z_below)
            const LayerD& lb = n.L[i - 1];
            GemmArgs g{};
            g.A = dz; g.sAe = (long long)B * l.out; g.sAm = l.out; g.sAk = 1;
            g.B = T->params + l.w_off; g.sBe = (long long)l.in * l.out; g.sBk = 1; g.sBn = l.out;
            g.C = bufs[which]; g.sCe = (long long)B * l.in; g.ldc = l.in;
            g.M = B; g.N = l.in; g.K = l.out;
            g.mode = kDgrad; g.act = lb.act; g.Z = n.Z[i - 1]; g.sZe = (long long)B * lb.out; g.ldz = lb.out;
            launch_gemm(T, g);
            dz = bufs[which];
            which ^= 1;
        } else if (dctx != nullptr) {   // d loss / d context = dZ_0 W_0[ctx rows]^T
            GemmArgs g{};
            g.A = dz; g.sAe = (long long)B * l.out; g.sAm = l.out; g.sAk = 1;
            g.B = T->params + l.w_off + (long long)ctx_col0 * l.out; g.sBe = (long long)l.in * l.out; g.sBk = 1; g.sBn = l.out;
            g.C = dctx; g.sCe = (long long)B * T->C; g.ldc = T->C;
            g.M = B; g.N = T->C; g.K = l.out;
            g.mode = kPlain; g.beta = beta;
            launch_gemm(T, g);
        }
    }
}

}  // namespace

extern "C" {

int cadm_train_create(const CadmTrainConfig* cfg, void** handle) {
    if (!cfg || !handle) return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: null argument");
    if (cfg->struct_size != (int32_t)sizeof(CadmTrainConfig)) return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: struct_size mismatch");
    if (cfg->ensemble < 1 || cfg->obs_dim < 1 || cfg->obs_dim > kMaxObs || cfg->act_dim < 1 || cfg->proc_obs_dim < 1 || cfg->hidden < 1 ||
        cfg->n_hidden < 1 || cfg->n_hidden > 7 || cfg->ctx_dim < 0)
        return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: bad dimensions");
    if (cfg->ctx_dim > 0 && (cfg->hist_len < 1 || cfg->enc_hidden[0] < 1)) return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: context encoder needs hist_len and enc_hidden");
    if (cfg->has_back && cfg->ctx_dim == 0) return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: the backward model belongs to the CaDM model (ctx_dim > 0)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return tfail(nullptr, CADM_ERR_CUDA, "cadm_train_create: no CUDA device (there is no CPU fallback)");
    Trainer* T = new (std::nothrow) Trainer();
    if (!T) return tfail(nullptr, CADM_ERR_ARG, "cadm_train_create: out of memory");
    T->cfg = *cfg;
    cudaGetDevice(&T->device);
    T->E = cfg->ensemble; T->D = cfg->obs_dim; T->P = cfg->proc_obs_dim; T->A = cfg->act_dim; T->C = cfg->ctx_dim; T->K = cfg->hist_len;
    T->has_enc = cfg->ctx_dim > 0;
    T->has_back = cfg->has_back != 0;
    // flat parameter layout: encoder | forward model | max_logvar, min_logvar | backward model
    if (T->has_enc) {
        int in = (T->D + T->A) * T->K;
        T->enc.in_dim = in;
        int li = 0;
        for (; li < 3 && cfg->enc_hidden[li] > 0; ++li) { add_layer(T, T->enc, in, cfg->enc_hidden[li], kRelu, cfg->context_weight_decays[li]); in = cfg->enc_hidden[li]; }
        add_layer(T, T->enc, in, T->C, kNone, cfg->context_weight_decays[li]);
    }
    const int In = T->P + T->A + T->C;
    auto dynamics = [&](NetD& n) {
        n.in_dim = In;
        int in = In;
        for (int i = 0; i < cfg->n_hidden; ++i) { add_layer(T, n, in, cfg->hidden, kSwish, cfg->weight_decays[i]); in = cfg->hidden; }
        add_layer(T, n, in, 2 * T->D, kNone, cfg->weight_decays[cfg->n_hidden]);      // [output_mu | output_logvar]
    };
    dynamics(T->fwd);
    T->maxlv_off = T->n_params; T->n_params += T->D;
    T->minlv_off = T->n_params; T->n_params += T->D;
    if (T->has_back) dynamics(T->back);
    cudaError_t e = cudaSuccess;
    if ((e = tmalloc(&T->params, T->n_params)) != cudaSuccess || (e = tmalloc(&T->grads, T->n_params)) != cudaSuccess ||
        (e = tmalloc(&T->adam_m, T->n_params)) != cudaSuccess || (e = tmalloc(&T->adam_v, T->n_params)) != cudaSuccess ||
        (e = tmalloc(&T->losses_dev, 8)) != cudaSuccess || (e = cudaMallocHost(reinterpret_cast<void**>(&T->losses_pin), 8 * sizeof(float))) != cudaSuccess) {
        std::string msg = std::string("cadm_train_create: ") + cudaGetErrorString(e);
        delete T;
        return tfail(nullptr, CADM_ERR_CUDA, msg);
    }
    const int sizes[12] = {T->P, T->P, T->A, T->A, T->D, T->D, T->D * T->K, T->D * T->K, T->A * T->K, T->A * T->K, T->D, T->D};
    for (int i = 0; i < 12; ++i)
        if ((e = tmalloc(&T->stats[i], std::max(sizes[i], 1))) != cudaSuccess) { delete T; return tfail(nullptr, CADM_ERR_CUDA, "cadm_train_create: cudaMalloc failed"); }
    *handle = T;
    return CADM_OK;
}

int cadm_train_destroy(void* handle) {
    Trainer* T = TH(handle);
    if (!T) return CADM_OK;
    DeviceGuard guard(T->device);
    cudaDeviceSynchronize();
    free_net_ws(T->enc); free_net_ws(T->fwd); free_net_ws(T->back);
    cudaFree(T->dz0); cudaFree(T->dz1); cudaFree(T->dctx); cudaFree(T->idx_dev);
    if (T->idx_pin) cudaFreeHost(T->idx_pin);
    cudaFree(T->params); cudaFree(T->grads); cudaFree(T->adam_m); cudaFree(T->adam_v); cudaFree(T->losses_dev);
    if (T->losses_pin) cudaFreeHost(T->losses_pin);
    for (float* p : T->stats) cudaFree(p);
    for (auto& d : T->ds) { cudaFree(d.obs); cudaFree(d.act); cudaFree(d.delta); cudaFree(d.obs_next); cudaFree(d.back_delta); cudaFree(d.cp_obs); cudaFree(d.cp_act); }
    delete T;
    return CADM_OK;
}

const char* cadm_train_last_error(const void* handle) {
    return handle ? reinterpret_cast<const Trainer*>(handle)->err.c_str() : g_train_create_error.c_str();
}

int64_t cadm_train_param_count(void* handle) { return handle ? TH(handle)->n_params : 0; }
int64_t cadm_train_launch_count(void* handle) { return handle ? TH(handle)->launches : 0; }

int cadm_train_set_params(void* handle, const float* flat_host, int64_t n) {
    Trainer* T = TH(handle);
    if (!T || !flat_host || n != T->n_params) return tfail(T, CADM_ERR_ARG, "cadm_train_set_params: wrong length");
    DeviceGuard guard(T->device);
    TCU(T, cudaMemcpy(T->params, flat_host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    return CADM_OK;
}

int cadm_train_get_params(void* handle, float* flat_host, int64_t n) {
    Trainer* T = TH(handle);
    if (!T || !flat_host || n != T->n_params) return tfail(T, CADM_ERR_ARG, "cadm_train_get_params: wrong length");
    DeviceGuard guard(T->device);
    TCU(T, cudaStreamSynchronize(T->stream));
    TCU(T, cudaMemcpy(flat_host, T->params, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return CADM_OK;
}

int cadm_train_get_grads(void* handle, float* flat_host, int64_t n) {
    Trainer* T = TH(handle);
    if (!T || !flat_host || n != T->n_params) return tfail(T, CADM_ERR_ARG, "cadm_train_get_grads: wrong length");
    DeviceGuard guard(T->device);
    TCU(T, cudaStreamSynchronize(T->stream));
    TCU(T, cudaMemcpy(flat_host, T->grads, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return CADM_OK;
}

int cadm_train_adam_state(void* handle, int32_t set, float* m_host, float* v_host, int64_t n, int64_t* t_inout) {
    Trainer* T = TH(handle);
    if (!T || !m_host || !v_host || !t_inout || n != T->n_params) return tfail(T, CADM_ERR_ARG, "cadm_train_adam_state: wrong length");
    DeviceGuard guard(T->device);
    TCU(T, cudaStreamSynchronize(T->stream));
    const size_t bytes = (size_t)n * sizeof(float);
    if (set) {
        TCU(T, cudaMemcpy(T->adam_m, m_host, bytes, cudaMemcpyHostToDevice));
        TCU(T, cudaMemcpy(T->adam_v, v_host, bytes, cudaMemcpyHostToDevice));
        T->t = *t_inout;
    } else {
        TCU(T, cudaMemcpy(m_host, T->adam_m, bytes, cudaMemcpyDeviceToHost));
        TCU(T, cudaMemcpy(v_host, T->adam_v, bytes, cudaMemcpyDeviceToHost));
        *t_inout = T->t;
    }
    return CADM_OK;
}

int cadm_train_set_norm(void* handle, const float* const* stats_host, int32_t count) {
    Trainer* T = TH(handle);
    if (!T || !stats_host) return tfail(T, CADM_ERR_ARG, "cadm_train_set_norm: null argument");
    DeviceGuard guard(T->device);
    const int need = T->has_back ? 12 : (T->has_enc ? 10 : 6);
    if (count < need) return tfail(T, CADM_ERR_ARG, "cadm_train_set_norm: not enough statistics for this model");
    const int sizes[12] = {T->P, T->P, T->A, T->A, T->D, T->D, T->D * T->K, T->D * T->K, T->A * T->K, T->A * T->K, T->D, T->D};
    for (int i = 0; i < need; ++i) {
        if (!stats_host[i]) return tfail(T, CADM_ERR_ARG, "cadm_train_set_norm: missing vector");
        TCU(T, cudaMemcpy(T->stats[i], stats_host[i], (size_t)sizes[i] * sizeof(float), cudaMemcpyHostToDevice));
    }
    T->have_norm = true;
    return CADM_OK;
}

int cadm_train_set_dataset(void* handle, int32_t which, int64_t rows, const float* obs_host, const float* act_host, const float* delta_host,
                           const float* obs_next_host, const float* back_delta_host, const float* cp_obs_host, const float* cp_act_host) {
    Trainer* T = TH(handle);
    if (!T || which < 0 || which > 1 || rows < 0) return tfail(T, CADM_ERR_ARG, "cadm_train_set_dataset: bad argument");
    DeviceGuard guard(T->device);
    if (rows > 0 && (!obs_host || !act_host || !delta_host)) return tfail(T, CADM_ERR_ARG, "cadm_train_set_dataset: obs / act / delta are required");
    if (rows > 0 && T->has_enc && (!cp_obs_host || !cp_act_host)) return tfail(T, CADM_ERR_ARG, "cadm_train_set_dataset: the CaDM model needs cp_obs / cp_act");
    if (rows > 0 && T->has_back && (!obs_next_host || !back_delta_host)) return tfail(T, CADM_ERR_ARG, "cadm_train_set_dataset: the backward model needs obs_next / back_delta");
    TCU(T, cudaStreamSynchronize(T->stream));
    Trainer::DS& d = T->ds[which];
    if (rows > d.cap) {
        cudaFree(d.obs); cudaFree(d.act); cudaFree(d.delta); cudaFree(d.obs_next); cudaFree(d.back_delta); cudaFree(d.cp_obs); cudaFree(d.cp_act);
        d = Trainer::DS();
        const size_t cap = (size_t)rows + (size_t)rows / 4;          // the reference's dataset grows every iteration
        TCU(T, tmalloc(&d.obs, cap * T->D));
        TCU(T, tmalloc(&d.act, cap * T->A));
        TCU(T, tmalloc(&d.delta, cap * T->D));
        if (T->has_back) { TCU(T, tmalloc(&d.obs_next, cap * T->D)); TCU(T, tmalloc(&d.back_delta, cap * T->D)); }
        if (T->has_enc) { TCU(T, tmalloc(&d.cp_obs, cap * T->D * T->K)); TCU(T, tmalloc(&d.cp_act, cap * T->A * T->K)); }
        d.cap = (long long)cap;
    }
    d.rows = rows;
    auto up = [&](float* dst, const float* src, size_t n) { return n == 0 ? cudaSuccess : cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyHostToDevice); };
    TCU(T, up(d.obs, obs_host, (size_t)rows * T->D));
    TCU(T, up(d.act, act_host, (size_t)rows * T->A));
    TCU(T, up(d.delta, delta_host, (size_t)rows * T->D));
    if (T->has_back) { TCU(T, up(d.obs_next, obs_next_host, (size_t)rows * T->D)); TCU(T, up(d.back_delta, back_delta_host, (size_t)rows * T->D)); }
    if (T->has_enc) { TCU(T, up(d.cp_obs, cp_obs_host, (size_t)rows * T->D * T->K)); TCU(T, up(d.cp_act, cp_act_host, (size_t)rows * T->A * T->K)); }
    return CADM_OK;
}

int cadm_train_step(void* handle, int32_t which, const int32_t* idx_host, int32_t B, int32_t train, float* losses_host) {
    Trainer* T = TH(handle);
    if (!T || which < 0 || which > 1 || !idx_host || B < 1 || !losses_host) return tfail(T, CADM_ERR_ARG, "cadm_train_step: bad argument");
    if (!T->have_norm) return tfail(T, CADM_ERR_STATE, "cadm_train_step: cadm_train_set_norm first");
    Trainer::DS& d = T->ds[which];
    if (d.rows < 1) return tfail(T, CADM_ERR_STATE, "cadm_train_step: cadm_train_set_dataset first");
    DeviceGuard guard(T->device);
    const long long R = (long long)T->E * B;
    for (long long i = 0; i < R; ++i)
        if (idx_host[i] < 0 || idx_host[i] >= d.rows) return tfail(T, CADM_ERR_ARG, "cadm_train_step: index outside the dataset");
    int rc = ensure_ws(T, B);
    if (rc != CADM_OK) return rc;
    cudaStream_t s = T->stream;
    std::memcpy(T->idx_pin, idx_host, (size_t)R * sizeof(int));
    TCU(T, cudaMemcpyAsync(T->idx_dev, T->idx_pin, (size_t)R * sizeof(int), cudaMemcpyHostToDevice, s));
    const int In = T->P + T->A + T->C;
    GatherArgs g{};
    g.env_id = T->cfg.env_id; g.D = T->D; g.P = T->P; g.A = T->A; g.C = T->C; g.K = T->K; g.E = T->E; g.B = B;
    g.idx = T->idx_dev;
    g.obs = d.obs; g.act = d.act; g.delta = d.delta; g.obs_next = d.obs_next; g.back_delta = d.back_delta; g.cp_obs = d.cp_obs; g.cp_act = d.cp_act;
    g.obs_mean = T->stats[0]; g.obs_std = T->stats[1]; g.act_mean = T->stats[2]; g.act_std = T->stats[3]; g.d_mean = T->stats[4]; g.d_std = T->stats[5];
    g.cpo_mean = T->stats[6]; g.cpo_std = T->stats[7]; g.cpa_mean = T->stats[8]; g.cpa_std = T->stats[9]; g.bd_mean = T->stats[10]; g.bd_std = T->stats[11];
    g.x_fwd = T->fwd.X0; g.t_fwd = T->fwd.T;
    g.x_back = T->has_back ? T->back.X0 : nullptr; g.t_back = T->has_back ? T->back.T : nullptr;
    g.x_enc = T->has_enc ? T->enc.X0 : nullptr;
    train_gather_kernel<<<(unsigned)R, 64, 0, s>>>(g);
    ++T->launches;
    if (T->has_enc) {
        net_forward(T, T->enc, B);
        const long long n = R * T->C;
        train_scatter_ctx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(T->enc.Aout.back(), R, T->C, In, T->P + T->A, T->fwd.X0,
                                                                              T->has_back ? T->back.X0 : nullptr);
        ++T->launches;
    }
    net_forward(T, T->fwd, B);
    float* maxlv = T->params + T->maxlv_off;
    float* minlv = T->params + T->minlv_off;
    train_loss_kernel<<<1, 1024, 0, s>>>(T->fwd.Aout.back(), T->fwd.T, maxlv, minlv, R, B, T->D, T->cfg.deterministic, 1.0f,
                                         train ? T->fwd.dHeads : nullptr, train ? T->grads + T->maxlv_off : nullptr,
                                         train ? T->grads + T->minlv_off : nullptr, T->losses_dev);
    ++T->launches;
    if (T->has_back) {
        net_forward(T, T->back, B);
        // built with deterministic=True whatever the forward model is (mlp_cadm_ensemble_cem_dynamics.py:228): mse only
        train_loss_kernel<<<1, 1024, 0, s>>>(T->back.Aout.back(), T->back.T, maxlv, minlv, R, B, T->D, 1, T->cfg.back_coeff,
                                             train ? T->back.dHeads : nullptr, nullptr, nullptr, T->losses_dev + 4);
        ++T->launches;
    }
    if (train) {
        net_backward(T, T->fwd, B, T->fwd.dHeads, T->has_enc ? T->dctx : nullptr, T->P + T->A, 0.f);
        if (T->has_back) net_backward(T, T->back, B, T->back.dHeads, T->dctx, T->P + T->A, 1.f);
        if (T->has_enc) net_backward(T, T->enc, B, T->dctx, nullptr, 0, 0.f);
        ++T->t;
        const double b1 = T->cfg.adam_beta1, b2 = T->cfg.adam_beta2;
        const float lr_t = (float)((double)T->cfg.learning_rate * std::sqrt(1.0 - std::pow(b2, (double)T->t)) / (1.0 - std::pow(b1, (double)T->t)));
        train_adam_kernel<<<(unsigned)((T->n_params + 255) / 256), 256, 0, s>>>(T->params, T->adam_m, T->adam_v, T->grads, T->n_params, lr_t,
                                                                                 T->cfg.adam_beta1, T->cfg.adam_beta2, T->cfg.adam_eps);
        ++T->launches;
    }
    TCU(T, cudaMemcpyAsync(T->losses_pin, T->losses_dev, 8 * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCU(T, cudaStreamSynchronize(s));
    TCU(T, cudaGetLastError());
    const float mse = T->losses_pin[0];
    const float back_mse = T->has_back ? T->losses_pin[4] : 0.f;
    float recon = T->cfg.deterministic ? mse : T->losses_pin[1] + T->losses_pin[2];
    if (T->has_back) recon += T->cfg.back_coeff * back_mse;
    losses_host[0] = mse;
    losses_host[1] = recon;
    losses_host[2] = back_mse;
    losses_host[3] = T->cfg.deterministic ? 0.f : T->losses_pin[1];
    return CADM_OK;
}

}  // extern "C"
