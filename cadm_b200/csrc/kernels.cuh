// Parameter blocks and launchers of the engine's kernels (definitions in cem_kernels.cu / rollout_*.cu).
#pragma once
#include "common.cuh"

namespace cadm {

struct SampleParams {
    int m, n_local, n_global, n_offset, hA, A;
    int env_offset;           // added to the environment index in the Philox counters
    int it;
    int mode;                 // 0: CEM truncated normal, 1: uniform(-1,1), 2: discrete uniform ints (hA = h)
    unsigned long long seed;
    const float* mean;        // [m, hA]
    const float* var;         // [m, hA]
    const float* z;           // nullable, [m, n_global, hA] of this iteration (CEM) / [m, n_local, hA] explicit u (RS)
    const int* u_int;         // nullable, RS discrete explicit draws [m, n_local, h]
    float* actions;           // [m, n_local, hA]
    int* actions_int;         // [m, n_local, h]
};

struct RefitParams {
    int m, n_local, n_global, n_offset, world;
    int h, A, k_elites, it;
    int env_offset;                 // added to the environment index in the Philox counters
    int npad;                       // power of two >= n_global
    float alpha;
    unsigned long long seed;
    const float* returns_buf;       // [world, m, n_local]
    const float* actions;           // this rank's candidates [m, n_local, hA]
    const float* z;                 // nullable: injected draws of this iteration [m, n_global, hA]
    float* mean;                    // [m, hA] in/out
    float* var;                     // [m, hA] in/out
    float* returns_log;             // nullable, [m, n_global] slot of this iteration
    int* elites_log;                // nullable, [m, k] slot of this iteration
    const float* ret_p;             // nullable: [m, n_global, p] particle returns (world == 1: particle mean folded in)
    int p;
    int stage_elites;               // set by the launcher: elite rows fit in shared memory
    // fused peer-memory all-gather (multi-GPU; nullable): the refit kernel itself averages this rank's particle returns,
    // stores the slice into every rank's exchange block and waits for the other ranks' slices (cem_kernels.cu)
    unsigned char* const* peers;    // device array [world]: base of every rank's exchange block as seen from this device
    long long slice_off;            // byte offset of this epoch parity's returns buffer [world, m, n_local] inside a block
    long long flag_off;             // byte offset of this epoch parity's flags [world, m_max] inside a block
    const float* ret_p_local;       // this rank's particle returns [m, n_local, p]
    int rank, m_max;
    int peer_epoch;                 // wait until flag (r, env) >= peer_epoch for every rank r
    long long timeout_cycles;       // bound of that wait (clock64 cycles)
    int* peer_timeout;              // host-mapped word set to 1 + rank-waited-for when the wait gives up (a peer died)
    // random shooting (mode_rs): argmax only
    int mode_rs;
    int* best;                      // [m]
};

struct EncoderParams {
    int m, E, D, A, K, C;
    int n_layers;                   // hidden + output
    int dims[6];                    // in, h0, h1, h2, C
    const float* W[5];              // [E, in, out]
    const float* b[5];              // [E, 1, out]
    const float* cp_obs;            // [m, D*K]
    const float* cp_act;            // [m, A*K]
    const float* cpo_mean; const float* cpo_std;   // [D*K]
    const float* cpa_mean; const float* cpa_std;   // [A*K]
    float* ctx;                     // [E, m, C]
};

cudaError_t launch_sample_actions(const SampleParams& S, cudaStream_t stream);
cudaError_t launch_particle_mean(const float* ret_p, float* out, int count, int p, cudaStream_t stream);
cudaError_t launch_refit(RefitParams R, cudaStream_t stream);
cudaError_t launch_rs_gather(const float* actions, const int* actions_int, const int* best, int m, int n_local, int h,
                             int A, float* action, int* action_int, cudaStream_t stream);
cudaError_t launch_encoder(const EncoderParams& Q, cudaStream_t stream);
// sampler-side state kept on the device (cadm/samplers/sampler.py:49-57,107-120,164-195)
cudaError_t launch_session_shift(const float* mean, float* prev_sol, float* action, int m, int h, int A, cudaStream_t stream);
cudaError_t launch_session_observe(const float* obs, const float* next_obs, const float* action, const unsigned char* done,
                                   float* hist_obs, float* hist_act, int* counts, float* prev_sol, int m, int D, int A, int K, int hA,
                                   int state_diff, cudaStream_t stream);
cudaError_t launch_session_reset(const unsigned char* mask, float* hist_obs, float* hist_act, int* counts, float* prev_sol, int m,
                                 int DK, int AK, int hA, cudaStream_t stream);
cudaError_t launch_pack_f32(float* dst, const float* src, int E, int in, int out, int Kp, int Np, int col0,
                            long long member_stride, long long layer_off, int clear, cudaStream_t stream);
cudaError_t launch_pack_bias(float* dst, const float* src, int E, int out, int col0, long long bias_stride, long long off,
                             cudaStream_t stream);
cudaError_t launch_rollout_f32(RolloutParams P, int num_sms, cudaStream_t stream, const char** name);
cudaError_t launch_rollout_tc(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms,
                              int num_sms, cudaStream_t stream, const char** name, long long* dbg);
cudaError_t launch_pack_tc(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad,
                           long long member_stride, long long layer_off, int clear, cudaStream_t stream);
cudaError_t launch_rollout_tcs(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms, int kps,
                               int rows_override, int skew, int num_sms, cudaStream_t stream, const char** name, long long* dbg);
// CTA-pair variant of the swapped-operand kernel (rollout_tcp.cu); reads the weight image of rollout_tcs.cu packed with kps = 4
bool tcp_supported(const RolloutParams& P, int kps);
cudaError_t launch_rollout_tcp(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms, int rows_override,
                               int num_sms, cudaStream_t stream, const char** name, long long* dbg);
cudaError_t launch_pack_tcs(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad, int kps,
                            long long member_stride, long long layer_off, int clear, cudaStream_t stream);
cudaError_t launch_tcs_gemm_selftest(const float* X, const unsigned char* wimg, int rows, int K, int Nout, int kps, int terms,
                                     float* out, cudaStream_t stream);
cudaError_t launch_tcs_mma_rate(int rows, int iters, int R, int kps, int mode, const unsigned char* src, long long* cycles,
                                cudaStream_t stream);
cudaError_t launch_tc_mma_rate(int N, int n_mma, int a_lbo, int n_acc, int swapped, int bg, const unsigned char* bg_src,
                               long long* cycles, cudaStream_t stream);
cudaError_t launch_tc_gemm_selftest(const float* X, const unsigned char* wimg, int K, int N, int terms, float* out,
                                    cudaStream_t stream);

}  // namespace cadm
