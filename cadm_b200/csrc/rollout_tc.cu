// Persistent tensor-core rollout kernel (CADM_PREC_TC_3X / CADM_PREC_TC_1X): tcgen05.mma with accumulators in TMEM.
//
// One CTA per SM; a CTA owns a tile of up to 128 rows (TMEM lanes) of ONE ensemble member and carries it through all
// h horizon steps of a CEM iteration -- state, return accumulator, activations and accumulators never leave the SM.
// Every layer is D[128 x N] = X[128 x K] * W^T[N x K] on the 5th-gen tensor cores (M = 128, N = 208 / 2D padded, K = 16
// per instruction, kind::f16, fp32 accumulate in TMEM).
//
// Precision (CADM_PREC_TC_3X): every fp32 operand is split as x = hi + lo with hi = fp16(x), lo = fp16(x - hi), after
// a power-of-two pre-scale (activations x 8, weights x 64) that keeps typical residuals out of the fp16 subnormal
// range; a product is 3 MMAs into ONE fp32 accumulator:   X_hi W_hi + X_hi W_lo + X_lo W_hi.
// The dropped lo*lo term and the rounding of lo are <= ~2^-21 relative -- fp32-class for the 1e-4 parity bar; the
// scale 512 is removed exactly in the epilogue.  (A and B of one tcgen05.mma must have the SAME format: an f16 x bf16
// instruction traps as illegal on sm_100a, which rules out an fp16-hi / bf16-lo split.)
// CADM_PREC_TC_1X issues only the first MMA (fp16 operands; fast, does not claim the bar).
//
// Warp roles (320 threads):
//   warp 0      weight producer: streams the member's packed weight image from L2 through a shared-memory ring with
//               1-D bulk async copies (TMA engine, mbarrier complete_tx), one stage = one K16 block (hi + lo)
//   warp 1      MMA issuer (one elected lane): waits "activation block ready" + "weight stage full", issues the MMAs,
//               tcgen05.commit frees the weight stage / publishes the accumulator
//   warps 2-9   epilogue: TMEM -> registers (tcgen05.ld, lane = row) -> bias + swish -> hi/lo split -> shared memory in
//               the UMMA K-major layout -> "block ready".  Two warps per TMEM lane quarter split the columns.  The next
//               layer's MMAs start per K16 block while the epilogue is still producing later blocks, and two TMEM
//               accumulators alternate, so the tensor pipe and the epilogue overlap inside one dependent chain.
//   The same warps run the per-step prologue (obs_preproc + normalise + concat, reward) and the final epilogue
//   (denormalise, bounded logvar, Gaussian sample, obs_postproc) -- core/utils.py:141-168 fused as in rollout_f32.cu.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace cadm {

constexpr int kEpiGroups = 4;                  // epilogue warp groups; each group = 4 warps = the 4 TMEM lane quarters
constexpr int kEpiThreads = kEpiGroups * 128;
constexpr int kTcThreads = 64 + kEpiThreads;
constexpr int kTileRows = 128;
constexpr int kMaxKB = 13;                     // K16 blocks of the widest layer (208 / 16)
constexpr int kXChunkBytes = 2048;             // one 8-wide k-chunk of all 128 rows: 128 rows x 16 B
constexpr int kXBytes = 2 * kMaxKB * kXChunkBytes;   // 53248 per operand half (hi or lo)
constexpr int kMaxStageBytes = 2 * 208 * 32;   // hi + lo block of one K16 step, N = 208
constexpr int kBlocksPerStage = 2;             // K16 blocks per ring slot (consecutive in consumption order): half the barrier waits
constexpr int kSlotBytes = kBlocksPerStage * kMaxStageBytes;

struct TcSmem {
    size_t off_xhi, off_xlo, off_w, off_s, off_bias, off_vec, off_rowi, off_feat, off_zero, off_act, off_ctx, off_nz, off_bar, total;
    int stages;
};

__host__ __device__ inline TcSmem tc_smem_layout(int D, int A, int C, int n_hidden, int Np, int NHp, int stages) {
    TcSmem L;
    size_t o = 0;
    L.off_xhi = o; o += kXBytes;
    L.off_xlo = o; o += kXBytes;
    L.off_w = o; o += (size_t)stages * kSlotBytes;
    L.off_s = o; o += (size_t)round_up(kTileRows * (D + 1), 4) * 4;
    o = (o + 15) / 16 * 16;
    L.off_bias = o; o += (size_t)round_up(n_hidden * Np + NHp, 4) * 4;
    L.off_vec = o; o += (size_t)(2 * kMaxObs + 2 * kMaxAct + 5 * kMaxObs) * 4;
    L.off_rowi = o; o += (size_t)kTileRows * 6 * 4;
    L.off_feat = o; o += (size_t)96 * 16 + 96 * 8;              // layer-0 feature tables (In <= 84, padded to 96)
    L.off_zero = o; o += 16;                                     // a zero word (padding features read it)
    L.off_act = o; o += (size_t)2 * kTileRows * A * 4;          // this step's / next step's actions of the tile rows
    L.off_ctx = o; o += (size_t)kTileRows * C * 4;              // context vector of every tile row
    L.off_nz = o; o += (size_t)kTileRows * ((D + 3) / 4 * 4) * 4;  // this step's N(0,1) draws, [row][4 * Philox blocks]
    o = (o + 15) / 16 * 16;
    L.off_bar = o; o += (size_t)(2 * 8 + kMaxKB + 2 + 2) * 8;     // w_full[8], w_empty[8], x_ready[13], acc_full[2], tmem slot
    L.total = o;
    L.stages = stages;
    return L;
}

// K16 blocks of a layer are split between the epilogue groups (group g owns blocks [first(g), first(g+1))) and the MMA
// warp consumes them round-robin over the groups -- the order in which they become ready.  blk_of_pos / pos_of_blk map
// between the natural block index kb and the position in that consumption order (also the order of the weight stream).
__host__ __device__ inline int grp_first(int nkb, int g) { return (nkb * g + kEpiGroups - 1) / kEpiGroups; }
__host__ __device__ inline int blk_group(int nkb, int kb) {
    int g = 0;
    while (g + 1 < kEpiGroups && kb >= grp_first(nkb, g + 1)) ++g;
    return g;
}
// The epilogue groups are staggered by one block (group g starts when group g-1 has published its first block), so
// block i of group g becomes ready at about time (g + i): consume in that order, ties by group.
__host__ __device__ inline int blk_key(int nkb, int kb) {
    const int g = blk_group(nkb, kb);
    return (g + (kb - grp_first(nkb, g))) * 16 + g;
}
__host__ __device__ inline int blk_of_pos(int nkb, int pos) {
    // the pos-th block in increasing key order (keys are distinct)
    for (int kb = 0; kb < nkb; ++kb) {
        int smaller = 0;
        for (int o = 0; o < nkb; ++o) smaller += blk_key(nkb, o) < blk_key(nkb, kb) ? 1 : 0;
        if (smaller == pos) return kb;
    }
    return -1;
}
__host__ __device__ inline int pos_of_blk(int nkb, int kb) {
    for (int pos = 0; pos < nkb; ++pos)
        if (blk_of_pos(nkb, pos) == kb) return pos;
    return -1;
}

// ------------------------------------------------------------------------------------------------
// weight packing: [E, in, out] fp32 -> per member, per K16 block: [hi block | lo block], each N x 16 in the UMMA
// K-major no-swizzle layout: byte(n, kk) = (kk / 8) * (16 N) + 16 n + 2 (kk % 8)
// ------------------------------------------------------------------------------------------------
struct PackPos { int pos[32]; };     // position of K16 block kb in the weight stream (pos_of_blk, computed once on the host:
                                     // evaluating it per element made a set_weights() cost 1 ms, 88 us per layer)
__global__ void pack_tc_kernel(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad,
                               long long member_stride, long long layer_off, int clear, float wscale, const PackPos pp) {
    const long long per_member = (long long)nkb * Npad * 16;
    const long long total = (long long)E * per_member;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / per_member);
        long long r = i - (long long)e * per_member;
        const int kb = (int)(r / (Npad * 16));
        r -= (long long)kb * Npad * 16;
        const int n = (int)(r / 16), kk = (int)(r % 16);
        const int k = kb * 16 + kk;
        const int ns = n - col0;
        const bool valid = ns >= 0 && ns < out && k < in;
        if (!valid && !clear) continue;
        const float w = valid ? src[((size_t)e * in + k) * out + ns] * wscale : 0.f;
        uint32_t hi, lo;
        tc::split2(w, 0.f, hi, lo);
        unsigned char* blk = dst + e * member_stride + layer_off + (long long)pp.pos[kb] * (2 * Npad * 32);
        const int off = (kk >> 3) * (16 * Npad) + 16 * n + 2 * (kk & 7);
        *reinterpret_cast<unsigned short*>(blk + off) = (unsigned short)(hi & 0xffffu);
        *reinterpret_cast<unsigned short*>(blk + Npad * 32 + off) = (unsigned short)(lo & 0xffffu);
    }
}

cudaError_t launch_pack_tc(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad,
                           long long member_stride, long long layer_off, int clear, cudaStream_t stream) {
    const float wscale = tc::kWScale;
    const long long total = (long long)E * nkb * Npad * 16;
    if (nkb > 32) return cudaErrorInvalidConfiguration;
    PackPos pp{};
    for (int kb = 0; kb < nkb; ++kb) pp.pos[kb] = pos_of_blk(nkb, kb);
    pack_tc_kernel<<<(int)min((total + 255) / 256, (long long)2368), 256, 0, stream>>>(dst, src, E, in, out, col0, nkb, Npad,
                                                                                       member_stride, layer_off, clear, wscale, pp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
struct Ring {
    int stage;
    uint32_t phase;
    int n;
    __device__ __forceinline__ void advance() {
        if (++stage == n) { stage = 0; phase ^= 1u; }
    }
};

// store 16 consecutive features (one K16 block `kb`) of row `row` as hi / lo halves in the UMMA A layout
__device__ __forceinline__ void store_block16(unsigned char* xhi, unsigned char* xlo, int kb, int row, const float (&y)[16]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::split2(y[2 * j], y[2 * j + 1], h[j], l[j]);
    const int o0 = (2 * kb) * kXChunkBytes + row * 16;
    const int o1 = o0 + kXChunkBytes;
    *reinterpret_cast<uint4*>(xhi + o0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(xhi + o1) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(xlo + o0) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(xlo + o1) = make_uint4(l[4], l[5], l[6], l[7]);
}

// publish a finished activation block to the MMA warp (generic-proxy writes -> async-proxy reads)
__device__ __forceinline__ void publish_block(uint64_t* bar, int lane) {
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar);
}

// 4-byte async copy global -> shared (LDGSTS) and its group fences
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct TcParams {
    RolloutParams R;
    const unsigned char* wimg;   // packed tensor-core weight image
    long long wimg_member_stride;
    int Np;                      // hidden width padded to 16 (N of the hidden layers, K of layers >= 1)
    int NHp;                     // 2D padded to 16
    int nkb0;                    // K16 blocks of layer 0
    int nkbH;                    // K16 blocks of the other layers
    int terms;                   // 3 (hi/lo split) or 1
    int stages;
    int tiles_per_member, total_tiles;
    long long* dbg;              // nullable: clock64 trace of CTA 0, [step][64]
    unsigned char order0[16];    // MMA consumption order of the K16 blocks of layer 0 ...
    unsigned char orderH[16];    // ... and of the other layers (kernel-parameter space -> uniform loads)
};

__global__ void __launch_bounds__(kTcThreads, 1) rollout_tc_kernel(const __grid_constant__ TcParams T) {
    griddep_launch();
    griddep_wait();            // programmatic dependent launch (common.cuh): the candidates' actions come from the kernel before
    const RolloutParams& P = T.R;
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcSmem L = tc_smem_layout(P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.stages);
    unsigned char* xhi = smem + L.off_xhi;
    unsigned char* xlo = smem + L.off_xlo;
    unsigned char* wring = smem + L.off_w;
    float* S = reinterpret_cast<float*>(smem + L.off_s);
    float* bias = reinterpret_cast<float*>(smem + L.off_bias);
    float* vec = reinterpret_cast<float*>(smem + L.off_vec);
    int* rowi = reinterpret_cast<int*>(smem + L.off_rowi);
    int4* feat_i = reinterpret_cast<int4*>(smem + L.off_feat);        // per layer-0 feature: {byte offset, row stride, flags}
    float2* feat_f = reinterpret_cast<float2*>(smem + L.off_feat + 96 * 16);   // {mean, 8 / (std + 1e-10)}
    float* act_s = reinterpret_cast<float*>(smem + L.off_act);
    float* ctx_s = reinterpret_cast<float*>(smem + L.off_ctx);
    float* nz_s = reinterpret_cast<float*>(smem + L.off_nz);
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* w_empty = w_full + 8;
    uint64_t* x_ready = w_empty + 8;
    uint64_t* acc_full = x_ready + kMaxKB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    volatile uint32_t* done = tmem_slot + 1;       // [2]: producer / MMA thread finished (the other lanes of their warps sleep on it)

    float* v_obs_mean = vec;
    float* v_obs_den = vec + kMaxObs;
    float* v_act_mean = vec + 2 * kMaxObs;
    float* v_act_den = v_act_mean + kMaxAct;
    float* v_dmean = v_act_den + kMaxAct;
    float* v_dscale = v_dmean + kMaxObs;
    float* v_2logstd = v_dscale + kMaxObs;
    float* v_maxlv = v_2logstd + kMaxObs;
    float* v_minlv = v_maxlv + kMaxObs;
    int* r_mi = rowi;
    int* r_src = rowi + kTileRows;
    int* r_pi = rowi + 2 * kTileRows;
    int* r_ctx = rowi + 3 * kTileRows;
    int* r_rid = rowi + 4 * kTileRows;
    int* r_eps = rowi + 5 * kTileRows;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nstage = T.stages;
    const int gemms_per_step = P.n_hidden + 1;
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int k = 0; k < kMaxKB; ++k) ptx::mbar_init(&x_ready[k], 4);
        ptx::mbar_init(&acc_full[0], 1);
        ptx::mbar_init(&acc_full[1], 1);
        done[0] = 0u;
        done[1] = 0u;
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t stage_bytes_h = 2u * T.Np * 32u;      // hi + lo of a hidden K16 block
    const uint32_t stage_bytes_o = 2u * T.NHp * 32u;     // heads

    // ======================= warp 0: weight producer =============================================
    if (warp == 0) {
        if (lane == 0) {
            Ring rp{0, 0, nstage};
            for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
                const int e = tile / T.tiles_per_member;
                const unsigned char* wsrc = T.wimg + (size_t)e * T.wimg_member_stride;
                for (int t = 0; t < P.h; ++t) {
                    size_t off = 0;
                    for (int g = 0; g < gemms_per_step; ++g) {
                        const int nkb = g == 0 ? T.nkb0 : T.nkbH;
                        const uint32_t bytes = g == P.n_hidden ? stage_bytes_o : stage_bytes_h;
                        for (int kb = 0; kb < nkb; kb += kBlocksPerStage) {          // the image is in consumption order: a stage
                            const uint32_t sb = bytes * (uint32_t)min(kBlocksPerStage, nkb - kb);   // is one contiguous copy
                            ptx::mbar_wait(&w_empty[rp.stage], rp.phase ^ 1u);
                            ptx::mbar_arrive_expect_tx(&w_full[rp.stage], sb);
                            ptx::bulk_g2s(wring + (size_t)rp.stage * kSlotBytes, wsrc + off, sb, &w_full[rp.stage]);
                            off += sb;
                            rp.advance();
                        }
                    }
                }
            }
            done[0] = 1u;
        } else {
            // lanes parked at a barrier keep competing for issue slots with the working lane of their own warp (measured in
            // rollout_tcs.cu); sleeping lanes do not
            while (!done[0]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warp 1: MMA issuer ==================================================
    // ONE elected thread walks the whole schedule inside a single elect.sync region: ptxas keeps the descriptors in uniform
    // registers and issues the three UTCHMMAs of a K block back to back.  (Electing per K block from a warp-uniform walk
    // cost ~2x the MMA time in instruction latency -- a lone thread has no latency hiding, so its loop has to be lean; see
    // rollout_tcs.cu for the measurements.)  The tcgen05 fence is needed once per GEMM (accumulator hand-over), not per block.
    else if (warp == 1) {
        if (ptx::elect_one()) {
            Ring rc{0, 0, nstage};
            uint32_t xphase = 0;          // bit kb = parity to wait for on x_ready[kb]
            uint32_t g_count = 0;         // running GEMM index: accumulator buffer = g & 1
            const uint32_t xhi_d = ptx::smem_u32(xhi) >> 4, xlo_d = ptx::smem_u32(xlo) >> 4, w_d = ptx::smem_u32(wring) >> 4;
            const uint64_t top = (uint64_t)((1u << 14) | 8u) << 32;      // version 1 (bit 46), SBO = 128 B
            const uint32_t a_lbo = (kXChunkBytes >> 4) << 16;             // A: LBO = one 8-wide k-chunk of all 128 rows
            for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
                for (int t = 0; t < P.h; ++t) {
                    const bool trace = T.dbg && blockIdx.x == 0 && tile == (int)blockIdx.x && t < 64;
                    for (int g = 0; g < gemms_per_step; ++g) {
                        const int nkb = g == 0 ? T.nkb0 : T.nkbH;
                        const uint32_t N = g == P.n_hidden ? (uint32_t)T.NHp : (uint32_t)T.Np;
                        const uint32_t d_tmem = tmem_base + (g_count & 1u) * 256u;
                        const uint32_t idesc = tc::idesc_f16(tc::kFmtF16, tc::kFmtF16, N);
                        const uint32_t b_lbo = N << 16;                   // B: LBO = 16 N bytes
                        const uint32_t b_lo_off = (N * 32u) >> 4;         // W_lo block behind W_hi
                        for (int pos = 0; pos < nkb; ++pos) {
                            const uint32_t kb = g == 0 ? T.order0[pos] : T.orderH[pos];
                            const int sub = pos % kBlocksPerStage;               // block within the ring slot
                            ptx::mbar_wait(&x_ready[kb], (xphase >> kb) & 1u);
                            xphase ^= 1u << kb;
                            if (sub == 0) ptx::mbar_wait(&w_full[rc.stage], rc.phase);
                            if (pos == 0) {
                                tc::fence_after_sync();
                                if (trace && g < 5) T.dbg[t * 64 + 32 + 4 * g] = clock64();
                            }
                            const uint32_t wb = w_d + (uint32_t)rc.stage * (kSlotBytes >> 4) + (uint32_t)sub * ((N * 64u) >> 4);
                            const uint64_t a_hi = top | ((xhi_d + kb * ((2u * kXChunkBytes) >> 4)) | a_lbo);
                            const uint64_t a_lo = top | ((xlo_d + kb * ((2u * kXChunkBytes) >> 4)) | a_lbo);
                            const uint64_t b_hi = top | (wb | b_lbo);
                            const uint64_t b_lo = top | ((wb + b_lo_off) | b_lbo);
                            tc::mma_f16_ss(d_tmem, a_hi, b_hi, idesc, pos > 0 ? 1u : 0u);
                            if (T.terms == 3) {
                                tc::mma_f16_ss(d_tmem, a_hi, b_lo, idesc, 1u);
                                tc::mma_f16_ss(d_tmem, a_lo, b_hi, idesc, 1u);
                            }
                            if (sub == kBlocksPerStage - 1 || pos == nkb - 1) {
                                tc::mma_commit(&w_empty[rc.stage]);
                                rc.advance();
                            }
                        }
                        tc::mma_commit(&acc_full[g_count & 1u]);
                        if (trace && g < 5) T.dbg[t * 64 + 33 + 4 * g] = clock64();
                        ++g_count;
                    }
                }
            }
            done[1] = 1u;
        } else {
            while (!done[1]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warps 2..9: prologue / epilogue ======================================
    else {
        const int et = tid - 64;                       // 0..255
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int group = (warp - 2) >> 2;             // owns the K16 blocks [grp_first(nkb, group), grp_first(nkb, group + 1))
        const int row = quarter * 32 + lane;           // tile row == TMEM lane
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t g_count = 0;
        const int D = P.D, A = P.A;
        const size_t eps_step_stride = (size_t)P.E * P.q * P.m * P.n_global * D;

        // one-time: normalisation vectors
        for (int i = et; i < P.P; i += kEpiThreads) { v_obs_mean[i] = P.obs_mean[i]; v_obs_den[i] = 1.0f / (P.obs_std[i] + 1e-10f); }
        for (int i = et; i < A; i += kEpiThreads) { v_act_mean[i] = P.act_mean[i]; v_act_den[i] = 1.0f / (P.act_std[i] + 1e-10f); }
        for (int i = et; i < D; i += kEpiThreads) {
            v_dmean[i] = P.delta_mean[i];
            v_dscale[i] = P.delta_std[i] + 1e-10f;
            v_2logstd[i] = 2.0f * logf(P.delta_std[i]);
            v_maxlv[i] = P.max_lv[i];
            v_minlv[i] = P.min_lv[i];
        }

        // layer-0 feature table (branch-free gather): feature k of row r reads the float at byte offset
        //   x + r * y + (z & parity-of-step ? A-buffer stride : 0)      of shared memory,     then (v - mean) * inv8
        // kinds: state element / action / context / zero word; sin, cos of the HalfCheetah angle are patched in by select
        for (int k = et; k < T.nkb0 * 16; k += kEpiThreads) {
            int base = (int)L.off_zero, stride = 0, flags = 0;      // flags: 1 = action double buffer, 2 = sin, 4 = cos
            float mean = 0.f, inv = 0.f;
            if (k < P.P) {
                int idx = k;
                if (P.env_id == CADM_ENV_HALFCHEETAH) {          // [o1, sin o2, cos o2, o3:]
                    if (k == 0) idx = 1;
                    else if (k == 1) { idx = 2; flags = 2; }
                    else if (k == 2) { idx = 2; flags = 4; }
                } else if (P.env_id == CADM_ENV_ANT) {
                    idx = k + 1;                                  // o[1:]
                }
                base = (int)L.off_s + idx * 4; stride = (D + 1) * 4;
                mean = P.obs_mean[k]; inv = 1.0f / (P.obs_std[k] + 1e-10f);
            } else if (k < P.P + A) {
                const int ai = k - P.P;
                base = (int)L.off_act + ai * 4; stride = A * 4; flags = 1;
                if (P.discrete) { mean = 0.f; inv = 1.f; }
                else { mean = P.act_mean[ai]; inv = 1.0f / (P.act_std[ai] + 1e-10f); }
            } else if (k < P.In) {
                base = (int)L.off_ctx + (k - P.P - A) * 4; stride = P.C * 4; mean = 0.f; inv = 1.f;
            }
            feat_i[k] = make_int4(base, stride, flags, 0);
            feat_f[k] = make_float2(mean, inv * tc::kXScale);
        }
        if (et == 0) *reinterpret_cast<float*>(smem + L.off_zero) = 0.f;

        for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
            const int e = tile / T.tiles_per_member;
            const int tile_row0 = P.row_lo + (tile - e * T.tiles_per_member) * P.rows_per_cta;
            const int nrows = min(P.rows_per_cta, P.row_hi - tile_row0);
            ptx::bar_sync(1, kEpiThreads);             // previous tile fully retired before its smem is reused
            for (int i = et; i < P.n_hidden * T.Np + T.NHp; i += kEpiThreads)      // hidden-layer biases pre-scaled by kXScale
                bias[i] = P.bpack[(size_t)e * P.bias_stride + i] * (i < P.n_hidden * T.Np ? tc::kXScale : 1.0f);
            if (et < kTileRows) {
                const int r = et;
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
                        rid = src; cidx = src; er = src;
                    }
                }
                r_mi[r] = mi; r_src[r] = src; r_pi[r] = pi; r_ctx[r] = cidx; r_rid[r] = rid; r_eps[r] = er;
            }
            ptx::bar_sync(1, kEpiThreads);
            for (int i = et; i < kTileRows * D; i += kEpiThreads) {
                const int r = i / D, d = i - r * D;
                float v = 0.f;
                if (r < nrows) v = (P.row_mode == kRowsPlanner) ? P.obs0[r_mi[r] * D + d] : P.obs0[(size_t)r_src[r] * D + d];
                S[r * (D + 1) + d] = v;
            }
            for (int i = et; i < kTileRows * P.C; i += kEpiThreads) {          // context of every row (constant over the steps)
                const int r = i / P.C, c = i - r * P.C;
                ctx_s[i] = r < nrows ? __ldg(P.ctx + (size_t)r_ctx[r] * P.C + c) : 0.f;
            }
            // actions of step 0 (later steps are prefetched one step ahead with cp.async)
            auto prefetch_actions = [&](int t) {
                if (P.discrete && P.row_mode == kRowsPlanner) return;
                float* dst = act_s + (t & 1) * kTileRows * A;
                for (int i = et; i < nrows * A; i += kEpiThreads) {
                    const int r = i / A, a = i - r * A;
                    cp_async4(dst + i, P.actions + ((size_t)r_src[r] * P.h + t) * A + a);
                }
                cp_async_commit();
            };
            prefetch_actions(0);
            cp_async_wait_all();
            ptx::bar_sync(1, kEpiThreads);

            float ret = 0.f;
            const bool valid = row < nrows;
#pragma unroll 1
            for (int t = 0; t < P.h; ++t) {
                long long* dbg = (T.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && tile == (int)blockIdx.x && t < 64) ? T.dbg + t * 64 : nullptr;
                if (dbg) dbg[0] = clock64();
                // ---------- prologue: reward of the current state; layer-0 input blocks -----------------
                const float* s = S + row * (D + 1);
                const float* arow = act_s + (t & 1) * kTileRows * A + row * A;       // this step's action of my row
                if (t + 1 < P.h) prefetch_actions(t + 1);                            // lands during this step
                if (dbg) dbg[13] = clock64();
                if (group == 0 && valid) {
                    if (env_reward_reads_next(P.env_id)) {
                        if (t > 0) ret += env_reward_next(P.env_id, s);
                    } else {
                        ret += env_reward_current(P.env_id, s, arow, A, P.max_torque);
                    }
                }
                if (dbg) dbg[14] = clock64();
                for (int kb = grp_first(T.nkb0, group); kb < grp_first(T.nkb0, group + 1); ++kb) {
                    float y[16];
                    float sn = 0.f, cs = 1.f;
                    if (P.env_id == CADM_ENV_HALFCHEETAH && kb == 0) sincosf(s[2], &sn, &cs);
                    const int par_off = (t & 1) * kTileRows * A * 4;
                    if (P.discrete && P.row_mode == kRowsPlanner) {
                        // one-hot of the integer action (random shooting with discrete actions): rare path, kept simple
                        const int onehot = __ldg(P.actions_int + (size_t)r_src[row] * P.h + t);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int k = kb * 16 + j;
                            const int4 fi = feat_i[k];
                            const float2 ff = feat_f[k];
                            float src = *reinterpret_cast<const float*>(smem + fi.x + row * fi.y);
                            if (fi.z & 1) src = (onehot == k - P.P) ? 1.f : 0.f;
                            y[j] = valid ? (src - ff.x) * ff.y : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int4 fi = feat_i[kb * 16 + j];
                            const float2 ff = feat_f[kb * 16 + j];
                            float src = *reinterpret_cast<const float*>(smem + fi.x + row * fi.y + ((fi.z & 1) ? par_off : 0));
                            src = (fi.z & 2) ? sn : src;
                            src = (fi.z & 4) ? cs : src;
                            y[j] = valid ? (src - ff.x) * ff.y : 0.f;
                        }
                    }
                    if (dbg) dbg[15] = clock64();
                    store_block16(xhi, xlo, kb, row, y);
                    publish_block(&x_ready[kb], lane);
                }
                if (dbg) dbg[1] = clock64();
                // ---------- hidden layers: accumulator -> bias + swish -> next layer's A operand ---------
#pragma unroll 1
                for (int l = 0; l < P.n_hidden; ++l) {
                    const uint32_t buf = g_count & 1u;
                    ptx::mbar_wait(&acc_full[buf], (g_count >> 1) & 1u);
                    tc::fence_after_sync();
                    if (dbg && l < 4) dbg[2 + 2 * l] = clock64();
                    ++g_count;
                    const uint32_t tcol = tmem_lane + buf * 256u;
                    const float* bl = bias + l * T.Np;
                    const int kb0 = grp_first(T.nkbH, group), kb1 = grp_first(T.nkbH, group + 1);
                    // stagger the groups by one block: group g starts when group g-1 has published its first block, so the
                    // MMA warp gets its first operand block ~3x sooner and the four warps of an SMSP run out of phase
                    // (MUFU-heavy and FMA-heavy stretches overlap instead of colliding)
                    if (group > 0) ptx::bar_sync(1 + group, 256);
                    if (kb0 >= kb1 && group + 1 < kEpiGroups) ptx::bar_arrive(2 + group, 256);
                    uint32_t v[16];
                    if (kb0 < kb1) tc::tmem_ld16(tcol + kb0 * 16, v);
#pragma unroll 1
                    for (int kb = kb0; kb < kb1; ++kb) {
                        tc::tmem_wait_ld();
                        float y[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(v[j]);
                        if (kb + 1 < kb1) tc::tmem_ld16(tcol + (kb + 1) * 16, v);      // prefetch the next block
                        float bb[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(&bb[j]) = *reinterpret_cast<const float4*>(bl + kb * 16 + j);
#pragma unroll
                        for (int j = 0; j < 16; ++j) y[j] = tc::swish8_fast(fmaf(y[j], 1.0f / tc::kWScale, bb[j]));
                        store_block16(xhi, xlo, kb, row, y);
                        tc::fence_before_sync();
                        publish_block(&x_ready[kb], lane);
                        if (kb == kb0 && group + 1 < kEpiGroups) ptx::bar_arrive(2 + group, 256);
                    }
                    if (dbg && l < 4) dbg[3 + 2 * l] = clock64();
                    // off the critical path (the warps wait for the next accumulator anyway): this step's Gaussian draws
                    if (l == 0 && !P.deterministic) {
                        const int nj = (D + 3) >> 2;
                        for (int i = et; i < kTileRows * nj; i += kEpiThreads) {
                            const int jb = i / kTileRows, r = i - jb * kTileRows;
                            float nz[4] = {0.f, 0.f, 0.f, 0.f};
                            if (r < nrows) {
                                if (P.eps != nullptr) {
                                    const float* ep = P.eps + (P.row_mode == kRowsPlanner ? (size_t)t * eps_step_stride : 0) + (size_t)r_eps[r] * D;
#pragma unroll
                                    for (int ii = 0; ii < 4; ++ii) nz[ii] = 4 * jb + ii < D ? __ldg(ep + 4 * jb + ii) : 0.f;
                                } else {
                                    normal4_fast(P.seed, (uint32_t)jb, (uint32_t)r_rid[r], (uint32_t)t, (uint32_t)P.it, nz);
                                }
                            }
                            *reinterpret_cast<float4*>(nz_s + r * (nj * 4) + 4 * jb) = make_float4(nz[0], nz[1], nz[2], nz[3]);
                        }
                    }
                }

                // ---------- heads -> Hd (aliases X_hi once the head GEMM has completed) ------------------
                {
                    const uint32_t buf = g_count & 1u;
                    ptx::mbar_wait(&acc_full[buf], (g_count >> 1) & 1u);
                    tc::fence_after_sync();
                    ++g_count;
                    if (dbg) dbg[10] = clock64();
                    float* Hd = reinterpret_cast<float*>(xhi);       // [NHp][128]  (column-major: conflict-free)
                    const uint32_t tcol = tmem_lane + buf * 256u;
                    const float* bl = bias + P.n_hidden * T.Np;
                    const int nb8 = T.NHp / 8;
                    for (int b = group; b < nb8; b += kEpiGroups) {
                        uint32_t v[8];
                        tc::tmem_ld8(tcol + b * 8, v);
                        tc::tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) Hd[(b * 8 + j) * kTileRows + row] = fmaf(__uint_as_float(v[j]), 1.0f / (tc::kWScale * tc::kXScale), bl[b * 8 + j]);
                    }
                    tc::fence_before_sync();
                }
                ptx::bar_sync(1, kEpiThreads);
                if (dbg) dbg[11] = clock64();

                // ---------- final epilogue: sample, next state ------------------------------------------
                // thread = (row, quarter of the state dims): equal work for all 512 threads, independent chains per dim
                {
                    const float* Hd = reinterpret_cast<const float*>(xhi);
                    const int r = et & (kTileRows - 1);
                    const int part = et >> 7;                                   // 0..3
                    const int dper = (D + kEpiGroups - 1) / kEpiGroups;
                    const int d0 = part * dper, d1 = min(D, d0 + dper);
                    if (r < nrows && d0 < d1) {
                        const float* nzr = nz_s + r * (((D + 3) >> 2) * 4);         // drawn while layer 1 was running
#pragma unroll
                        for (int i = 0; i < 12; ++i) {
                            const int d = d0 + i;
                            if (d >= d1) break;
                            const float mu = Hd[d * kTileRows + r];
                            float lv = Hd[(D + d) * kTileRows + r];
                            const float dmu = mu * v_dscale[d] + v_dmean[d];
                            float delta = dmu;
                            if (!P.deterministic) {
                                lv = fast_bounded_logvar(lv, v_maxlv[d], v_minlv[d]);
                                delta = dmu + nzr[d] * fast_exp((lv + v_2logstd[d]) * 0.5f);
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
                cp_async_wait_all();                       // next step's actions have landed (issued in the prologue)
                ptx::bar_sync(1, kEpiThreads);
                if (dbg) dbg[12] = clock64();
            }
            if (P.row_mode == kRowsPlanner && group == 0 && valid) {
                if (env_reward_reads_next(P.env_id)) ret += env_reward_next(P.env_id, S + row * (D + 1));
                P.ret_p[(size_t)r_src[row] * P.p + r_pi[row]] = ret;
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
static int g_tc_smem = 0;

cudaError_t launch_rollout_tc(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms,
                              int num_sms, cudaStream_t stream, const char** name, long long* dbg) {
    TcParams T{};
    T.dbg = dbg;
    T.wimg = wimg;
    T.wimg_member_stride = wimg_member_stride;
    T.Np = round_up(P.H, 16);
    T.NHp = round_up(2 * P.D, 16);
    T.nkb0 = round_up(P.In, 16) / 16;
    T.nkbH = T.Np / 16;
    T.terms = terms;
    if (P.row_hi <= 0) { P.row_lo = 0; P.row_hi = P.rows_per_member; }
    const int span = P.row_hi - P.row_lo;                                 // rows of every member this launch covers
    if (span < 1 || P.row_lo < 0 || P.row_hi > P.rows_per_member) return cudaErrorInvalidValue;
    int tiles = (span + kTileRows - 1) / kTileRows;
    P.rows_per_cta = (span + tiles - 1) / tiles;                          // balance the rows over the tiles
    tiles = (span + P.rows_per_cta - 1) / P.rows_per_cta;
    T.tiles_per_member = tiles;
    T.total_tiles = tiles * P.E;
    for (int i = 0; i < T.nkb0; ++i) T.order0[i] = (unsigned char)blk_of_pos(T.nkb0, i);
    for (int i = 0; i < T.nkbH; ++i) T.orderH[i] = (unsigned char)blk_of_pos(T.nkbH, i);
    // shared-memory ring: as many stages as fit in 227 KB
    int stages = 4;
    while (stages > 2 && tc_smem_layout(P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, stages).total > 226 * 1024) --stages;
    T.stages = stages;
    const TcSmem L = tc_smem_layout(P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, stages);
    if (L.total > 226 * 1024) return cudaErrorInvalidConfiguration;
    T.R = P;
    if ((int)L.total > g_tc_smem) {
        cudaError_t e = cudaFuncSetAttribute(rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        if (e != cudaSuccess) return e;
        g_tc_smem = (int)L.total;
    }
    if (name) *name = terms == 3 ? "rollout_tc_kernel(fp16 hi/lo x3)" : "rollout_tc_kernel(f16 x1)";
    const int grid = min(T.total_tiles, num_sms);
    return launch_chain(rollout_tc_kernel, dim3(grid), dim3(kTcThreads), L.total, stream, T);
}


// ------------------------------------------------------------------------------------------------
// Self-test: ONE 128 x N x K product with exactly the operand layouts, descriptors and split arithmetic of the rollout
// kernel (device diagnostic behind cadm_selftest_tc_gemm; tests compare it with an fp64 matmul).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_gemm_selftest_kernel(const float* __restrict__ X, const unsigned char* wimg,
                                                                   int K, int N, int terms, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* xhi = smem;
    unsigned char* xlo = smem + kXBytes;
    unsigned char* wst = smem + 2 * kXBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kXBytes + kMaxStageBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (K + 15) / 16;
    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {                      // a fully converged warp (warp 0 just diverged on tid == 0)
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 256);
        tc::tmem_relinquish();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = kb * 16 + j;
            y[j] = k < K ? X[(size_t)tid * K + k] * tc::kXScale : 0.f;
        }
        store_block16(xhi, xlo, kb, tid, y);
    }
    ptx::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        const uint32_t bytes = 2u * N * 32u;
        const uint32_t xhi_a = ptx::smem_u32(xhi), xlo_a = ptx::smem_u32(xlo), w_a = ptx::smem_u32(wst);
        uint32_t ph = 0;
        for (int pos = 0; pos < nkb; ++pos) {
            const int kb = blk_of_pos(nkb, pos);            // the weight stream is stored in consumption order
            ptx::mbar_arrive_expect_tx(&bars[0], bytes);
            ptx::bulk_g2s(wst, wimg + (size_t)pos * bytes, bytes, &bars[0]);
            ptx::mbar_wait(&bars[0], ph);
            tc::fence_after_sync();
            const uint64_t a_hi = tc::smem_desc(xhi_a + 2 * kb * kXChunkBytes, kXChunkBytes, 128);
            const uint64_t a_lo = tc::smem_desc(xlo_a + 2 * kb * kXChunkBytes, kXChunkBytes, 128);
            const uint64_t b_hi = tc::smem_desc(w_a, 16 * N, 128);
            const uint64_t b_lo = tc::smem_desc(w_a + N * 32, 16 * N, 128);
            const uint32_t idesc = tc::idesc_f16(tc::kFmtF16, tc::kFmtF16, N);
            tc::mma_f16_ss(tmem_base, a_hi, b_hi, idesc, pos > 0 ? 1u : 0u);
            if (terms == 3) {
                tc::mma_f16_ss(tmem_base, a_hi, b_lo, idesc, 1u);
                tc::mma_f16_ss(tmem_base, a_lo, b_hi, idesc, 1u);
            }
            tc::mma_commit(&bars[1]);
            ptx::mbar_wait(&bars[1], ph);          // serialise: the single weight slot is reused
            ph ^= 1u;
        }
    }
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tc::tmem_ld8(tl + c, v);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]) * (1.0f / (tc::kWScale * tc::kXScale));
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 256);
    }
}

cudaError_t launch_tc_gemm_selftest(const float* X, const unsigned char* wimg, int K, int N, int terms, float* out,
                                    cudaStream_t stream) {
    const int smem_bytes = 2 * kXBytes + kMaxStageBytes + 64;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    tc_gemm_selftest_kernel<<<1, 128, smem_bytes, stream>>>(X, wimg, K, N, terms, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark: `n_mma` back-to-back tcgen05.mma (M = 128, N, K = 16, kind::f16) on resident shared-memory operands
// in the rollout kernels' layouts, nothing else running on the SM.  `n_acc` accumulators are used round-robin (1 = every
// MMA depends on the previous one); `swapped` selects the operand layouts of rollout_tcs.cu (A = 128 weight rows K-major,
// B = N batch rows MN-major).  Reports clock64 cycles (issue of first -> commit seen).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int N, int n_mma, int a_lbo, int n_acc, int swapped, int bg, const unsigned char* bg_src,
                                                              long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kXBytes + 65536);      // [0] MMA done, [1] stop flag, [2..5] bg copies
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(bars + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (2 * kXBytes + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1);
        for (int i = 2; i < 6; ++i) ptx::mbar_init(&bars[i], 1);
        *stop = 0u;
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    ptx::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        const uint32_t x_a = ptx::smem_u32(smem), w_a = ptx::smem_u32(smem + 2 * kXBytes);
        const uint32_t idesc = tc::idesc_f16(tc::kFmtF16, tc::kFmtF16, N) | (swapped ? (1u << 16) : 0u);
        const uint32_t acc_stride = 512u / (uint32_t)n_acc;
        const int style = a_lbo >> 16;           // issue-loop style (diagnostic): 0 = elect per MMA, 1 = one elected region
        a_lbo &= 0xffff;
        const long long t0 = clock64();
        if (style == 0) {
            // the whole warp walks the loop with warp-uniform operands; one elected lane issues each MMA
            for (int i = 0; i < n_mma; ++i) {
                const int kb = i % kMaxKB;
                uint64_t a, b;
                if (swapped) {
                    a = tc::smem_desc(w_a + (kb & 1) * 4096, 2048, 128);            // 128 weight rows, K-major
                    b = tc::smem_desc(x_a + kb * 256, 128, 3328);                    // N batch rows, MN-major
                } else {
                    a = tc::smem_desc(x_a + 2 * kb * kXChunkBytes, a_lbo, 128);
                    b = tc::smem_desc(w_a, 16 * N, 128);
                }
                if (ptx::elect_one()) tc::mma_f16_ss(tmem_base + (uint32_t)(i % n_acc) * acc_stride, a, b, idesc, i >= n_acc ? 1u : 0u);
                __syncwarp();
            }
        } else if (ptx::elect_one()) {
            // one elected thread runs the whole loop: groups of 3 MMAs (the hi/lo terms of one K16 block), descriptors
            // advanced by constant adds, 13 K blocks per accumulator pass
            const uint64_t a0 = swapped ? tc::smem_desc(w_a, 2048, 128) : tc::smem_desc(x_a, a_lbo, 128);
            const uint64_t b0 = swapped ? tc::smem_desc(x_a, 128, 3328) : tc::smem_desc(w_a, 16 * N, 128);
            const uint64_t a_inc = swapped ? (256u >> 4) : ((2u * kXChunkBytes) >> 4);
            const uint64_t b_inc = swapped ? (256u >> 4) : 0u;
            const uint64_t lo_off = 8192u >> 4;
            int issued = 0, pass = 0;
            while (issued < n_mma) {
                const uint32_t d = tmem_base + (uint32_t)(pass % n_acc) * acc_stride;
                uint64_t a = a0, b = b0;
#pragma unroll
                for (int kb = 0; kb < kMaxKB; ++kb) {
                    tc::mma_f16_ss(d, a, b, idesc, (kb > 0 || pass >= n_acc) ? 1u : 0u);
                    tc::mma_f16_ss(d, a, b + lo_off, idesc, 1u);
                    tc::mma_f16_ss(d, a + lo_off, b, idesc, 1u);
                    a += a_inc;
                    b += b_inc;
                    if ((bg & 2) && (kb & 1)) tc::mma_commit(&bars[5]);       // a commit every 6 MMAs, like a weight stage
                }
                issued += 3 * kMaxKB;
                ++pass;
            }
        }
        __syncwarp();
        const long long t1 = clock64();
        if (ptx::elect_one()) tc::mma_commit(&bars[0]);
        __syncwarp();
        ptx::mbar_wait(&bars[0], 0);
        const long long t2 = clock64();
        if (tid == 0) {
            cycles[0] = t1 - t0;
            cycles[1] = t2 - t0;
            *stop = 1u;
        }
    } else if (warp == 2 && (bg & 1)) {
        // background weight stream: 8 KB bulk copies global -> shared (4 in flight) into a region the MMAs do not read
        if (tid == 64) {
            unsigned char* dst = smem + 2 * kXBytes + 32768;
            uint32_t ph[4] = {0u, 0u, 0u, 0u};
            long long n = 0;
            for (int i = 0; i < 4; ++i) {
                ptx::mbar_arrive_expect_tx(&bars[2 + i], 8192);
                ptx::bulk_g2s(dst + i * 8192, bg_src + ((n++ * 8192) & 0xfffff), 8192, &bars[2 + i]);
            }
            while (!*stop) {
                for (int i = 0; i < 4; ++i) {
                    ptx::mbar_wait(&bars[2 + i], ph[i]);
                    ph[i] ^= 1u;
                    ptx::mbar_arrive_expect_tx(&bars[2 + i], 8192);
                    ptx::bulk_g2s(dst + i * 8192, bg_src + ((n++ * 8192) & 0xfffff), 8192, &bars[2 + i]);
                }
            }
            for (int i = 0; i < 4; ++i) ptx::mbar_wait(&bars[2 + i], ph[i]);
            cycles[2] = n;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

cudaError_t launch_tc_mma_rate(int N, int n_mma, int a_lbo, int n_acc, int swapped, int bg, const unsigned char* bg_src,
                               long long* cycles, cudaStream_t stream) {
    const int smem_bytes = 2 * kXBytes + 65536 + 128;
    cudaError_t e = cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    tc_mma_rate_kernel<<<1, 128, smem_bytes, stream>>>(N, n_mma, a_lbo, n_acc, swapped, bg, bg_src, cycles);
    return cudaGetLastError();
}

}  // namespace cadm
