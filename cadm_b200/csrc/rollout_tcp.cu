// EXPERIMENT (opt-in: cadm_set_option "tc_variant" = 3 / CADM_TC_VARIANT=3; never picked automatically).
// Persistent tensor-core rollout kernel for a CTA PAIR (thread-block cluster of 2): the swapped-operand scheme of
// rollout_tcs.cu with the hidden units split ACROSS the two CTAs and two row tiles ping-ponged through them.
//
// The idea.  rollout_tcs.cu gives one CTA 32 rows and both M tiles of every hidden layer (H = 200 -> 128 + 80 units).  Its layer
// time is a dependent chain: MMAs of tile 0 and tile 1 (2 x 13 K blocks x ~89 cycles, paced by the 4 KB weight fetch per MMA, not
// by the tensor pipe), the epilogue of tile 0 under tile 1's MMAs, the epilogue of tile 1 exposed, then the next layer: ~3.05k
// cycles of which the tensor pipe works 2.3k (profiles/r2_tcs_warptrace.log).  The epilogue cannot shrink -- it is MUFU bound (2
// per activation at 8 cycles per warp instruction, tools/probes/pipe_rates2.cu) -- so give the tensor pipe something independent
// to do while an epilogue runs:
//   - a pair owns TWO subtiles (A, B) of N rows of one ensemble member;
//   - CTA r holds only M tile r of the hidden layers (its half of the weight stream) and computes it for BOTH subtiles, one after
//     the other: while its 16 epilogue warps turn subtile A's accumulator into the next layer's operand, its MMA thread is
//     already issuing subtile B's MMAs, and vice versa;
//   - a layer's operand X^T[K = all hidden units x rows] is needed by both CTAs, so every epilogue warp stores its 16-byte rows
//     twice: into its own shared memory and, over distributed shared memory, into the peer's (st.shared::cluster), then arrives
//     on the "operand ready" mbarrier of both CTAs; an accumulator is handed to the epilogue with a MULTICAST tcgen05.commit to
//     both CTAs, because a CTA may overwrite the peer's operand buffer only when the peer's MMAs have read it too;
//   - the heads (one M = 64 tile) and the final epilogue (sample, state update, reward, next step's layer-0 operand) of subtile s
//     run on CTA s, so that serial tail is shared out as well.
//
// The outcome (B200, C2: 512 us per launch against 239 us for rollout_tcs.cu; profiles/r2_tcp_trace.log).  It is correct -- bit
// for bit the results of rollout_tcs.cu, the same parity suite -- and it loses, for one reason the trace shows directly: the
// SM-to-SM path.  Every (layer, subtile) moves 16 KB of activations from CTA 0 to CTA 1 (10 KB the other way), and a CTA pair
// exchanges distributed shared memory at 11 B/clk per direction with st.shared::cluster and 14.5 B/clk with cp.async.bulk
// shared::cta -> shared::cluster (tools/probes/dsmem_bw_probe.cu, profiles/r2_dsmem_bw.log; the B300 notes quote 17-21 B/clk):
// ~1.1k cycles per exchange, as long as the 13 K blocks of MMAs it was meant to overlap, and it sits on the dependent chain
// (epilogue -> peer's operand -> peer's MMAs).  The same arithmetic rules out the cta_group::2 form of the idea (M = 256 across
// the pair, B operand split by rows): its epilogue still sends half of every CTA's activations across, 26 KB per 64-row layer =
// 1.1k cycles against 1.35k of MMA time.  Splitting the hidden units over SMs cannot pay on this part; rows are the only axis
// that shards for free.  Kept as a tested record of that measurement (and of the cluster plumbing: DSMEM addressing, remote
// mbarrier arrives, multicast commits, matrix descriptors inside a cluster).
//
// Scope: the reference architecture (hidden width 193..208 -> two M tiles, layer-0 input <= 48 features, at most 32 state
// dimensions), kps = 4 weight image of rollout_tcs.cu (read in place: CTA r streams the stages of its tile).  Arithmetic, operand
// layouts, scales and summation orders are those of rollout_tcs.cu.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "tcs_common.cuh"

namespace cadm {

constexpr int kPEpiThreads = 512;
constexpr int kPThreads = 64 + kPEpiThreads;
constexpr int kPMaxRows = 64;                  // rows per subtile (multiple of 16)
constexpr int kPSubStride = 128;               // TMEM columns between the accumulators of subtile A and B (2 x rows <= 128)
constexpr int kPHeadCol = 256;                 // TMEM column of the head accumulator
constexpr int kPMaxStages = 16;
constexpr int kPMaxEnt = 64;                   // weight stages per horizon step and CTA

struct TcpSmem {
    size_t off_x, off_x0, off_w, off_s, off_hd, off_bias, off_vec, off_rowi, off_feat, off_zero, off_act, off_ctx, off_nz, off_bar, total;
    int xbytes, x0bytes, slot_bytes;
};

__host__ __device__ inline TcpSmem tcp_smem_layout(int N, int D, int A, int C, int n_hidden, int Np, int NHp, int Kcap, int nkb0, int stages) {
    TcpSmem L;
    size_t o = 0;
    L.xbytes = (N / 8) * (Kcap / 8) * 128;       // one operand half (hi or lo) of one subtile's layer input
    L.slot_bytes = 4 * 8192;
    L.off_x = o; o += (size_t)4 * L.xbytes;      // subtile s: [hi | lo] at s * 2 xbytes (single-buffered: see the header comment)
    L.x0bytes = (N / 8) * (nkb0 * 2) * 128;      // layer-0 input (its own buffer, as in rollout_tcs.cu)
    L.off_x0 = o; o += (size_t)4 * L.x0bytes;
    o = (o + 127) / 128 * 128;
    L.off_w = o; o += (size_t)stages * L.slot_bytes;
    L.off_s = o; o += (size_t)round_up(N * (D + 3), 4) * 4;
    L.off_hd = o; o += (size_t)NHp * (N + 4) * 4;
    o = (o + 15) / 16 * 16;
    L.off_bias = o; o += (size_t)round_up(n_hidden * Np + NHp, 4) * 4;
    o = (o + 15) / 16 * 16;
    L.off_vec = o; o += (size_t)(8 * kMaxObs) * 4;
    L.off_rowi = o; o += (size_t)N * 6 * 4;
    L.off_feat = o; o += (size_t)96 * 16 + 96 * 8;
    L.off_zero = o; o += 16;
    L.off_act = o; o += (size_t)2 * N * A * 4;
    L.off_ctx = o; o += (size_t)N * (C > 0 ? C : 1) * 4;
    L.off_nz = o; o += (size_t)N * ((D + 3) / 4 * 4) * 4;
    o = (o + 15) / 16 * 16;
    L.off_bar = o; o += (size_t)(2 * kPMaxStages + 2 + 2 + 2 + 1 + 2) * 8;   // w_full, w_empty, xr[2], x0r[2], acc_full[2], head_full, tmem slot
    L.total = o;
    return L;
}

struct TcpParams {
    RolloutParams R;
    const unsigned char* wimg;
    long long wimg_member_stride;
    int Np, NHp, nkb0, nkbH, Kcap;
    int terms, stages;
    int N;                       // rows per subtile
    int pairs_per_member, total_pairs;
    int nent[2];                 // weight stages per horizon step of CTA rank 0 / 1
    uint2 tab[2][kPMaxEnt];      // {byte offset inside the member's weight image, bytes}
    long long* dbg;              // nullable: clock64 trace of pair 0 at step h / 2 (DIAG instantiation only), [rank][128]
};

namespace tcp {

// The MMAs of ONE M tile of one GEMM for one subtile: NKB K16 blocks in weight stages of 4, fully unrolled.
template <int TERMS, int NKB, int MM>
__device__ __forceinline__ void issue_tile(uint32_t d_tmem, uint32_t R, uint32_t w16, uint32_t slot16, uint32_t nstage, tcs::RingPos& rp,
                                           uint32_t xg16, uint32_t rows, uint32_t xsbo, uint64_t* w_full, uint64_t* w_empty) {
    const uint32_t hi32 = (1u << 14);
    const uint64_t a_top = (uint64_t)(hi32 | (128u >> 4)) << 32;
    const uint64_t b_top = (uint64_t)(hi32 | (xsbo >> 4)) << 32;
    const uint32_t b_lbo = (128u >> 4) << 16;
    const uint32_t idesc1 = tcs::idesc(rows, MM), idesc2 = tcs::idesc(2u * rows, MM);
#pragma unroll
    for (int s0 = 0; s0 < NKB; s0 += 4) {
        const int kbs = (NKB - s0) < 4 ? (NKB - s0) : 4;
        ptx::mbar_wait(&w_full[rp.slot], rp.phase);
        const uint32_t slot = w16 + rp.slot * slot16;
        const uint64_t a_hi = a_top | (slot | (R << 16));
        const uint64_t a_lo = a_top | ((slot + (uint32_t)kbs * 2u * R) | (R << 16));
        const uint64_t b = b_top | ((xg16 + (uint32_t)s0 * 16u) | b_lbo);
        const uint32_t accum = s0 > 0 ? 1u : 0u;
        if (kbs == 4) tcs::issue_blocks<TERMS, 4>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
        else if (kbs == 3) tcs::issue_blocks<TERMS, 3>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
        else if (kbs == 2) tcs::issue_blocks<TERMS, 2>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
        else tcs::issue_blocks<TERMS, 1>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
        tc::mma_commit(&w_empty[rp.slot]);
        if (++rp.slot == nstage) { rp.slot = 0; rp.phase ^= 1u; }
    }
}

}  // namespace tcp

// TERMS: 3 (fp16 hi/lo split) or 1; NKB0: K16 blocks of layer 0 (2 or 3); NHP: padded width of the heads (48 or 64).
template <int TERMS, int NKB0, int NHP, bool DIAG>
__global__ void __launch_bounds__(kPThreads, 1) rollout_tcp_kernel(const __grid_constant__ TcpParams T) {
    const RolloutParams& P = T.R;
    long long* const dbgbase = DIAG ? T.dbg : nullptr;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int N = T.N;
    const TcpSmem L = tcp_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, T.stages);
    unsigned char* xbuf = smem + L.off_x;          // subtile s: X_hi at s * 2 xbytes, X_lo directly behind it
    unsigned char* x0buf = smem + L.off_x0;        // subtile s: layer-0 X_hi at s * 2 x0bytes, X_lo behind it
    unsigned char* wring = smem + L.off_w;
    float* S = reinterpret_cast<float*>(smem + L.off_s);
    float* Hd = reinterpret_cast<float*>(smem + L.off_hd);
    float* bias = reinterpret_cast<float*>(smem + L.off_bias);
    float* vec = reinterpret_cast<float*>(smem + L.off_vec);
    int* rowi = reinterpret_cast<int*>(smem + L.off_rowi);
    int4* feat_i = reinterpret_cast<int4*>(smem + L.off_feat);
    float2* feat_f = reinterpret_cast<float2*>(smem + L.off_feat + 96 * 16);
    float* act_s = reinterpret_cast<float*>(smem + L.off_act);
    float* ctx_s = reinterpret_cast<float*>(smem + L.off_ctx);
    float* nz_s = reinterpret_cast<float*>(smem + L.off_nz);
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* w_empty = w_full + kPMaxStages;
    uint64_t* xr = w_empty + kPMaxStages;          // [subtile]: a hidden layer's output = the next GEMM's operand is complete (both CTAs' warps)
    uint64_t* x0r = xr + 2;                        // [subtile]: the layer-0 operand of the next step is complete (the owner CTA's warps)
    uint64_t* acc_full = x0r + 2;                  // [subtile]: BOTH CTAs' MMAs of this GEMM have completed (multicast commits)
    uint64_t* head_full = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(head_full + 1);
    volatile uint32_t* done = tmem_slot + 1;

    float4* fc = reinterpret_cast<float4*>(vec);
    int* r_mi = rowi;
    int* r_src = rowi + N;
    int* r_pi = rowi + 2 * N;
    int* r_ctx = rowi + 3 * N;
    int* r_rid = rowi + 4 * N;
    int* r_eps = rowi + 5 * N;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = ptx::cluster_ctarank();          // which M tile of the hidden layers, and which subtile's tail
    const uint32_t peer = rank ^ 1u;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nstage = T.stages;
    const int xsbo = (T.Kcap / 8) * 128;
    const int xsbo0 = T.nkb0 * 2 * 128;
    const uint32_t Rme = rank == 0 ? 128u : (uint32_t)(T.Np - 128);      // rows (hidden units) of this CTA's weight tile
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&xr[s], 32);            // 16 epilogue warps of each CTA
            ptx::mbar_init(&x0r[s], 16);           // the 16 epilogue warps of the subtile's owner
            ptx::mbar_init(&acc_full[s], 2);       // one commit per CTA
        }
        ptx::mbar_init(head_full, 1);
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
    ptx::cluster_sync();                           // the peer's barriers exist before anything is sent to them
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    // distance from an address in this CTA's shared memory to the same location in the peer's (shared::cluster window)
    const uint32_t smem0 = ptx::smem_u32(smem);
    const uint32_t rdelta = ptx::mapa(smem0, peer) - smem0;

    // ======================= warp 0: weight producer =============================================
    if (warp == 0) {
        if (lane == 0) {
            tcs::Ring rp{0, 0, nstage};
            const uint2* tab = T.tab[rank];
            const int nent = T.nent[rank];
            for (int pt = pair_id; pt < T.total_pairs; pt += n_pairs) {
                const int e = pt / T.pairs_per_member;
                const unsigned char* wsrc = T.wimg + (size_t)e * T.wimg_member_stride;
                for (int t = 0; t < P.h; ++t) {
                    for (int i = 0; i < nent; ++i) {
                        const uint2 en = tab[i];
                        ptx::mbar_wait(&w_empty[rp.stage], rp.phase ^ 1u);
                        ptx::mbar_arrive_expect_tx(&w_full[rp.stage], en.y);
                        ptx::bulk_g2s(wring + (size_t)rp.stage * L.slot_bytes, wsrc + en.x, en.y, &w_full[rp.stage]);
                        rp.advance();
                    }
                }
            }
            done[0] = 1u;
        } else {
            while (!done[0]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warp 1: MMA issuer ==================================================
    else if (warp == 1) {
        if (ptx::elect_one()) {
            // In a cluster the 32-bit shared::cta address carries the CTA's rank above bit 24 (measured: 0x01000400 for rank 1,
            // tools/probes/dsmem_probe.cu); the matrix descriptors take the 18-bit offset inside the CTA's own window
            auto desc16 = [](uint32_t saddr) { return (saddr & 0x3FFFFu) >> 4; };
            const uint32_t w16 = desc16(ptx::smem_u32(wring)), slot16 = (uint32_t)L.slot_bytes >> 4;
            const uint32_t xs16[2] = {desc16(ptx::smem_u32(xbuf)), desc16(ptx::smem_u32(xbuf) + 2u * (uint32_t)L.xbytes)};
            const uint32_t x0s16[2] = {desc16(ptx::smem_u32(x0buf)), desc16(ptx::smem_u32(x0buf) + 2u * (uint32_t)L.x0bytes)};
            tcs::RingPos rp{0u, 0u};
            uint32_t ph_x0[2] = {0u, 0u}, ph_xr[2] = {0u, 0u};
            for (int pt = pair_id; pt < T.total_pairs; pt += n_pairs) {
                for (int t = 0; t < P.h; ++t) {
                    long long* md = (dbgbase && pair_id == 0 && pt == pair_id && t == P.h / 2) ? dbgbase + rank * 128 + 64 : nullptr;
                    if (md) md[0] = clock64();
                    for (int g = 0; g < P.n_hidden; ++g) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)s * kPSubStride;
                            if (md && g < 5) md[1 + (g * 2 + s) * 3] = clock64();
                            if (g == 0) {
                                ptx::mbar_wait_cluster(&x0r[s], ph_x0[s]);
                                ph_x0[s] ^= 1u;
                                ptx::fence_proxy_async_all();
                                tc::fence_after_sync();
                                if (md && g < 5) md[2 + (g * 2 + s) * 3] = clock64();
                                tcp::issue_tile<TERMS, NKB0, 128>(d_tmem, Rme, w16, slot16, (uint32_t)nstage, rp, x0s16[s], (uint32_t)N, (uint32_t)xsbo0,
                                                                  w_full, w_empty);
                            } else {
                                ptx::mbar_wait_cluster(&xr[s], ph_xr[s]);
                                ph_xr[s] ^= 1u;
                                ptx::fence_proxy_async_all();
                                tc::fence_after_sync();
                                if (md && g < 5) md[2 + (g * 2 + s) * 3] = clock64();
                                tcp::issue_tile<TERMS, 13, 128>(d_tmem, Rme, w16, slot16, (uint32_t)nstage, rp, xs16[s], (uint32_t)N, (uint32_t)xsbo,
                                                                w_full, w_empty);
                            }
                            tc::mma_commit_multicast(&acc_full[s], (uint16_t)3);
                            if (md && g < 5) md[3 + (g * 2 + s) * 3] = clock64();
                        }
                    }
                    // heads of this CTA's own subtile (M = 64: NHP <= 64 outputs)
                    if (md) md[40] = clock64();
                    ptx::mbar_wait_cluster(&xr[rank], ph_xr[rank]);
                    if (md) md[41] = clock64();
                    ph_xr[rank] ^= 1u;
                    ph_xr[peer] ^= 1u;             // the peer subtile's last hidden output also lands here; nobody on this CTA reads it
                    ptx::fence_proxy_async_all();
                    tc::fence_after_sync();
                    tcp::issue_tile<TERMS, 13, 64>(tmem_base + kPHeadCol, (uint32_t)NHP, w16, slot16, (uint32_t)nstage, rp, xs16[rank], (uint32_t)N,
                                                   (uint32_t)xsbo, w_full, w_empty);
                    tc::mma_commit(head_full);
                    if (md) md[42] = clock64();
                }
            }
            done[1] = 1u;
        } else {
            while (!done[1]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warps 2..17: prologue / epilogue ======================================
    else {
        const int et = tid - 64;
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int cslice = ew >> 2;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int nchunks = N >> 3;
        const int nq_head = (NHP + 15) >> 4;                                     // M = 64 heads: 16 outputs per TMEM lane quarter
        const bool helper = quarter >= nq_head;
        const int n_help = 128 * (4 - nq_head);
        const int ht = ((quarter - nq_head) + (4 - nq_head) * cslice) * 32 + lane;
        const int qlane = quarter * 32 + lane;
        const bool act_me = quarter * 32 < (int)Rme;                             // the warp has hidden units of this CTA's tile
        const bool ok_me = qlane < (int)Rme;
        const int unit = (int)rank * 128 + qlane;                                // this thread's hidden unit = K index of the next layer
        const int so = cslice * xsbo + (unit >> 3) * 128 + (unit & 7) * 16;
        const uint32_t tc0 = tmem_lane + (uint32_t)cslice * 8u;
        const uint32_t xbuf_s = ptx::smem_u32(xbuf);
        const uint32_t hd_s = ptx::smem_u32(Hd);
        const uint32_t x0_own = ptx::smem_u32(x0buf) + rank * 2u * (uint32_t)L.x0bytes;       // layer-0 operand of this CTA's own subtile
        uint64_t* const xr_peer0 = nullptr;
        (void)xr_peer0;
        const uint32_t xr_rem[2] = {ptx::smem_u32(&xr[0]) + rdelta, ptx::smem_u32(&xr[1]) + rdelta};
        const uint32_t x0r_rem = ptx::smem_u32(&x0r[rank]) + rdelta;
        uint32_t cnt_acc[2] = {0u, 0u};                // completions of acc_full[s]
        uint32_t cnt_head = 0u;
        const int D = P.D, A = P.A;
        const size_t eps_step_stride = (size_t)P.E * P.q * P.m * P.n_global * D;
        const int nj = (D + 3) >> 2;

        // layer-0 feature table and the per-dimension constants of the final epilogue: as in rollout_tcs.cu
        const int K0 = T.nkb0 * 16;
        for (int k = et; k < K0; k += kPEpiThreads) {
            int base = (int)L.off_zero, stride = 0, flags = 0;
            float mean = 0.f, inv = 0.f;
            if (k < P.P) {
                int idx = k;
                if (P.env_id == CADM_ENV_HALFCHEETAH) {
                    if (k == 0) idx = 1;
                    else if (k == 1) idx = D + 1;
                    else if (k == 2) idx = D + 2;
                } else if (P.env_id == CADM_ENV_ANT) {
                    idx = k + 1;
                }
                base = (int)L.off_s + idx * 4; stride = (D + 3) * 4;
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
        ptx::bar_sync(1, kPEpiThreads);
        for (int d = et; d < D; d += kPEpiThreads) {
            int kf = d;
            if (P.env_id == CADM_ENV_HALFCHEETAH) kf = d <= 1 ? d - 1 : (d == 2 ? -1 : d);
            else if (P.env_id == CADM_ENV_ANT) kf = d - 1;
            const float ds = P.delta_std[d];
            float finv = 0.f, fofs = 0.f;
            if (kf >= 0) { const float2 ff = feat_f[kf]; finv = ff.y; fofs = -ff.x * ff.y; }
            fc[2 * d] = make_float4(ds + 1e-10f, P.delta_mean[d], P.max_lv[d] * tc::kLog2e, ds * ds * expf(P.max_lv[d]));
            fc[2 * d + 1] = make_float4(ds * ds * expf(P.min_lv[d]), finv, fofs, __int_as_float(kf >= 0 ? (kf >> 3) * 128 + (kf & 7) * 16 : -1));
        }

        for (int pt = pair_id; pt < T.total_pairs; pt += n_pairs) {
            const int e = pt / T.pairs_per_member;
            const int pair_row0 = (pt - e * T.pairs_per_member) * P.rows_per_cta;               // rows_per_cta = rows of one PAIR tile
            const int pair_rows = min(P.rows_per_cta, P.rows_per_member - pair_row0);
            const int rows_a = min(N, pair_rows);                                               // subtile A takes the first N rows
            const int tile_row0 = pair_row0 + (rank ? rows_a : 0);                              // this CTA's own subtile
            const int nrows = rank ? pair_rows - rows_a : rows_a;
            ptx::bar_sync(1, kPEpiThreads);
            for (int i = et; i < P.n_hidden * T.Np + T.NHp; i += kPEpiThreads)
                bias[i] = P.bpack[(size_t)e * P.bias_stride + i] * (i < P.n_hidden * T.Np ? -tc::kLog2e : 1.0f);
            if (et < N) {
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
            ptx::bar_sync(1, kPEpiThreads);
            for (int i = et; i < N * D; i += kPEpiThreads) {
                const int r = i / D, d = i - r * D;
                float v = 0.f;
                if (r < nrows) v = (P.row_mode == kRowsPlanner) ? P.obs0[r_mi[r] * D + d] : P.obs0[(size_t)r_src[r] * D + d];
                S[r * (D + 3) + d] = v;
            }
            for (int i = et; i < N * P.C; i += kPEpiThreads) {
                const int r = i / P.C, c = i - r * P.C;
                ctx_s[i] = r < nrows ? __ldg(P.ctx + (size_t)r_ctx[r] * P.C + c) : 0.f;
            }
            auto prefetch_actions = [&](int t, int first, int stride) {
                if (P.discrete && P.row_mode == kRowsPlanner) return;
                float* dst = act_s + (t & 1) * N * A;
                for (int i = first; i < nrows * A; i += stride) {
                    const int r = i / A, a = i - r * A;
                    tcs::cp_async4(dst + i, P.actions + ((size_t)r_src[r] * P.h + t) * A + a);
                }
                tcs::cp_async_commit();
            };
            for (int i = et; i < 2 * N * A; i += kPEpiThreads) act_s[i] = 0.f;
            ptx::bar_sync(1, kPEpiThreads);
            if (P.env_id == CADM_ENV_HALFCHEETAH && et < N) sincosf(S[et * (D + 3) + 2], &S[et * (D + 3) + D + 1], &S[et * (D + 3) + D + 2]);
            griddep_wait();                            // the actions come from the kernel launched before this one (common.cuh)
            prefetch_actions(0, et, kPEpiThreads);
            tcs::cp_async_wait_all();
            ptx::bar_sync(1, kPEpiThreads);

            float ret = 0.f;
            const int invN = ((1 << 20) + N - 1) / N;
            const int ndp = (D + 1) >> 1;
            const int act_slot = P.n_hidden > 1 ? 1 : 0;
            const bool onehot_mode = P.discrete && P.row_mode == kRowsPlanner;
            const bool is_hc = P.env_id == CADM_ENV_HALFCHEETAH;
            const bool want_out = P.states != nullptr || P.next_obs != nullptr || P.mu_out != nullptr || P.lv_out != nullptr;
            // layer-0 features k in [k_lo, k_hi) of step t for this CTA's OWN subtile, written to both CTAs' copy of its operand
            auto build_features = [&](int t, int k_lo, int k_hi, int first, int stride) {
                const int par_off = (t & 1) * N * A * 4;
                const int nk = k_hi - k_lo;
                for (int i = first; i < nk * (N >> 1); i += stride) {
                    const int rp = i / nk, k = k_lo + (i - rp * nk);
                    const int4 fi = feat_i[k];
                    const float2 ff = feat_f[k];
                    float y[2];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int r = rp * 2 + jj;
                        float src = *reinterpret_cast<const float*>(smem + fi.x + r * fi.y + ((fi.z & 1) ? par_off : 0));
                        if (onehot_mode && (fi.z & 1)) {
                            const int oh = r < nrows ? __ldg(P.actions_int + (size_t)r_src[r] * P.h + t) : -1;
                            src = (oh == k - P.P) ? 1.f : 0.f;
                        }
                        y[jj] = r < nrows ? (src - ff.x) * ff.y : 0.f;
                    }
                    uint32_t hq, lq;
                    tc::split2(y[0], y[1], hq, lq);
                    const uint32_t o = x0_own + (uint32_t)((rp >> 2) * xsbo0 + (k >> 3) * 128 + (k & 7) * 16 + (rp & 3) * 4);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(o), "r"(hq) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(o + (uint32_t)L.x0bytes), "r"(lq) : "memory");
                    ptx::sts32_cluster(o + rdelta, hq);
                    ptx::sts32_cluster(o + rdelta + (uint32_t)L.x0bytes, lq);
                }
            };
            // hand this subtile's layer-0 operand to the MMA threads of both CTAs
            auto publish_input = [&]() {
                ptx::fence_proxy_async_all();
                __syncwarp();
                if (lane == 0) { ptx::mbar_arrive(&x0r[rank]); ptx::mbar_arrive_cluster(x0r_rem); }
            };
            auto add_reward = [&](int t) {
                if (et < nrows && !env_reward_reads_next(P.env_id))
                    ret += env_reward_current(P.env_id, S + et * (D + 3), act_s + (t & 1) * N * A + et * A, A, P.max_torque);
            };
            build_features(0, 0, K0, et, kPEpiThreads);
            publish_input();
            add_reward(0);
#pragma unroll 1
            for (int t = 0; t < P.h; ++t) {
                long long* ed = (dbgbase && pair_id == 0 && pt == pair_id && t == P.h / 2 && ew == 0 && lane == 0) ? dbgbase + rank * 128 : nullptr;
                if (ed) ed[0] = clock64();
                // ---------- hidden layers: for each subtile, this CTA's units: accumulator -> bias + swish -> operand of both CTAs
#pragma unroll 1
                for (int l = 0; l < P.n_hidden; ++l) {
                    const float su = l == 0 ? -tc::kLog2e / (tc::kWScale * tc::kXScale) : 1.0f / (tc::kWScale * 8.0f);
                    const float2 su2 = make_float2(su, su);
                    const float bu = ok_me ? bias[l * T.Np + unit] : 0.f;
                    const float2 bu2 = make_float2(bu, bu);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        uint32_t tcol = tc0 + (uint32_t)s * kPSubStride;
                        uint32_t xo = xbuf_s + (uint32_t)(s * 2 * L.xbytes + so);
                        if (ed && l < 5) ed[1 + (l * 2 + s) * 5] = clock64();
                        ptx::mbar_wait_cluster(&acc_full[s], cnt_acc[s] & 1u);
                        ++cnt_acc[s];
                        tc::fence_after_sync();
                        if (ed && l < 5) ed[2 + (l * 2 + s) * 5] = clock64();
                        if (act_me) {
                            for (int c = cslice; c < nchunks; c += 4, tcol += 32u, xo += 4u * (uint32_t)xsbo) {
                                uint32_t v[8];
                                float2 a[4];
                                tc::tmem_ld8(tcol, v);
                                if (TERMS == 3) {
                                    uint32_t v1[8];
                                    tc::tmem_ld8(tcol + N, v1);
                                    tc::tmem_wait_ld();
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        a[j] = tc::fadd2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                                         make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])));
                                } else {
                                    tc::tmem_wait_ld();
#pragma unroll
                                    for (int j = 0; j < 4; ++j) a[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                                }
                                float2 y[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) y[j] = tc::swish_pair_u(tc::ffma2(a[j], su2, bu2));
                                if (ok_me) {
                                    uint32_t hq[4], lq[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) tc::split2(y[j].x, y[j].y, hq[j], lq[j]);
                                    ptx::sts128(xo, hq[0], hq[1], hq[2], hq[3]);
                                    ptx::sts128(xo + (uint32_t)L.xbytes, lq[0], lq[1], lq[2], lq[3]);
                                    ptx::sts128_cluster(xo + rdelta, hq[0], hq[1], hq[2], hq[3]);
                                    ptx::sts128_cluster(xo + rdelta + (uint32_t)L.xbytes, lq[0], lq[1], lq[2], lq[3]);
                                }
                            }
                        }
                        if (ed && l < 5) ed[3 + (l * 2 + s) * 5] = clock64();
                        tc::fence_before_sync();
                        ptx::fence_proxy_async_all();
                        if (ed && l < 5) ed[4 + (l * 2 + s) * 5] = clock64();
                        __syncwarp();
                        if (lane == 0) { ptx::mbar_arrive(&xr[s]); ptx::mbar_arrive_cluster(xr_rem[s]); }
                        if (ed && l < 5) ed[5 + (l * 2 + s) * 5] = clock64();
                    }
                    // ---------- off the critical path (the warps now wait for the next layer's first accumulator)
                    if (l == 0) {
                        if (t + 1 < P.h) {
                            if (n_help == 0) prefetch_actions(t + 1, et, kPEpiThreads);
                            else if (helper) prefetch_actions(t + 1, ht, n_help);
                        }
                        if (!P.deterministic) {
                            for (int i = et; i < N * nj; i += kPEpiThreads) {
                                const int jb = (i * invN) >> 20, r = i - jb * N;
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
                    if (l == act_slot && t + 1 < P.h) {
                        // the next step's action features of the own subtile: both CTAs' layer-0 MMAs of this step have read the
                        // operand (every warp has passed acc_full[rank] of layer 0, which counts both commits)
                        if (n_help == 0) {
                            tcs::cp_async_wait_all();
                            ptx::bar_sync(1, kPEpiThreads);
                            build_features(t + 1, P.P, P.P + A, et, kPEpiThreads);
                        } else if (helper) {
                            tcs::cp_async_wait_all();
                            ptx::bar_sync(2, n_help);
                            build_features(t + 1, P.P, P.P + A, ht, n_help);
                        }
                    }
                }

                // ---------- tail of the OWN subtile: heads -> Hd, sample, next state, next step's state features ----------
                const int it0 = et;
                const int dp0 = (it0 * invN) >> 20, r0 = it0 - dp0 * N, d00 = 2 * dp0, d01 = d00 + 1;
                const bool item0 = d00 < D && r0 < nrows, two0 = d01 < D;
                float4 ca0[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
                float4 cb0[2] = {make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), make_float4(0.f, 0.f, 0.f, __int_as_float(-1))};
                float s0[2] = {0.f, 0.f};
                if (item0) {
                    ca0[0] = fc[2 * d00]; cb0[0] = fc[2 * d00 + 1];
                    s0[0] = S[r0 * (D + 3) + d00];
                    if (two0) { ca0[1] = fc[2 * d01]; cb0[1] = fc[2 * d01 + 1]; s0[1] = S[r0 * (D + 3) + d01]; }
                }
                {
                    const uint32_t par0 = cnt_head & 1u;
                    ++cnt_head;
                    if (quarter < nq_head) {
                        const int jh = quarter * 16 + (lane & 15);                 // M = 64: accumulator row j in TMEM lane 32 (j / 16) + j % 16
                        const bool lane_ok = lane < 16 && jh < NHP;
                        const float bj = jh < NHP ? bias[P.n_hidden * T.Np + jh] : 0.f;
                        const float sc = 1.0f / (tc::kWScale * tc::kActScale);
                        const float2 sc2 = make_float2(sc, sc), bj2 = make_float2(bj, bj);
                        uint32_t hd_a = hd_s + (uint32_t)((jh * (N + 4) + cslice * 8) * 4);
                        uint32_t tcol = tc0 + kPHeadCol;
                        if (ed) ed[52] = clock64();
                        ptx::mbar_wait(head_full, par0);
                        tc::fence_after_sync();
                        if (ed) ed[53] = clock64();
                        for (int c = cslice; c < nchunks; c += 4, tcol += 32u, hd_a += 128u) {
                            uint32_t v[8];
                            float2 a[4];
                            tc::tmem_ld8(tcol, v);
                            if (TERMS == 3) {
                                uint32_t v1[8];
                                tc::tmem_ld8(tcol + N, v1);
                                tc::tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    a[j] = tc::fadd2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                                     make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])));
                            } else {
                                tc::tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 4; ++j) a[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                            }
                            if (lane_ok) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) a[j] = tc::ffma2(a[j], sc2, bj2);
                                ptx::sts128(hd_a, __float_as_uint(a[0].x), __float_as_uint(a[0].y), __float_as_uint(a[1].x), __float_as_uint(a[1].y));
                                ptx::sts128(hd_a + 16u, __float_as_uint(a[2].x), __float_as_uint(a[2].y), __float_as_uint(a[3].x), __float_as_uint(a[3].y));
                            }
                        }
                        tc::fence_before_sync();
                    }
                }
                ptx::bar_sync(1, kPEpiThreads);

                // final epilogue: one item = the state dimensions (d0, d0 + 1) of row r (core/utils.py:84-90, 162; see rollout_tcs.cu)
                auto final_pair = [&](int r, int d0, bool two, const float4 (&ca)[2], const float4 (&cb)[2], const float (&s_old)[2]) {
                    float mu[2] = {0.f, 0.f}, lv[2] = {0.f, 0.f}, nz[2] = {0.f, 0.f}, sn[2];
                    mu[0] = Hd[d0 * (N + 4) + r];
                    lv[0] = Hd[(D + d0) * (N + 4) + r];
                    if (two) { mu[1] = Hd[(d0 + 1) * (N + 4) + r]; lv[1] = Hd[(D + d0 + 1) * (N + 4) + r]; }
                    if (!P.deterministic) {
                        const float2 z2 = *reinterpret_cast<const float2*>(nz_s + r * (nj * 4) + d0);
                        nz[0] = z2.x; nz[1] = z2.y;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float delta = fmaf(mu[j], ca[j].x, ca[j].y);
                        if (!P.deterministic) {
                            const float q = tc::ex2_approx(fmaf(lv[j], -tc::kLog2e, ca[j].z));
                            const float var = fmaf(ca[j].w, tc::rcp_approx(1.0f + q), cb[j].x);
                            float sd;
                            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(var));
                            delta = fmaf(nz[j], sd, delta);
                        }
                        sn[j] = env_postproc(P.env_id, s_old[j], delta, d0 + j);
                    }
                    uint32_t hq[2] = {0u, 0u}, lq[2] = {0u, 0u}, hs = 0u, ls = 0u;
                    const bool feat = t + 1 < P.h;
                    const bool angle = is_hc && d0 == 2;
                    if (feat) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) tc::split2(fmaf(sn[j], cb[j].y, cb[j].z), 0.f, hq[j], lq[j]);
                        if (angle) {
                            float f0, f1;
                            tcs::sincos_reduced(sn[0], f0, f1);
                            const float2 fa = feat_f[1], fb = feat_f[2];
                            tc::split2((f0 - fa.x) * fa.y, (f1 - fb.x) * fb.y, hs, ls);
                        }
                    }
                    S[r * (D + 3) + d0] = sn[0];
                    if (two) S[r * (D + 3) + d0 + 1] = sn[1];
                    if (feat) {
                        const uint32_t xr0 = x0_own + (uint32_t)((r >> 3) * xsbo0 + (r & 7) * 2);
                        const uint32_t lo_off = (uint32_t)L.x0bytes;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int ko = __float_as_int(cb[j].w);
                            if (ko >= 0 && (j == 0 || two)) {
                                ptx::sts16(xr0 + ko, hq[j]);
                                ptx::sts16(xr0 + lo_off + ko, lq[j]);
                                ptx::sts16_cluster(xr0 + rdelta + ko, hq[j]);
                                ptx::sts16_cluster(xr0 + rdelta + lo_off + ko, lq[j]);
                            }
                        }
                        if (angle) {
                            ptx::sts16(xr0 + 16, hs);
                            ptx::sts16(xr0 + lo_off + 16, ls);
                            ptx::sts16(xr0 + 32, hs >> 16);
                            ptx::sts16(xr0 + lo_off + 32, ls >> 16);
                            ptx::sts16_cluster(xr0 + rdelta + 16, hs);
                            ptx::sts16_cluster(xr0 + rdelta + lo_off + 16, ls);
                            ptx::sts16_cluster(xr0 + rdelta + 32, hs >> 16);
                            ptx::sts16_cluster(xr0 + rdelta + lo_off + 32, ls >> 16);
                        }
                    }
                    if (want_out) {
                        for (int j = 0; j < (two ? 2 : 1); ++j) {
                            const int d = d0 + j;
                            if (P.row_mode == kRowsPlanner) {
                                if (P.states != nullptr)
                                    P.states[(((size_t)t * P.m * P.n_local + r_src[r]) * P.p + r_pi[r]) * D + d] = sn[j];
                            } else {
                                const size_t o = (size_t)r_src[r] * D + d;
                                if (P.next_obs) P.next_obs[o] = sn[j];
                                if (P.mu_out) P.mu_out[o] = mu[j];
                                if (P.lv_out) P.lv_out[o] = P.deterministic ? lv[j] : fast_bounded_logvar(lv[j], P.max_lv[d], P.min_lv[d]);
                            }
                        }
                    }
                };
                if (item0) final_pair(r0, d00, two0, ca0, cb0, s0);
                for (int i = et + kPEpiThreads; i < N * ndp; i += kPEpiThreads) {
                    const int dp = (i * invN) >> 20, r = i - dp * N;
                    if (r >= nrows) continue;
                    const int d0 = 2 * dp;
                    const bool two = d0 + 1 < D;
                    const float4 ca[2] = {fc[2 * d0], two ? fc[2 * d0 + 2] : make_float4(0.f, 0.f, 0.f, 0.f)};
                    const float4 cb[2] = {fc[2 * d0 + 1], two ? fc[2 * d0 + 3] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1))};
                    const float so2[2] = {S[r * (D + 3) + d0], two ? S[r * (D + 3) + d0 + 1] : 0.f};
                    final_pair(r, d0, two, ca, cb, so2);
                }
                if (ed) ed[54] = clock64();
                if (t + 1 < P.h) publish_input();
                if (ed) ed[55] = clock64();
                ptx::bar_sync(1, kPEpiThreads);
                if (ed) ed[56] = clock64();
                if (env_reward_reads_next(P.env_id)) {
                    if (et < nrows) ret += env_reward_next(P.env_id, S + et * (D + 3));
                } else if (t + 1 < P.h) {
                    add_reward(t + 1);
                }
            }
            if (P.row_mode == kRowsPlanner && et < nrows) P.ret_p[(size_t)r_src[et] * P.p + r_pi[et]] = ret;
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    ptx::cluster_sync();                           // the peer may still be storing into / arriving on this CTA's shared memory
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
static int g_tcp_smem = 0;

bool tcp_supported(const RolloutParams& P, int kps) {
    const int Np = round_up(P.H, 16), NHp = round_up(2 * P.D, 16), nkb0 = round_up(P.In, 16) / 16;
    return kps == 4 && Np == 208 && P.n_hidden >= 1 && (nkb0 == 2 || nkb0 == 3) && (NHp == 48 || NHp == 64) && !(nkb0 == 2 && NHp == 64);
}

cudaError_t launch_rollout_tcp(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms, int rows_override,
                               int num_sms, cudaStream_t stream, const char** name, long long* dbg) {
    if (!tcp_supported(P, 4)) return cudaErrorInvalidConfiguration;
    TcpParams T{};
    T.wimg = wimg;
    T.dbg = dbg;
    T.wimg_member_stride = wimg_member_stride;
    T.Np = round_up(P.H, 16);
    T.NHp = round_up(2 * P.D, 16);
    T.nkb0 = round_up(P.In, 16) / 16;
    T.nkbH = T.Np / 16;
    T.Kcap = max(T.Np, T.nkb0 * 16);
    T.terms = terms;
    const int max_pairs = num_sms / 2;
    int N = rows_override > 0 ? rows_override : 32;
    if (rows_override <= 0)        // the smallest subtile whose pair tiles fit the SM pairs in one wave
        while (N < kPMaxRows && P.E * ((P.rows_per_member + 2 * N - 1) / (2 * N)) > max_pairs) N += 16;
    N = min(kPMaxRows, max(16, round_up(N, 16)));
    while (N > 16 && tcp_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, 2).total > 226 * 1024) N -= 16;
    int pairs = (P.rows_per_member + 2 * N - 1) / (2 * N);
    const int rows_per_pair = (P.rows_per_member + pairs - 1) / pairs;          // balanced
    N = min(N, round_up((rows_per_pair + 1) / 2, 16));
    pairs = (P.rows_per_member + 2 * N - 1) / (2 * N);
    P.rows_per_cta = min(2 * N, (P.rows_per_member + pairs - 1) / pairs);
    pairs = (P.rows_per_member + P.rows_per_cta - 1) / P.rows_per_cta;
    T.N = N;
    T.pairs_per_member = pairs;
    T.total_pairs = pairs * P.E;
    int stages = kPMaxStages;
    while (stages > 2 && tcp_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, stages).total > 226 * 1024) --stages;
    T.stages = stages;
    const TcpSmem L = tcp_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, stages);
    if (L.total > 226 * 1024) return cudaErrorInvalidConfiguration;
    T.R = P;
    // The weight stages of one horizon step per CTA rank, in consumption order: hidden GEMM g, this rank's M tile, once for
    // subtile A and once for subtile B; then the heads (both ranks: each computes them for its own subtile).  Offsets follow the
    // image of rollout_tcs.cu: per GEMM [tile 0: nkb x 64 x 128 bytes][tile 1], a stage of kbs K blocks = kbs x 64 x R bytes.
    for (int rank = 0; rank < 2; ++rank) {
        int n = 0;
        long long goff = 0;
        for (int g = 0; g <= P.n_hidden; ++g) {
            const int nkb = g == 0 ? T.nkb0 : T.nkbH;
            const bool head = g == P.n_hidden;
            const int Npad = head ? T.NHp : T.Np;
            const int R = head ? T.NHp : (rank == 0 ? 128 : T.Np - 128);
            const long long toff = goff + ((!head && rank == 1) ? (long long)nkb * 64 * 128 : 0);
            for (int rep = 0; rep < (head ? 1 : 2); ++rep)
                for (int s0 = 0; s0 < nkb; s0 += 4) {
                    const int kbs = std::min(4, nkb - s0);
                    if (n >= kPMaxEnt) return cudaErrorInvalidConfiguration;
                    T.tab[rank][n++] = make_uint2((unsigned)(toff + (long long)s0 * 64 * R), (unsigned)(kbs * 64 * R));
                }
            goff += (long long)nkb * Npad * 64;
        }
        T.nent[rank] = n;
    }
    if ((int)L.total > g_tcp_smem) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); };
#define CADM_TCP_SET(TE, K0, HP) set((const void*)rollout_tcp_kernel<TE, K0, HP, false>); set((const void*)rollout_tcp_kernel<TE, K0, HP, true>)
        CADM_TCP_SET(3, 2, 48); CADM_TCP_SET(1, 2, 48); CADM_TCP_SET(3, 3, 48); CADM_TCP_SET(1, 3, 48); CADM_TCP_SET(3, 3, 64); CADM_TCP_SET(1, 3, 64);
#undef CADM_TCP_SET
        if (e != cudaSuccess) return e;
        g_tcp_smem = (int)L.total;
    }
    if (name) *name = terms == 3 ? "rollout_tcp_kernel(CTA pairs, swapped operands, fp16 hi/lo x3)" : "rollout_tcp_kernel(CTA pairs, swapped operands, f16 x1)";
    const int grid = 2 * min(T.total_pairs, max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPThreads);
    cfg.dynamicSmemBytes = L.total;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 2 : 1;
    cudaError_t le;
    const bool diag = dbg != nullptr;
#define CADM_TCP_LAUNCH(TE, K0, HP) (diag ? cudaLaunchKernelEx(&cfg, rollout_tcp_kernel<TE, K0, HP, true>, T) : cudaLaunchKernelEx(&cfg, rollout_tcp_kernel<TE, K0, HP, false>, T))
    if (T.nkb0 == 2) le = terms == 3 ? CADM_TCP_LAUNCH(3, 2, 48) : CADM_TCP_LAUNCH(1, 2, 48);
    else if (T.NHp == 48) le = terms == 3 ? CADM_TCP_LAUNCH(3, 3, 48) : CADM_TCP_LAUNCH(1, 3, 48);
    else le = terms == 3 ? CADM_TCP_LAUNCH(3, 3, 64) : CADM_TCP_LAUNCH(1, 3, 64);
#undef CADM_TCP_LAUNCH
    return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace cadm
