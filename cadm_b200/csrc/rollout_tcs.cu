// Persistent tensor-core rollout kernel, SWAPPED operands ("small-tile" variant of rollout_tc.cu).
//
// The planner's batch is small: HalfCheetah PE-TS has 800 rows per ensemble member, i.e. 35 tiles of 128 rows for a
// 148-SM part.  This variant puts the HIDDEN UNITS on the MMA's M axis and the rows on its N axis,
//
//      D^T[hidden (TMEM lanes) x rows (TMEM columns)] = W^T[hidden x K] * X^T[K x rows],
//
// so a CTA can own as few as 16 rows and the tile count follows the SM count (C2: 85 CTAs x 48 rows or 125 x 32).
//   A operand  = weights, K-major no-swizzle core matrices (the layout the weight image is packed in); hidden widths
//                above 128 use two M = 128 tiles (the second one reads past its 80 rows into finite weight bytes; those
//                accumulator lanes are never read back)
//   B operand  = activations, MN-major (row index contiguous) no-swizzle core matrices: an epilogue thread owns ONE
//                hidden unit (TMEM lane) and writes 8 consecutive rows as one 16-byte store.  X_lo is laid out directly
//                behind X_hi at the same 8-row-group stride, so ONE MMA with N = 2 x rows multiplies a weight block by
//                [X_hi ; X_lo] (the weight block -- 4 KB, the dominant shared-memory fetch of a small-N MMA -- is read once)
//   precision  = fp16 hi/lo split, the same three products as rollout_tc.cu, kept in two accumulator column ranges
//                [W_hi X_hi + W_lo X_hi | W_hi X_lo] that the epilogue adds in fp32 (2 MMAs per K16 block instead of 3)
//
// Measured on B200 (tools/tc_rate.py): a kind::f16 MMA with M = 128 costs max(N / 2, 32 + N / 4) cycles -- below N = 128
// the 4 KB A fetch from shared memory (128 B/clk) sets the pace, not the tensor pipe; and it needs a lean issue loop (one
// elected thread inside a single elect.sync region: 40 cycles per N = 32 MMA, against ~100 when each MMA is elected
// separately and ~300 with per-MMA index arithmetic).
//
// Warp roles (576 threads): warp 0 streams the member's weight image from L2 through a shared-memory ring (1-D bulk
// async copies, one stage = up to `kps` K16 blocks of one M tile, hi + lo); warp 1 = one elected thread issuing the MMAs
// from a compile-time schedule (reference architecture) or a per-step stage table; warps 2-17 are the epilogue: all 16
// take M tile 0 and then M tile 1 (quarter = TMEM lane quarter, four column slices).  The layer input is double-buffered,
// so tile 0's epilogue stores while tile 1's MMAs still read the previous buffer, and the next layer's first 8 K blocks
// (produced by tile 0) are issued straight behind tile 1's MMAs while tile 1's epilogue runs.
// The per-step serial part is kept short: the step's Gaussian draws and the next step's action prefetch run while the
// warps wait for layer 1; the final epilogue (bounded logvar, sample, obs_postproc -- cadm/dynamics/core/utils.py:141-168)
// is one thread per (row, state dim) and writes the next step's state features straight into the layer-0 operand, the
// warps without head outputs build its action / context features meanwhile; the reward reads the CURRENT state (quirk Q4).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "tcs_common.cuh"

namespace cadm {

constexpr int kSEpiThreads = 512;
constexpr int kSThreads = 64 + kSEpiThreads;
constexpr int kSMaxRows = 64;                  // rows per tile (multiple of 16)
constexpr int kSAccStride = 256;               // TMEM columns between the accumulators of the two M tiles (3 x rows <= 256)
constexpr int kSMaxStages = 16;
constexpr int kSMaxKps = 4;
constexpr int kSMaxEnt = 96;                   // stage-table entries per step

struct TcsSmem {
    size_t off_x, off_x0, off_w, off_s, off_hd, off_bias, off_vec, off_rowi, off_feat, off_zero, off_act, off_ctx, off_nz, off_tab, off_bar, total;
    int xbytes, x0bytes, slot_bytes;
};

__host__ __device__ inline TcsSmem tcs_smem_layout(int N, int D, int A, int C, int n_hidden, int Np, int NHp, int Kcap, int nkb0, int kps,
                                                   int stages) {
    TcsSmem L;
    size_t o = 0;
    L.xbytes = (N / 8) * (Kcap / 8) * 128;       // one operand half (hi or lo) of one layer-input buffer
    L.slot_bytes = kps * 8192;
    L.off_x = o; o += (size_t)4 * L.xbytes;      // two buffers x [hi | lo]: the hidden layers' outputs
    // the layer-0 input has a buffer of its own (its K0 = 16 nkb0 features only, row-group stride K0 / 8 * 128 bytes): the hidden
    // layers never overwrite it, so the features that do not change over the horizon (context, padding, rows beyond the tile) are
    // written once, and the next step's action features can be written while this step's layers are still running
    L.x0bytes = (N / 8) * (nkb0 * 2) * 128;
    L.off_x0 = o; o += (size_t)2 * L.x0bytes;
    o = (o + 127) / 128 * 128;
    L.off_w = o; o += (size_t)stages * L.slot_bytes;
    L.off_s = o; o += (size_t)round_up(N * (D + 3), 4) * 4;          // state rows: D values, pad, sin / cos of the angle
    L.off_hd = o; o += (size_t)NHp * (N + 4) * 4;                    // head outputs [unit][row], rows padded by 4 (bank spread)
    o = (o + 15) / 16 * 16;
    L.off_bias = o; o += (size_t)round_up(n_hidden * Np + NHp, 4) * 4;
    o = (o + 15) / 16 * 16;
    L.off_vec = o; o += (size_t)(8 * kMaxObs) * 4;                   // per state dimension: two float4 of epilogue constants
    L.off_rowi = o; o += (size_t)N * 6 * 4;
    L.off_feat = o; o += (size_t)96 * 16 + 96 * 8;
    L.off_zero = o; o += 16;
    L.off_act = o; o += (size_t)2 * N * A * 4;
    L.off_ctx = o; o += (size_t)N * (C > 0 ? C : 1) * 4;
    L.off_nz = o; o += (size_t)N * ((D + 3) / 4 * 4) * 4;           // this step's N(0,1) draws, [row][4 * blocks]
    o = (o + 15) / 16 * 16;
    L.off_tab = o; o += (size_t)kSMaxEnt * 32;                        // stage table (copied from the kernel parameters)
    L.off_bar = o; o += (size_t)(2 * kSMaxStages + 2 + 2 + 4) * 8;   // w_full, w_empty, xr[2], acc_full[2], tmem slot
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------------
// weight image: per member, per GEMM, per M tile (rows R = min(128, Npad - 128 mt)), per stage of `kps` K16 blocks:
//   [hi: (2 kbs) k-chunks x (R rows x 16 B)] [lo: same]        byte(row, chunk c, kk) = c 16 R + 16 row + 2 (kk % 8)
// ------------------------------------------------------------------------------------------------
__global__ void pack_tcs_kernel(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad,
                                int kps, long long member_stride, long long layer_off, int clear, float wscale) {
    const long long per_member = (long long)nkb * 16 * Npad;
    const long long total = (long long)E * per_member;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / per_member);
        long long r = i - (long long)e * per_member;
        const int n = (int)(r / (nkb * 16));
        const int k = (int)(r - (long long)n * (nkb * 16));
        const int ns = n - col0;
        const bool valid = ns >= 0 && ns < out && k < in;
        if (!valid && !clear) continue;
        const float w = valid ? src[((size_t)e * in + k) * out + ns] * wscale : 0.f;
        uint32_t hi, lo;
        tc::split2(w, 0.f, hi, lo);
        const int mt = n >> 7, row = n & 127;
        const int R = min(128, Npad - 128 * mt);
        const int kb = k >> 4, kk = k & 15;
        const int s0 = (kb / kps) * kps;
        const int kbs = min(kps, nkb - s0);
        const int c = 2 * (kb - s0) + (kk >> 3);
        const long long off = (mt ? (long long)nkb * 64 * 128 : 0) + (long long)s0 * 64 * R + (long long)c * 16 * R + 16 * row + 2 * (kk & 7);
        unsigned char* base = dst + e * member_stride + layer_off;
        *reinterpret_cast<unsigned short*>(base + off) = (unsigned short)(hi & 0xffffu);
        *reinterpret_cast<unsigned short*>(base + off + (long long)kbs * 32 * R) = (unsigned short)(lo & 0xffffu);
    }
}

cudaError_t launch_pack_tcs(unsigned char* dst, const float* src, int E, int in, int out, int col0, int nkb, int Npad, int kps,
                            long long member_stride, long long layer_off, int clear, cudaStream_t stream) {
    const long long total = (long long)E * nkb * 16 * Npad;
    pack_tcs_kernel<<<(int)min((total + 255) / 256, (long long)2368), 256, 0, stream>>>(dst, src, E, in, out, col0, nkb, Npad, kps,
                                                                                        member_stride, layer_off, clear, tc::kWScale);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
namespace tcs {

// All stages of ONE GEMM with every loop bound a compile-time constant (the hot configurations): NKB K16 blocks, output
// width NPAD (one or two M tiles), KPS blocks per stage, KB_SPLIT = first K block produced by the M-tile-1 epilogue warps
// of the previous layer (0: wait for both groups up front).  Fully unrolled: per stage one mbarrier wait, 2 KPS
// back-to-back UTCHMMAs whose descriptors differ by immediates, one or two commits -- about 30 instructions, so the lone
// issuing thread keeps ahead of the tensor pipe (the table-driven loop needs ~150 per stage and does not).
template <int TERMS, int NKB, int NPAD, int KPS, int KB_SPLIT, int MM = 128>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_base, uint32_t w16, uint32_t slot16, int nstage, RingPos& rp, uint32_t xg16,
                                           uint32_t rows, uint32_t xsbo, uint64_t* w_full, uint64_t* w_empty, uint64_t* xr,
                                           uint64_t* acc_full, uint32_t ev, long long* dbg = nullptr) {
    const uint32_t hi32 = (1u << 14);
    const uint64_t a_top = (uint64_t)(hi32 | (128u >> 4)) << 32;
    const uint64_t b_top = (uint64_t)(hi32 | (xsbo >> 4)) << 32;
    const uint32_t b_lbo = (128u >> 4) << 16;
    const uint32_t idesc1 = idesc(rows, MM), idesc2 = idesc(2u * rows, MM);
    constexpr int NMT = (NPAD + 127) / 128;
    bool waited1 = false;
    ptx::mbar_wait(&xr[0], ev);
    if (KB_SPLIT == 0) { ptx::mbar_wait(&xr[1], ev); waited1 = true; }
    tc::fence_after_sync();
    if (dbg) dbg[0] = clock64();
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
        constexpr int R0 = NPAD < 128 ? NPAD : 128;
        const uint32_t R = mt == 0 ? (uint32_t)R0 : (uint32_t)(NPAD - 128);
        const uint32_t d_tmem = tmem_base + (uint32_t)mt * kSAccStride;
#pragma unroll
        for (int s0 = 0; s0 < NKB; s0 += KPS) {
            constexpr int dummy = 0;
            (void)dummy;
            const int kbs = (NKB - s0) < KPS ? (NKB - s0) : KPS;
            const bool last = mt == NMT - 1 && s0 + kbs >= NKB;
            if (!waited1 && (s0 + kbs > KB_SPLIT || last)) {
                ptx::mbar_wait(&xr[1], ev);
                tc::fence_after_sync();
                waited1 = true;
                if (dbg) dbg[2] = clock64();
            }
            ptx::mbar_wait(&w_full[rp.slot], rp.phase);
            const uint32_t slot = w16 + rp.slot * slot16;
            const uint64_t a_hi = a_top | (slot | (R << 16));
            const uint64_t a_lo = a_top | ((slot + (uint32_t)kbs * 2u * R) | (R << 16));
            const uint64_t b = b_top | ((xg16 + (uint32_t)s0 * 16u) | b_lbo);
            const uint32_t accum = s0 > 0 ? 1u : 0u;
            if (kbs == 4) issue_blocks<TERMS, 4>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
            else if (kbs == 3) issue_blocks<TERMS, 3>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
            else if (kbs == 2) issue_blocks<TERMS, 2>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
            else issue_blocks<TERMS, 1>(d_tmem, a_hi, a_lo, b, 2u * R, rows, idesc1, idesc2, accum);
            tc::mma_commit(&w_empty[rp.slot]);
            if (s0 + kbs >= NKB) tc::mma_commit(&acc_full[mt]);
            if (dbg && mt == 0 && s0 + kbs >= NKB) dbg[3] = clock64();
            if (++rp.slot == (uint32_t)nstage) { rp.slot = 0; rp.phase ^= 1u; }
        }
    }
}

}  // namespace tcs

struct TcsParams {
    RolloutParams R;
    const unsigned char* wimg;
    long long wimg_member_stride;
    int Np, NHp, nkb0, nkbH, Kcap;
    int terms, stages, kps;
    int N;                       // rows per tile (multiple of 16, <= kSMaxRows)
    int nmt;                     // M tiles of a hidden GEMM: 1 (Np <= 128) or 2
    int tiles_per_member, total_tiles;
    int debug;                   // diagnostic bits: 1 no weight copies, 2 no hidden-layer epilogue work (results are garbage)
    int skew;                    // start delay step in cycles (CTA i waits (i % 8) * skew)
    int nent;                    // stages per horizon step
    uint4 tab[2 * kSMaxEnt];     // the stage table (kernel-parameter space: uniform loads straight into uniform registers)
    long long* dbg;              // nullable: clock64 trace of CTA 0, [step][64]
};

// NKB0 > 0 selects the compile-time schedule of the MMA thread for the reference architecture (4 hidden layers of width
// 193..208, KPS = 4): NKB0 = K16 blocks of layer 0, NHP = padded width of the heads.  NKB0 == 0: table-driven schedule.
// DIAG = true compiles the diagnostic switches (TcsParams::debug) and the clock64 trace in; the production instantiations carry
// neither: every instruction between "accumulator ready" and the mbarrier arrival of an epilogue warp is on the dependent chain of
// the layers, and the runtime checks alone cost 17 us of a 290 us launch when they were always compiled in.
template <int TERMS, int NKB0, int NHP, bool DIAG>
__global__ void __launch_bounds__(kSThreads, 1) rollout_tcs_kernel(const __grid_constant__ TcsParams T) {
    const RolloutParams& P = T.R;
    const int dbgbits = DIAG ? T.debug : 0;
    long long* const dbgbase = DIAG ? T.dbg : nullptr;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int N = T.N;
    const TcsSmem L = tcs_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, T.kps, T.stages);
    unsigned char* xbuf = smem + L.off_x;          // buffer b: X_hi at b * 2 xbytes, X_lo directly behind it
    unsigned char* x0buf = smem + L.off_x0;        // layer-0 input: X_hi, X_lo directly behind it (x0bytes each)
    unsigned char* wring = smem + L.off_w;
    float* S = reinterpret_cast<float*>(smem + L.off_s);
    float* Hd = reinterpret_cast<float*>(smem + L.off_hd);             // [NHp][N] head outputs
    float* bias = reinterpret_cast<float*>(smem + L.off_bias);
    float* vec = reinterpret_cast<float*>(smem + L.off_vec);
    int* rowi = reinterpret_cast<int*>(smem + L.off_rowi);
    int4* feat_i = reinterpret_cast<int4*>(smem + L.off_feat);
    float2* feat_f = reinterpret_cast<float2*>(smem + L.off_feat + 96 * 16);
    float* act_s = reinterpret_cast<float*>(smem + L.off_act);
    float* ctx_s = reinterpret_cast<float*>(smem + L.off_ctx);
    float* nz_s = reinterpret_cast<float*>(smem + L.off_nz);
    uint4* tab = reinterpret_cast<uint4*>(smem + L.off_tab);
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* w_empty = w_full + kSMaxStages;
    uint64_t* xr = w_empty + kSMaxStages;          // [2]: layer input produced by the M-tile-0 / M-tile-1 warps
    uint64_t* acc_full = xr + 2;                   // [M tile]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    volatile uint32_t* done = tmem_slot + 1;       // [2]: the producer / MMA thread has finished (its warp's other lanes sleep on it)

    float4* fc = reinterpret_cast<float4*>(vec);      // [2 d]: {dscale, dmean, max_lv log2e, std^2 e^max}, [2 d + 1]: {std^2 e^min, finv, fofs, feature byte offset}
    int* r_mi = rowi;
    int* r_src = rowi + N;
    int* r_pi = rowi + 2 * N;
    int* r_ctx = rowi + 3 * N;
    int* r_rid = rowi + 4 * N;
    int* r_eps = rowi + 5 * N;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nstage = T.stages;
    const int xsbo = (T.Kcap / 8) * 128;
    const int xsbo0 = T.nkb0 * 2 * 128;            // row-group stride of the layer-0 input buffer
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        ptx::mbar_init(&xr[0], 16);
        ptx::mbar_init(&xr[1], 16);
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
    for (int i = tid; i < 2 * T.nent; i += kSThreads) tab[i] = T.tab[i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nent = T.nent;
    griddep_launch();                              // the refit kernel may become resident; it blocks until this grid is complete
    if (T.skew > 0) {                              // de-phase the CTAs' weight-stream bursts (see launch_rollout_tcs)
        const long long until = clock64() + (long long)(blockIdx.x % 8) * T.skew;
        while (clock64() < until) __nanosleep(200);
    }

    // ======================= warp 0: weight producer =============================================
    if (warp == 0) {
        if (lane == 0) {
            tcs::Ring rp{0, 0, nstage};
            for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
                const int e = tile / T.tiles_per_member;
                const unsigned char* wsrc = T.wimg + (size_t)e * T.wimg_member_stride;
                for (int t = 0; t < P.h; ++t) {
                    size_t off = 0;
                    for (int i = 0; i < nent; ++i) {
                        const uint32_t bytes = tab[2 * i].y;
                        ptx::mbar_wait(&w_empty[rp.stage], rp.phase ^ 1u);
                        if (dbgbits & 1) {                          // diagnostic: no weight traffic at all
                            ptx::mbar_arrive(&w_full[rp.stage]);
                        } else {
                            ptx::mbar_arrive_expect_tx(&w_full[rp.stage], bytes);
                            ptx::bulk_g2s(wring + (size_t)rp.stage * L.slot_bytes, wsrc + off, bytes, &w_full[rp.stage]);
                        }
                        off += bytes;
                        rp.advance();
                    }
                }
            }
            done[0] = 1u;
        } else {
            // A lane parked at a warp barrier keeps competing for issue slots with the working lane of its own warp
            // (measured: the single-thread loops run ~2x slower); sleeping lanes do not.
            while (!done[0]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warp 1: MMA issuer ==================================================
    // ONE elected thread walks the stage table: inside a single elect.sync region ptxas keeps the descriptors in uniform
    // registers and the UTCHMMAs issue back to back.  A lone thread has no latency hiding, and its ~400 cycles of
    // per-stage bookkeeping (table decode, descriptor arithmetic, mbarrier try_wait, ring advance) cost as much as the
    // stage's MMAs take to execute -- so the NEXT stage is prepared in the middle of the current stage's MMA sequence,
    // while the first half of its MMAs keeps the tensor pipe fed.
    else if (warp == 1) {
        if (NKB0 > 0) {
            if (ptx::elect_one()) {
                const uint32_t x16 = ptx::smem_u32(xbuf) >> 4, w16 = ptx::smem_u32(wring) >> 4, slot16 = (uint32_t)L.slot_bytes >> 4;
                const uint32_t xb1 = x16 + ((uint32_t)(2 * L.xbytes) >> 4);     // layer-input buffer 1
                const uint32_t x0_16 = ptx::smem_u32(x0buf) >> 4;
                tcs::RingPos rp{0u, 0u};
                uint32_t ev = 0;
                for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
                    for (int t = 0; t < P.h; ++t) {
                        long long* dbg = (dbgbase && blockIdx.x == 0 && tile == (int)blockIdx.x && t < 64) ? dbgbase + t * 64 + 32 : nullptr;
                        if (dbg) dbg[0] = clock64();
                        tcs::issue_gemm<TERMS, NKB0, 208, 4, 0>(tmem_base, w16, slot16, nstage, rp, x0_16, (uint32_t)N, (uint32_t)xsbo0, w_full,
                                                                w_empty, xr, acc_full, ev, dbg ? dbg + 0 : nullptr);
                        ev ^= 1u;
                        if (dbg) { dbg[1] = clock64(); dbg[4] = dbg[1]; }
                        tcs::issue_gemm<TERMS, 13, 208, 4, 8>(tmem_base, w16, slot16, nstage, rp, xb1, (uint32_t)N, (uint32_t)xsbo, w_full,
                                                              w_empty, xr, acc_full, ev, dbg ? dbg + 4 : nullptr);
                        ev ^= 1u;
                        if (dbg) { dbg[5] = clock64(); dbg[8] = dbg[5]; }
                        tcs::issue_gemm<TERMS, 13, 208, 4, 8>(tmem_base, w16, slot16, nstage, rp, x16, (uint32_t)N, (uint32_t)xsbo, w_full,
                                                              w_empty, xr, acc_full, ev, dbg ? dbg + 8 : nullptr);
                        ev ^= 1u;
                        if (dbg) { dbg[9] = clock64(); dbg[12] = dbg[9]; }
                        tcs::issue_gemm<TERMS, 13, 208, 4, 8>(tmem_base, w16, slot16, nstage, rp, xb1, (uint32_t)N, (uint32_t)xsbo, w_full,
                                                              w_empty, xr, acc_full, ev, dbg ? dbg + 12 : nullptr);
                        ev ^= 1u;
                        if (dbg) { dbg[13] = clock64(); dbg[16] = dbg[13]; }
                        tcs::issue_gemm<TERMS, 13, NHP, 4, 8, (NHP <= 64 ? 64 : 128)>(tmem_base, w16, slot16, nstage, rp, x16, (uint32_t)N, (uint32_t)xsbo, w_full,
                                                              w_empty, xr, acc_full, ev, dbg ? dbg + 16 : nullptr);
                        ev ^= 1u;
                        if (dbg) dbg[17] = clock64();
                    }
                }
                done[1] = 1u;
            } else {
                while (!done[1]) __nanosleep(2000);
            }
        } else if (ptx::elect_one()) {
            const uint32_t x16 = ptx::smem_u32(xbuf) >> 4, w16 = ptx::smem_u32(wring) >> 4, slot16 = (uint32_t)L.slot_bytes >> 4;
            const uint32_t hi32 = (1u << 14);                                   // descriptor version 1 (bit 46)
            const uint64_t a_top = (uint64_t)(hi32 | (128u >> 4)) << 32;        // A: SBO = 128 B between 8-row groups
            const uint64_t b_top = (uint64_t)(hi32 | ((uint32_t)xsbo >> 4)) << 32;   // B: SBO = stride between 8-row (N) groups
            const uint64_t b_top0 = (uint64_t)(hi32 | ((uint32_t)xsbo0 >> 4)) << 32;  // the same for the layer-0 input buffer
            const uint32_t x0_16 = ptx::smem_u32(x0buf) >> 4;
            const uint32_t b_lbo = (128u >> 4) << 16;                           // B: LBO = 128 B between the two k-groups
            const uint32_t rows = (uint32_t)N;
            const uint32_t idesc1_h = tcs::idesc(rows), idesc2_h = tcs::idesc(2u * rows);            // hidden layers: M = 128
            const uint32_t hM = T.NHp <= 64 ? 64u : 128u;                                                // heads
            const uint32_t idesc1_o = tcs::idesc(rows, hM), idesc2_o = tcs::idesc(2u * rows, hM);
            int my_tiles = 0;
            for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) ++my_tiles;
            long long remaining = (long long)my_tiles * P.h * nent;
            const bool trace = dbgbase != nullptr && blockIdx.x == 0;
            // state of the stage being issued
            uint32_t ev = 0;                       // parity of the layer-input event the next GEMM waits for
            int stage = 0, ti = 0, tstep = 0;      // ring slot, table index, horizon step (trace only)
            uint32_t phase = 0;
            uint4 e0 = tab[0], e1 = tab[1];
            // the first stage: blocking waits
            if (remaining > 0) {
                ptx::mbar_wait(&xr[0], ev);
                ptx::mbar_wait(&xr[1], ev);
                ptx::mbar_wait(&w_full[0], 0);
                tc::fence_after_sync();
                if (trace) dbgbase[32] = clock64();
            }
            while (remaining > 0) {
                --remaining;
                const uint32_t R = e1.x, a_step = 2u * R;
                const int kbs = (int)e1.y;
                const uint32_t d_tmem = tmem_base + e0.w;
                const uint32_t slot = w16 + (uint32_t)stage * slot16;
                const uint64_t a_hi = a_top | (slot | (R << 16));
                const uint64_t a_lo = a_top | ((slot + (uint32_t)kbs * a_step) | (R << 16));
                const uint64_t b = e1.w == 0u ? (b_top0 | ((x0_16 + e0.z) | b_lbo)) : (b_top | ((x16 + e0.z) | b_lbo));
                const bool skip = (dbgbits & 4) != 0;
                const bool is_head = e1.w == (uint32_t)P.n_hidden;
                const uint32_t idesc1 = is_head ? idesc1_o : idesc1_h, idesc2 = is_head ? idesc2_o : idesc2_h;
                // ---- first half of the stage's MMAs
                if (!skip) {
                    tcs::issue_blocks<TERMS, 1>(d_tmem, a_hi, a_lo, b, a_step, rows, idesc1, idesc2, e1.z);
                    if (kbs > 1) tcs::issue_blocks<TERMS, 1>(d_tmem, a_hi + a_step, a_lo + a_step, b + 16u, a_step, rows, idesc1, idesc2, 1u);
                }
                // ---- prepare the next stage under them
                int nti = ti + 1, nstep = tstep;
                if (nti == nent) { nti = 0; ++nstep; }
                int nstg = stage + 1;
                uint32_t nphase = phase;
                if (nstg == nstage) { nstg = 0; nphase ^= 1u; }
                const uint4 f0 = tab[2 * nti], f1 = tab[2 * nti + 1];
                const uint32_t nev = (e0.x & 8u) ? (ev ^ 1u) : ev;
                bool ready = false;
                if (remaining > 0 && !(f0.x & 3u)) ready = ptx::mbar_try_wait(&w_full[nstg], nphase);
                // ---- second half
                if (!skip) {
                    if (kbs > 2) tcs::issue_blocks<TERMS, 1>(d_tmem, a_hi + 2u * a_step, a_lo + 2u * a_step, b + 32u, a_step, rows, idesc1, idesc2, 1u);
                    if (kbs > 3) tcs::issue_blocks<TERMS, 1>(d_tmem, a_hi + 3u * a_step, a_lo + 3u * a_step, b + 48u, a_step, rows, idesc1, idesc2, 1u);
                }
                tc::mma_commit(&w_empty[stage]);
                if (e0.x & 4u) tc::mma_commit(&acc_full[(e0.x >> 4) & 1u]);
                if (trace && (e0.x & 8u) && tstep < 64 && e1.w < 5) dbgbase[tstep * 64 + 32 + 4 * e1.w + 1] = clock64();
                // ---- whatever the next stage still has to wait for
                if (remaining > 0 && !ready) {
                    if (f0.x & 1u) ptx::mbar_wait(&xr[0], nev);
                    if (f0.x & 2u) ptx::mbar_wait(&xr[1], nev);
                    ptx::mbar_wait(&w_full[nstg], nphase);
                    // the accumulator / layer-input hand-over from the epilogue warps needs the tcgen05 fence; a weight
                    // stage alone does not (async-proxy writes completed on the mbarrier)
                    if (f0.x & 3u) tc::fence_after_sync();
                    if (trace && (f0.x & 1u) && nstep < 64 && f1.w < 5) dbgbase[nstep * 64 + 32 + 4 * f1.w] = clock64();
                }
                e0 = f0; e1 = f1; ev = nev; stage = nstg; phase = nphase; ti = nti; tstep = nstep;
            }
            done[1] = 1u;
        } else {
            while (!done[1]) __nanosleep(2000);
        }
        __syncwarp();
    }
    // ======================= warps 2..17: prologue / epilogue ======================================
    else {
        const int et = tid - 64;                       // 0..511
        const int ew = warp - 2;                       // 0..15
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int cslice = ew >> 2;                    // the warp takes the 8-column (8 batch rows) chunks cslice, cslice + 4, ...
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int nchunks = N >> 3;
        // the warps whose TMEM lane quarter holds no head output are idle in the head epilogue: they prefetch the next step's
        // actions and build its action / context features there, off the critical path
        const bool head64 = T.NHp <= 64;                 // heads issued as M = 64 MMAs: 16 outputs per TMEM lane quarter
        const int nq_head = head64 ? (T.NHp + 15) >> 4 : (T.NHp + 31) >> 5;
        const bool helper = quarter >= nq_head;
        const int n_help = 128 * (4 - nq_head);                                  // 0 when the heads fill all four quarters
        const int ht = ((quarter - nq_head) + (4 - nq_head) * cslice) * 32 + lane;    // index among the helper threads
        // per-thread constants of the hidden-layer epilogue
        const int qlane = quarter * 32 + lane;                                     // this thread's hidden unit within an M tile
        const bool act0 = quarter * 32 < T.Np, act1 = T.nmt == 2 && 128 + quarter * 32 < T.Np;       // the warp has units in tile 0 / 1
        const bool ok0 = qlane < T.Np, ok1 = 128 + qlane < T.Np;
        const int so0 = cslice * xsbo + (qlane >> 3) * 128 + (qlane & 7) * 16;     // operand-store offset: chunk cslice, unit qlane
        const int so1 = so0 + 16 * 128;                                            // unit 128 + qlane
        const uint32_t tc0 = tmem_lane + (uint32_t)cslice * 8u;
        const uint32_t xbuf_s = ptx::smem_u32(xbuf);
        const uint32_t hd_s = ptx::smem_u32(Hd);
        const uint32_t x0_s = ptx::smem_u32(x0buf);
        uint32_t g_count = 0;                          // GEMMs so far   = completions of acc_full[0]
        uint32_t c1cnt = 0;                            // 2-tile GEMMs   = completions of acc_full[1]
        const int D = P.D, A = P.A;
        const size_t eps_step_stride = (size_t)P.E * P.q * P.m * P.n_global * D;
        const int nj = (D + 3) >> 2;                   // Philox blocks (4 state dims each) per row

        // layer-0 feature table: feature k of row r reads the float at smem byte offset x + r * y (+ action double-buffer
        // offset when z & 1), then (v - mean) * (8 / (std + 1e-10)); sin / cos of the HalfCheetah angle by select
        const int K0 = T.nkb0 * 16;
        for (int k = et; k < K0; k += kSEpiThreads) {
            int base = (int)L.off_zero, stride = 0, flags = 0;
            float mean = 0.f, inv = 0.f;
            if (k < P.P) {
                int idx = k;
                if (P.env_id == CADM_ENV_HALFCHEETAH) {          // [o1, sin o2, cos o2, o3:]
                    if (k == 0) idx = 1;
                    else if (k == 1) idx = D + 1;                 // sin o2, kept beside the state (computed once per step)
                    else if (k == 2) idx = D + 2;                 // cos o2
                } else if (P.env_id == CADM_ENV_ANT) {
                    idx = k + 1;                                  // o[1:]
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
        ptx::bar_sync(1, kSEpiThreads);
        // per state dimension d: the constants of the final epilogue, two 16-byte loads per dimension.  State dimension d is the
        // processed-observation feature kf (obs_preproc: HalfCheetah [o1, sin o2, cos o2, o3:], Ant o[1:], identity otherwise):
        // its normalisation (x - mean) * inv is folded into one FMA, and the byte offset of feature kf inside a row group is stored
        // as an integer (-1: the dimension is no feature; HalfCheetah's angle, d = 2, is handled apart)
        for (int d = et; d < D; d += kSEpiThreads) {
            int kf = d;
            if (P.env_id == CADM_ENV_HALFCHEETAH) kf = d <= 1 ? d - 1 : (d == 2 ? -1 : d);
            else if (P.env_id == CADM_ENV_ANT) kf = d - 1;
            const float ds = P.delta_std[d];
            float finv = 0.f, fofs = 0.f;
            if (kf >= 0) { const float2 ff = feat_f[kf]; finv = ff.y; fofs = -ff.x * ff.y; }
            fc[2 * d] = make_float4(ds + 1e-10f, P.delta_mean[d], P.max_lv[d] * tc::kLog2e, ds * ds * expf(P.max_lv[d]));
            fc[2 * d + 1] = make_float4(ds * ds * expf(P.min_lv[d]), finv, fofs, __int_as_float(kf >= 0 ? (kf >> 3) * 128 + (kf & 7) * 16 : -1));
        }

        for (int tile = blockIdx.x; tile < T.total_tiles; tile += gridDim.x) {
            const int e = tile / T.tiles_per_member;
            const int tile_row0 = P.row_lo + (tile - e * T.tiles_per_member) * P.rows_per_cta;
            const int nrows = min(P.rows_per_cta, P.row_hi - tile_row0);
            ptx::bar_sync(1, kSEpiThreads);             // previous tile fully retired before its smem is reused
            for (int i = et; i < P.n_hidden * T.Np + T.NHp; i += kSEpiThreads)      // hidden-layer biases pre-scaled by kXScale
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
            ptx::bar_sync(1, kSEpiThreads);
            for (int i = et; i < N * D; i += kSEpiThreads) {
                const int r = i / D, d = i - r * D;
                float v = 0.f;
                if (r < nrows) v = (P.row_mode == kRowsPlanner) ? P.obs0[r_mi[r] * D + d] : P.obs0[(size_t)r_src[r] * D + d];
                S[r * (D + 3) + d] = v;
            }
            for (int i = et; i < N * P.C; i += kSEpiThreads) {
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
            for (int i = et; i < 2 * N * A; i += kSEpiThreads) act_s[i] = 0.f;   // rows >= nrows stay finite
            ptx::bar_sync(1, kSEpiThreads);
            if (P.env_id == CADM_ENV_HALFCHEETAH && et < N) sincosf(S[et * (D + 3) + 2], &S[et * (D + 3) + D + 1], &S[et * (D + 3) + D + 2]);
            // everything above reads inputs that are constant over the decision; the candidates' actions come from the kernel
            // launched just before this one (sample_actions), and the particle returns written at the end are still being read
            // by the refit before it: wait for the predecessor grid here (programmatic dependent launch, common.cuh)
            griddep_wait();
            prefetch_actions(0, et, kSEpiThreads);
            tcs::cp_async_wait_all();
            ptx::bar_sync(1, kSEpiThreads);

            float ret = 0.f;                               // return of row `et` (threads et < nrows)
            const int invN = ((1 << 20) + N - 1) / N;       // i / N for i < 2^20 / 64 as a multiply and a shift
            const int ndp = (D + 1) >> 1;                   // pairs of state dimensions = items per row of the final epilogue
            const int act_slot = P.n_hidden > 1 ? 1 : 0;    // the layer after which the next step's action features are written
            const bool onehot_mode = P.discrete && P.row_mode == kRowsPlanner;
            const bool is_hc = P.env_id == CADM_ENV_HALFCHEETAH;
            const int kf_shift = P.env_id == CADM_ENV_ANT ? -1 : 0;       // obs_preproc: Ant drops o[0]
            const bool want_out = P.states != nullptr || P.next_obs != nullptr || P.mu_out != nullptr || P.lv_out != nullptr;
            // layer-0 input features k in [k_lo, K0) of step t for every row, from the state / the step's actions / the
            // context; one work item = (feature, pair of rows), 4-byte stores into the MN-major layout
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
                    const int o = (rp >> 2) * xsbo0 + (k >> 3) * 128 + (k & 7) * 16 + (rp & 3) * 4;
                    *reinterpret_cast<uint32_t*>(x0buf + o) = hq;
                    *reinterpret_cast<uint32_t*>(x0buf + L.x0bytes + o) = lq;
                }
            };
            // hand the layer-0 input to the MMA thread (generic-proxy writes -> async-proxy reads)
            auto publish_input = [&]() {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) { ptx::mbar_arrive(&xr[0]); ptx::mbar_arrive(&xr[1]); }
            };
            // reward terms of step t that read the CURRENT observation and the action (quirk Q4); threads et < nrows
            auto add_reward = [&](int t) {
                if (et < nrows && !env_reward_reads_next(P.env_id))
                    ret += env_reward_current(P.env_id, S + et * (D + 3), act_s + (t & 1) * N * A + et * A, A, P.max_torque);
            };
            // step 0: everything from the start state
            build_features(0, 0, K0, et, kSEpiThreads);       // every feature once; per step only the state and the actions change
            publish_input();
            add_reward(0);
#pragma unroll 1
            for (int t = 0; t < P.h; ++t) {
                long long* dbg = (dbgbase && blockIdx.x == 0 && ew == 2 && lane == 0 && tile == (int)blockIdx.x && t < 64) ? dbgbase + t * 64 : nullptr;
                if (dbg) dbg[0] = clock64();
                if (dbg) dbg[1] = clock64();
                // per-warp detail of one step (rows 32.. of the trace buffer): for each (layer, M tile) six time stamps
                long long* wdbg = (dbgbase && blockIdx.x == 0 && lane == 0 && tile == (int)blockIdx.x && t == T.R.h / 2) ? dbgbase + (32 + ew) * 64 : nullptr;
                // ---------- hidden layers: accumulator -> bias + swish -> next layer's B operand ----------
#pragma unroll 1
                for (int l = 0; l < P.n_hidden; ++l) {
                    // GEMM l read buffer l & 1; its output (the input of GEMM l + 1) goes to the other buffer
                    const uint32_t xo_s = xbuf_s + (uint32_t)(((l + 1) & 1) * 2 * L.xbytes);
                    // All 16 warps take each M tile in turn (quarter = TMEM lane quarter, cslice = column chunks): tile 0's
                    // epilogue runs under tile 1's MMAs, tile 1's under the next layer's first K blocks.
                    // u = -log2(e) (W . x + b): the accumulator holds kWScale * (input scale) * (W . x); the input scale is kXScale for
                    // layer 0 and kActScale = -8 log2(e) afterwards (tc::swish_pair_u)
                    const float su = l == 0 ? -tc::kLog2e / (tc::kWScale * tc::kXScale) : 1.0f / (tc::kWScale * 8.0f);
                    const float2 su2 = make_float2(su, su);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (mt < T.nmt) {
                            // everything that does not need the accumulator comes BEFORE the wait: the bias (a shared-memory load),
                            // the TMEM address and the store offset of this thread's hidden unit
                            const bool active = mt ? act1 : act0;
                            const bool ok = mt ? ok1 : ok0;
                            const float bu = ok ? bias[l * T.Np + mt * 128 + qlane] : 0.f;
                            const float2 bu2 = make_float2(bu, bu);
                            uint32_t tcol = tc0 + (uint32_t)mt * kSAccStride;
                            uint32_t xo = xo_s + (uint32_t)(mt ? so1 : so0);              // 32-bit shared address: no generic-pointer conversion
                            if (wdbg && l < 4) wdbg[(l * 2 + mt) * 6 + 0] = clock64();
                            ptx::mbar_wait(&acc_full[mt], (mt ? c1cnt : g_count) & 1u);
                            tc::fence_after_sync();
                            if (wdbg && l < 4) wdbg[(l * 2 + mt) * 6 + 1] = clock64();
                            if (dbg && l < 4 && mt == 0) dbg[2 + 2 * l] = clock64();
                            if (active && !(dbgbits & 2)) {
                                for (int c = cslice; c < nchunks; c += 4, tcol += 32u, xo += 4u * (uint32_t)xsbo) {
                                    uint32_t v[8];
                                    float2 a[4];
                                    tc::tmem_ld8(tcol, v);
                                    if (TERMS == 3) {
                                        uint32_t v1[8];
                                        tc::tmem_ld8(tcol + N, v1);
                                        tc::tmem_wait_ld();
                                        if (wdbg && l < 4 && c == cslice) wdbg[(l * 2 + mt) * 6 + 2] = clock64();
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
                                    if (ok) {
                                        uint32_t hq[4], lq[4];
#pragma unroll
                                        for (int j = 0; j < 4; ++j) tc::split2(y[j].x, y[j].y, hq[j], lq[j]);
                                        ptx::sts128(xo, hq[0], hq[1], hq[2], hq[3]);
                                        ptx::sts128(xo + (uint32_t)L.xbytes, lq[0], lq[1], lq[2], lq[3]);
                                    }
                                }
                            }
                            if (wdbg && l < 4) wdbg[(l * 2 + mt) * 6 + 3] = clock64();
                            tc::fence_before_sync();
                            ptx::fence_proxy_async();
                            if (wdbg && l < 4) wdbg[(l * 2 + mt) * 6 + 4] = clock64();
                        }
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&xr[mt]);
                        if (wdbg && l < 4) wdbg[(l * 2 + mt) * 6 + 5] = clock64();
                        if (dbg && l < 4 && mt == 0) dbg[3 + 2 * l] = clock64();
                    }
                    ++g_count;
                    if (T.nmt == 2) ++c1cnt;
                    // ---------- off the critical path: the warps are idle from here until the next layer's first accumulator is
                    // ready: next step's actions, this step's noise
                    if (l == 0) {
                        if (t + 1 < P.h) {
                            if (n_help == 0) prefetch_actions(t + 1, et, kSEpiThreads);
                            else if (helper) prefetch_actions(t + 1, ht, n_help);
                        }
                        if (!P.deterministic) {
                            for (int i = et; i < N * nj; i += kSEpiThreads) {
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
                        // the next step's action features: layer 0 of this step has consumed its input buffer (every warp has waited
                        // for its accumulators) and the prefetch was issued a layer ago by the same threads
                        if (n_help == 0) {
                            tcs::cp_async_wait_all();
                            ptx::bar_sync(1, kSEpiThreads);
                            build_features(t + 1, P.P, P.P + A, et, kSEpiThreads);
                        } else if (helper) {
                            tcs::cp_async_wait_all();
                            ptx::bar_sync(2, n_help);
                            build_features(t + 1, P.P, P.P + A, ht, n_help);
                        }
                    }
                }

                // ---------- heads -> Hd[unit][row] ----------------------------------------------------
                // this thread's item of the final epilogue: state dimensions (2 dp, 2 dp + 1) of row r.  Its constants and its
                // current state are loaded BEFORE the waits below, so that only the head outputs and the noise are loaded on the
                // dependent chain
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
                    const uint32_t par0 = g_count & 1u;
                    ++g_count;
                    if (quarter < nq_head) {
                        const bool lane_ok = !head64 || lane < 16;
                        const int jh = head64 ? quarter * 16 + (lane & 15) : qlane;
                        const float bj = jh < T.NHp ? bias[P.n_hidden * T.Np + jh] : 0.f;
                        const float sc = 1.0f / (tc::kWScale * tc::kActScale);
                        const float2 sc2 = make_float2(sc, sc), bj2 = make_float2(bj, bj);
                        uint32_t hd_a = hd_s + (uint32_t)((jh * (N + 4) + cslice * 8) * 4);
                        uint32_t tcol = tc0;
                        if (wdbg) wdbg[48] = clock64();
                        ptx::mbar_wait(&acc_full[0], par0);
                        tc::fence_after_sync();
                        if (wdbg) wdbg[49] = clock64();
                        if (dbg) dbg[10] = clock64();
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
                            if (jh < T.NHp && lane_ok) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) a[j] = tc::ffma2(a[j], sc2, bj2);
                                ptx::sts128(hd_a, __float_as_uint(a[0].x), __float_as_uint(a[0].y), __float_as_uint(a[1].x), __float_as_uint(a[1].y));
                                ptx::sts128(hd_a + 16u, __float_as_uint(a[2].x), __float_as_uint(a[2].y), __float_as_uint(a[3].x), __float_as_uint(a[3].y));
                            }
                        }
                        tc::fence_before_sync();
                    }
                }
                if (wdbg) wdbg[50] = clock64();
                ptx::bar_sync(1, kSEpiThreads);
                if (wdbg) wdbg[51] = clock64();
                if (dbg) dbg[11] = clock64();

                // ---------- final epilogue: sample, next state, next step's state features -------------------------------
                // (core/utils.py:84-90, 162): delta = mu std + mean + eps sqrt(std^2 e^min + std^2 e^max / (1 + e^(max - lv))) -- the two
                // soft bounds of the log-variance in closed form: with lv1 = max - softplus(max - lv), lv2 = min + softplus(lv1 - min),
                // e^lv2 = e^min + e^lv1 = e^min + e^max / (1 + e^(max - lv)): ex2, rcp and sqrt on a three-deep chain
                // One item = the state dimensions (d0, d0 + 1) of row r.  Every load comes first (the compiler cannot move a load
                // above a store to shared memory it cannot tell apart), then the arithmetic of both dimensions side by side, then
                // the stores.
                auto final_pair = [&](int r, int d0, bool two, const float4 (&ca)[2], const float4 (&cb)[2], const float (&s_old)[2]) {
                    float mu[2] = {0.f, 0.f}, lv[2] = {0.f, 0.f}, nz[2] = {0.f, 0.f}, sn[2];
                    mu[0] = Hd[d0 * (N + 4) + r];
                    lv[0] = Hd[(D + d0) * (N + 4) + r];
                    if (two) { mu[1] = Hd[(d0 + 1) * (N + 4) + r]; lv[1] = Hd[(D + d0 + 1) * (N + 4) + r]; }
                    if (!P.deterministic) {
                        const float2 z2 = *reinterpret_cast<const float2*>(nz_s + r * (nj * 4) + d0);      // d0 is even: 8-byte aligned
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
                    const bool angle = is_hc && d0 == 2;           // HalfCheetah: features 1 and 2 are the sine and cosine of o2
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
                        const uint32_t xr0 = x0_s + (uint32_t)((r >> 3) * xsbo0 + (r & 7) * 2);      // 32-bit shared address
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int ko = __float_as_int(cb[j].w);
                            if (ko >= 0 && (j == 0 || two)) {
                                ptx::sts16(xr0 + ko, hq[j]);
                                ptx::sts16(xr0 + L.x0bytes + ko, lq[j]);
                            }
                        }
                        if (angle) {
                            ptx::sts16(xr0 + 16, hs);
                            ptx::sts16(xr0 + L.x0bytes + 16, ls);
                            ptx::sts16(xr0 + 32, hs >> 16);
                            ptx::sts16(xr0 + L.x0bytes + 32, ls >> 16);
                        }
                    }
                    if (want_out) {                              // trajectories / one-step outputs were asked for (tests, predict)
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
                for (int i = et + kSEpiThreads; i < N * ndp; i += kSEpiThreads) {      // tiles of more than 512 items: the rest
                    const int dp = (i * invN) >> 20, r = i - dp * N;
                    if (r >= nrows) continue;
                    const int d0 = 2 * dp;
                    const bool two = d0 + 1 < D;
                    const float4 ca[2] = {fc[2 * d0], two ? fc[2 * d0 + 2] : make_float4(0.f, 0.f, 0.f, 0.f)};
                    const float4 cb[2] = {fc[2 * d0 + 1], two ? fc[2 * d0 + 3] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1))};
                    const float so[2] = {S[r * (D + 3) + d0], two ? S[r * (D + 3) + d0 + 1] : 0.f};
                    final_pair(r, d0, two, ca, cb, so);
                }
                if (dbg) dbg[13] = clock64();
                if (wdbg) wdbg[52] = clock64();
                if (t + 1 < P.h) publish_input();
                if (dbg) dbg[14] = clock64();
                if (wdbg) wdbg[53] = clock64();
                ptx::bar_sync(1, kSEpiThreads);            // the new state is complete and visible
                if (wdbg) wdbg[54] = clock64();
                if (dbg) dbg[15] = clock64();
                if (env_reward_reads_next(P.env_id)) {
                    if (et < nrows) ret += env_reward_next(P.env_id, S + et * (D + 3));
                } else if (t + 1 < P.h) {
                    add_reward(t + 1);
                }
                if (dbg) dbg[12] = clock64();
            }
            if (P.row_mode == kRowsPlanner && et < nrows) P.ret_p[(size_t)r_src[et] * P.p + r_pi[et]] = ret;
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
static int g_tcs_smem = 0;

// rows per tile: the smallest multiple of 16 from 32 up whose tiles fit the SMs in one wave.  Below N = 128 an MMA costs
// 32 + N / 4 cycles (tools/tc_rate.py: the 4 KB weight block is fetched from shared memory for every MMA), so a CTA's
// time per step is nearly independent of N and the way to go faster is to use more SMs with fewer rows each; 16-row tiles
// would need two waves for the BASELINE workloads and are only taken on request.
int tcs_pick_rows(int rows_per_member, int E, int num_sms) {
    int N = 32;
    while (N < kSMaxRows && E * ((rows_per_member + N - 1) / N) > num_sms) N += 16;
    return N;
}

cudaError_t launch_rollout_tcs(RolloutParams P, const unsigned char* wimg, long long wimg_member_stride, int terms, int kps,
                               int rows_override, int skew, int num_sms, cudaStream_t stream, const char** name, long long* dbg) {
    TcsParams T{};
    T.dbg = dbg;
    T.wimg = wimg;
    T.wimg_member_stride = wimg_member_stride;
    T.Np = round_up(P.H, 16);
    T.NHp = round_up(2 * P.D, 16);
    T.nkb0 = round_up(P.In, 16) / 16;
    T.nkbH = T.Np / 16;
    T.Kcap = max(T.Np, T.nkb0 * 16);
    T.terms = terms;
    T.kps = kps;
    T.skew = skew & 0xfffff;
    T.debug = (skew >> 20) & 0xff;
    const bool generic = (skew >> 28) & 1;     // diagnostic: force the table-driven MMA schedule
    T.nmt = T.Np > 128 ? 2 : 1;
    if (T.Np > 256 || T.NHp > 128 || kps < 1 || kps > kSMaxKps) return cudaErrorInvalidConfiguration;
    if (P.row_hi <= 0) { P.row_lo = 0; P.row_hi = P.rows_per_member; }
    const int span = P.row_hi - P.row_lo;                       // rows of every member this launch covers
    if (span < 1 || P.row_lo < 0 || P.row_hi > P.rows_per_member) return cudaErrorInvalidValue;
    int N = rows_override > 0 ? rows_override : tcs_pick_rows(span, P.E, num_sms);
    N = min(kSMaxRows, max(16, round_up(N, 16)));
    // wide state vectors with 64-row tiles do not leave room for two weight stages: take narrower tiles (more of them)
    while (N > 16 && tcs_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, kps, 2).total > 226 * 1024) N -= 16;
    int tiles = (span + N - 1) / N;
    N = min(N, round_up((span + tiles - 1) / tiles, 16));      // balance the rows over the tiles
    P.rows_per_cta = min(N, (span + tiles - 1) / tiles);
    tiles = (span + P.rows_per_cta - 1) / P.rows_per_cta;
    T.N = N;
    T.tiles_per_member = tiles;
    T.total_tiles = tiles * P.E;
    int stages = kSMaxStages;
    while (stages > 2 && tcs_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, kps, stages).total > 226 * 1024) --stages;
    T.stages = stages;
    const TcsSmem L = tcs_smem_layout(N, P.D, P.A, P.C, P.n_hidden, T.Np, T.NHp, T.Kcap, T.nkb0, kps, stages);
    if (L.total > 226 * 1024) return cudaErrorInvalidConfiguration;
    T.R = P;
    // The stage table of one horizon step, in weight-stream order: GEMM g, M tile mt, K blocks [s0, s0 + kbs).  Two 16-byte
    // words per stage, everything the MMA thread needs precomputed:
    //   e0 = { flags, bytes in the weight image, X offset (16-byte units: input buffer g & 1, K block s0), TMEM column of the
    //          M tile's accumulator }          e1 = { R, kbs, accumulate (s0 > 0), g }
    //   flags: 1 wait xr[0] first (first stage of the GEMM), 2 wait xr[1] first (first stage that needs a K block produced
    //          by the M-tile-1 warps), 4 commit acc_full[mt] after, 8 last stage of the GEMM, 16 M tile 1
    {
        int n = 0;
        for (int g = 0; g <= P.n_hidden; ++g) {
            const int nkb = g == 0 ? T.nkb0 : T.nkbH;
            const int Npad = g == P.n_hidden ? T.NHp : T.Np;
            const int nmt = (Npad + 127) >> 7;
            const int kb_split = (g == 0 || T.nmt == 1) ? 0 : 8;       // the prologue spreads layer 0 over all warps
            bool w1 = false;
            for (int mt = 0; mt < nmt; ++mt) {
                const int R = std::min(128, Npad - 128 * mt);
                for (int s0 = 0; s0 < nkb; s0 += kps) {
                    const int kbs = std::min(kps, nkb - s0);
                    const bool last = mt == nmt - 1 && s0 + kbs >= nkb;
                    uint32_t flags = mt ? 16u : 0u;
                    if (mt == 0 && s0 == 0) flags |= 1u;
                    if (!w1 && (s0 + kbs > kb_split || last)) { flags |= 2u; w1 = true; }
                    if (s0 + kbs >= nkb) flags |= 4u;
                    if (last) flags |= 8u;
                    if (n >= kSMaxEnt) return cudaErrorInvalidConfiguration;
                    T.tab[2 * n] = make_uint4(flags, (uint32_t)kbs * 64u * (uint32_t)R,
                                              (g == 0 ? 0u : (uint32_t)(g & 1) * ((uint32_t)(2 * L.xbytes) >> 4)) + (uint32_t)s0 * 16u, (uint32_t)mt * kSAccStride);
                    T.tab[2 * n + 1] = make_uint4((uint32_t)R, (uint32_t)kbs, s0 > 0 ? 1u : 0u, (uint32_t)g);
                    ++n;
                }
            }
        }
        T.nent = n;
    }
    if ((int)L.total > g_tcs_smem) {
        cudaError_t e = cudaSuccess;
        auto set = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); };
#define CADM_TCS_SET(TE, K0, HP) set((const void*)rollout_tcs_kernel<TE, K0, HP, false>); set((const void*)rollout_tcs_kernel<TE, K0, HP, true>)
        CADM_TCS_SET(3, 0, 0); CADM_TCS_SET(1, 0, 0); CADM_TCS_SET(3, 2, 48); CADM_TCS_SET(1, 2, 48);
        CADM_TCS_SET(3, 3, 48); CADM_TCS_SET(1, 3, 48); CADM_TCS_SET(3, 3, 64); CADM_TCS_SET(1, 3, 64);
#undef CADM_TCS_SET
        if (e != cudaSuccess) return e;
        g_tcs_smem = (int)L.total;
    }
    if (name) *name = terms == 3 ? "rollout_tcs_kernel(swapped operands, fp16 hi/lo x3)" : "rollout_tcs_kernel(swapped operands, f16 x1)";
    const int grid = min(T.total_tiles, num_sms);
    // compile-time MMA schedule for the reference architecture, table-driven otherwise
    const bool spec = !generic && P.n_hidden == 4 && T.Np == 208 && kps == 4 && (T.nkb0 == 2 || T.nkb0 == 3) && (T.NHp == 48 || T.NHp == 64) &&
                      !(T.nkb0 == 2 && T.NHp == 64);
    const bool diag = T.debug != 0 || T.dbg != nullptr;
#define CADM_TCS_LAUNCH(TE, K0, HP)                                                                  \
    do {                                                                                             \
        if (diag) le = launch_chain(rollout_tcs_kernel<TE, K0, HP, true>, dim3(grid), dim3(kSThreads), L.total, stream, T);    \
        else le = launch_chain(rollout_tcs_kernel<TE, K0, HP, false>, dim3(grid), dim3(kSThreads), L.total, stream, T);         \
    } while (0)
    cudaError_t le = cudaSuccess;
    if (!spec) { if (terms == 3) CADM_TCS_LAUNCH(3, 0, 0); else CADM_TCS_LAUNCH(1, 0, 0); }
    else if (T.nkb0 == 2) { if (terms == 3) CADM_TCS_LAUNCH(3, 2, 48); else CADM_TCS_LAUNCH(1, 2, 48); }
    else if (T.NHp == 48) { if (terms == 3) CADM_TCS_LAUNCH(3, 3, 48); else CADM_TCS_LAUNCH(1, 3, 48); }
    else { if (terms == 3) CADM_TCS_LAUNCH(3, 3, 64); else CADM_TCS_LAUNCH(1, 3, 64); }
#undef CADM_TCS_LAUNCH
    return le != cudaSuccess ? le : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Self-test: out[rows, Nout] = X[rows, K] * W[K, Nout] with exactly the operand layouts, descriptors, staging and split
// arithmetic of rollout_tcs_kernel (device diagnostic behind cadm_selftest_tcs_gemm; tests compare with fp64).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tcs_gemm_selftest_kernel(const float* __restrict__ X, const unsigned char* wimg, int rows,
                                                                    int K, int Nout, int kps, int terms, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int nkb = (K + 15) / 16;
    const int Npad = round_up(Nout, 16);
    const int Kcap = nkb * 16;
    const int xsbo = (Kcap / 8) * 128;
    const int xbytes = (rows / 8) * xsbo;
    unsigned char* xhi = smem;
    unsigned char* xlo = smem + xbytes;                 // directly behind X_hi: [X_hi ; X_lo] is one B operand of 2 x rows
    unsigned char* wst = smem + (2 * xbytes + 127) / 128 * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wst + kps * 8192);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    for (int i = tid; i < Kcap * (rows / 8); i += 128) {
        const int rg = i / Kcap, k = i - rg * Kcap;
        float y[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) y[jj] = k < K ? X[(size_t)(rg * 8 + jj) * K + k] * tc::kXScale : 0.f;
        tcs::store_rows8(xhi, xlo, xsbo, rg, k, y);
    }
    ptx::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nmt = (Npad + 127) >> 7;
    if (warp == 0) {
        if (ptx::elect_one()) {
            const uint32_t xhi_d = ptx::smem_u32(xhi) >> 4, w_a = ptx::smem_u32(wst);
            uint32_t ph = 0;
            size_t off = 0;
            for (int mt = 0; mt < nmt; ++mt) {
                const int R = min(128, Npad - 128 * mt);
                for (int s0 = 0; s0 < nkb; s0 += kps) {
                    const int kbs = min(kps, nkb - s0);
                    const uint32_t bytes = (uint32_t)kbs * 64u * (uint32_t)R;
                    ptx::mbar_arrive_expect_tx(&bars[0], bytes);
                    ptx::bulk_g2s(wst, wimg + off, bytes, &bars[0]);
                    off += bytes;
                    ptx::mbar_wait(&bars[0], ph);
                    tc::fence_after_sync();
                    if (terms == 3)
                        tcs::issue_stage<3>(tmem_base + mt * kSAccStride, w_a >> 4, (uint32_t)R, kbs, s0 > 0 ? 1u : 0u, xhi_d + s0 * 16u,
                                            (uint32_t)rows, (uint32_t)xsbo);
                    else
                        tcs::issue_stage<1>(tmem_base + mt * kSAccStride, w_a >> 4, (uint32_t)R, kbs, s0 > 0 ? 1u : 0u, xhi_d + s0 * 16u,
                                            (uint32_t)rows, (uint32_t)xsbo);
                    tc::mma_commit(&bars[1]);
                    ptx::mbar_wait(&bars[1], ph);          // serialise: the single weight slot is reused
                    ph ^= 1u;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int mt = 0; mt < nmt; ++mt) {
        const int j = mt * 128 + tid;
        for (int c = 0; c < rows; c += 8) {
            uint32_t v[8], v1[8];
            tc::tmem_ld8(tl + mt * kSAccStride + c, v);
            tc::tmem_ld8(tl + mt * kSAccStride + rows + c, v1);
            tc::tmem_wait_ld();
            if (j < Nout) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float acc = __uint_as_float(v[q]);
                    if (terms == 3) acc += __uint_as_float(v1[q]);
                    out[(size_t)(c + q) * Nout + j] = acc * (1.0f / (tc::kWScale * tc::kXScale));
                }
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

cudaError_t launch_tcs_gemm_selftest(const float* X, const unsigned char* wimg, int rows, int K, int Nout, int kps, int terms,
                                     float* out, cudaStream_t stream) {
    const int nkb = (K + 15) / 16;
    const int xbytes = (rows / 8) * (nkb * 16 / 8) * 128;
    const int smem_bytes = (2 * xbytes + 127) / 128 * 128 + kps * 8192 + 64;
    cudaError_t e = cudaFuncSetAttribute(tcs_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    tcs_gemm_selftest_kernel<<<1, 128, smem_bytes, stream>>>(X, wimg, rows, K, Nout, kps, terms, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark of the swapped kernel's MMA stream in isolation: one elected thread issues `iters` rounds of 4 weight
// stages (kps K16 blocks each, R weight rows, 2 MMAs per block as in rollout_tcs_kernel<3>) on resident shared memory.
// mode bit 0: a second thread streams bulk copies INTO the ring slots while they are read (the real kernel's traffic);
// mode bit 1: alternate the two accumulator regions per round; mode bit 2: 16 epilogue-like warps hammer shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSThreads, 1) tcs_mma_rate_kernel(int rows, int iters, int R, int kps, int mode,
                                                                    const unsigned char* src, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int xsbo = (208 / 8) * 128;
    const int xbytes = (rows / 8) * xsbo;
    unsigned char* xb = smem;
    unsigned char* ring = smem + (4 * xbytes + 127) / 128 * 128;
    const int slot_bytes = kps * 8192;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 4 * slot_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(bars + 7);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4 * xbytes + 4 * slot_bytes + 128) / 4; i += kSThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) ptx::mbar_init(&bars[i], 1);
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
    if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t x16 = ptx::smem_u32(xb) >> 4, w16 = ptx::smem_u32(ring) >> 4;
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const uint32_t d = tmem_base + ((mode & 2) ? (uint32_t)(it & 1) * kSAccStride : 0u);
                for (int sl = 0; sl < 4; ++sl) {
                    tcs::issue_stage<3>(d, w16 + (uint32_t)sl * ((uint32_t)slot_bytes >> 4), (uint32_t)R, kps, sl > 0 ? 1u : 0u,
                                        x16 + (uint32_t)((sl * kps) % 13) * 16u, (uint32_t)rows, (uint32_t)xsbo);
                    tc::mma_commit(&bars[1]);
                }
            }
            const long long t1 = clock64();
            tc::mma_commit(&bars[0]);
            ptx::mbar_wait(&bars[0], 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) {
                cycles[0] = t1 - t0;
                cycles[1] = t2 - t0;
            }
            *stop = 1u;
        }
        __syncwarp();
    } else if (warp == 0 && (mode & 1)) {
        if (tid == 0) {
            uint32_t ph[4] = {0u, 0u, 0u, 0u};
            long long n = 0;
            const uint32_t bytes = (uint32_t)kps * 64u * (uint32_t)R;
            for (int i = 0; i < 4; ++i) {
                ptx::mbar_arrive_expect_tx(&bars[2 + i], bytes);
                ptx::bulk_g2s(ring + i * slot_bytes, src + ((n++ * 32768) & 0xfffff), bytes, &bars[2 + i]);
            }
            while (!*stop) {
                for (int i = 0; i < 4; ++i) {
                    ptx::mbar_wait(&bars[2 + i], ph[i]);
                    ph[i] ^= 1u;
                    ptx::mbar_arrive_expect_tx(&bars[2 + i], bytes);
                    ptx::bulk_g2s(ring + i * slot_bytes, src + ((n++ * 32768) & 0xfffff), bytes, &bars[2 + i]);
                }
            }
            for (int i = 0; i < 4; ++i) ptx::mbar_wait(&bars[2 + i], ph[i]);
            if (blockIdx.x == 0) cycles[2] = n * bytes;
        }
    } else if (warp >= 2 && (mode & 4)) {
        // epilogue-like traffic: TMEM loads + 16-byte shared-memory stores into the second X buffer
        const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        unsigned char* dst = xb + 2 * xbytes;
        const int lane = tid & 31;
        while (!*stop) {
            uint32_t v[8];
            tc::tmem_ld8(tl + 8 * (warp & 7), v);
            tc::tmem_wait_ld();
            *reinterpret_cast<uint4*>(dst + ((warp - 2) & 3) * xsbo + lane * 16) = make_uint4(v[0], v[1], v[2], v[3]);
            __nanosleep(100);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

cudaError_t launch_tcs_mma_rate(int rows, int iters, int R, int kps, int mode, const unsigned char* src, long long* cycles,
                                cudaStream_t stream) {
    const int xbytes = (rows / 8) * (208 / 8) * 128;
    const int smem_bytes = (4 * xbytes + 127) / 128 * 128 + 4 * kps * 8192 + 256;
    cudaError_t e = cudaFuncSetAttribute(tcs_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    tcs_mma_rate_kernel<<<(mode & 8) ? 148 : ((mode & 16) ? 74 : 1), kSThreads, smem_bytes, stream>>>(rows, iters, R, kps, mode, src, cycles);
    return cudaGetLastError();
}

}  // namespace cadm
