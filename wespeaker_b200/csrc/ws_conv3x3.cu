// Halo-resident 3x3 stride-1 convolution on tcgen05 (sm_100a) for the low-channel 2-D stages:
// BasicBlock convs of ResNet layers 1-3 (wespeaker/models/resnet.py:35-69) and of CAM++'s FCM head
// (wespeaker/models/campplus.py:245-330), 16-bit activations, channels-last [B][F][T][C], C in {32, 64, 128}.
//
// The generic conv-GEMM kernel (ws_gemm_tc2/3.cu) fetches one A tile PER TAP: every input element crosses L2 -> SM nine
// times and a C = 32 layer issues 9 x (8 KB + 2 KB) TMA loads per 128 x 32 tile (round-1 profile: 221 us for an 18.9
// GFLOP conv whose HBM floor is ~25 us).  Here each input row of F is loaded ONCE into a shared-memory ring:
//   * one ring slot = one input row f of one utterance (or of `nb` short utterances side by side) including its two halo
//     columns t = -1 and t = T (TMA out-of-bounds zero fill = the conv's zero padding), stored as K-major swizzled operand
//     rows (one row per time step);
//   * output row f needs slots f-1, f, f+1; tap (df, dt) is a tcgen05.mma operand read of slot f+df starting (1 + dt) rows
//     into it — a descriptor start shifted by whole rows (the swizzle is a function of the absolute shared-memory address,
//     as in ws_res2_fused.cu), so there is no im2col and no per-tap reload;
//   * the CTA walks down F: each step loads one new row, retires one, and runs 9 * npan k-blocks of MMAs into a
//     double-buffered TMEM accumulator; weights are resident in shared memory (C <= 64) or streamed through a TMA ring;
//   * epilogue: folded-BN bias, optional residual (direct 16-byte global loads issued before the accumulator wait), ReLU,
//     swizzled staging + TMA store (which also clips the halo / padding rows).
// Steps (output rows over all utterances) are split contiguously over one persistent CTA per SM.
//
// Warp roles (384 threads): w0 input-row producer, w1 MMA issuer, w2 TMEM allocator, w3 weight producer, w4..w11 epilogue.
#include "ws_tc_common.cuh"

namespace {
using namespace ws_tcdev;

constexpr int kC3Threads = 384;
constexpr int kC3MaxSmem = 222 * 1024;

__device__ __forceinline__ void umma_f16_c3(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// ---- 2-CTA cluster mode (p.cl == 2): the two CTAs walk different output rows but stream the SAME weight blocks, so each
// fetches half of every block and TMA-multicasts it to both; a ring stage is reused once both CTAs' MMAs have read it
__device__ __forceinline__ uint32_t c3_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void c3_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void c3_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void c3_tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
        "%4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_c3_mc(uint32_t bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_c3(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// clock64 phase profile (p.prof != null): accumulate the cycles one elected thread of each role spends in each wait
#define C3_T0() const long long t0__ = p.prof ? clock64() : 0
#define C3_ACC(var) do { if (p.prof) var += clock64() - t0__; } while (0)

// Walk over this CTA's contiguous range of output-row steps.  The (f, b group, t tile, n tile) decomposition costs integer
// divisions only once; every role advances it incrementally (a 64-bit div/mod per step and role was ~5k cycles per step of
// pure overhead in the first version of this kernel).
struct C3Step {
    int f, bg, tt, nt;
    bool first, last;
};
struct C3Iter {
    int s, s_beg, s_end, f, bg, tt, nt;
    __device__ __forceinline__ C3Iter(const WsC3Params& p, int beg, int end) : s(beg), s_beg(beg), s_end(end) {
        f = beg % p.F;
        int img = beg / p.F;
        bg = img % p.n_bg; img /= p.n_bg;
        tt = img % p.n_tt;
        nt = img / p.n_tt;
    }
    __device__ __forceinline__ bool done() const { return s >= s_end; }
    __device__ __forceinline__ C3Step cur(const WsC3Params& p) const {
        C3Step x;
        x.f = f; x.bg = bg; x.tt = tt; x.nt = nt;
        x.first = (s == s_beg) || (f == 0);
        x.last = (s + 1 == s_end) || (f == p.F - 1);
        return x;
    }
    __device__ __forceinline__ void next(const WsC3Params& p) {
        ++s;
        if (++f == p.F) {
            f = 0;
            if (++bg == p.n_bg) { bg = 0; if (++tt == p.n_tt) { tt = 0; ++nt; } }
        }
    }
};

template <int DT, int ROWB, int NPAN, int NMT>
__global__ void __launch_bounds__(kC3Threads, 1) ws_conv3x3_kernel(const __grid_constant__ WsC3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 11];
    __shared__ uint32_t s_tmem;

    // warp index through a shuffle: the compiler then knows the role branches are warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int slot_bytes = p.npan * p.slot_rows * p.row_bytes;
    const int wblk_bytes = p.N * p.row_bytes;
    const int nwblk = 9 * p.npan;
    const uint32_t ring = base;
    const uint32_t wbuf = ring + (uint32_t)(p.R * slot_bytes);
    const uint32_t stg = (wbuf + (uint32_t)((p.w_resident ? nwblk : p.w_stages) * wblk_bytes) + 1023u) & ~1023u;
    const int npanels_out = p.N / p.panel_cols;
    const int panel_stride = p.stg_rows * p.panel_bytes;
    const int tile_bytes = npanels_out * panel_stride;
    const int obuf_bytes = p.n_mt * tile_bytes;                    // one staging buffer = the n_mt output tiles of a step
    const uint32_t s_par = stg + (uint32_t)(p.stg_bufs * obuf_bytes);
    const uint32_t bar_afull = smem_u32(&s_bar[0]);
    const uint32_t bar_aempty = smem_u32(&s_bar[WS_C3_MAX_RING]);
    const uint32_t bar_wfull = smem_u32(&s_bar[2 * WS_C3_MAX_RING]);
    const uint32_t bar_wempty = smem_u32(&s_bar[2 * WS_C3_MAX_RING + WS_C3_MAX_WSTAGES]);
    const uint32_t bar_wres = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES]);
    const uint32_t bar_tfull = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 1]);   // [2]
    const uint32_t bar_tempty = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 3]);  // [2]
    const uint32_t bar_rfull = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 5]);   // [set][buffer] residual panels landed
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * p.n_mt * p.N) tmem_cols <<= 1;

    int s_beg = (int)((long long)p.total_steps * blockIdx.x / gridDim.x);
    int s_end = (int)((long long)p.total_steps * (blockIdx.x + 1) / gridDim.x);
    int w_rounds = s_end - s_beg;          // weight passes this CTA's ring goes through (one per step)
    const bool mc = p.cl == 2;
    const uint32_t crank = mc ? c3_cluster_rank() : 0u;
    if (mc) {   // the cluster takes a contiguous range of steps and splits it in halves; both CTAs consume ceil(n / 2) weight passes
        const int cid = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
        const int c_beg = (int)((long long)p.total_steps * cid / ncl), c_end = (int)((long long)p.total_steps * (cid + 1) / ncl);
        const int n0 = (c_end - c_beg + 1) >> 1;
        s_beg = crank == 0 ? c_beg : c_beg + n0;
        s_end = crank == 0 ? c_beg + n0 : c_end;
        w_rounds = n0;
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.amap); prefetch_tmap(&p.amap_tail); prefetch_tmap(&p.wmap); prefetch_tmap(&p.omap); prefetch_tmap(&p.rmap);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.R; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1); }
        for (int i = 0; i < WS_C3_MAX_WSTAGES; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, mc ? 2 : 1); }
        mbar_init(bar_wres, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
        for (int i = 0; i < 6; ++i) mbar_init(bar_rfull + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (mc) c3_cluster_sync();   // the peer's barriers are initialised before any multicast load / remote arrive reaches them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================================ input-row producer ================================
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(p.npan * p.rows_loaded * p.row_bytes);
            int j = 0;
            long long pw_aempty = 0;
            const long long tstart = p.prof ? clock64() : 0;
            int slot = 0;
            uint32_t sphase = 1;                             // waits on the "previous" phase of a fresh barrier pass at once
            auto load_row = [&](int frow, int bg, int tt) {
                { C3_T0(); mbar_wait(bar_aempty + 8 * slot, sphase); C3_ACC(pw_aempty); }
                if (p.dbg & 2) { mbar_arrive(bar_afull + 8 * slot); if (++slot == p.R) { slot = 0; sphase ^= 1u; } ++j; return; }
                mbar_expect_tx(bar_afull + 8 * slot, tx);
                const int tc = tt * p.tb - 1, b0 = bg * p.nb;
                for (int kp = 0; kp < p.npan; ++kp) {
                    const uint32_t dst = ring + (uint32_t)((slot * p.npan + kp) * p.slot_rows * p.row_bytes);
                    if (p.st == 2) {      // even plane from output column t0, odd plane shifted by one (O'[k] = x[2k - 1])
                        tma_load_4d(dst, &p.amap, bar_afull + 8 * slot, kp * p.kc, tc + 1, frow, b0);
                        tma_load_4d(dst + (uint32_t)(p.sub_rows * p.row_bytes), &p.amap_tail, bar_afull + 8 * slot, kp * p.kc, tc, frow, b0);
                    } else if (p.single_box) {
                        tma_load_4d(dst, &p.amap, bar_afull + 8 * slot, kp * p.kc, tc, frow, b0);
                    } else {
                        for (int m = 0; m < p.n_mt; ++m)
                            tma_load_4d(dst + (uint32_t)(m * 128 * p.row_bytes), &p.amap, bar_afull + 8 * slot, kp * p.kc,
                                        tc + 128 * m, frow, b0);
                        tma_load_4d(dst + (uint32_t)(p.n_mt * 128 * p.row_bytes), &p.amap_tail, bar_afull + 8 * slot,
                                    kp * p.kc, tc + 128 * p.n_mt, frow, b0);
                    }
                }
                if (++slot == p.R) { slot = 0; sphase ^= 1u; }
                ++j;
            };
            for (C3Iter it(p, s_beg, s_end); !it.done(); it.next(p)) {
                const C3Step st = it.cur(p);
                const int fi = p.sf * st.f;                  // input row under the centre tap
                if (st.first) load_row(fi - 1, st.bg, st.tt);
                if (st.first || p.sf == 2) load_row(fi, st.bg, st.tt);
                load_row(fi + 1, st.bg, st.tt);
            }
            if (p.prof) { p.prof[blockIdx.x * 16 + 0] = pw_aempty; p.prof[blockIdx.x * 16 + 1] = clock64() - tstart; }
        }
    } else if (warp == 3) {
        // ================================ weight producer ================================
        if (lane == 0 && w_rounds > 0) {
            if (p.w_resident) {
                mbar_expect_tx(bar_wres, (uint32_t)(nwblk * wblk_bytes));
                for (int tap = 0; tap < 9; ++tap)
                    for (int kp = 0; kp < p.npan; ++kp)
                        tma_load_2d(wbuf + (uint32_t)((tap * p.npan + kp) * wblk_bytes), &p.wmap, bar_wres,
                                    tap * p.Cin + kp * p.kc, 0);
            } else {
                int ws = 0;
                uint32_t wphase = 1;
                C3Iter it(p, s_beg, s_end);
                for (int r = 0; r < w_rounds; ++r) {
                    const int nt = mc ? 0 : it.cur(p).nt;      // cluster mode needs one channel tile (host guarantees n_nt == 1)
                    for (int tap = 0; tap < 9; ++tap)
                        for (int kp = 0; kp < p.npan; ++kp) {
                            mbar_wait(bar_wempty + 8 * ws, wphase);
                            mbar_expect_tx(bar_wfull + 8 * ws, (uint32_t)wblk_bytes);
                            if (mc)   // this CTA's half of the block (N/2 rows), to both CTAs
                                c3_tma_load_2d_mc(wbuf + (uint32_t)(ws * wblk_bytes) + crank * (uint32_t)(wblk_bytes >> 1), &p.wmap,
                                                  bar_wfull + 8 * ws, tap * p.Cin + kp * p.kc, (int)crank * (p.N >> 1), (unsigned short)3);
                            else
                                tma_load_2d(wbuf + (uint32_t)(ws * wblk_bytes), &p.wmap, bar_wfull + 8 * ws,
                                            tap * p.Cin + kp * p.kc, nt * p.N);
                            if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
                        }
                    if (!mc) it.next(p);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // One elected lane issues every tcgen05.mma of the CTA, but the WHOLE warp runs this loop with warp-uniform values:
        // inside an `if (lane == 0)` region the compiler has to move every descriptor into uniform registers through an
        // ELECT / R2UR.BROADCAST convergence loop in front of each UTCHMMA (~20 instructions, 65-80 cycles per MMA measured
        // with tools/experimental/mma_rate_probe.cu — more than a 128 x N x 16 MMA with N <= 128 takes).  Descriptor
        // arithmetic is hoisted to once per step / k-block; the k loop is a compile-time unroll of 64-bit adds.
        if (s_beg < s_end) {
            uint32_t elected;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            constexpr int KPER = ROWB / 32;
            const uint64_t dhi = umma_desc(0u, ROWB);                                 // descriptor without the address field
            const uint32_t dt_off = (uint32_t)ROWB >> 4;                              // one operand row, in 16-byte units
            // operand start of the three time taps inside a slot: consecutive rows, or (stride 2 along t) the shifted-odd,
            // even, shifted-odd+1 planes
            const uint32_t sub_off = (uint32_t)(p.sub_rows * ROWB) >> 4;
            const uint32_t tap_off0 = p.st == 2 ? sub_off : 0u, tap_off1 = p.st == 2 ? 0u : dt_off,
                           tap_off2 = p.st == 2 ? sub_off + dt_off : 2u * dt_off;    // (scalars: a runtime-indexed array would live in local memory)
            const uint32_t kp_off = (uint32_t)(p.slot_rows * ROWB) >> 4;
            const uint32_t mt_off = (uint32_t)(128 * ROWB) >> 4;
            const uint32_t wb_off = (uint32_t)wblk_bytes >> 4;
            const uint64_t wdesc0 = dhi | (uint64_t)((wbuf & 0x3FFFFu) >> 4);
            const int R = p.R, N = p.N;
            int j = 0, wst = 0, step = 0;
            uint32_t wph = 0;
            long long mw_afull = 0, mw_tempty = 0, mw_wfull = 0;
            const long long tstart = p.prof ? clock64() : 0;
            if (p.w_resident) mbar_wait(bar_wres, 0);
            int jslot = 0;                                   // == j % R and (j / R) & 1, kept incrementally
            uint32_t jphase = 0;
            for (C3Iter it(p, s_beg, s_end); !it.done(); it.next(p), ++step) {
                const C3Step st = it.cur(p);
                const int nnew = st.first ? 3 : p.sf;        // a new image segment starts with rows f-1, f, f+1
                j += nnew;                                   // rows f-1, f, f+1 are loads j-3, j-2, j-1
                {   // wait for the newly loaded rows: slot / phase of load i follow incrementally from (jslot, jphase)
                    C3_T0();
                    for (int i = 0; i < nnew; ++i) {
                        mbar_wait(bar_afull + 8 * jslot, jphase);
                        if (++jslot == R) { jslot = 0; jphase ^= 1u; }
                    }
                    C3_ACC(mw_afull);
                }
                const int buf = step & 1;
                { C3_T0(); mbar_wait(bar_tempty + 8 * buf, ((((uint32_t)step) >> 1) & 1u) ^ 1u); C3_ACC(mw_tempty); }
                tc_fence_after();
                // compact tap loop (a full 9-way unroll made the issuing warp run ~9 KB of straight-line code once per step,
                // far beyond the 6 KB L0 instruction cache): df / dt advance incrementally, only the k loop is unrolled
                int sl0 = jslot - 3 < 0 ? jslot - 3 + R : jslot - 3;        // slot of row f-1 (= load j-3)
                const int slot_first = sl0;
                uint64_t a_df = dhi | (uint64_t)(((ring + (uint32_t)(sl0 * slot_bytes)) & 0x3FFFFu) >> 4);
                const uint64_t a_ring0 = dhi | (uint64_t)((ring & 0x3FFFFu) >> 4);
                const uint32_t slot_off = (uint32_t)slot_bytes >> 4;
                const uint32_t tacc0 = tmem_base + (uint32_t)(buf * NMT * N);
                uint64_t bd = wdesc0;
                uint32_t accum = 0;
                // runtime loop over the three input rows, compile-time unroll of everything inside (time taps, K panels, M
                // tiles, k steps): the scalar work between two MMAs is a couple of uniform adds, no selects or branches
#pragma unroll 1
                for (int df = 0; df < 3; ++df) {
#pragma unroll
                    for (int dt = 0; dt < 3; ++dt) {
                        const uint64_t a_dt = a_df + (uint64_t)(dt == 0 ? tap_off0 : (dt == 1 ? tap_off1 : tap_off2));
#pragma unroll
                        for (int kp = 0; kp < NPAN; ++kp) {
                            if (!p.w_resident) {
                                { C3_T0(); mbar_wait(bar_wfull + 8 * wst, wph); C3_ACC(mw_wfull); }
                                tc_fence_after();
                                bd = wdesc0 + (uint64_t)(wst * wb_off);
                            }
                            if (elected && !(p.dbg & 8)) {
#pragma unroll
                                for (int mt = 0; mt < NMT; ++mt) {
#pragma unroll
                                    for (int k = 0; k < KPER; ++k)
                                        umma_f16_c3(tacc0 + (uint32_t)(mt * N), a_dt + (uint64_t)(kp * kp_off + mt * mt_off + 2 * k),
                                                    bd + (uint64_t)(2 * k), p.idesc, (kp == 0 && k == 0) ? accum : 1u);
                                }
                            }
                            accum = 1u;
                            if (p.w_resident) {
                                bd += wb_off;
                            } else {
                                if (elected) { if (mc) umma_commit_c3_mc(bar_wempty + 8 * wst, (unsigned short)3); else umma_commit_c3(bar_wempty + 8 * wst); }
                                if (++wst == p.w_stages) { wst = 0; wph ^= 1u; }
                            }
                        }
                    }
                    if (++sl0 == R) { sl0 = 0; a_df = a_ring0; } else a_df += slot_off;   // next input row (ring wrap)
                }
                int sl[3];
                sl[0] = slot_first;
                sl[1] = slot_first + 1 == R ? 0 : slot_first + 1;
                sl[2] = sl[1] + 1 == R ? 0 : sl[1] + 1;
                if (elected) {
                    umma_commit_c3(bar_tfull + 8 * buf);
                    umma_commit_c3(bar_aempty + 8 * sl[0]);              // row f-1 is not needed by later steps
                    if (st.last || p.sf == 2) umma_commit_c3(bar_aempty + 8 * sl[1]);   // stride 2: neither is the centre row
                    if (st.last) umma_commit_c3(bar_aempty + 8 * sl[2]);                 // end of this image segment
                }
                __syncwarp();
            }
            // cluster mode: a CTA with one step fewer than its peer still has to release the weight stages of the last pass
            if (mc && !p.w_resident) {
                for (int r = s_end - s_beg; r < w_rounds; ++r)
                    for (int i = 0; i < 9 * NPAN; ++i) {
                        mbar_wait(bar_wfull + 8 * wst, wph);
                        if (elected) { mbar_arrive(bar_wempty + 8 * wst); c3_arrive_remote(bar_wempty + 8 * wst, crank ^ 1u); }
                        __syncwarp();
                        if (++wst == p.w_stages) { wst = 0; wph ^= 1u; }
                    }
            }
            if (p.prof && lane == 0) {
                p.prof[blockIdx.x * 16 + 2] = mw_afull; p.prof[blockIdx.x * 16 + 3] = mw_tempty;
                p.prof[blockIdx.x * 16 + 4] = mw_wfull; p.prof[blockIdx.x * 16 + 5] = clock64() - tstart;
            }
            (void)j;
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        // Two independent warp sets (w4..w7, w8..w11), each covering the 128 accumulator rows.  The store units of a step
        // ((M tile, column panel) pairs = one TMA box each) alternate between the sets; a set has its own named barrier,
        // its own elected TMA thread (stores, residual loads and bulk-group waits are per thread) and its own staging
        // panels, so the two halves of a step's epilogue never wait for each other.
        const int q = warp & 3, r = q * 32 + lane, set = (warp - 4) >> 2, et = threadIdx.x - 128;
        const bool tma_thread = (et == 128 * set);
        float* spar = reinterpret_cast<float*>(smem_raw + (s_par - smem_u32(smem_raw)));
        const int nunits = p.n_mt * npanels_out, cpp = p.panel_cols / 32;      // store units per step, 32-column chunks per panel
        auto set_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + set) : "memory"); };
        int step = 0, last_nt = -1;
        long long ew_store = 0, ew_tfull = 0, ew_res = 0, ew_body = 0;
        const long long tstart = p.prof ? clock64() : 0;
        const bool has_res = p.res != nullptr && !(p.dbg & 1);
        const int nbuf = p.stg_bufs;
        // The residual panel of a unit is TMA-loaded INTO that unit's output staging panel (same box geometry and swizzle);
        // the epilogue reads it from there and overwrites it in place.  With three staging buffers the load for step s+1 is
        // issued at the START of step s (its buffer was last read by the store of step s-2, long drained): a full step of
        // lead time.  With two it can only be issued at the end of step s, with one in line.
        const uint32_t unit_tx = (uint32_t)(p.panel_bytes * (p.case_b ? p.P * p.nb : 128));
        int nmine = 0;
        for (int u = set; u < nunits; u += 2) ++nmine;
        auto issue_res = [&](const C3Step& x, int ob) {
            mbar_expect_tx(bar_rfull + 8 * (3 * set + ob), unit_tx * (uint32_t)nmine);
            for (int u = set; u < nunits; u += 2) {
                const int mt = u / npanels_out, pn = u - mt * npanels_out;
                tma_load_4d(stg + (uint32_t)(ob * obuf_bytes + mt * tile_bytes + pn * panel_stride), &p.rmap,
                            bar_rfull + 8 * (3 * set + ob), x.nt * p.N + pn * p.panel_cols, p.case_b ? 0 : x.tt * p.tb + 128 * mt,
                            x.f, x.bg * p.nb);
            }
        };
        if (has_res && nbuf >= 2 && tma_thread && nmine > 0 && s_beg < s_end) issue_res(C3Iter(p, s_beg, s_end).cur(p), 0);
        int ob = 0;                                          // == step % nbuf
        uint32_t rph = 0;                                    // == (step / nbuf) & 1
        for (C3Iter it(p, s_beg, s_end); !it.done(); it.next(p), ++step) {
            const C3Step st = it.cur(p);
            const int n0 = st.nt * p.N, t0 = st.tt * p.tb, b0 = st.bg * p.nb, buf = step & 1;
            const uint32_t sbuf = stg + (uint32_t)(ob * obuf_bytes);
            const int ob_next = ob + 1 == nbuf ? 0 : ob + 1;
            if (nbuf == 1) {   // single staging buffer: free once this set's previous store has read it
                C3_T0();
                if (tma_thread) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    if (has_res && nmine > 0) issue_res(st, 0);
                }
                C3_ACC(ew_store);
            } else if (nbuf == 3 && tma_thread) {   // buffer of step s+1 == buffer of step s-2: drained unless 2 stores pend
                C3_T0();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                C3_ACC(ew_store);
                if (has_res && nmine > 0) {
                    C3Iter nx = it;
                    nx.next(p);
                    if (!nx.done()) issue_res(nx.cur(p), ob_next);
                }
            }
            if (st.nt != last_nt) {     // (all 256 epilogue threads: both sets walk the same steps)
                if (last_nt >= 0) epi_bar_sync();
                for (int c = et; c < p.N; c += 256) spar[c] = p.bias ? __ldg(p.bias + n0 + c) : 0.f;
                last_nt = st.nt;
                epi_bar_sync();
            }
            set_bar();
            { C3_T0(); mbar_wait(bar_tfull + 8 * buf, (((uint32_t)step) >> 1) & 1u); C3_ACC(ew_tfull); }
            tc_fence_after();
            if (has_res && nmine > 0) {
                C3_T0();
                mbar_wait(bar_rfull + 8 * (3 * set + ob), rph);
                C3_ACC(ew_res);
            }
            const long long tbody = p.prof ? clock64() : 0;
            for (int u = set; u < nunits && !(p.dbg & 1); u += 2) {
                const int mt = u / npanels_out, pn = u - mt * npanels_out;
                const uint32_t pbase = sbuf + (uint32_t)(mt * tile_bytes + pn * panel_stride);
                for (int cc = 0; cc < cpp; ++cc) {
                    const int c = pn * p.panel_cols + cc * 32;
                    uint32_t raw[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.n_mt + mt) * p.N + c), raw);
                    tmem_ld_wait();
                    float v[32];
                    const float4* sb = reinterpret_cast<const float4*>(spar + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = sb[i];
                        v[4 * i] = __uint_as_float(raw[4 * i]) + b4.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b4.y;
                        v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b4.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b4.w;
                    }
                    if (has_res) {
                        float rin[32];
                        stage_load32(pbase, r, p.panel_bytes, cc * 32, DT, rin);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += rin[i];
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        if (p.relu == 2) {   // Hardtanh(0, 20): the Res2Net / ERes2Net "ReLU" (eres2net.py:43-52)
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fminf(v[i], 20.f);
                        }
                    }
                    if (p.lens != nullptr) {   // length-masked batch: frames behind the utterance's end are the next conv's padding
                        const int i2 = mt * 128 + r, u = i2 / p.P, tti = i2 - u * p.P, bb = min(b0 + u, p.B - 1);
                        if (t0 + tti >= p.lens[bb]) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0.f;
                        }
                    }
                    // rows that are no output position (halo columns, rows past T or B) hold junk: the TMA store clips them
                    stage_store32(pbase, r, p.panel_bytes, cc * 32, DT, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            set_bar();
            if (tma_thread) {
                if (!(p.dbg & 1)) {
                    for (int u = set; u < nunits; u += 2) {
                        const int mt = u / npanels_out, pn = u - mt * npanels_out;
                        tma_store_4d(&p.omap, sbuf + (uint32_t)(mt * tile_bytes + pn * panel_stride), n0 + pn * p.panel_cols,
                                     p.case_b ? 0 : t0 + 128 * mt, st.f, b0);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (nbuf == 2) {   // the other buffer (this set's store of step s-1) must be drained before anything lands in it
                    C3_T0();
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    C3_ACC(ew_store);
                    if (has_res && nmine > 0) {
                        C3Iter nx = it;
                        nx.next(p);
                        if (!nx.done()) issue_res(nx.cur(p), ob_next);
                    }
                }
            }
            ob = ob_next;
            if (ob == 0) rph ^= 1u;
            if (p.prof) ew_body += clock64() - tbody;
        }
        if (tma_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.prof && et == 0) {
            p.prof[blockIdx.x * 16 + 6] = ew_store; p.prof[blockIdx.x * 16 + 7] = ew_tfull; p.prof[blockIdx.x * 16 + 8] = ew_res;
            p.prof[blockIdx.x * 16 + 9] = ew_body; p.prof[blockIdx.x * 16 + 10] = clock64() - tstart;
            p.prof[blockIdx.x * 16 + 11] = s_end - s_beg;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (mc) c3_cluster_sync();   // the peer may still multicast into this CTA's ring / arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

namespace {
template <int DT, int ROWB, int NPAN, int NMT>
const char* c3_launch_one(const WsC3Params* p, cudaStream_t s, bool attr_only) {
    if (attr_only) {
        cudaError_t e = cudaFuncSetAttribute(ws_conv3x3_kernel<DT, ROWB, NPAN, NMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC3MaxSmem);
        return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
    }
    if (p->cl == 2) {   // 2-CTA clusters (weight multicast): cluster dimension as a launch attribute
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)p->grid); cfg.blockDim = dim3(kC3Threads); cfg.dynamicSmemBytes = (size_t)p->smem_bytes; cfg.stream = s;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, ws_conv3x3_kernel<DT, ROWB, NPAN, NMT>, *p);
        return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
    }
    ws_conv3x3_kernel<DT, ROWB, NPAN, NMT><<<p->grid, kC3Threads, p->smem_bytes, s>>>(*p);
    return nullptr;
}
// (operand row bytes, K panels per tap, M tiles per step): 64-byte rows only exist with one panel (C = 32)
template <int DT>
const char* c3_dispatch(const WsC3Params* p, cudaStream_t s, bool attr_only, int rowb, int npan, int nmt) {
    if (rowb == 64 && npan == 1 && nmt == 1) return c3_launch_one<DT, 64, 1, 1>(p, s, attr_only);
    if (rowb == 64 && npan == 1 && nmt == 2) return c3_launch_one<DT, 64, 1, 2>(p, s, attr_only);
    if (rowb == 128 && npan == 1 && nmt == 1) return c3_launch_one<DT, 128, 1, 1>(p, s, attr_only);
    if (rowb == 128 && npan == 1 && nmt == 2) return c3_launch_one<DT, 128, 1, 2>(p, s, attr_only);
    if (rowb == 128 && npan == 2 && nmt == 1) return c3_launch_one<DT, 128, 2, 1>(p, s, attr_only);
    if (rowb == 128 && npan == 2 && nmt == 2) return c3_launch_one<DT, 128, 2, 2>(p, s, attr_only);
    return "conv3x3: unsupported (row bytes, K panels, M tiles) combination";
}
}  // namespace

extern "C" const char* ws_c3_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    const int combos[6][3] = {{64, 1, 1}, {64, 1, 2}, {128, 1, 1}, {128, 1, 2}, {128, 2, 1}, {128, 2, 2}};
    for (auto& c : combos) {
        const char* m = c3_dispatch<WS_BF16>(nullptr, nullptr, true, c[0], c[1], c[2]);
        if (!m) m = c3_dispatch<WS_F16>(nullptr, nullptr, true, c[0], c[1], c[2]);
        if (m) { cudaGetLastError(); return m; }
    }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" int ws_c3_max_smem(void) { return kC3MaxSmem; }

extern "C" const char* ws_c3_launch(const WsC3Params* p, cudaStream_t s) {
    const char* m;
    if (p->dtype == WS_BF16) m = c3_dispatch<WS_BF16>(p, s, false, p->row_bytes, p->npan, p->n_mt);
    else if (p->dtype == WS_F16) m = c3_dispatch<WS_F16>(p, s, false, p->row_bytes, p->npan, p->n_mt);
    else return "conv3x3: 16-bit activations only";
    if (m) return m;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
