// 2-CTA (cta_group::2) variant of the persistent conv/GEMM (ws_gemm_tc2.cu): a cluster of two CTAs on one TPC computes a
// 256-position x bn-channel tile with UMMA M=256.  Each CTA loads its own 128-row activation tile and HALF of the bn-row
// weight tile; the leader CTA issues tcgen05.mma.cta_group::2, which reads both halves across the pair, so L2 operand
// traffic per FLOP drops by a third versus two independent 128 x bn tiles (the round-1 profile showed the 1-CTA kernel is
// L2-operand-bandwidth bound).  Completion is multicast to both CTAs' mbarriers; each CTA runs its own epilogue from its
// own TMEM; accumulator buffers are handed back through the leader's barrier with remote (mapa) arrives.
//
// CL = 4: a cluster of TWO such pairs on adjacent position tiles and the same channel tile.  The CTAs with the same rank
// inside their pair need the same half of the weight tile: each loads a QUARTER of the tile and TMA-multicasts it to both,
// so a CTA pulls 16 KB (activations) + 8 KB (weights) per k-block from L2 instead of 16 + 16 (the round-1/2 profiles show
// the mainloop pinned at the L2 -> SM operand rate, 9-11 TB/s).  Stage reuse then needs BOTH pairs to have consumed the
// stage: each pair's MMA leader multicasts its tcgen05.commit to all four CTAs and the empty barriers count two arrivals.
//
// ---- original v2 header:
// Persistent warp-specialised conv/GEMM on tcgen05 + TMEM + TMA (sm_100a), version 2.
//
// Same operator as ws_gemm_tc.cu (taps over rank-4 TMA maps, 128 positions x bn channels per tile, K-major swizzled
// operands) with the three things the round-1 profile asked for:
//   * persistent CTAs (one per SM) walking a static tile schedule, TMEM accumulator double-buffered (2 x bn columns)
//     so the epilogue of tile i overlaps the MMAs of tile i+1, no per-tile prologue (TMEM alloc, barrier init);
//   * bn up to 256: A tile (16 KB) is reused against a 256-row W tile -> 25 % less L2 operand traffic per FLOP;
//   * coalesced, asynchronous epilogue: per-channel parameters staged in smem, residual / Res2 "add2" tiles fetched by
//     TMA into swizzled smem while the mainloop runs, outputs packed into 128-B-swizzled smem panels and written with
//     TMA stores (cp.async.bulk.tensor ... bulk_group), which also clips out-of-range rows/positions for free.
//
// Warp roles (256 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w3 idle, w4..w7 epilogue.
#include "ws_tc_common.cuh"

#ifdef WS_TC3_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_T(x) const long long x = clock64()
#define PROF_ADD(i, v) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_prof[i], (unsigned long long)(v)); } while (0)
#else
#define PROF_T(x)
#define PROF_ADD(i, v)
#endif

namespace {
using namespace ws_tcdev;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
        : "memory");
}
// 2-SM TMA loads: executed by both CTAs, data lands in the issuing CTA's smem, transaction bytes are credited to the
// LEADER CTA's mbarrier (peer bit of the barrier address cleared, cute::Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// completion of all prior MMAs -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit(uint32_t bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// 2-SM TMA load multicast to the CTAs in `mask` (same smem offset in each; the bytes are credited to each destination's
// pair-leader barrier, cute SM100_TMA_2SM_LOAD_MULTICAST)
__device__ __forceinline__ void tma2_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
        "{%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
constexpr int kThreads3 = 384;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4..w11 epilogue (2 warps per TMEM lane quadrant)
constexpr int kMaxDynSmem3 = 220 * 1024;

// LEAN != 0 (= WsDType of the activations + 1) compiles only the common epilogue: bias, ReLU, BN scale/shift, residual
// tile, ReLU, one output (+ its 3xTF32 low twin for fp32).  The generic epilogue (row bias, gates, Res2 second output,
// tanh/sigmoid, every dtype) is ~100 KB of SASS of which a given layer executes a few KB scattered between never-taken
// branches; ncu showed its warps stalled on instruction fetch (stall_no_inst) for half of their samples, which made the
// short-K 1x1 convs epilogue-bound.  The lean chunk body is ~4 KB and stays in the 6 KB L0 instruction cache.
template <int KIND, int LEAN, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kThreads3, 1) ws_conv_gemm_tc3_kernel(const __grid_constant__ WsTc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_TC_MAX_STAGES + 6];
    __shared__ uint32_t s_tmem;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t rank = cluster_ctarank();            // 0..CL-1; pairs are (0,1) and (2,3)
    const uint32_t r2 = rank & 1u, pr = rank >> 1;      // rank inside the pair, pair inside the cluster
    const bool leader = r2 == 0;
    const int ncl = (int)gridDim.x / CL, cid = (int)blockIdx.x / CL;
    const unsigned short pair_mask = (unsigned short)(3u << (2 * pr)), all_mask = (unsigned short)((1u << CL) - 1u);
    const int a_bytes = 128 * p.bk_bytes, b_bytes = (p.bn / 2) * p.bk_bytes;   // own A tile + own HALF of the W tile
    const int stage_bytes = a_bytes + b_bytes;
    const int es = KIND == 0 ? 4 : 2;
    const int panel_bytes = p.panel_bytes;               // row bytes of one staging panel (128 / 64)
    const int panel_cols = panel_bytes / es;
    const int npanels = p.bn / panel_cols;
    const int tile_out_bytes = 128 * p.bn * es;          // one full output tile in staging
    const uint32_t ring = base;
    const uint32_t stg_out = ring + (uint32_t)(p.nstages * stage_bytes);                  // p.nout output tiles
    const uint32_t stg_in = stg_out + (uint32_t)(p.nout * tile_out_bytes);                // only if p.has_epin
    const uint32_t s_par = stg_in + (uint32_t)(p.has_epin ? tile_out_bytes : 0);          // 3 * bn floats
    const uint32_t bar_full = smem_u32(&s_bar[0]);
    const uint32_t bar_empty = smem_u32(&s_bar[WS_TC_MAX_STAGES]);
    const uint32_t bar_tfull = smem_u32(&s_bar[2 * WS_TC_MAX_STAGES]);       // [2]
    const uint32_t bar_tempty = smem_u32(&s_bar[2 * WS_TC_MAX_STAGES + 2]);  // [2]
    const uint32_t bar_ifull = smem_u32(&s_bar[2 * WS_TC_MAX_STAGES + 4]);
    const uint32_t bar_iempty = smem_u32(&s_bar[2 * WS_TC_MAX_STAGES + 5]);
    const uint32_t tmem_cols = (uint32_t)(2 * p.bn < 32 ? 32 : 2 * p.bn);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < WS_MAX_SRC; ++i) prefetch_tmap(&p.amap[i]);
        prefetch_tmap(&p.wmap);
        for (int i = 0; i < p.nout; ++i) prefetch_tmap(&p.omap[i]);
        if (p.has_epin) prefetch_tmap(&p.imap);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, CL / 2);   // one commit per pair
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 16);  // one arrive per epilogue warp of BOTH CTAs (leader's copy is the live one)
        }
        mbar_init(bar_ifull, 1);
        mbar_init(bar_iempty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();   // both CTAs' barriers initialised and TMEM allocated before any cross-CTA traffic
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            const int bk_elems = p.bk_bytes / es;
            int it = 0, j = 0;
            for (int tile = cid; tile < p.num_tiles; tile += ncl, ++j) {
                int tt = (tile / p.tiles_n) * CL + (int)rank;           // this CTA's position tile inside the pair
                const int n0 = (tile % p.tiles_n) * p.bn;
                const int t0 = (tt % p.tiles_t) << p.bt_log2; tt /= p.tiles_t;
                const int f0 = (tt % p.tiles_f) << p.bf_log2; tt /= p.tiles_f;
                const int b0 = tt << p.bb_log2;                            // >= B for the odd tail tile: all OOB
                for (int tp = 0; tp < p.ntaps; ++tp) {
                    const WsTcTap tap = p.taps[tp];
                    for (int kb = 0; kb < tap.nkb; ++kb) {
                        // 3xTF32: three passes per k-block, small terms first: x_lo*W, x*W_lo, x*W (fp32 operands are
                        // truncated to tf32 by the tensor core, so x and W themselves serve as the high parts)
                        for (int ps = 3 - p.nsplit; ps < 3; ++ps, ++it) {
                            const int s = it % p.nstages;
                            const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                            if (leader) mbar_expect_tx(bar_full + 8 * s, (uint32_t)(2 * stage_bytes));  // both CTAs' bytes
                            const uint32_t sa = ring + (uint32_t)(s * stage_bytes);
                            tma2_load_4d(sa, ps == 0 ? &p.amap_lo[tap.map] : &p.amap[tap.map], bar_full + 8 * s,
                                         tap.c0 + kb * bk_elems, t0 + tap.dt, f0 + tap.df, b0);
                            if (CL == 2)
                                tma2_load_2d(sa + (uint32_t)a_bytes, ps == 1 ? &p.wmap_lo : &p.wmap, bar_full + 8 * s,
                                             tap.wk + kb * bk_elems, n0 + (int)r2 * (p.bn / 2));
                            else   // this CTA's quarter of the weight tile, to itself and to the same-rank CTA of the other pair
                                tma2_load_2d_mc(sa + (uint32_t)(a_bytes + (int)pr * (b_bytes / 2)), ps == 1 ? &p.wmap_lo : &p.wmap,
                                                bar_full + 8 * s, tap.wk + kb * bk_elems,
                                                n0 + (int)r2 * (p.bn / 2) + (int)pr * (p.bn / 4),
                                                (unsigned short)((1u << r2) | (4u << r2)));
                        }
                    }
                }
                // epilogue-input tile (residual / add2): issued after this tile's operand loads so that waiting for
                // the previous tile's epilogue to release the buffer never starves the MMA pipe
                if (p.has_epin) {
                    mbar_wait(bar_iempty, ((uint32_t)j & 1u) ^ 1u);
                    mbar_expect_tx(bar_ifull, (uint32_t)tile_out_bytes);
                    for (int pn = 0; pn < npanels; ++pn)
                        tma_load_4d(stg_in + (uint32_t)(pn * 128 * panel_bytes), &p.imap, bar_ifull, n0 + pn * panel_cols,
                                    t0, f0, b0);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ================================
        // all 32 lanes run the loop with warp-uniform values; only the tcgen05 instructions are predicated on the elected
        // lane (see elect_one() in ws_tc_common.cuh)
        if (leader) {
            const uint32_t elected = elect_one();
            const int kper = p.bk_bytes / 32;
            int s = 0, j = 0;
            uint32_t ph = 0;
            for (int tile = cid; tile < p.num_tiles; tile += ncl, ++j) {
                const int buf = j & 1;
                PROF_T(m0);
                mbar_wait(bar_tempty + 8 * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);  // epilogue drained this buffer
                PROF_T(m1);
                PROF_ADD(8, m1 - m0);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * p.bn);
                const int nkit = p.nk_total * p.nsplit;
                for (int kit = 0; kit < nkit; ++kit) {
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = ring + (uint32_t)(s * stage_bytes);
                    const uint64_t adesc = umma_desc(sa, p.bk_bytes);
                    const uint64_t bdesc = umma_desc(sa + (uint32_t)a_bytes, p.bk_bytes);
                    if (elected) {
                        for (int k = 0; k < kper; ++k)
                            umma<KIND>(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                       (uint32_t)((kit | k) != 0));
                        umma_commit(bar_empty + 8 * s, all_mask);
                    }
                    if (++s == p.nstages) { s = 0; ph ^= 1u; }
                }
                if (elected) umma_commit(bar_tfull + 8 * buf, pair_mask);
                __syncwarp();
                PROF_T(m2);
                PROF_ADD(9, m2 - m1);
                PROF_ADD(10, 1);
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        // 8 epilogue warps: warps w and w+4 share TMEM lane quadrant (w & 3) and split the tile's columns in halves, so the
        // epilogue keeps up with short-K mainloops (1x1 convs with K = C were epilogue-bound with 4 warps)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 128;  // 0..255
        const int half_cols = p.bn >= 64 ? p.bn / 2 : p.bn;
        const int c_beg = ((warp - 4) >> 2) * half_cols;
        const int c_end = p.bn >= 64 ? c_beg + half_cols : (warp < 8 ? p.bn : 0);
        const WsEpi& e = p.epi;
        float* spar = reinterpret_cast<float*>(smem_raw + (s_par - smem_u32(smem_raw)));
        int j = 0, last_n0 = -1;
        for (int tile = cid; tile < p.num_tiles; tile += ncl, ++j) {
            int tt = (tile / p.tiles_n) * CL + (int)rank;
            const int n0 = (tile % p.tiles_n) * p.bn;
            const int t0 = (tt % p.tiles_t) << p.bt_log2; tt /= p.tiles_t;
            const int f0 = (tt % p.tiles_f) << p.bf_log2; tt /= p.tiles_f;
            const int b0 = tt << p.bb_log2;
            const int buf = j & 1;
            // staging buffers are free once the previous tile's TMA stores have finished reading them
            PROF_T(e0);
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            PROF_T(e1);
            if (n0 != last_n0) {
                epi_stage_params(p, spar, n0, et);
                last_n0 = n0;
            }
            epi_bar_sync();
            PROF_T(e2);
            // row -> output position (for the per-row epilogue inputs that are not tile-shaped)
            const int t = t0 + (r & ((1 << p.bt_log2) - 1));
            const int f = f0 + ((r >> p.bt_log2) & ((1 << p.bf_log2) - 1));
            const int b = b0 + (r >> (p.bt_log2 + p.bf_log2));
            const bool valid = (t < p.T) && (f < p.F) && (b < p.B);
            const long long pos = ((long long)b * p.F + f) * p.T + t;
            int eb = 0, etm = 0;
            if (valid && (e.rowbias != nullptr || e.gate != nullptr)) {
                eb = (int)(pos / e.FT);
                etm = (int)(pos % e.T);
            }
            mbar_wait(bar_tfull + 8 * buf, ((uint32_t)j >> 1) & 1u);
            PROF_T(e3);
            tc_fence_after();
            if (p.has_epin) mbar_wait(bar_ifull, (uint32_t)j & 1u);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.bn);
            // (the two warps of a lane quadrant interleave on their scheduler: one computes while the other waits on LDTM)
            const EpiTile xt{spar, stg_in, stg_out, tile_out_bytes, panel_bytes, panel_cols, n0, r, valid, eb, etm};
#pragma unroll 1
            for (int c = c_beg; c < c_end; c += 32) {
                uint32_t raw[32];
                tmem_ld32(trow + (uint32_t)c, raw);
                tmem_ld_wait();
                epi_chunk<LEAN>(p, xt, raw, c);
            }
            // accumulator buffer (and the epilogue-input tile) are drained: hand them back
            PROF_T(e4);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_remote(bar_tempty + 8 * buf, rank & ~1u);   // the MMA issuer lives in the pair's leader CTA
                if (p.has_epin) mbar_arrive(bar_iempty);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            epi_bar_sync();
            if (et == 0) {
                for (int o = 0; o < p.nout; ++o)
                    for (int pn = 0; pn < npanels; ++pn)
                        tma_store_4d(&p.omap[o], stg_out + (uint32_t)(o * tile_out_bytes + pn * 128 * panel_bytes),
                                     n0 + pn * panel_cols, t0, f0, b0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#ifdef WS_TC3_PROFILE
                PROF_T(e5);
                PROF_ADD(0, e1 - e0); PROF_ADD(1, e2 - e1); PROF_ADD(2, e3 - e2); PROF_ADD(3, e4 - e3); PROF_ADD(4, e5 - e4);
                PROF_ADD(5, 1);
#endif
            }
            if constexpr (LEAN == WS_BF16 + 1 || LEAN == WS_F16 + 1) epi_colsum<LEAN>(p, stg_out, et, n0, t0, b0);
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer can still touch its smem / TMEM / barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

#ifdef WS_TC3_PROFILE
extern "C" void ws_tc3_prof_read(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
}
#endif

namespace {
template <int KIND, int LEAN, int CL>
inline cudaError_t tc3_attr1() {
    return cudaFuncSetAttribute(ws_conv_gemm_tc3_kernel<KIND, LEAN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem3);
}
template <int KIND, int LEAN>
inline cudaError_t tc3_attr() {
    cudaError_t e = tc3_attr1<KIND, LEAN, 2>();
    return e == cudaSuccess ? tc3_attr1<KIND, LEAN, 4>() : e;
}
template <int KIND, int LEAN>
inline void tc3_go(const WsTc2Params* p, cudaStream_t s) {
    if (p->cl == 4) ws_conv_gemm_tc3_kernel<KIND, LEAN, 4><<<p->grid, kThreads3, p->smem_bytes, s>>>(*p);
    else ws_conv_gemm_tc3_kernel<KIND, LEAN, 2><<<p->grid, kThreads3, p->smem_bytes, s>>>(*p);
}
}  // namespace

extern "C" const char* ws_tc3_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = tc3_attr<0, 0>();
    if (e == cudaSuccess) e = tc3_attr<1, 0>();
    if (e == cudaSuccess) e = tc3_attr<0, WS_F32 + 1>();
    if (e == cudaSuccess) e = tc3_attr<1, WS_BF16 + 1>();
    if (e == cudaSuccess) e = tc3_attr<1, WS_F16 + 1>();
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" int ws_tc3_max_smem(void) { return kMaxDynSmem3; }

// how many clusters of `cl` CTAs (at the kernel's maximum shared memory) the device keeps resident at once: GPCs whose SM
// count is not a multiple of cl leave SMs idle, so this can be below num_sms / cl
extern "C" int ws_tc3_max_clusters(int cl) {
    static int cache[64][2];
    static bool have[64][2];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || (cl != 2 && cl != 4)) return 0;
    const int k = cl == 4;
    if (have[dev][k]) return cache[dev][k];
    if (ws_tc3_init() != nullptr) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ws_num_sms() / cl * cl));
    cfg.blockDim = dim3(kThreads3);
    cfg.dynamicSmemBytes = kMaxDynSmem3;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = (unsigned)cl; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cl == 4 ? cudaOccupancyMaxActiveClusters(&n, ws_conv_gemm_tc3_kernel<1, WS_BF16 + 1, 4>, &cfg)
                            : cudaOccupancyMaxActiveClusters(&n, ws_conv_gemm_tc3_kernel<1, WS_BF16 + 1, 2>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    cache[dev][k] = n; have[dev][k] = true;
    return n;
}

extern "C" const char* ws_tc3_launch(const WsTc2Params* p, cudaStream_t s) {
    const int lean = p->epi_generic ? 0 : ws_tc_lean_kind(p);
    if (const char* m = ws_tc_colsum_check(p, lean)) return m;
    if (p->cl != 2 && p->cl != 4) return "ws_tc3_launch: cluster size must be 2 or 4";
    if (lean == WS_F32 + 1) tc3_go<0, WS_F32 + 1>(p, s);
    else if (lean == WS_BF16 + 1) tc3_go<1, WS_BF16 + 1>(p, s);
    else if (lean == WS_F16 + 1) tc3_go<1, WS_F16 + 1>(p, s);
    else if (p->kind == 0) tc3_go<0, 0>(p, s);
    else tc3_go<1, 0>(p, s);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
