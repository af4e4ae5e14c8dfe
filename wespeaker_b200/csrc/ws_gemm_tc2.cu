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

namespace {
using namespace ws_tcdev;

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
constexpr int kThreads2 = 384;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4..w11 epilogue (2 warps per TMEM lane quadrant)
constexpr int kMaxDynSmem2 = 220 * 1024;

template <int KIND, int LEAN>
__global__ void __launch_bounds__(kThreads2, 1) ws_conv_gemm_tc2_kernel(const __grid_constant__ WsTc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_TC_MAX_STAGES + 6];
    __shared__ uint32_t s_tmem;

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int a_bytes = 128 * p.bk_bytes, b_bytes = p.bn * p.bk_bytes;
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
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8);  // one arrive per epilogue warp
        }
        mbar_init(bar_ifull, 1);
        mbar_init(bar_iempty, 8);
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
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            const int bk_elems = p.bk_bytes / es;
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
                int tt = tile;
                const int n0 = (tt % p.tiles_n) * p.bn; tt /= p.tiles_n;
                const int t0 = (tt % p.tiles_t) << p.bt_log2; tt /= p.tiles_t;
                const int f0 = (tt % p.tiles_f) << p.bf_log2; tt /= p.tiles_f;
                const int b0 = tt << p.bb_log2;
                for (int tp = 0; tp < p.ntaps; ++tp) {
                    const WsTcTap tap = p.taps[tp];
                    for (int kb = 0; kb < tap.nkb; ++kb) {
                        // 3xTF32: three passes per k-block, small terms first: x_lo*W, x*W_lo, x*W (fp32 operands are
                        // truncated to tf32 by the tensor core, so x and W themselves serve as the high parts)
                        for (int ps = 3 - p.nsplit; ps < 3; ++ps, ++it) {
                            const int s = it % p.nstages;
                            const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                            mbar_expect_tx(bar_full + 8 * s, (uint32_t)stage_bytes);
                            const uint32_t sa = ring + (uint32_t)(s * stage_bytes);
                            tma_load_4d(sa, ps == 0 ? &p.amap_lo[tap.map] : &p.amap[tap.map], bar_full + 8 * s,
                                        tap.c0 + kb * bk_elems, t0 + tap.dt - (p.dbg_shift >= 0 ? 1 : 0), f0 + tap.df, b0);
                            tma_load_2d(sa + (uint32_t)a_bytes, ps == 1 ? &p.wmap_lo : &p.wmap, bar_full + 8 * s,
                                        tap.wk + kb * bk_elems, n0);
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
        // ================================ MMA issuer ================================
        // all 32 lanes run the loop with warp-uniform values; only the tcgen05 instructions are predicated on the elected
        // lane (see elect_one() in ws_tc_common.cuh)
        {
            const uint32_t elected = elect_one();
            const int kper = p.bk_bytes / 32;
            int s = 0, j = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
                const int buf = j & 1;
                mbar_wait(bar_tempty + 8 * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * p.bn);
                const int nkit = p.nk_total * p.nsplit;
                for (int kit = 0; kit < nkit; ++kit) {
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = ring + (uint32_t)(s * stage_bytes);
                    uint64_t adesc = umma_desc(sa, p.bk_bytes);
                    if (p.dbg_shift >= 0)  // row-shifted operand read: start one row in, base_offset field [49,52)
                        adesc = umma_desc(sa + (uint32_t)p.bk_bytes, p.bk_bytes) | ((uint64_t)(p.dbg_shift & 7) << 49);
                    const uint64_t bdesc = umma_desc(sa + (uint32_t)a_bytes, p.bk_bytes);
                    if (elected) {
                        for (int k = 0; k < kper; ++k)
                            umma<KIND>(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                       (uint32_t)((kit | k) != 0));
                        umma_commit(bar_empty + 8 * s);
                    }
                    if (++s == p.nstages) { s = 0; ph ^= 1u; }
                }
                if (elected) umma_commit(bar_tfull + 8 * buf);
                __syncwarp();
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
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
            int tt = tile;
            const int n0 = (tt % p.tiles_n) * p.bn; tt /= p.tiles_n;
            const int t0 = (tt % p.tiles_t) << p.bt_log2; tt /= p.tiles_t;
            const int f0 = (tt % p.tiles_f) << p.bf_log2; tt /= p.tiles_f;
            const int b0 = tt << p.bb_log2;
            const int buf = j & 1;
            // staging buffers are free once the previous tile's TMA stores have finished reading them
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (n0 != last_n0) {
                epi_stage_params(p, spar, n0, et);
                last_n0 = n0;
            }
            epi_bar_sync();
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_tempty + 8 * buf);
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
            }
            if constexpr (LEAN == WS_BF16 + 1 || LEAN == WS_F16 + 1) epi_colsum<LEAN>(p, stg_out, et, n0, t0, b0);
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

namespace {
template <int KIND, int LEAN>
inline cudaError_t tc2_attr() {
    return cudaFuncSetAttribute(ws_conv_gemm_tc2_kernel<KIND, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem2);
}
}  // namespace

extern "C" const char* ws_tc2_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = tc2_attr<0, 0>();
    if (e == cudaSuccess) e = tc2_attr<1, 0>();
    if (e == cudaSuccess) e = tc2_attr<0, WS_F32 + 1>();
    if (e == cudaSuccess) e = tc2_attr<1, WS_BF16 + 1>();
    if (e == cudaSuccess) e = tc2_attr<1, WS_F16 + 1>();
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" int ws_tc2_max_smem(void) { return kMaxDynSmem2; }

extern "C" const char* ws_tc2_launch(const WsTc2Params* p, cudaStream_t s) {
    const int lean = p->epi_generic ? 0 : ws_tc_lean_kind(p);
    if (const char* m = ws_tc_colsum_check(p, lean)) return m;
    if (lean == WS_F32 + 1) ws_conv_gemm_tc2_kernel<0, WS_F32 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (lean == WS_BF16 + 1) ws_conv_gemm_tc2_kernel<1, WS_BF16 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (lean == WS_F16 + 1) ws_conv_gemm_tc2_kernel<1, WS_F16 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (p->kind == 0) ws_conv_gemm_tc2_kernel<0, 0><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else ws_conv_gemm_tc2_kernel<1, 0><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
