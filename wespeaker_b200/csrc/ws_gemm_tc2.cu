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
#include "ws_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int bk_bytes) {
    const uint64_t layout = bk_bytes == 128 ? 2ull : (bk_bytes == 64 ? 4ull : 6ull);
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * bk_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte chunk index after the TMA/UMMA swizzle for a row of `row_bytes` (128/64/32) inside a 1024-B aligned panel
__device__ __forceinline__ int swz_chunk(int chunk, int row, int row_bytes) {
    return row_bytes == 128 ? (chunk ^ (row & 7)) : (row_bytes == 64 ? (chunk ^ ((row >> 1) & 3)) : (chunk ^ ((row >> 2) & 1)));
}

constexpr int kThreads2 = 384;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4..w11 epilogue (2 warps per TMEM lane quadrant)
constexpr int kMaxDynSmem2 = 220 * 1024;

// 32 consecutive fp32 values -> packed activation dtype -> swizzled staging panel row
__device__ __forceinline__ void stage_store32(uint32_t panel_base, int row, int row_bytes, int col_in_panel, int dt,
                                              const float* v) {
    const uint32_t rbase = panel_base + (uint32_t)(row * row_bytes);
    if (dt == WS_F32) {
        const int c0 = (col_in_panel * 4) >> 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t a = rbase + (uint32_t)(swz_chunk(c0 + i, row, row_bytes) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                         "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                         : "memory");
        }
    } else {
        const int c0 = (col_in_panel * 2) >> 4;
        const bool bf = dt == WS_BF16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                w[j] = bf ? ws_pack2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1], WS_BF16)
                          : ws_pack2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1], WS_F16);
            const uint32_t a = rbase + (uint32_t)(swz_chunk(c0 + i, row, row_bytes) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                         : "memory");
        }
    }
}
// inverse: read 32 consecutive values of a TMA-loaded swizzled panel row
__device__ __forceinline__ void stage_load32(uint32_t panel_base, int row, int row_bytes, int col_in_panel, int dt,
                                             float* v) {
    const uint32_t rbase = panel_base + (uint32_t)(row * row_bytes);
    if (dt == WS_F32) {
        const int c0 = (col_in_panel * 4) >> 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t a = rbase + (uint32_t)(swz_chunk(c0 + i, row, row_bytes) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                         : "r"(a));
        }
    } else {
        const int c0 = (col_in_panel * 2) >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
            const uint32_t a = rbase + (uint32_t)(swz_chunk(c0 + i, row, row_bytes) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * i + 2 * j] = ws_16_to_f(w[j] & 0xffffu, dt);
                v[8 * i + 2 * j + 1] = ws_16_to_f(w[j] >> 16, dt);
            }
        }
    }
}

// LEAN != 0 (= WsDType of the activations + 1) compiles only the common epilogue: bias, ReLU, BN scale/shift, residual
// tile, ReLU, one output (+ its 3xTF32 low twin for fp32).  The generic epilogue (row bias, gates, Res2 second output,
// tanh/sigmoid, every dtype) is ~100 KB of SASS of which a given layer executes a few KB scattered between never-taken
// branches; ncu (2-CTA twin, ws_gemm_tc3.cu) showed its warps stalled on instruction fetch (stall_no_inst) for half of their samples, which made the
// short-K 1x1 convs epilogue-bound.  The lean chunk body is ~4 KB and stays in the 6 KB L0 instruction cache.
template <int KIND, int LEAN>
__global__ void __launch_bounds__(kThreads2, 1) ws_conv_gemm_tc2_kernel(const __grid_constant__ WsTc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_TC_MAX_STAGES + 6];
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        if (lane == 0) {
            const int kper = p.bk_bytes / 32;
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++j) {
                const int buf = j & 1;
                mbar_wait(bar_tempty + 8 * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);  // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * p.bn);
                for (int kit = 0; kit < p.nk_total * p.nsplit; ++kit, ++it) {
                    const int s = it % p.nstages;
                    const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = ring + (uint32_t)(s * stage_bytes);
                    uint64_t adesc = umma_desc(sa, p.bk_bytes);
                    if (p.dbg_shift >= 0)  // row-shifted operand read: start one row in, base_offset field [49,52)
                        adesc = umma_desc(sa + (uint32_t)p.bk_bytes, p.bk_bytes) | ((uint64_t)(p.dbg_shift & 7) << 49);
                    const uint64_t bdesc = umma_desc(sa + (uint32_t)a_bytes, p.bk_bytes);
                    for (int k = 0; k < kper; ++k)
                        umma<KIND>(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                   (uint32_t)((kit | k) != 0));
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_tfull + 8 * buf);
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
                for (int c = et; c < p.bn; c += 256) {
                    // all three loads are issued before the first store: interleaved load/store pairs were serialised by
                    // the compiler (possible aliasing) and cost three L2 round trips per tile on the epilogue's critical path
                    const float pb = e.bias ? __ldg(e.bias + n0 + c) : 0.f;
                    const float ps = e.scale ? __ldg(e.scale + n0 + c) : 1.f;
                    const float ph = e.scale ? __ldg(e.shift + n0 + c) : 0.f;
                    spar[c] = pb;
                    spar[p.bn + c] = ps;
                    spar[2 * p.bn + c] = ph;
                }
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
#pragma unroll 1
            for (int c = c_beg; c < c_end; c += 32) {
                uint32_t raw[32];
                tmem_ld32(trow + (uint32_t)c, raw);
                tmem_ld_wait();
                float v[32];
                if constexpr (LEAN != 0) {
                    constexpr int DT = LEAN - 1;                          // activation dtype of this instantiation
                    constexpr int PCOLS = DT == WS_F32 ? 32 : 64;         // columns of one 128-byte staging panel
                    const float4* sb = reinterpret_cast<const float4*>(spar + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x = sb[i];
                        v[4 * i] = __uint_as_float(raw[4 * i]) + x.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + x.y;
                        v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + x.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + x.w;
                    }
                    if (e.act1 == WS_ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (e.scale != nullptr) {
                        const float4* ss = reinterpret_cast<const float4*>(spar + p.bn + c);
                        const float4* sh = reinterpret_cast<const float4*>(spar + 2 * p.bn + c);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 a = ss[i], d = sh[i];
                            v[4 * i] = fmaf(v[4 * i], a.x, d.x); v[4 * i + 1] = fmaf(v[4 * i + 1], a.y, d.y);
                            v[4 * i + 2] = fmaf(v[4 * i + 2], a.z, d.z); v[4 * i + 3] = fmaf(v[4 * i + 3], a.w, d.w);
                        }
                    }
                    const uint32_t poff = (uint32_t)((c / PCOLS) * 128 * 128);
                    if (p.has_epin) {  // residual
                        float rin[32];
                        stage_load32(stg_in + poff, r, 128, c % PCOLS, DT, rin);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += rin[i];
                    }
                    if (e.act2 == WS_ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    stage_store32(stg_out + poff, r, 128, c % PCOLS, DT, v);
                    if constexpr (DT == WS_F32) {
                        if (p.nsplit == 3) {  // 3xTF32: the low twin of the output feeds the next layer's x_lo pass
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = ws_tf32_lo(v[i]);
                            stage_store32(stg_out + (uint32_t)tile_out_bytes + poff, r, 128, c % PCOLS, DT, v);
                        }
                    }
                    continue;
                }
                {
                    const float4* sb = reinterpret_cast<const float4*>(spar + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x = sb[i];
                        v[4 * i] = __uint_as_float(raw[4 * i]) + x.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + x.y;
                        v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + x.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + x.w;
                    }
                }
                if (e.rowbias != nullptr && valid) {
                    const float4* rb = reinterpret_cast<const float4*>(e.rowbias + (long long)eb * e.rowbias_ld + n0 + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x = __ldg(rb + i);
                        v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
                    }
                }
                ws_act_vec<32>(v, e.act1);
                if (e.scale != nullptr) {
                    const float4* ss = reinterpret_cast<const float4*>(spar + p.bn + c);
                    const float4* sh = reinterpret_cast<const float4*>(spar + 2 * p.bn + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a = ss[i], d = sh[i];
                        v[4 * i] = fmaf(v[4 * i], a.x, d.x); v[4 * i + 1] = fmaf(v[4 * i + 1], a.y, d.y);
                        v[4 * i + 2] = fmaf(v[4 * i + 2], a.z, d.z); v[4 * i + 3] = fmaf(v[4 * i + 3], a.w, d.w);
                    }
                }
                if (e.gate != nullptr && valid) {
                    const float4* g = reinterpret_cast<const float4*>(
                        e.gate + ((long long)eb * e.gate_nseg + etm / e.gate_seg) * e.gate_ld + n0 + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x = __ldg(g + i);
                        v[4 * i] *= x.x; v[4 * i + 1] *= x.y; v[4 * i + 2] *= x.z; v[4 * i + 3] *= x.w;
                    }
                }
                const int pn = c / panel_cols, cin = c % panel_cols;
                float rin[32];
                if (p.has_epin) stage_load32(stg_in + (uint32_t)(pn * 128 * panel_bytes), r, panel_bytes, cin, e.dtype, rin);
                if (p.has_epin && !p.has_out2) {  // residual
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += rin[i];
                }
                ws_act_vec<32>(v, e.act2);
                const uint32_t poff = (uint32_t)(pn * 128 * panel_bytes);
                int ob = 0;
                stage_store32(stg_out + poff, r, panel_bytes, cin, e.dtype, v);
                if (p.has_out2) {  // Res2: next conv's input = this output + the next channel group
#pragma unroll
                    for (int i = 0; i < 32; ++i) rin[i] += v[i];
                }
                if (p.nsplit == 3) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = ws_tf32_lo(v[i]);
                    stage_store32(stg_out + (uint32_t)(++ob * tile_out_bytes) + poff, r, panel_bytes, cin, e.dtype, v);
                }
                if (p.has_out2) {
                    stage_store32(stg_out + (uint32_t)(++ob * tile_out_bytes) + poff, r, panel_bytes, cin, e.dtype, rin);
                    if (p.nsplit == 3) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) rin[i] = ws_tf32_lo(rin[i]);
                        stage_store32(stg_out + (uint32_t)(++ob * tile_out_bytes) + poff, r, panel_bytes, cin, e.dtype, rin);
                    }
                }
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
            if constexpr (LEAN == WS_BF16 + 1 || LEAN == WS_F16 + 1) {
                // fused SE squeeze (WsEpi::colsum): column sums of the staged (already rounded) tile, read back from the
                // swizzled staging panels while the TMA store drains them; thread = (column pair, 64-row half)
                if (e.colsum != nullptr && b0 < p.B && t0 + (et >> 7) * 64 < p.T) {   // units past the last position do not exist
                    const int cpair = et & 127, rh = et >> 7;
                    const int pos0 = t0 + rh * 64;
                    const int nvalid = min(64, p.T - pos0);
                    const int split = min(nvalid, (pos0 / e.colsum_T + 1) * e.colsum_T - pos0);
                    const int col = 2 * cpair;
                    const uint32_t cbase = stg_out + (uint32_t)((col >> 6) * 128 * 128 + (col & 7) * 2);
                    const int chunk = (col & 63) >> 3;
                    float a0 = 0.f, a1 = 0.f, b0s = 0.f, b1s = 0.f;
#pragma unroll 4
                    for (int rr = 0; rr < split; ++rr) {
                        const int row = rh * 64 + rr;
                        uint32_t w;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(cbase + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4))));
                        a0 += ws_16_to_f(w & 0xffffu, LEAN - 1); a1 += ws_16_to_f(w >> 16, LEAN - 1);
                    }
#pragma unroll 4
                    for (int rr = max(split, 0); rr < nvalid; ++rr) {
                        const int row = rh * 64 + rr;
                        uint32_t w;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(cbase + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4))));
                        b0s += ws_16_to_f(w & 0xffffu, LEAN - 1); b1s += ws_16_to_f(w >> 16, LEAN - 1);
                    }
                    const long long unit = (long long)(t0 >> 6) + rh;
                    float* dst = e.colsum + unit * 2 * (long long)(p.tiles_n * p.bn) + n0 + col;
                    *reinterpret_cast<float2*>(dst) = make_float2(a0, a1);
                    *reinterpret_cast<float2*>(dst + (long long)(p.tiles_n * p.bn)) = make_float2(b0s, b1s);
                }
            }
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
// 0 = generic epilogue, else activation dtype + 1 (see the LEAN template parameter)
inline int tc2_lean(const WsTc2Params* p) {
    const WsEpi& e = p->epi;
    const bool common = !p->has_out2 && e.rowbias == nullptr && e.gate == nullptr &&
                        (e.act1 == WS_ACT_NONE || e.act1 == WS_ACT_RELU) &&
                        (e.act2 == WS_ACT_NONE || e.act2 == WS_ACT_RELU) && p->panel_bytes == 128;
    if (!common) return 0;
    if (p->kind == 1 && p->nsplit == 1 && p->nout == 1 && p->bn >= 64 && (e.dtype == WS_BF16 || e.dtype == WS_F16))
        return e.dtype + 1;
    if (p->kind == 0 && e.dtype == WS_F32 && p->bn >= 32 &&
        ((p->nsplit == 1 && p->nout == 1) || (p->nsplit == 3 && p->nout == 2)))
        return WS_F32 + 1;
    return 0;
}
template <int KIND, int LEAN>
inline cudaError_t tc2_attr() {
    return cudaFuncSetAttribute(ws_conv_gemm_tc2_kernel<KIND, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem2);
}
}  // namespace

extern "C" const char* ws_tc2_init(void) {
    static bool done = false;
    if (done) return nullptr;
    cudaError_t e = tc2_attr<0, 0>();
    if (e == cudaSuccess) e = tc2_attr<1, 0>();
    if (e == cudaSuccess) e = tc2_attr<0, WS_F32 + 1>();
    if (e == cudaSuccess) e = tc2_attr<1, WS_BF16 + 1>();
    if (e == cudaSuccess) e = tc2_attr<1, WS_F16 + 1>();
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    done = true;
    return nullptr;
}

extern "C" int ws_tc2_max_smem(void) { return kMaxDynSmem2; }

extern "C" const char* ws_tc2_launch(const WsTc2Params* p, cudaStream_t s) {
    const int lean = p->epi_generic ? 0 : tc2_lean(p);
    if (p->epi.colsum != nullptr && (lean != WS_BF16 + 1 && lean != WS_F16 + 1 || p->tiles_f != 1 || p->tiles_b != 1 ||
                                     p->bt_log2 != 7 || p->epi.colsum_T < 128))
        return "conv colsum: needs the lean 16-bit epilogue on a dense 1x1 conv (flat positions) and >= 128 frames per utterance";
    if (lean == WS_F32 + 1) ws_conv_gemm_tc2_kernel<0, WS_F32 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (lean == WS_BF16 + 1) ws_conv_gemm_tc2_kernel<1, WS_BF16 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (lean == WS_F16 + 1) ws_conv_gemm_tc2_kernel<1, WS_F16 + 1><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else if (p->kind == 0) ws_conv_gemm_tc2_kernel<0, 0><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    else ws_conv_gemm_tc2_kernel<1, 0><<<p->grid, kThreads2, p->smem_bytes, s>>>(*p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
