// Fused Res2Conv1dReluBn chain (wespeaker/models/ecapa_tdnn.py:29-78) for 16-bit activations on tcgen05.
//
// The reference runs 7 *dependent* dilated k=3 convs on w-channel groups:  sp_i = BN_i(ReLU(conv_i(sp_{i-1} + x_i))).
// As separate launches each is latency-bound (27 us for 5 GFLOP on B200).  Here ONE persistent CTA owns a whole
// utterance and walks the chain on-chip:
//   * the conv input s_i = sp_{i-1} + x_i lives in shared memory as a K-major 128B-swizzled operand buffer with zero rows
//     around it; the three dilated taps are three tcgen05.mma operand reads of the SAME buffer at row offsets -d, 0, +d
//     (the UMMA descriptor start address is shifted by whole 128-B rows; the swizzle is a function of the absolute
//     smem address, verified on B200 with tools/shift_probe.py), so there is no im2col, no halo exchange, and the
//     intermediate never goes to HBM;
//   * weights stream through a TMA ring (they are L2-resident, shared by all CTAs); accumulators sit in TMEM;
//   * the epilogue (bias, ReLU, BN affine) writes sp_i to the output buffer through a swizzled staging tile + TMA store
//     and writes s_{i+1} = sp_i + x_{i+1} straight back into the operand buffer for the next conv.
//   * utterances longer than 256 frames are walked in time tiles of 256 rows that overlap by 2 x 32: the chain's receptive
//     field is 7 x dilation <= 28 rows per side, so the inner 192 rows of a tile are exact and only those are stored (three
//     64-row TMA boxes per channel panel); rows before t = 0 and behind the utterance's end are forced to zero in the operand
//     buffer - they are the convs' zero padding.
// Warp roles: w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4.. epilogue.
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int EW>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory"); }

// K-major SWIZZLE_128B operand descriptor (rows 128 B apart, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
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

constexpr int kPadRows = 8;      // zero rows before t=0 (>= max dilation, multiple of 8 keeps the swizzle phase = t & 7)
constexpr int kSRows = 256 + 16; // operand buffer rows per K panel
constexpr int kWStages = 4;
constexpr int kNumConv = 7;

// DT = activation dtype (compile-time so the epilogue carries one conversion path); 384 threads: w0 producer, w1 MMA,
// w2 TMEM alloc, w4..w11 epilogue (warps w and w+4 share a TMEM lane quadrant and split the w8 columns).  The epilogue
// sits on the serial conv_i -> conv_{i+1} chain, so its latency (not throughput) is what the kernel time is made of.
// EW = epilogue warps: 8 (384 threads, one CTA per SM: the 128-wide groups of ECAPA-1024 fill the shared memory), or 4 (256
// threads at <= 128 registers, ~100 KB of shared memory: TWO CTAs per SM for the 64-wide groups of ECAPA-512, so that the MMAs
// of one utterance run under the epilogue of another - inside one chain the two strictly alternate).
template <int DT, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, EW == 4 ? 2 : 1) ws_res2_fused_kernel(const __grid_constant__ WsRes2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * kWStages + 5];
    __shared__ __align__(16) float s_par[3][128];
    __shared__ uint32_t s_tmem;
    // warp index through a shuffle: role branches are warp-uniform for the compiler (see the MMA issuer below)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int npan = p.w8 >> 6;                       // K panels of 64 channels (128 B) per tap
    const int nmt = p.ntile > 1 ? 2 : (p.T + 127) >> 7;   // M tiles (128 rows) per work unit
    // work unit = (utterance, time tile); tile k covers rows [k * tstride - halo, +256); single-tile utterances: halo 0
    const int halo = p.ntile > 1 ? 32 : 0, tstride = 256 - 2 * halo, nunits = p.B * p.ntile;
    const int wblk_bytes = p.w8 * 128;                // one weight k-block: w8 output rows x 128 B
    const uint32_t sS = base;                                             // [npan][kSRows][128 B]
    const uint32_t sW = sS + (uint32_t)(npan * kSRows * 128);             // [kWStages][wblk_bytes]
    const uint32_t sO = sW + (uint32_t)(kWStages * wblk_bytes);           // XO: [npan][256 rows][128 B]: x_{i+1} in, sp_i out
    const uint32_t bar_wfull = smem_u32(&s_bar[0]);
    const uint32_t bar_wempty = smem_u32(&s_bar[kWStages]);
    const uint32_t bar_x0 = smem_u32(&s_bar[2 * kWStages]);
    const uint32_t bar_sready = smem_u32(&s_bar[2 * kWStages + 1]);
    const uint32_t bar_acc = smem_u32(&s_bar[2 * kWStages + 2]);
    const uint32_t bar_sfree = smem_u32(&s_bar[2 * kWStages + 3]);
    const uint32_t bar_xn = smem_u32(&s_bar[2 * kWStages + 4]);   // x_{i+1} tile landed in the XO buffer
    const uint32_t tmem_cols = (uint32_t)(nmt * p.w8 < 32 ? 32 : (nmt * p.w8 <= 64 ? 64 : (nmt * p.w8 <= 128 ? 128 : 256)));

    // zero the whole operand buffer once: the pad rows (t < 0, t >= 128*nmt) are never written again
    {
        const int n16 = npan * kSRows * 128 / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sS + (uint32_t)(i * 16)), "r"(0u) : "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.xmap); prefetch_tmap(&p.wmap); prefetch_tmap(&p.omap); prefetch_tmap(&p.omap64);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWStages; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        mbar_init(bar_x0, 1);
        mbar_init(bar_sready, EW);
        mbar_init(bar_acc, 1);
        mbar_init(bar_sfree, 1);
        mbar_init(bar_xn, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // zeros visible to the TMA / MMA proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ================================ producer: x_0 tiles and the weight ring ================================
        if (lane == 0) {
            int wit = 0, u = 0;
            for (int un = blockIdx.x; un < nunits; un += gridDim.x, ++u) {
                const int b = un / p.ntile, t0 = (un % p.ntile) * tstride - halo;
                mbar_wait(bar_sfree, ((uint32_t)u & 1u) ^ 1u);   // previous unit's last MMA no longer reads S
                mbar_expect_tx(bar_x0, (uint32_t)(npan * nmt * 128 * 128));
                for (int kp = 0; kp < npan; ++kp)
                    for (int mt = 0; mt < nmt; ++mt)
                        tma_load_3d(sS + (uint32_t)((kp * kSRows + kPadRows + mt * 128) * 128), &p.xmap, bar_x0, kp * 64,
                                    t0 + mt * 128, b);
                for (int i = 0; i < kNumConv; ++i)
                    for (int tap = 0; tap < 3; ++tap)
                        for (int kp = 0; kp < npan; ++kp, ++wit) {
                            const int s = wit % kWStages;
                            mbar_wait(bar_wempty + 8 * s, (((uint32_t)(wit / kWStages)) & 1u) ^ 1u);
                            mbar_expect_tx(bar_wfull + 8 * s, (uint32_t)wblk_bytes);
                            tma_load_2d(sW + (uint32_t)(s * wblk_bytes), &p.wmap, bar_wfull + 8 * s, tap * p.w8 + kp * 64,
                                        i * p.w8);
                        }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // All 32 lanes run the loop with warp-uniform values and only the tcgen05 instructions are predicated on one elected
        // lane: inside an `if (lane == 0)` region every UTCHMMA is preceded by an ELECT / R2UR.BROADCAST convergence loop
        // (~20 SASS instructions, 65-80 cycles per MMA: tools/experimental/mma_rate_probe.cu), longer than these N = 64 / 128
        // MMAs occupy the tensor pipe.
        {
            uint32_t elected;
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
            int ws = 0, u = 0, g = 0;
            uint32_t wph = 0;
            for (int un = blockIdx.x; un < nunits; un += gridDim.x, ++u) {
                for (int i = 0; i < kNumConv; ++i, ++g) {
                    if (g > 0) mbar_wait(bar_sready, ((uint32_t)(g - 1)) & 1u);  // s_i written, TMEM drained
                    if (i == 0) mbar_wait(bar_x0, (uint32_t)u & 1u);
                    tc_fence_after();
                    for (int tap = 0; tap < 3; ++tap)
                        for (int kp = 0; kp < npan; ++kp) {
                            mbar_wait(bar_wfull + 8 * ws, wph);
                            tc_fence_after();
                            const uint64_t bdesc = umma_desc128(sW + (uint32_t)(ws * wblk_bytes));
                            for (int mt = 0; mt < nmt; ++mt) {
                                const int row = kPadRows + mt * 128 + (tap - 1) * p.dil;   // dilated tap = row shift
                                const uint64_t adesc = umma_desc128(sS + (uint32_t)((kp * kSRows + row) * 128));
                                const uint32_t tacc = tmem_base + (uint32_t)(mt * p.w8);
                                if (elected) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_f16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                                 (uint32_t)((tap | kp | k) != 0));
                                }
                            }
                            if (elected) umma_commit(bar_wempty + 8 * ws);
                            if (++ws == kWStages) { ws = 0; wph ^= 1u; }
                        }
                    if (elected) umma_commit(bar_acc);
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        // The XO buffer holds the next channel group x_{i+1} (TMA-loaded while the MMAs run); each thread reads its chunk,
        // overwrites it with sp_i (stored to HBM by TMA afterwards) and writes s_{i+1} = sp_i + x_{i+1} into the operand
        // buffer.  The elected thread re-arms XO with the following group as soon as the store has read it.
        const int q = warp & 3, r = q * 32 + lane, et = threadIdx.x - 128;   // et 0..255
        const int c_beg = EW == 8 ? ((warp - 4) >> 2) * (p.w8 >> 1) : 0, c_end = EW == 8 ? c_beg + (p.w8 >> 1) : p.w8;
        const uint32_t xo_bytes = (uint32_t)(npan * nmt * 128 * 128);
        auto load_xn = [&](int un_, int grp) {
            const int bb = un_ / p.ntile, tt0 = (un_ % p.ntile) * tstride - halo;
            mbar_expect_tx(bar_xn, xo_bytes);
            for (int kp = 0; kp < npan; ++kp)
                for (int mt = 0; mt < nmt; ++mt)
                    tma_load_3d(sO + (uint32_t)((kp * 256 + mt * 128) * 128), &p.xmap, bar_xn, grp * p.w8 + kp * 64, tt0 + mt * 128, bb);
        };
        if (et == 0 && (int)blockIdx.x < nunits) load_xn(blockIdx.x, 1);
        int g = 0, xc = 0;
        for (int un = blockIdx.x; un < nunits; un += gridDim.x) {
            const int b = un / p.ntile, t0 = (un % p.ntile) * tstride - halo;
            const int Tb = p.lens != nullptr ? min(p.T, p.lens[b]) : p.T;   // rows t >= Tb (and t < 0) are conv padding: stay zero everywhere
            for (int i = 0; i < kNumConv; ++i, ++g) {
                if (et < p.w8) {
                    s_par[0][et] = p.bias[i * p.w8 + et];
                    s_par[1][et] = p.scale[i * p.w8 + et];
                    s_par[2][et] = p.shift[i * p.w8 + et];
                }
                epi_bar_sync<EW>();
                mbar_wait(bar_acc, (uint32_t)g & 1u);
                tc_fence_after();
                if (i == kNumConv - 1 && et == 0) mbar_arrive(bar_sfree);  // conv 6 done: S may take the next utterance
                const bool has_next = i < kNumConv - 1;
                if (has_next) { mbar_wait(bar_xn, (uint32_t)xc & 1u); ++xc; }
#pragma unroll 1
                for (int mt = 0; mt < nmt; ++mt) {
                    const int t = mt * 128 + r;                       // row inside the tile (addresses the buffers)
                    const bool valid = t0 + t >= 0 && t0 + t < Tb;    // frame of the utterance
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.w8);
#pragma unroll 1
                    for (int c = c_beg; c < c_end; c += 32) {
                        uint32_t raw[32];
                        tmem_ld32(trow + (uint32_t)c, raw);
                        tmem_ld_wait();
                        float v[32];
                        const float4* pb = reinterpret_cast<const float4*>(&s_par[0][c]);
                        const float4* ps = reinterpret_cast<const float4*>(&s_par[1][c]);
                        const float4* ph = reinterpret_cast<const float4*>(&s_par[2][c]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = pb[j], s4 = ps[j], h4 = ph[j];
                            v[4 * j] = fmaf(fmaxf(__uint_as_float(raw[4 * j]) + b4.x, 0.f), s4.x, h4.x);
                            v[4 * j + 1] = fmaf(fmaxf(__uint_as_float(raw[4 * j + 1]) + b4.y, 0.f), s4.y, h4.y);
                            v[4 * j + 2] = fmaf(fmaxf(__uint_as_float(raw[4 * j + 2]) + b4.z, 0.f), s4.z, h4.z);
                            v[4 * j + 3] = fmaf(fmaxf(__uint_as_float(raw[4 * j + 3]) + b4.w, 0.f), s4.w, h4.w);
                        }
                        const int pn = c >> 6, c16 = (c & 63) >> 3;      // panel, first 16-B chunk inside the 128-B row
                        const uint32_t orow = sO + (uint32_t)((pn * 256 + t) * 128);
                        const uint32_t srow = sS + (uint32_t)((pn * kSRows + kPadRows + t) * 128);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t off = (uint32_t)(((c16 + j) ^ (t & 7)) << 4);
                            uint32_t xs[4] = {0u, 0u, 0u, 0u};
                            if (has_next)
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(xs[0]), "=r"(xs[1]), "=r"(xs[2]), "=r"(xs[3]) : "r"(orow + off));
                            uint32_t w[4], o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float a0 = valid ? v[8 * j + 2 * k] : 0.f, a1 = valid ? v[8 * j + 2 * k + 1] : 0.f;
                                w[k] = ws_pack2(a0, a1, DT);
                                o[k] = ws_pack2(a0 + ws_16_to_f(xs[k] & 0xffffu, DT), a1 + ws_16_to_f(xs[k] >> 16, DT), DT);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + off), "r"(w[0]), "r"(w[1]),
                                         "r"(w[2]), "r"(w[3]) : "memory");
                            if (has_next && valid)   // rows >= T stay zero = conv padding
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + off), "r"(o[0]), "r"(o[1]),
                                             "r"(o[2]), "r"(o[3]) : "memory");
                        }
                    }
                }
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sready);
                epi_bar_sync<EW>();
                if (et == 0) {
                    for (int pn = 0; pn < npan; ++pn) {
                        if (p.ntile > 1) {   // only the exact inner rows [halo, 256 - halo) of the tile: three 64-row boxes
                            for (int q3 = 0; q3 < 3; ++q3)
                                tma_store_3d(&p.omap64, sO + (uint32_t)((pn * 256 + halo + 64 * q3) * 128), i * p.w8 + pn * 64,
                                             t0 + halo + 64 * q3, b);
                        } else {
                            for (int mt = 0; mt < nmt; ++mt)
                                tma_store_3d(&p.omap, sO + (uint32_t)((pn * 256 + mt * 128) * 128), i * p.w8 + pn * 64, mt * 128, b);
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // XO has been read by the store
                    if (i < kNumConv - 2) load_xn(un, i + 2);                         // group for the next conv's epilogue
                    else if (i == kNumConv - 1 && un + (int)gridDim.x < nunits) load_xn(un + gridDim.x, 1);
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

extern "C" const char* ws_res2_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_BF16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_F16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    // the two-CTAs-per-SM variant: ~100 KB each, and the carveout must leave room for both
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_BF16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_F16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_BF16, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_res2_fused_kernel<WS_F16, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" const char* ws_res2_launch(const WsRes2Params* p, cudaStream_t s) {
    if (p->ew == 4) {
        if (p->dtype == WS_BF16) ws_res2_fused_kernel<WS_BF16, 4><<<p->grid, 256, p->smem_bytes, s>>>(*p);
        else if (p->dtype == WS_F16) ws_res2_fused_kernel<WS_F16, 4><<<p->grid, 256, p->smem_bytes, s>>>(*p);
        else return "res2_fused: 16-bit activations only";
    } else if (p->dtype == WS_BF16) ws_res2_fused_kernel<WS_BF16, 8><<<p->grid, 384, p->smem_bytes, s>>>(*p);
    else if (p->dtype == WS_F16) ws_res2_fused_kernel<WS_F16, 8><<<p->grid, 384, p->smem_bytes, s>>>(*p);
    else return "res2_fused: 16-bit activations only";
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
