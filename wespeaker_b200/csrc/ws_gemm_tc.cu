// Conv1d/Conv2d/Linear as implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   D[128 positions x bn channels] (fp32, TMEM) = sum over taps, k-blocks  A_tap[128 x bk] * W[bn x bk]^T
//
// * A tiles (activations, channels-last) are fetched by TMA from a rank-4 tensor map (C, T, F, B) with a
//   box of (bk, 2^bt, 2^bf, 2^bb) = 128 rows; the conv tap offset is added to the T/F box coordinate and
//   out-of-bounds rows are zero-filled by TMA — that *is* the conv zero padding, so no im2col and no halo
//   logic.  W tiles come from a rank-2 map (Ktot, Cout).  Both land K-major with the 128/64/32-byte TMA
//   swizzle that the UMMA shared-memory descriptor names.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane issues tcgen05.mma, completion is
//   committed to mbarriers), warps 2..5 = epilogue (tcgen05.ld 32x32b from the TMEM lane quadrant
//   warp_idx % 4, fused bias/activation/BN-affine/gate/residual epilogue, vectorised global stores).
// * multi-stage smem ring (full/empty mbarriers); the accumulator never leaves TMEM until the epilogue.
//
// Replaces the library-dispatched cuDNN conv + elementwise launches of
// wespeaker/models/ecapa_tdnn.py:85-106, resnet.py:35-69, campplus.py:55-83,138-170.
#include "ws_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2=SW128,4=SW64,6=SW32)
// rows are bk_bytes apart, 8-row groups SBO = 8*bk_bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int bk_bytes) {
    uint64_t layout = bk_bytes == 128 ? 2ull : (bk_bytes == 64 ? 4ull : 6ull);
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major)
    d |= (uint64_t)((8 * bk_bytes) >> 4) << 32;               // SBO
    d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int kMaxDynSmem = 200 * 1024;  // dynamic + static (barriers) must stay under the 227 KB opt-in limit
constexpr int kThreads = 192;  // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2..5 epilogue

template <int KIND>
__global__ void __launch_bounds__(kThreads) ws_conv_gemm_tc_kernel(const __grid_constant__ WsTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_TC_MAX_STAGES + 1];
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms need 1024-B aligned bases
    const int a_bytes = 128 * p.bk_bytes, b_bytes = p.bn * p.bk_bytes;
    const int stage_bytes = a_bytes + b_bytes;
    const uint32_t bar_full = smem_u32(&s_bar[0]);
    const uint32_t bar_empty = smem_u32(&s_bar[WS_TC_MAX_STAGES]);
    const uint32_t bar_acc = smem_u32(&s_bar[2 * WS_TC_MAX_STAGES]);

    // tile decode: output-channel tile fastest so CTAs sharing an activation tile are co-resident (L2 reuse)
    int tile = blockIdx.x;
    const int n0 = (tile % p.tiles_n) * p.bn; tile /= p.tiles_n;
    const int t0 = (tile % p.tiles_t) << p.bt_log2; tile /= p.tiles_t;
    const int f0 = (tile % p.tiles_f) << p.bf_log2; tile /= p.tiles_f;
    const int b0 = tile << p.bb_log2;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < WS_MAX_SRC; ++i) prefetch_tmap(&p.amap[i]);
        prefetch_tmap(&p.wmap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < p.nstages; ++s) {
                mbar_init(bar_full + 8 * s, 1);
                mbar_init(bar_empty + 8 * s, 1);
            }
            mbar_init(bar_acc, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        uint32_t ncols = p.bn < 32 ? 32u : (uint32_t)p.bn;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int tp = 0; tp < p.ntaps; ++tp) {
                const WsTcTap tap = p.taps[tp];
                const int bk_elems = p.bk_bytes >> (p.kind == 0 ? 2 : 1);
                for (int kb = 0; kb < tap.nkb; ++kb, ++it) {
                    const int s = it % p.nstages;
                    const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    mbar_expect_tx(bar_full + 8 * s, (uint32_t)stage_bytes);
                    const uint32_t sa = tiles + (uint32_t)(s * stage_bytes);
                    tma_load_4d(sa, &p.amap[tap.map], bar_full + 8 * s, tap.c0 + kb * bk_elems, t0 + tap.dt,
                                f0 + tap.df, b0);
                    tma_load_2d(sa + (uint32_t)a_bytes, &p.wmap, bar_full + 8 * s, tap.wk + kb * bk_elems, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const int kper = p.bk_bytes / 32;  // UMMA_K is 32 bytes of K for both tf32 (8) and bf16/f16 (16)
            for (int it = 0; it < p.nk_total; ++it) {
                const int s = it % p.nstages;
                const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t sa = tiles + (uint32_t)(s * stage_bytes);
                const uint64_t adesc = umma_desc(sa, p.bk_bytes);
                const uint64_t bdesc = umma_desc(sa + (uint32_t)a_bytes, p.bk_bytes);
                for (int k = 0; k < kper; ++k) {
                    // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                    umma<KIND>(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                               (uint32_t)((it | k) != 0));
                }
                umma_commit(bar_empty + 8 * s);  // frees the smem stage once these MMAs retire
            }
            umma_commit(bar_acc);  // accumulator complete -> epilogue
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;  // accumulator row == TMEM lane == position within the tile
        const int t = t0 + (r & ((1 << p.bt_log2) - 1));
        const int f = f0 + ((r >> p.bt_log2) & ((1 << p.bf_log2) - 1));
        const int b = b0 + (r >> (p.bt_log2 + p.bf_log2));
        const bool valid = (t < p.T) && (f < p.F) && (b < p.B);
        const long long pos = ((long long)b * p.F + f) * p.T + t;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        if (p.bn >= 32) {
            for (int c = 0; c < p.bn; c += 32) {
                uint32_t raw[32];
                tmem_ld32(trow + (uint32_t)c, raw);
                tmem_ld_wait();
                if (valid) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    ws_epilogue<32>(p.epi, pos, n0 + c, v);
                }
            }
        } else {
            uint32_t raw[16];
            tmem_ld16(trow, raw);
            tmem_ld_wait();
            if (valid) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
                ws_epilogue<16>(p.epi, pos, n0, v);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        uint32_t ncols = p.bn < 32 ? 32u : (uint32_t)p.bn;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace

extern "C" const char* ws_tc_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = cudaFuncSetAttribute(ws_conv_gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ws_conv_gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" const char* ws_tc_launch(const WsTcParams* p, cudaStream_t s) {
    const int stage_bytes = (128 + p->bn) * p->bk_bytes;
    const int smem = p->nstages * stage_bytes + 1024;
    const int grid = p->tiles_n * p->tiles_t * p->tiles_f * p->tiles_b;
    if (p->kind == 0) ws_conv_gemm_tc_kernel<0><<<grid, kThreads, smem, s>>>(*p);
    else ws_conv_gemm_tc_kernel<1><<<grid, kThreads, smem, s>>>(*p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
