// ASTP tail in one launch: attention logits + softmax over time + weighted mean / std (pooling_layers.py:119-144,
//   alpha = softmax(linear2(tanh(linear1(x))), dim=T);  mean = sum alpha x;  std = sqrt(clamp(sum alpha x^2 - mean^2, 1e-7))).
// linear1 (+tanh) stays a conv-GEMM launch that leaves H [B][T][128]; this kernel replaces the linear2 launch (which wrote
// the [B][T][C] logits, 157 MB at B = 256) and the statistics launch (which read them back together with x):
//
//   * the logits tile is computed TRANSPOSED: D[channel, t] = W2[channel, :] . H[t, :]  (A = a 128-channel block of W2,
//     B = the utterance's H rows, N = up to 256 frames), so a TMEM lane is a channel and its columns are time: the softmax
//     over time and the weighted sums are a per-thread loop over the accumulator row - no cross-thread reduction, and the
//     logits never leave the SM (they also stay fp32 instead of being rounded to 16 bits);
//   * linear2's bias is constant over time and cancels in the softmax: it is not applied;
//   * utterances longer than 256 frames run as chunks with an online (running-max) softmax, so any T works;
//   * a work unit is (utterance, g consecutive channel blocks): H is fetched once per unit and reused for its g blocks;
//   * x reaches the statistics through a TMA ring of [128 frames x 128 channels] tiles in shared memory (un-swizzled, so
//     thread = channel reads a conflict-free 2-byte column): per-thread 2-byte global loads made the epilogue
//     latency-bound (first version: 541 us; straight-line code with one group of register prefetch: 120 us).
//
// Warp roles (384 threads): w0 TMA producer (H chunks, W2 blocks), w1 MMA issuer, w2 TMEM allocator, w4..w11 two epilogue
// sets of 4 warps; consecutive channel blocks alternate between the sets and between the two 256-column TMEM buffers, so
// the exp-heavy epilogue of block j overlaps the MMAs and the epilogue of block j+1.
#include "ws_tc_common.cuh"

namespace {
using namespace ws_tcdev;

constexpr int kThreadsA = 384;
constexpr int kHPanelBytes = 256 * 128;   // one K panel (64 of the 128 hidden units) of a 256-frame H chunk
constexpr int kWPanelBytes = 128 * 128;   // one K panel of a 128-channel W2 block
constexpr int kXTileBytes = 128 * 256;   // 128 frames x 128 channels of x, row pitch 256 B, no swizzle
constexpr int kXSlots = 3;
constexpr int kSmemA = 2 * kHPanelBytes + 2 * 2 * kWPanelBytes + kXSlots * kXTileBytes + 1024;   // H chunk + two W2 blocks + x ring

__device__ __forceinline__ void umma1(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit1(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float ex2f(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
template <int DT>
__device__ __forceinline__ float cvt16(uint32_t v) {
    if (DT == WS_BF16) return __uint_as_float(v << 16);
    return __half2float(__ushort_as_half((unsigned short)v));
}

#define AP_T(x) const long long x = p.prof ? clock64() : 0
#define AP_ADD(i, v) do { if (p.prof && lane == 0) p.prof[(size_t)blockIdx.x * 16 + (i)] += (v); } while (0)

template <int DT>
__global__ void __launch_bounds__(kThreadsA, 1) ws_astp_fused_kernel(const __grid_constant__ WsAstpParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[10 + 2 * kXSlots];
    __shared__ uint32_t s_tmem;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t hbuf = base, wbuf = base + 2u * kHPanelBytes, xbuf = wbuf + 4u * kWPanelBytes;
    const uint32_t bar_hfull = smem_u32(&s_bar[0]), bar_hempty = smem_u32(&s_bar[1]);
    const uint32_t bar_wfull = smem_u32(&s_bar[2]), bar_wempty = smem_u32(&s_bar[4]);     // [2] each
    const uint32_t bar_tfull = smem_u32(&s_bar[6]), bar_tempty = smem_u32(&s_bar[8]);
    const uint32_t bar_xfull = smem_u32(&s_bar[10]), bar_xempty = smem_u32(&s_bar[10 + kXSlots]);   // [kXSlots] each

    if (warp == 0 && lane == 0) { prefetch_tmap(&p.hmap); prefetch_tmap(&p.wmap); prefetch_tmap(&p.xmap); }
    if (warp == 1 && lane == 0) {
        mbar_init(bar_hfull, 1); mbar_init(bar_hempty, 1);
        for (int i = 0; i < kXSlots; ++i) { mbar_init(bar_xfull + 8 * i, 1); mbar_init(bar_xempty + 8 * i, 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, 1);
            mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4);   // one arrive per epilogue warp of the set
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const int nblk = p.C / 128, upb = nblk / p.g, nunits = p.B * upb;
    AP_T(k0);

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t h_it = 0, w_it = 0, x_it = 0;
            for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
                const int b = u / upb, blk0 = (u % upb) * p.g;
                const int Tb = p.lens ? max(1, min(p.T, p.lens[b])) : p.T;
                const int nch = (Tb + 255) >> 8;
                auto load_h = [&](int ch) {
                    AP_T(a0);
                    mbar_wait(bar_hempty, (h_it & 1u) ^ 1u);
                    AP_T(a1); AP_ADD(0, a1 - a0);
                    mbar_expect_tx(bar_hfull, 2u * kHPanelBytes);
                    tma_load_3d(hbuf, &p.hmap, bar_hfull, 0, ch * 256, b);
                    tma_load_3d(hbuf + kHPanelBytes, &p.hmap, bar_hfull, 64, ch * 256, b);
                    ++h_it;
                };
                if (nch == 1) load_h(0);
                for (int j = 0; j < p.g; ++j) {
                    for (int ch = 0; ch < nch; ++ch) {
                        if (nch > 1) load_h(ch);
                        if (ch == 0) {
                            const uint32_t ws = w_it & 1u;
                            AP_T(a2);
                            mbar_wait(bar_wempty + 8 * ws, ((w_it >> 1) & 1u) ^ 1u);
                            AP_T(a3); AP_ADD(1, a3 - a2);
                            mbar_expect_tx(bar_wfull + 8 * ws, 2u * kWPanelBytes);
                            const uint32_t dst = wbuf + ws * 2u * kWPanelBytes;
                            tma_load_2d(dst, &p.wmap, bar_wfull + 8 * ws, 0, (blk0 + j) * 128);
                            tma_load_2d(dst + kWPanelBytes, &p.wmap, bar_wfull + 8 * ws, 64, (blk0 + j) * 128);
                            ++w_it;
                        }
                        // x tiles of this (block, chunk): 128 frames x 128 channels each, in the order the epilogue consumes them
                        const int rows = min(256, Tb - ch * 256);
                        for (int hf = 0; hf * 128 < rows; ++hf, ++x_it) {
                            const uint32_t xs = x_it % kXSlots;
                            mbar_wait(bar_xempty + 8 * xs, ((x_it / kXSlots) & 1u) ^ 1u);
                            mbar_expect_tx(bar_xfull + 8 * xs, (uint32_t)kXTileBytes);
                            tma_load_3d(xbuf + xs * kXTileBytes, &p.xmap, bar_xfull + 8 * xs, (blk0 + j) * 128, ch * 256 + hf * 128, b);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const uint32_t elected = elect_one();
        const uint32_t fmt = DT == WS_BF16 ? 1u : 0u;
        const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 4) << 24);
        uint32_t h_it = 0, w_it = 0, q = 0, nset[2] = {0u, 0u};
        for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int b = u / upb;
            const int Tb = p.lens ? max(1, min(p.T, p.lens[b])) : p.T;
            const int nch = (Tb + 255) >> 8;
            AP_T(m0);
            if (nch == 1) { mbar_wait(bar_hfull, h_it & 1u); ++h_it; }
            AP_T(m1); AP_ADD(2, m1 - m0);
            for (int j = 0; j < p.g; ++j, ++q) {
                const uint32_t set = q & 1u, ws = w_it & 1u;
                for (int ch = 0; ch < nch; ++ch) {
                    if (nch > 1) { mbar_wait(bar_hfull, h_it & 1u); ++h_it; }
                    AP_T(m2);
                    if (ch == 0) mbar_wait(bar_wfull + 8 * ws, (w_it >> 1) & 1u);
                    AP_T(m3); AP_ADD(3, m3 - m2);
                    const uint32_t ns = set ? nset[1] : nset[0];
                    mbar_wait(bar_tempty + 8 * set, (ns & 1u) ^ 1u);
                    AP_T(m4); AP_ADD(4, m4 - m3); AP_ADD(5, 1);
                    tc_fence_after();
                    const int rows = min(256, Tb - ch * 256);
                    const uint32_t N = (uint32_t)((rows + 15) & ~15);
                    const uint32_t idesc = idesc0 | ((N >> 3) << 17);
                    const uint32_t tacc = tmem_base + set * 256u;
                    const uint32_t wa = wbuf + ws * 2u * kWPanelBytes;
                    const uint64_t a0 = umma_desc(wa, 128), a1 = umma_desc(wa + kWPanelBytes, 128);
                    const uint64_t b0 = umma_desc(hbuf, 128), b1 = umma_desc(hbuf + kHPanelBytes, 128);
                    if (elected) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma1(tacc, a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma1(tacc, a1 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc, 1u);
                        umma_commit1(bar_tfull + 8 * set);
                        if (nch > 1) umma_commit1(bar_hempty);
                        if (ch == nch - 1) umma_commit1(bar_wempty + 8 * ws);
                    }
                    __syncwarp();
                    if (set) ++nset[1]; else ++nset[0];
                }
                ++w_it;
            }
            if (nch == 1 && elected) umma_commit1(bar_hempty);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ================================ epilogue: per-channel online softmax statistics over time ================================
        const uint32_t set = (uint32_t)((warp - 4) >> 2);
        const int quad = warp & 3;
        const float kLog2e = 1.4426950408889634f;
        uint32_t q = 0, ns = 0, x_it = 0;
        const uint32_t xcol = (uint32_t)((quad * 32 + lane) * 2);
        for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int b = u / upb, blk0 = (u % upb) * p.g;
            const int Tb = p.lens ? max(1, min(p.T, p.lens[b])) : p.T;
            const int nch = (Tb + 255) >> 8;
            const uint32_t x_per_block = (uint32_t)((Tb + 127) >> 7);   // x tiles of one channel block of this utterance
            for (int j = 0; j < p.g; ++j, ++q) {
                if ((q & 1u) != set) { x_it += x_per_block; continue; }
                const int c = (blk0 + j) * 128 + quad * 32 + lane;
                float m = -INFINITY, s0 = 0.f, s1 = 0.f, s2 = 0.f;
                for (int ch = 0; ch < nch; ++ch, ++ns) {
                    const int rows = min(256, Tb - ch * 256);
                    AP_T(e0);
                    mbar_wait(bar_tfull + 8 * set, ns & 1u);
                    AP_T(e1);
                    if (quad == 0) AP_ADD(6 + 2 * set, e1 - e0);
                    tc_fence_after();
                    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + set * 256u;
                    for (int hf = 0; hf * 128 < rows; ++hf, ++x_it) {
                        const uint32_t xs = x_it % kXSlots;
                        mbar_wait(bar_xfull + 8 * xs, (x_it / kXSlots) & 1u);
                        const uint32_t xt = xbuf + xs * kXTileBytes + xcol;
                        const int hrows = min(128, rows - hf * 128);
                        // straight-line code per group of 32 frames (a branch per element made this loop latency-bound: ISETP ->
                        // BRA -> FADD chains with two warps per scheduler): frames behind the end get logit -inf (weight 0), x 0
                        auto group = [&](uint32_t* lr, int g0, int nv) {   // 32 frames: nv >= 32 valid, else the first nv
                            uint32_t xr[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(xr[i]) : "r"(xt + (uint32_t)((g0 + i) * 256)));
                            if (nv < 32) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) { lr[i] = i < nv ? lr[i] : 0xff800000u; xr[i] = i < nv ? xr[i] : 0u; }
                            }
                            float g4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                            for (int i = 0; i < 32; ++i) g4[i & 3] = fmaxf(g4[i & 3], __uint_as_float(lr[i]));
                            const float gm = fmaxf(fmaxf(g4[0], g4[1]), fmaxf(g4[2], g4[3]));
                            const float mn = fmaxf(m, gm);
                            const float sc = ex2f((m - mn) * kLog2e);   // m = -inf on the first group: ex2(-inf) = 0
                            m = mn;
                            const float moff = -mn * kLog2e;
                            float a0[2] = {0.f, 0.f}, a1[2] = {0.f, 0.f}, a2[2] = {0.f, 0.f};
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float e = ex2f(fmaf(__uint_as_float(lr[i]), kLog2e, moff));
                                const float xv = cvt16<DT>(xr[i]);
                                const float ex = e * xv;
                                a0[i & 1] += e;
                                a1[i & 1] += ex;
                                a2[i & 1] = fmaf(ex, xv, a2[i & 1]);
                            }
                            s0 = fmaf(s0, sc, a0[0] + a0[1]);
                            s1 = fmaf(s1, sc, a1[0] + a1[1]);
                            s2 = fmaf(s2, sc, a2[0] + a2[1]);
                        };
                        // two groups per TMEM wait: the second group's tcgen05.ld completes behind the first group's arithmetic
#pragma unroll 1
                        for (int g0 = 0; g0 < hrows; g0 += 64) {
                            const int nva = hrows - g0, nvb = nva - 32;
                            uint32_t la[32], lb[32];
                            tmem_ld32(trow + (uint32_t)(hf * 128 + g0), la);
                            if (nvb > 0) tmem_ld32(trow + (uint32_t)(hf * 128 + g0 + 32), lb);
                            tmem_ld_wait();
                            group(la, g0, nva);
                            if (nvb > 0) group(lb, g0 + 32, nvb);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_xempty + 8 * xs);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * set);
                    AP_T(e2);
                    if (quad == 0) AP_ADD(7 + 2 * set, e2 - e1);
                }
                const float mean = s1 / s0;
                const float var = s2 / s0 - mean * mean;
                p.out[(long long)b * 2 * p.C + c] = mean;
                p.out[(long long)b * 2 * p.C + p.C + c] = sqrtf(fmaxf(var, 1e-7f));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { AP_T(k1); AP_ADD(10, k1 - k0); }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace

extern "C" const char* ws_astp_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = cudaFuncSetAttribute(ws_astp_fused_kernel<WS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemA);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ws_astp_fused_kernel<WS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemA);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" int ws_astp_smem(void) { return kSmemA; }

extern "C" const char* ws_astp_launch(const WsAstpParams* p, cudaStream_t s) {
    if (p->dtype == WS_BF16) ws_astp_fused_kernel<WS_BF16><<<p->grid, kThreadsA, kSmemA, s>>>(*p);
    else if (p->dtype == WS_F16) ws_astp_fused_kernel<WS_F16><<<p->grid, kThreadsA, kSmemA, s>>>(*p);
    else return "ws_astp_launch: 16-bit activations only";
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
