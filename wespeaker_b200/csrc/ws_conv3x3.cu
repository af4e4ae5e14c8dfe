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
__device__ __forceinline__ void umma_commit_c3(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// clock64 phase profile (p.prof != null): accumulate the cycles one elected thread of each role spends in each wait
#define C3_T0() const long long t0__ = p.prof ? clock64() : 0
#define C3_ACC(var) do { if (p.prof) var += clock64() - t0__; } while (0)

struct C3Step {
    int f, bg, tt, nt;
    bool first, last;
};
__device__ __forceinline__ C3Step c3_step(const WsC3Params& p, long long s, long long s_beg, long long s_end) {
    C3Step x;
    x.f = (int)(s % p.F);
    long long img = s / p.F;
    x.bg = (int)(img % p.n_bg); img /= p.n_bg;
    x.tt = (int)(img % p.n_tt);
    x.nt = (int)(img / p.n_tt);
    x.first = (s == s_beg) || (x.f == 0);
    x.last = (s + 1 == s_end) || (x.f == p.F - 1);
    return x;
}

template <int DT>
__global__ void __launch_bounds__(kC3Threads, 1) ws_conv3x3_kernel(const __grid_constant__ WsC3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 5];
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    const uint32_t s_par = stg + (uint32_t)(p.n_mt * tile_bytes);
    const uint32_t bar_afull = smem_u32(&s_bar[0]);
    const uint32_t bar_aempty = smem_u32(&s_bar[WS_C3_MAX_RING]);
    const uint32_t bar_wfull = smem_u32(&s_bar[2 * WS_C3_MAX_RING]);
    const uint32_t bar_wempty = smem_u32(&s_bar[2 * WS_C3_MAX_RING + WS_C3_MAX_WSTAGES]);
    const uint32_t bar_wres = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES]);
    const uint32_t bar_tfull = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 1]);   // [2]
    const uint32_t bar_tempty = smem_u32(&s_bar[2 * WS_C3_MAX_RING + 2 * WS_C3_MAX_WSTAGES + 3]);  // [2]
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * p.n_mt * p.N) tmem_cols <<= 1;

    const long long s_beg = (long long)p.total_steps * blockIdx.x / gridDim.x;
    const long long s_end = (long long)p.total_steps * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.amap); prefetch_tmap(&p.amap_tail); prefetch_tmap(&p.wmap); prefetch_tmap(&p.omap);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.R; ++i) { mbar_init(bar_afull + 8 * i, 1); mbar_init(bar_aempty + 8 * i, 1); }
        for (int i = 0; i < WS_C3_MAX_WSTAGES; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, 1); }
        mbar_init(bar_wres, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
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
        // ================================ input-row producer ================================
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(p.npan * p.rows_loaded * p.row_bytes);
            int j = 0;
            long long pw_aempty = 0;
            const long long tstart = p.prof ? clock64() : 0;
            auto load_row = [&](int frow, int bg, int tt) {
                const int slot = j % p.R;
                { C3_T0(); mbar_wait(bar_aempty + 8 * slot, (((uint32_t)(j / p.R)) & 1u) ^ 1u); C3_ACC(pw_aempty); }
                mbar_expect_tx(bar_afull + 8 * slot, tx);
                const int tc = tt * p.tb - 1, b0 = bg * p.nb;
                for (int kp = 0; kp < p.npan; ++kp) {
                    const uint32_t dst = ring + (uint32_t)((slot * p.npan + kp) * p.slot_rows * p.row_bytes);
                    if (p.single_box) {
                        tma_load_4d(dst, &p.amap, bar_afull + 8 * slot, kp * p.kc, tc, frow, b0);
                    } else {
                        for (int m = 0; m < p.n_mt; ++m)
                            tma_load_4d(dst + (uint32_t)(m * 128 * p.row_bytes), &p.amap, bar_afull + 8 * slot, kp * p.kc,
                                        tc + 128 * m, frow, b0);
                        tma_load_4d(dst + (uint32_t)(p.n_mt * 128 * p.row_bytes), &p.amap_tail, bar_afull + 8 * slot,
                                    kp * p.kc, tc + 128 * p.n_mt, frow, b0);
                    }
                }
                ++j;
            };
            for (long long s = s_beg; s < s_end; ++s) {
                const C3Step st = c3_step(p, s, s_beg, s_end);
                if (st.first) { load_row(st.f - 1, st.bg, st.tt); load_row(st.f, st.bg, st.tt); }
                load_row(st.f + 1, st.bg, st.tt);
            }
            if (p.prof) { p.prof[blockIdx.x * 16 + 0] = pw_aempty; p.prof[blockIdx.x * 16 + 1] = clock64() - tstart; }
        }
    } else if (warp == 3) {
        // ================================ weight producer ================================
        if (lane == 0 && s_beg < s_end) {
            if (p.w_resident) {
                mbar_expect_tx(bar_wres, (uint32_t)(nwblk * wblk_bytes));
                for (int tap = 0; tap < 9; ++tap)
                    for (int kp = 0; kp < p.npan; ++kp)
                        tma_load_2d(wbuf + (uint32_t)((tap * p.npan + kp) * wblk_bytes), &p.wmap, bar_wres,
                                    tap * p.Cin + kp * p.kc, 0);
            } else {
                int wit = 0;
                for (long long s = s_beg; s < s_end; ++s) {
                    const C3Step st = c3_step(p, s, s_beg, s_end);
                    for (int tap = 0; tap < 9; ++tap)
                        for (int kp = 0; kp < p.npan; ++kp, ++wit) {
                            const int ws = wit % p.w_stages;
                            mbar_wait(bar_wempty + 8 * ws, (((uint32_t)(wit / p.w_stages)) & 1u) ^ 1u);
                            mbar_expect_tx(bar_wfull + 8 * ws, (uint32_t)wblk_bytes);
                            tma_load_2d(wbuf + (uint32_t)(ws * wblk_bytes), &p.wmap, bar_wfull + 8 * ws,
                                        tap * p.Cin + kp * p.kc, st.nt * p.N);
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0 && s_beg < s_end) {
            const int kper = p.row_bytes / 32;
            int j = 0, wit = 0, step = 0;
            long long mw_afull = 0, mw_tempty = 0, mw_wfull = 0;
            const long long tstart = p.prof ? clock64() : 0;
            if (p.w_resident) mbar_wait(bar_wres, 0);
            for (long long s = s_beg; s < s_end; ++s, ++step) {
                const C3Step st = c3_step(p, s, s_beg, s_end);
                const int nnew = st.first ? 3 : 1;           // a new image segment starts with rows f-1, f, f+1
                j += nnew;                                   // rows f-1, f, f+1 are loads j-3, j-2, j-1
                { C3_T0(); for (int i = j - nnew; i < j; ++i) mbar_wait(bar_afull + 8 * (i % p.R), ((uint32_t)(i / p.R)) & 1u); C3_ACC(mw_afull); }
                const int buf = step & 1;
                { C3_T0(); mbar_wait(bar_tempty + 8 * buf, ((((uint32_t)step) >> 1) & 1u) ^ 1u); C3_ACC(mw_tempty); }
                tc_fence_after();
                for (int tap = 0; tap < 9; ++tap) {
                    const int df = tap / 3, dt = tap % 3;               // 0..2 (= offset + 1)
                    const int slot = (j - 3 + df) % p.R;
                    for (int kp = 0; kp < p.npan; ++kp) {
                        uint32_t wb;
                        int ws = 0;
                        if (p.w_resident) {
                            wb = wbuf + (uint32_t)((tap * p.npan + kp) * wblk_bytes);
                        } else {
                            ws = wit % p.w_stages;
                            { C3_T0(); mbar_wait(bar_wfull + 8 * ws, ((uint32_t)(wit / p.w_stages)) & 1u); C3_ACC(mw_wfull); }
                            tc_fence_after();
                            wb = wbuf + (uint32_t)(ws * wblk_bytes);
                        }
                        const uint64_t bdesc = umma_desc(wb, p.row_bytes);
                        for (int mt = 0; mt < p.n_mt; ++mt) {
                            const uint32_t arow = ring + (uint32_t)(((slot * p.npan + kp) * p.slot_rows + dt + mt * 128) * p.row_bytes);
                            const uint64_t adesc = umma_desc(arow, p.row_bytes);
                            const uint32_t tacc = tmem_base + (uint32_t)((buf * p.n_mt + mt) * p.N);
                            for (int k = 0; k < kper; ++k)
                                umma_f16_c3(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                            (uint32_t)((tap | kp | k) != 0));
                        }
                        if (!p.w_resident) { umma_commit_c3(bar_wempty + 8 * ws); ++wit; }
                    }
                }
                umma_commit_c3(bar_tfull + 8 * buf);
                umma_commit_c3(bar_aempty + 8 * ((j - 3) % p.R));       // row f-1 is not needed by later steps
                if (st.last) {                                           // end of this image segment: release f and f+1 too
                    umma_commit_c3(bar_aempty + 8 * ((j - 2) % p.R));
                    umma_commit_c3(bar_aempty + 8 * ((j - 1) % p.R));
                }
            }
            if (p.prof) {
                p.prof[blockIdx.x * 16 + 2] = mw_afull; p.prof[blockIdx.x * 16 + 3] = mw_tempty;
                p.prof[blockIdx.x * 16 + 4] = mw_wfull; p.prof[blockIdx.x * 16 + 5] = clock64() - tstart;
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        const int q = warp & 3, r = q * 32 + lane, set = (warp - 4) >> 2, et = threadIdx.x - 128;
        float* spar = reinterpret_cast<float*>(smem_raw + (s_par - smem_u32(smem_raw)));
        const int nchunks = p.N / 32, njobs = p.n_mt * nchunks;
        constexpr int MAXJ = 4;                                          // n_mt * N <= 256  =>  <= 8 jobs, 4 per warp set
        int step = 0, last_nt = -1;
        long long ew_store = 0, ew_tfull = 0, ew_res = 0, ew_body = 0;
        const long long tstart = p.prof ? clock64() : 0;
        for (long long s = s_beg; s < s_end; ++s, ++step) {
            const C3Step st = c3_step(p, s, s_beg, s_end);
            const int n0 = st.nt * p.N, t0 = st.tt * p.tb, b0 = st.bg * p.nb, buf = step & 1;
            { C3_T0(); if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); C3_ACC(ew_store); }   // staging free again
            if (st.nt != last_nt) {
                for (int c = et; c < p.N; c += 256) spar[c] = p.bias ? __ldg(p.bias + n0 + c) : 0.f;
                last_nt = st.nt;
            }
            epi_bar_sync();
            // residual rows are fetched before the accumulator wait (their latency overlaps the MMAs of this step)
            uint4 rr[MAXJ][4];
            const long long tres = p.prof ? clock64() : 0;
#pragma unroll
            for (int jj = 0; jj < MAXJ; ++jj) {
                const int job = set + 2 * jj;
#pragma unroll
                for (int i = 0; i < 4; ++i) rr[jj][i] = make_uint4(0u, 0u, 0u, 0u);
                if (job < njobs) {
                    const int mt = job / nchunks, c = (job % nchunks) * 32;
                    const int i = mt * 128 + r, u = i / p.P, tti = i - u * p.P;
                    const bool ok = u < p.nb && tti < p.tb && (t0 + tti) < p.T && (b0 + u) < p.B;
                    if (ok && p.res != nullptr) {
                        const long long pos = ((long long)(b0 + u) * p.F + st.f) * p.T + t0 + tti;
                        const uint4* src = reinterpret_cast<const uint4*>((const unsigned short*)p.res + pos * p.res_ld + n0 + c);
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2) rr[jj][i2] = __ldg(src + i2);
                    }
                }
            }
            if (p.prof) ew_res += clock64() - tres;
            { C3_T0(); mbar_wait(bar_tfull + 8 * buf, (((uint32_t)step) >> 1) & 1u); C3_ACC(ew_tfull); }
            tc_fence_after();
            const long long tbody = p.prof ? clock64() : 0;
#pragma unroll
            for (int jj = 0; jj < MAXJ; ++jj) {
                const int job = set + 2 * jj;
                if (job < njobs) {
                    const int mt = job / nchunks, c = (job % nchunks) * 32;
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
                    if (p.res != nullptr) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float x8[8];
                            ws_unpack8(rr[jj][i], DT, x8);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[8 * i + k] += x8[k];
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    // rows that are no output position (halo columns, rows past T or B) hold junk: the TMA store clips them
                    stage_store32(stg + (uint32_t)(mt * tile_bytes + (c / p.panel_cols) * panel_stride), r, p.panel_bytes,
                                  c % p.panel_cols, DT, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            epi_bar_sync();
            if (et == 0) {
                for (int mt = 0; mt < p.n_mt; ++mt)
                    for (int pn = 0; pn < npanels_out; ++pn)
                        tma_store_4d(&p.omap, stg + (uint32_t)(mt * tile_bytes + pn * panel_stride), n0 + pn * p.panel_cols,
                                     p.case_b ? 0 : t0 + 128 * mt, st.f, b0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (p.prof) ew_body += clock64() - tbody;
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.prof && et == 0) {
            p.prof[blockIdx.x * 16 + 6] = ew_store; p.prof[blockIdx.x * 16 + 7] = ew_tfull; p.prof[blockIdx.x * 16 + 8] = ew_res;
            p.prof[blockIdx.x * 16 + 9] = ew_body; p.prof[blockIdx.x * 16 + 10] = clock64() - tstart;
            p.prof[blockIdx.x * 16 + 11] = s_end - s_beg;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace

extern "C" const char* ws_c3_init(void) {
    static unsigned long long done = 0;
    int dev = 0;
    if (!ws_dev_needs_init(&done, &dev)) return nullptr;
    cudaError_t e = cudaFuncSetAttribute(ws_conv3x3_kernel<WS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC3MaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ws_conv3x3_kernel<WS_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC3MaxSmem);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaGetErrorString(e); }
    ws_dev_mark_init(&done, dev);
    return nullptr;
}

extern "C" int ws_c3_max_smem(void) { return kC3MaxSmem; }

extern "C" const char* ws_c3_launch(const WsC3Params* p, cudaStream_t s) {
    if (p->dtype == WS_BF16) ws_conv3x3_kernel<WS_BF16><<<p->grid, kC3Threads, p->smem_bytes, s>>>(*p);
    else if (p->dtype == WS_F16) ws_conv3x3_kernel<WS_F16><<<p->grid, kC3Threads, p->smem_bytes, s>>>(*p);
    else return "conv3x3: 16-bit activations only";
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
