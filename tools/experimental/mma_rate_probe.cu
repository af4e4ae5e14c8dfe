// Probe: cycles per tcgen05.mma (kind::f16, M=128, cta_group::1) as a function of N, of the number of independent TMEM
// accumulators the MMAs rotate over, of a row-shifted A start address, and of how the issuing loop is written.
// Operands are whatever is in shared memory (zeros); only timing matters.  Build: nvcc -arch=sm_100a -o mma_rate_probe ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc128(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// mode 0: naive loop (descriptor arithmetic per MMA); mode 1: 4 MMAs per iteration with precomputed descriptors
__global__ void __launch_bounds__(128, 1) probe(int N, int nacc, int shift_rows, int iters, int mode, int nkb, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0u) : "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const int uwarp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (mode == 2 && uwarp == 0) {
        // warp-uniform issue: every lane runs the loop, only the tcgen05.mma itself is predicated on an elected lane, so
        // the descriptors stay in uniform registers (no per-MMA ELECT / R2UR.BROADCAST convergence loop)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = base + (uint32_t)(shift_rows * 128), b0 = base + 96 * 1024;
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int kb = it % nkb;
            const uint64_t ad = umma_desc128(a0 + kb * 18432), bd = umma_desc128(b0 + (kb & 1) * 32768);
            const uint32_t tacc = tmem + (uint32_t)((it % nacc) * N);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (elected) umma(tacc, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        const long long t1 = clock64();
        if (elected) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    if (mode != 2 && warp == 0 && lane == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // A k-blocks: nkb blocks of 144 rows x 128 B (18 KB) from base; B k-blocks behind them (N rows x 128 B)
        const uint32_t a0 = base + (uint32_t)(shift_rows * 128), b0 = base + 96 * 1024;
        const long long t0 = clock64();
        if (mode == 0) {
            for (int it = 0; it < iters; ++it) {
                const int kb = it % nkb;
                const uint64_t ad = umma_desc128(a0 + kb * 18432), bd = umma_desc128(b0 + (kb & 1) * 32768);
                const uint32_t tacc = tmem + (uint32_t)((it % nacc) * N);
                for (int k = 0; k < 4; ++k) umma(tacc, ad + 2 * k, bd + 2 * k, idesc, 1u);
            }
        } else {
            uint64_t ad[4], bd[4];
            for (int k = 0; k < 4; ++k) { ad[k] = umma_desc128(a0) + 2 * k; bd[k] = umma_desc128(b0) + 2 * k; }
            for (int it = 0; it < iters; ++it) {
                const uint32_t tacc = tmem + (uint32_t)((it % nacc) * N);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(tacc, ad[k], bd[k], idesc, 1u);
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("N nacc shift mode nkb grid : issue_cyc/MMA total_cyc/MMA (ideal N/2)\n");
    for (int grid : {148})
        for (int mode : {1, 2})
            for (int N : {32, 64, 128, 256})
                for (int nacc : {1, 2, 4})
                    for (int shift : {1}) {
                        if (nacc * N > 512) continue;
                        const int nkb = 4;
                        probe<<<grid, 128, 180 * 1024>>>(N, nacc, shift, iters, mode, nkb, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        long long h[2];
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        printf("%3d %d %d %d %d %3d : %7.1f %7.1f (%d)\n", N, nacc, shift, mode, nkb, grid, (double)h[0] / (iters * 4),
                               (double)h[1] / (iters * 4), N / 2);
                    }
    return 0;
}
