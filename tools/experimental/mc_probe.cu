// Probe for the next optimisation of the conv-GEMM mainloop (DESIGN.md "what comes next"): is the operand delivery rate
// (37.6 B/clk/SM measured by ncu on ws_conv_gemm_tc3_kernel) a per-SM delivery cap, or an L2-read cap that TMA multicast
// inside a thread-block cluster would lift?  NOT part of the library; build and run on a B200:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mc_probe tools/experimental/mc_probe.cu -lcuda
//   ./gpurun_out/mc_probe
//
// Every CTA streams 32 KB stages through an 8-deep smem ring with no math: one 16 KB "A" tile that only this CTA reads
// and one 16 KB "B" tile that all CTAs of the cluster read.  Variants: cluster size 1 / 2 / 4, B fetched by every CTA
// (unicast) or fetched in 1/cluster slices and multicast.  Output: delivered bytes per clock per SM and aggregate TB/s.
// Read next to `ncu --metrics lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum`.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kStages = 8, kTileBytes = 16 * 1024, kRowsPerTile = 128;   // 128 rows x 128 B

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
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
                 "[%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CL = cluster size, MC = multicast the shared tile in 1/CL slices
template <int CL, bool MC>
__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ CUtensorMap map_full,
                                                      const __grid_constant__ CUtensorMap map_slice, int tiles_total, int iters,
                                                      unsigned long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kStages];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kStages]);
    const uint32_t rank = CL > 1 ? cluster_rank() : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, MC ? CL : 1);      // with multicast every CTA of the cluster writes into this stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) cluster_sync();
    const long long t_begin = clock64();
    const int cluster_id = blockIdx.x / CL;
    if (warp == 0 && lane == 0) {                       // producer
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            mbar_expect_tx(full0 + 8 * s, 2 * kTileBytes);
            const uint32_t dst = base + (uint32_t)(s * 2 * kTileBytes);
            const int a_tile = (int)(((long long)blockIdx.x * 7919 + (long long)it * 131) % tiles_total);       // private tile
            const int b_tile = (int)(((long long)cluster_id * 104729 + (long long)it * 257 + 13) % tiles_total); // cluster-shared tile
            tma_load_2d(dst, &map_full, full0 + 8 * s, 0, a_tile * kRowsPerTile);
            if (MC) {
                const int rows = kRowsPerTile / CL;
                tma_load_2d_mc(dst + kTileBytes + rank * (uint32_t)(rows * 128), &map_slice, full0 + 8 * s, 0,
                               b_tile * kRowsPerTile + (int)rank * rows, (uint16_t)((1u << CL) - 1u));
            } else {
                tma_load_2d(dst + kTileBytes, &map_full, full0 + 8 * s, 0, b_tile * kRowsPerTile);
            }
        }
    } else if (warp == 1 && lane == 0) {                // consumer: release the stage as soon as it has landed
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            if (MC) {
                for (uint32_t r = 0; r < (uint32_t)CL; ++r) mbar_arrive_cluster(empty0 + 8 * s, r);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
            }
        }
    }
    __syncthreads();
    const long long t_end = clock64();
    if (CL > 1) cluster_sync();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = (unsigned long long)(t_end - t_begin);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, void* ptr, long long rows, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

template <int CL, bool MC>
static void run(EncodeFn enc, void* buf, long long rows, int iters, double ghz, int sms) {
    CUtensorMap full = make_map(enc, buf, rows, kRowsPerTile);
    CUtensorMap slice = make_map(enc, buf, rows, kRowsPerTile / CL);
    const int grid = sms / CL * CL;
    const size_t smem = (size_t)kStages * 2 * kTileBytes + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<CL, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long* cyc;
    CK(cudaMalloc(&cyc, grid * sizeof(unsigned long long)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int tiles_total = (int)(rows / kRowsPerTile);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, probe_kernel<CL, MC>, full, slice, tiles_total, iters, cyc));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double delivered = (double)grid * iters * 2.0 * kTileBytes;           // bytes landing in shared memory
    const double l2_reads = MC ? (double)grid * iters * (kTileBytes + (double)kTileBytes / CL) : delivered;
    printf("cluster %d %-9s grid %3d: %8.1f us  delivered %6.2f TB/s = %5.1f B/clk/SM @%.2f GHz   L2 requests %6.2f TB/s\n", CL,
           MC ? "multicast" : "unicast", grid, ms * 1e3, delivered / (ms * 1e-3) / 1e12,
           delivered / grid / (ms * 1e-3 * ghz * 1e9), ghz, l2_reads / (ms * 1e-3) / 1e12);
    CK(cudaFree(cyc));
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    const long long mb = argc > 2 ? atoll(argv[2]) : 64;                         // working set (MB): <= 126 MB stays L2-resident
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const long long rows = mb * 1024 * 1024 / 128 / kRowsPerTile * kRowsPerTile;
    void* buf;
    CK(cudaMalloc(&buf, (size_t)rows * 128));
    CK(cudaMemset(buf, 1, (size_t)rows * 128));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn enc = (EncodeFn)fn;
    const double ghz = khz * 1e-6;
    printf("%d SMs, %d iterations of 2 x 16 KB per CTA, %lld MB working set\n", sms, iters, mb);
    run<1, false>(enc, buf, rows, iters, ghz, sms);
    run<2, false>(enc, buf, rows, iters, ghz, sms);
    run<2, true>(enc, buf, rows, iters, ghz, sms);
    run<4, false>(enc, buf, rows, iters, ghz, sms);
    run<4, true>(enc, buf, rows, iters, ghz, sms);
    return 0;
}
