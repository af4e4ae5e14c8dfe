// fp32-FFMA implicit-GEMM conv (CUDA cores).  This is the exact-fp32 path used for the <=1e-4 parity mode
// ("fp32" precision) and as the on-GPU cross-check of the tcgen05 kernel: identical tap/epilogue semantics
// (ws_common.cuh), no tensor cores, every product and sum in IEEE fp32 like the reference's CPU/cuDNN fp32.
//
//   block tile: 64 output positions x 64 output channels, 256 threads, 4x4 register tile per thread,
//   K streamed through shared memory in chunks of 16 channels of one tap.
#include "ws_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) ws_conv_gemm_simt_kernel(const __grid_constant__ WsSimtParams p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> channels, ty -> positions
    const long long npos = (long long)p.B * p.F * p.T;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // loader assignment: 64 rows x 16 k: thread loads 4 consecutive k of one row
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long lpos = m0 + lrow;
    int lb = 0, lf = 0, lt = 0;
    const bool lvalid = lpos < npos;
    if (lvalid) {
        lt = (int)(lpos % p.T);
        long long r = lpos / p.T;
        lf = (int)(r % p.F);
        lb = (int)(r / p.F);
    }
    const int wrow = n0 + lrow;  // weight row loaded by this thread
    const bool wvalid = wrow < p.Cout;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tp = 0; tp < p.ntaps; ++tp) {
        const WsTap tap = p.taps[tp];
        const WsSrc& src = p.src[tap.src];
        const int it = lt + tap.dt, jf = lf + tap.df;
        const bool rvalid = lvalid && it >= 0 && it < src.T && jf >= 0 && jf < src.F && lb < src.B;
        const long long abase = rvalid ? ((long long)lb * src.sB + (long long)jf * src.sF + (long long)it * src.sT + tap.c0) : 0;
        for (int k0 = 0; k0 < tap.nch; k0 += BK) {
            float a[4], w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lk + j;
                a[j] = (rvalid && k < tap.nch) ? ws_ld(src.ptr, p.dtype, abase + k) : 0.f;
                w[j] = (wvalid && k < tap.nch) ? ws_ld(p.W, p.dtype, (long long)wrow * p.Ktot + tap.wk + k) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                As[lk + j][lrow] = a[j];
                Bs[lk + j][lrow] = w[j];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float aa[4] = {av.x, av.y, av.z, av.w};
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            }
        }
    }
    const int col0 = n0 + tx * 4;
    if (col0 < p.Cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long pos = m0 + ty * 4 + i;
            if (pos < npos) ws_epilogue<4>(p.epi, pos, col0, acc[i]);
        }
    }
}

}  // namespace

extern "C" const char* ws_simt_launch(const WsSimtParams* p, cudaStream_t s) {
    const long long npos = (long long)p->B * p->F * p->T;
    dim3 grid((unsigned)((npos + BM - 1) / BM), (unsigned)((p->Cout + BN - 1) / BN));
    ws_conv_gemm_simt_kernel<<<grid, 256, 0, s>>>(*p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
