// Two-covariance PLDA log-likelihood-ratio scoring in fp64 (the reference is numpy fp64:
// wespeaker/utils/plda/two_cov_plda.py:156-184).  The per-trial formula is refactored into a GEMM:
//
//   LLR(i,j) = sum_d t_jd * (m_id / v_id)  -  1/2 sum_d t_jd^2 / v_id  +  r_i  +  c_j
//     m_i = n_i psi/(n_i psi + 1) * e_i ,  v_i = 1 + psi/(n_i psi + 1)
//     r_i = -1/2 (sum log v_i + sum m_i^2 / v_i) ,  c_j = 1/2 (sum log(psi+1) + sum t_j^2/(psi+1))
//
// With a constant enroll count n the second term only depends on j and moves into c_j (K = D); with per-speaker
// counts it is a second K-block (K = 2D): P_i = [m_i/v_i , -1/(2 v_i)], Q_j = [t_j , t_j^2].
// fp64 keeps |error| ~1e-13, far inside the 1e-5 budget which fp32 accumulators cannot guarantee (sums of
// magnitude ~256 have ulp 3e-5).  B200 runs DFMA at half the FFMA rate, so this is DFMA-bound, not HBM-bound:
// 128x64 block tile, 8x4 register tile per thread, K streamed through shared memory.
#include "ws_kernels.cuh"

namespace {

constexpr int TM = 128, TN = 64, TK = 8;

__global__ void __launch_bounds__(256) dgemm_nt_kernel(const double* __restrict__ A, const double* __restrict__ Bm,
                                                       const double* __restrict__ rowc, const double* __restrict__ colc,
                                                       void* __restrict__ out, int out_is_f64, long long M, long long N,
                                                       int K, long long out_ld) {
    __shared__ double As[TK][TM + 2];
    __shared__ double Bs[TK][TN + 2];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> 4 columns, ty -> 8 rows
    const long long m0 = (long long)blockIdx.y * TM, n0 = (long long)blockIdx.x * TN;
    const int arow = tid >> 1, ak = (tid & 1) * 4;
    const int brow = tid >> 2, bk = (tid & 3) * 2;
    const bool av = (m0 + arow) < M, bv = (n0 + brow) < N;
    const double* ap = A + (m0 + arow) * (long long)K + ak;
    const double* bp = Bm + (n0 + brow) * (long long)K + bk;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < K; k0 += TK) {
        double a[4] = {0, 0, 0, 0}, bb[2] = {0, 0};
        if (av) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + ak + j < K) a[j] = ap[k0 + j];
        }
        if (bv) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (k0 + bk + j < K) bb[j] = bp[k0 + j];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) As[ak + j][arow] = a[j];
#pragma unroll
        for (int j = 0; j < 2; ++j) Bs[bk + j][brow] = bb[j];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            double av8[8], bv4[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) av8[i] = As[k][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv4[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av8[i], bv4[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long r = m0 + ty * 8 + i;
        if (r >= M) continue;
        const double rc = rowc != nullptr ? rowc[r] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long c = n0 + tx * 4 + j;
            if (c >= N) continue;
            const double v = acc[i][j] + rc + (colc != nullptr ? colc[c] : 0.0);
            if (out_is_f64) ((double*)out)[r * out_ld + c] = v;
            else ((float*)out)[r * out_ld + c] = (float)v;
        }
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (double)in[i];
}

// one warp per row
__global__ void center_norm_kernel(double* __restrict__ x, const double* __restrict__ mean_vec, long long N, int D,
                                   int do_norm) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    double* p = x + r * D;
    double ss = 0.0;
    for (int d = lane; d < D; d += 32) {
        double v = p[d] - (mean_vec != nullptr ? mean_vec[d] : 0.0);
        p[d] = v;
        ss = fma(v, v, ss);
    }
    if (do_norm) {
        ss = warp_sum_d(ss);
        const double sc = sqrt((double)D) / sqrt(ss);
        for (int d = lane; d < D; d += 32) p[d] *= sc;
    }
}

__global__ void prep_enroll_kernel(const double* __restrict__ e, const int* __restrict__ counts, int const_n,
                                   const double* __restrict__ psi, long long N, int D, int K, double* __restrict__ P,
                                   double* __restrict__ rowc) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const double n = counts != nullptr ? (double)counts[r] : (double)const_n;
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double ps = psi[d];
        const double den = n * ps + 1.0;
        const double m = n * ps / den * e[r * D + d];
        const double v = 1.0 + ps / den;
        P[r * K + d] = m / v;
        if (K == 2 * D) P[r * K + D + d] = -0.5 / v;
        acc += log(v) + m * m / v;
    }
    acc = warp_sum_d(acc);
    if (lane == 0) rowc[r] = -0.5 * acc;
}

__global__ void prep_test_kernel(const double* __restrict__ t, const double* __restrict__ psi, int const_n,
                                 long long M, int D, int K, double* __restrict__ Q, double* __restrict__ colc) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double ps = psi[d];
        const double tv = t[r * D + d];
        Q[r * K + d] = tv;
        acc += 0.5 * (log(ps + 1.0) + tv * tv / (ps + 1.0));
        if (K == 2 * D) {
            Q[r * K + D + d] = tv * tv;
        } else {
            const double n = (double)const_n;
            const double v = 1.0 + ps / (n * ps + 1.0);
            acc -= 0.5 * tv * tv / v;
        }
    }
    acc = warp_sum_d(acc);
    if (lane == 0) colc[r] = acc;
}

__global__ void trials_kernel(const double* __restrict__ P, const double* __restrict__ rowc,
                              const double* __restrict__ Q, const double* __restrict__ colc,
                              const long long* __restrict__ ei, const long long* __restrict__ ti, long long ntrials,
                              int K, double* __restrict__ out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= ntrials) return;
    const long long i = ei[r], j = ti[r];
    double acc = 0.0;
    for (int k = lane; k < K; k += 32) acc = fma(P[i * K + k], Q[j * K + k], acc);
    acc = warp_sum_d(acc);
    if (lane == 0) out[r] = acc + rowc[i] + colc[j];
}

inline const char* last_err() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
inline unsigned rows_grid(long long rows) { return (unsigned)((rows + 7) / 8); }

}  // namespace

const char* ws_launch_f32_to_f64(const float* in, double* out, long long n, cudaStream_t s) {
    long long g = (n + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    if (g < 1) g = 1;
    f32_to_f64_kernel<<<(unsigned)g, 256, 0, s>>>(in, out, n);
    return last_err();
}

const char* ws_launch_dgemm_nt(const double* A, const double* Bm, const double* rowc, const double* colc, void* out,
                               int out_is_f64, long long M, long long N, int K, long long out_ld, cudaStream_t s) {
    if (M <= 0 || N <= 0) return nullptr;
    const long long gy = (M + TM - 1) / TM, gx = (N + TN - 1) / TN;
    if (gy > 65535) return "dgemm_nt: M too large for one launch (tile the enroll rows)";
    dim3 grid((unsigned)gx, (unsigned)gy);
    dgemm_nt_kernel<<<grid, 256, 0, s>>>(A, Bm, rowc, colc, out, out_is_f64, M, N, K, out_ld);
    return last_err();
}

const char* ws_launch_plda_center_norm(double* x, const double* mean_vec, long long N, int D, int do_norm,
                                       cudaStream_t s) {
    if (N <= 0) return nullptr;
    center_norm_kernel<<<rows_grid(N), 256, 0, s>>>(x, mean_vec, N, D, do_norm);
    return last_err();
}
const char* ws_launch_plda_rownorm(double* x, long long N, int D, cudaStream_t s) {
    return ws_launch_plda_center_norm(x, nullptr, N, D, 1, s);
}
const char* ws_launch_plda_prep_enroll(const double* e, const int* counts, int const_n, const double* psi,
                                       long long N, int D, int K, double* P, double* rowc, cudaStream_t s) {
    if (N <= 0) return nullptr;
    prep_enroll_kernel<<<rows_grid(N), 256, 0, s>>>(e, counts, const_n, psi, N, D, K, P, rowc);
    return last_err();
}
const char* ws_launch_plda_prep_test(const double* t, const double* psi, int const_n, long long M, int D, int K,
                                     double* Q, double* colc, cudaStream_t s) {
    if (M <= 0) return nullptr;
    prep_test_kernel<<<rows_grid(M), 256, 0, s>>>(t, psi, const_n, M, D, K, Q, colc);
    return last_err();
}
const char* ws_launch_plda_trials(const double* P, const double* rowc, const double* Q, const double* colc,
                                  const long long* ei, const long long* ti, long long ntrials, int K, double* out,
                                  cudaStream_t s) {
    if (ntrials <= 0) return nullptr;
    trials_kernel<<<rows_grid(ntrials), 256, 0, s>>>(P, rowc, Q, colc, ei, ti, ntrials, K, out);
    return last_err();
}
