// C ABI of the two-covariance PLDA scorer (kernels in ws_plda.cu).  Mirrors
// wespeaker/utils/plda/two_cov_plda.py:156-184 (transform_embedding, log_likelihood_ratio) and the
// preparation steps of eval_sv (:218-244); arithmetic in fp64 like the numpy reference.
#include <cmath>
#include <memory>

#include "../../include/wespeaker_b200.h"
#include "ws_host.h"

using namespace ws;

struct ws_plda {
    int dim = 0, device = 0, normalize_length = 0;
    double *A = nullptr, *offset = nullptr, *psi = nullptr, *meanv = nullptr;
    double *x64 = nullptr;      // staging for fp32 -> fp64 rows
    size_t x64_rows = 0;
    double *P = nullptr, *rowc = nullptr, *Q = nullptr, *colc = nullptr;
    size_t p_elems = 0, q_elems = 0, r_elems = 0, c_elems = 0;
    ~ws_plda() {
        cudaFree(A); cudaFree(offset); cudaFree(psi); cudaFree(meanv); cudaFree(x64);
        cudaFree(P); cudaFree(rowc); cudaFree(Q); cudaFree(colc);
    }
};

namespace {
int ensure(double** p, size_t* have, size_t need) {
    if (need <= *have) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    WS_CK(cudaMalloc((void**)p, need * sizeof(double)));
    *have = need;
    return 0;
}
int prepare(ws_plda* p, const double* enroll, const int* counts, int const_n, long long N, const double* test,
            long long M, int* Kout, cudaStream_t s) {
    const int D = p->dim;
    const int K = counts != nullptr ? 2 * D : D;
    if (ensure(&p->P, &p->p_elems, (size_t)N * K) || ensure(&p->rowc, &p->r_elems, (size_t)N) ||
        ensure(&p->Q, &p->q_elems, (size_t)M * K) || ensure(&p->colc, &p->c_elems, (size_t)M))
        return 1;
    WS_CKS(ws_launch_plda_prep_enroll(enroll, counts, const_n, p->psi, N, D, K, p->P, p->rowc, s));
    WS_CKS(ws_launch_plda_prep_test(test, p->psi, const_n, M, D, K, p->Q, p->colc, s));
    *Kout = K;
    return 0;
}
}  // namespace

extern "C" {

int ws_plda_create(int dim, const double* mu, const double* transform, const double* psi, const double* offset,
                   int normalize_length, int device, ws_plda** out) {
    if (!transform || !psi || !offset || !out || dim <= 0) { set_err("ws_plda_create: bad argument"); return 1; }
    (void)mu;  // mu only enters through offset = -transform @ mu (two_cov_plda.py:153,346)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_err("ws_plda_create: no CUDA device (this scorer has no CPU fallback)");
        return 1;
    }
    WS_CK(cudaSetDevice(device));
    std::unique_ptr<ws_plda> p(new ws_plda());
    p->dim = dim; p->device = device; p->normalize_length = normalize_length;
    WS_CK(cudaMalloc((void**)&p->A, (size_t)dim * dim * 8));
    WS_CK(cudaMalloc((void**)&p->offset, (size_t)dim * 8));
    WS_CK(cudaMalloc((void**)&p->psi, (size_t)dim * 8));
    WS_CK(cudaMalloc((void**)&p->meanv, (size_t)dim * 8));
    WS_CK(cudaMemcpy(p->A, transform, (size_t)dim * dim * 8, cudaMemcpyHostToDevice));
    WS_CK(cudaMemcpy(p->offset, offset, (size_t)dim * 8, cudaMemcpyHostToDevice));
    WS_CK(cudaMemcpy(p->psi, psi, (size_t)dim * 8, cudaMemcpyHostToDevice));
    *out = p.release();
    return 0;
}

int ws_plda_transform(ws_plda* p, const float* x_dev, long long N, const double* mean_vec_host, int pre_norm,
                      double* y_dev, void* stream) {
    if (!p || !x_dev || !y_dev || N < 0) { set_err("ws_plda_transform: bad argument"); return 1; }
    if (N == 0) return 0;
    WS_CK(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int D = p->dim;
    if (ensure(&p->x64, &p->x64_rows, (size_t)N * D)) return 1;
    WS_CKS(ws_launch_f32_to_f64(x_dev, p->x64, N * D, s));
    const double* mv = nullptr;
    if (mean_vec_host) {
        WS_CK(cudaMemcpyAsync(p->meanv, mean_vec_host, (size_t)D * 8, cudaMemcpyHostToDevice, s));
        mv = p->meanv;
    }
    if (mv || pre_norm) WS_CKS(ws_launch_plda_center_norm(p->x64, mv, N, D, pre_norm != 0, s));
    // y = A x + offset, tiled over rows to respect the grid limit
    const long long chunk = 65535LL * 128;
    for (long long r = 0; r < N; r += chunk) {
        const long long n = std::min(chunk, N - r);
        WS_CKS(ws_launch_dgemm_nt(p->x64 + r * D, p->A, nullptr, p->offset, y_dev + r * D, 1, n, D, D, D, s));
    }
    if (p->normalize_length) WS_CKS(ws_launch_plda_rownorm(y_dev, N, D, s));
    return 0;
}

// fp64-input twin of ws_plda_transform: eval_sv (two_cov_plda.py:218-233) averages a speaker's sessions before the
// transform and the reference keeps that mean in fp64; x64_dev rows are (already mean-subtracted) fp64 vectors.
int ws_plda_transform64(ws_plda* p, const double* x64_dev, long long N, int pre_norm, double* y_dev, void* stream) {
    if (!p || !x64_dev || !y_dev || N < 0) { set_err("ws_plda_transform64: bad argument"); return 1; }
    if (N == 0) return 0;
    WS_CK(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int D = p->dim;
    const double* src = x64_dev;
    if (pre_norm) {
        if (ensure(&p->x64, &p->x64_rows, (size_t)N * D)) return 1;
        WS_CK(cudaMemcpyAsync(p->x64, x64_dev, (size_t)N * D * 8, cudaMemcpyDeviceToDevice, s));
        WS_CKS(ws_launch_plda_center_norm(p->x64, nullptr, N, D, 1, s));
        src = p->x64;
    }
    const long long chunk = 65535LL * 128;
    for (long long r = 0; r < N; r += chunk) {
        const long long n = std::min(chunk, N - r);
        WS_CKS(ws_launch_dgemm_nt(src + r * D, p->A, nullptr, p->offset, y_dev + r * D, 1, n, D, D, D, s));
    }
    if (p->normalize_length) WS_CKS(ws_launch_plda_rownorm(y_dev, N, D, s));
    return 0;
}

int ws_plda_score_matrix(ws_plda* p, const double* enroll_t_dev, const int* counts_dev, int const_n, long long N,
                         const double* test_t_dev, long long M, void* out_dev, int out_is_f64, long long out_ld,
                         void* stream) {
    if (!p || !enroll_t_dev || !test_t_dev || !out_dev) { set_err("ws_plda_score_matrix: null argument"); return 1; }
    if (N == 0 || M == 0) return 0;
    if (out_ld < M) { set_err("ws_plda_score_matrix: out_ld < M"); return 1; }
    WS_CK(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    int K = 0;
    if (prepare(p, enroll_t_dev, counts_dev, const_n, N, test_t_dev, M, &K, s)) return 1;
    const long long chunk = 65535LL * 128;
    for (long long r = 0; r < N; r += chunk) {
        const long long n = std::min(chunk, N - r);
        void* o = out_is_f64 ? (void*)((double*)out_dev + r * out_ld) : (void*)((float*)out_dev + r * out_ld);
        WS_CKS(ws_launch_dgemm_nt(p->P + r * K, p->Q, p->rowc + r, p->colc, o, out_is_f64, n, M, K, out_ld, s));
    }
    return 0;
}

int ws_plda_score_trials(ws_plda* p, const double* enroll_t_dev, const int* counts_dev, int const_n, long long N,
                         const double* test_t_dev, long long M, const long long* ei_dev, const long long* ti_dev,
                         long long ntrials, double* out_dev, void* stream) {
    if (!p || !enroll_t_dev || !test_t_dev || !out_dev || !ei_dev || !ti_dev) { set_err("ws_plda_score_trials: null argument"); return 1; }
    if (ntrials == 0) return 0;
    WS_CK(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    int K = 0;
    if (prepare(p, enroll_t_dev, counts_dev, const_n, N, test_t_dev, M, &K, s)) return 1;
    WS_CKS(ws_launch_plda_trials(p->P, p->rowc, p->Q, p->colc, ei_dev, ti_dev, ntrials, K, out_dev, s));
    return 0;
}

void ws_plda_destroy(ws_plda* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    delete p;
}

}  // extern "C"
