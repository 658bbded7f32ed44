// selftest.cu — device-side self-tests exported through the C ABI (ab_selftest_*): checks of kernels' internal variants
// against each other on caller-supplied inputs.  Not on any product path.
#include <cuda_runtime.h>

#include "../bdf_core.cuh"
#include "../registry.h"

namespace amici_b200 {

// The three forms of cvSLdet (all orders per lane / three lanes with shuffles / three lanes with shared-memory exchange)
// on the same ring buffers; one warp per case for the 32-lane form, one 16-lane and one 4-lane group for the exchange form.
__global__ void selftest_sldet_kernel(int n, const double* ssdat, const int* head, const int* q, int* out) {
    __shared__ double sd[5][8];
    __shared__ double xb[2][40];
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < 40; i += blockDim.x) (&sd[0][0])[i] = ssdat[(size_t)c * 40 + i];
        __syncthreads();
        double s[6][4], ratc[5][4];
        for (int i = 1; i <= 5; ++i) {
            int idx = head[c] + i - 1;
            idx = idx >= 5 ? idx - 5 : idx;
            for (int kk = 1; kk <= 3; ++kk) {
                s[i][kk] = sd[idx][kk];
                if (i <= 4) ratc[i][kk] = sd[idx][4 + kk];
            }
        }
        const int serial = bdf_sldet(s, ratc, q[c]);
        Group<32> g32;
        const int lanes = bdf_sldet_lanes<32>(g32, sd, head[c], q[c]);
        Group<16> g16;
        const int x16 = bdf_sldet_xchg<16>(g16, sd, xb[(threadIdx.x >> 4) & 1], head[c], q[c]);
        __syncwarp();
        int x4 = 0;
        if (threadIdx.x < 4) {
            Group<4> g4;
            x4 = bdf_sldet_xchg<4>(g4, sd, xb[0], head[c], q[c]);
        }
        __syncwarp();
        if (threadIdx.x == 0) { out[4 * c] = serial; out[4 * c + 1] = lanes; out[4 * c + 2] = x16; out[4 * c + 3] = x4; }
        if (threadIdx.x == 17 && x16 != serial) out[4 * c + 2] = -1000;   // the upper half must agree too
    }
}

}  // namespace amici_b200

extern "C" int ab_selftest_sldet(int n, const double* ssdat, const int* head, const int* q, int* out) {
    using namespace amici_b200;
    if (n <= 0 || !ssdat || !head || !q || !out) return AB_E_BAD_ARGUMENT;
    double* d_s = nullptr;
    int *d_h = nullptr, *d_q = nullptr, *d_o = nullptr;
    AB_CUDA(cudaMalloc(&d_s, sizeof(double) * 40 * n));
    AB_CUDA(cudaMalloc(&d_h, sizeof(int) * n));
    AB_CUDA(cudaMalloc(&d_q, sizeof(int) * n));
    AB_CUDA(cudaMalloc(&d_o, sizeof(int) * 4 * n));
    AB_CUDA(cudaMemcpy(d_s, ssdat, sizeof(double) * 40 * n, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(d_h, head, sizeof(int) * n, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(d_q, q, sizeof(int) * n, cudaMemcpyHostToDevice));
    selftest_sldet_kernel<<<n < 1024 ? n : 1024, 32>>>(n, d_s, d_h, d_q, d_o);
    g_launches++;
    AB_CUDA(cudaGetLastError());
    AB_CUDA(cudaMemcpy(out, d_o, sizeof(int) * 4 * n, cudaMemcpyDeviceToHost));
    cudaFree(d_s); cudaFree(d_h); cudaFree(d_q); cudaFree(d_o);
    return AB_OK;
}

// ---- general Hermite spline (hermite_spline.cuh) on caller-supplied configurations ---------------------------------------
#include "../hermite_spline.cuh"
namespace amici_b200 {
constexpr int ST_SPL_NN = 8, ST_SPL_NP = 4;

__global__ void selftest_spline_kernel(int ncases, const int* cfg, const double* nodes, const double* values, const double* slopes, int npar,
                                       const double* dvaluesdp, const double* dslopesdp, int nt, const double* ts, double* value_out,
                                       double* sens_out, int* err_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncases) return;
    const int* k = cfg + 8 * c;
    const SplineCfg sc{k[1], k[2], k[3], k[4], k[5] != 0, k[6] != 0, k[7] != 0};
    HermiteSplineData<ST_SPL_NN, ST_SPL_NP> H;
    hs_build(H, sc, k[0], nodes + (size_t)c * ST_SPL_NN, values + (size_t)c * ST_SPL_NN, slopes + (size_t)c * ST_SPL_NN, npar,
             dvaluesdp + (size_t)c * ST_SPL_NN * npar, dslopesdp + (size_t)c * ST_SPL_NN * npar);
    for (int i = 0; i < nt; ++i) {
        bool err = false;
        const double t = ts[(size_t)c * nt + i];
        value_out[(size_t)c * nt + i] = hs_value(H, t, &err);
        for (int ip = 0; ip < npar; ++ip) sens_out[((size_t)c * nt + i) * npar + ip] = hs_sensitivity(H, t, ip, &err);
        err_out[(size_t)c * nt + i] = err ? 1 : 0;
    }
}
}  // namespace amici_b200

extern "C" int ab_selftest_spline(int ncases, const int* cfg, const double* nodes, const double* values, const double* slopes, int npar,
                                  const double* dvaluesdp, const double* dslopesdp, int nt, const double* ts, double* value_out,
                                  double* sens_out, int* err_out) {
    using namespace amici_b200;
    if (ncases <= 0 || nt <= 0 || npar < 0 || npar > ST_SPL_NP || !cfg || !nodes || !values || !slopes || !ts || !value_out || !err_out)
        return AB_E_BAD_ARGUMENT;
    for (int c = 0; c < ncases; ++c)
        if (cfg[8 * c] < 2 || cfg[8 * c] > ST_SPL_NN) return AB_E_BAD_ARGUMENT;
    const size_t nn = (size_t)ncases * ST_SPL_NN, np = npar > 0 ? npar : 1;
    int *d_cfg = nullptr, *d_err = nullptr;
    double *d_x = nullptr, *d_v = nullptr, *d_m = nullptr, *d_dv = nullptr, *d_dm = nullptr, *d_t = nullptr, *d_val = nullptr, *d_s = nullptr;
    AB_CUDA(cudaMalloc(&d_cfg, sizeof(int) * 8 * ncases));
    AB_CUDA(cudaMalloc(&d_err, sizeof(int) * ncases * nt));
    AB_CUDA(cudaMalloc(&d_x, sizeof(double) * nn));
    AB_CUDA(cudaMalloc(&d_v, sizeof(double) * nn));
    AB_CUDA(cudaMalloc(&d_m, sizeof(double) * nn));
    AB_CUDA(cudaMalloc(&d_dv, sizeof(double) * nn * np));
    AB_CUDA(cudaMalloc(&d_dm, sizeof(double) * nn * np));
    AB_CUDA(cudaMalloc(&d_t, sizeof(double) * ncases * nt));
    AB_CUDA(cudaMalloc(&d_val, sizeof(double) * ncases * nt));
    AB_CUDA(cudaMalloc(&d_s, sizeof(double) * ncases * nt * np));
    AB_CUDA(cudaMemcpy(d_cfg, cfg, sizeof(int) * 8 * ncases, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(d_x, nodes, sizeof(double) * nn, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(d_v, values, sizeof(double) * nn, cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(d_m, slopes, sizeof(double) * nn, cudaMemcpyHostToDevice));
    if (npar > 0) {
        AB_CUDA(cudaMemcpy(d_dv, dvaluesdp, sizeof(double) * nn * npar, cudaMemcpyHostToDevice));
        AB_CUDA(cudaMemcpy(d_dm, dslopesdp, sizeof(double) * nn * npar, cudaMemcpyHostToDevice));
    }
    AB_CUDA(cudaMemcpy(d_t, ts, sizeof(double) * ncases * nt, cudaMemcpyHostToDevice));
    selftest_spline_kernel<<<(ncases + 63) / 64, 64>>>(ncases, d_cfg, d_x, d_v, d_m, npar, d_dv, d_dm, nt, d_t, d_val, d_s, d_err);
    g_launches++;
    AB_CUDA(cudaGetLastError());
    AB_CUDA(cudaMemcpy(value_out, d_val, sizeof(double) * ncases * nt, cudaMemcpyDeviceToHost));
    if (npar > 0 && sens_out) AB_CUDA(cudaMemcpy(sens_out, d_s, sizeof(double) * ncases * nt * npar, cudaMemcpyDeviceToHost));
    AB_CUDA(cudaMemcpy(err_out, d_err, sizeof(int) * ncases * nt, cudaMemcpyDeviceToHost));
    cudaFree(d_cfg); cudaFree(d_err); cudaFree(d_x); cudaFree(d_v); cudaFree(d_m); cudaFree(d_dv); cudaFree(d_dm); cudaFree(d_t); cudaFree(d_val); cudaFree(d_s);
    return AB_OK;
}
