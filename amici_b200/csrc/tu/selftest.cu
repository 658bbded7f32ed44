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
