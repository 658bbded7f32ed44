// engine.cu — model registry, objective reduction kernel and the C ABI (include/amici_b200.h) of the batched BDF engine.
//
// Kernel mapping (sm_100a, 148 SMs): one condition per lane-group of G lanes of a warp, G*C >= 1+nplist
// columns; the grid is persistent (occupancy-sized, a multiple of the SM count) and groups pull
// condition indices from a device-side queue head, so conditions that need many steps do not stall
// the lanes that finished early (per-trajectory step control without lock-step batching).
// The per-condition Nordsieck history lives in dynamic shared memory for the whole solve; HBM sees
// only the inputs and the ReturnData fields.  The solve kernel itself is in launch.cuh and is
// instantiated once per model in tu/<model>.cu.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "model_common.cuh"
#include "registry.h"

namespace amici_b200 {

std::atomic<long> g_launches{0};
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
#define g_err last_error()

// sum of llh / sllh over successful conditions; one block per output column, tree reduction in shared memory
__global__ void __launch_bounds__(256) reduce_objective_kernel(int N, int nplist, const int* status, const double* llh,
                                                               const double* sllh, double* out) {
    __shared__ double sh[256];
    const int colid = blockIdx.x;  // 0: llh, 1..nplist: sllh, nplist+1: failure count
    double acc = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const bool ok = status ? (status[n] == 0) : true;
        if (colid == nplist + 1) acc += ok ? 0.0 : 1.0;
        else if (ok) acc += (colid == 0) ? llh[n] : sllh[(size_t)n * nplist + (colid - 1)];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[colid] = sh[0];
}

// FP64 FMA throughput probe: 8 independent dependent-FMA chains per thread, enough warps to fill every SM.
// Used by bench.py as the denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 figure).
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 12345.678) sink[0] = r;  // keeps the chains alive
}

static const ModelEntry* registry(int* count) {
    static const ModelEntry entries[] = {
        entry_jakstat_adjoint(),
        entry_steadystate(),
        entry_robertson_ode(),
    };
    *count = (int)(sizeof(entries) / sizeof(entries[0]));
    return entries;
}

static const ModelEntry* find_model(const char* name) {
    int n = 0;
    const ModelEntry* e = registry(&n);
    for (int i = 0; i < n; ++i)
        if (name && std::strcmp(e[i].name, name) == 0) return &e[i];
    g_err = std::string("unknown model '") + (name ? name : "(null)") + "'";
    return nullptr;
}

static int check_args(const ab_solver_options* o, const ab_batch* b) {
    if (!o || !b) { g_err = "null argument"; return AB_E_BAD_ARGUMENT; }
    if (b->N < 0 || b->nt < 0 || b->nplist < 0) { g_err = "negative size"; return AB_E_BAD_ARGUMENT; }
    if (!b->ts && b->nt > 0) { g_err = "ts is NULL"; return AB_E_BAD_ARGUMENT; }
    if (!b->p || !b->status) { g_err = "p/status is NULL"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi >= 1 && b->nplist > 0 && !b->plist) { g_err = "plist is NULL"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi >= 1 && o->ism != 1) { g_err = "only simultaneous forward sensitivities are implemented"; return AB_E_UNSUPPORTED; }
    if (!(o->rtol >= 0) || !(o->atol >= 0)) { g_err = "negative tolerance"; return AB_E_BAD_ARGUMENT; }
    return AB_OK;
}

}  // namespace amici_b200

using namespace amici_b200;

extern "C" {

int ab_model_count(void) {
    int n = 0;
    registry(&n);
    return n;
}

const char* ab_model_name(int index) {
    int n = 0;
    const ModelEntry* e = registry(&n);
    return (index >= 0 && index < n) ? e[index].name : nullptr;
}

int ab_get_dims(const char* model, ab_dims* out) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    if (!out) return AB_E_BAD_ARGUMENT;
    *out = e->dims;
    return AB_OK;
}

int ab_get_nominal(const char* model, double* p, double* k) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    double pt[64], kt[64];
    e->nominal(pt, kt);
    if (p) std::memcpy(p, pt, sizeof(double) * e->dims.np);
    if (k) std::memcpy(k, kt, sizeof(double) * e->dims.nk);
    return AB_OK;
}

int ab_solve_device(const char* model, const ab_solver_options* opt, const ab_batch* b, int* work_counter, void* stream) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    int rc = check_args(opt, b);
    if (rc != AB_OK) return rc;
    if (!work_counter) { g_err = "work_counter is NULL"; return AB_E_BAD_ARGUMENT; }
    if (b->N == 0) return AB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int npl = (opt->sensi >= 1) ? b->nplist : 0;
    const size_t N = b->N, nt = b->nt, ny = e->dims.ny;
    if (b->res) AB_CUDA(cudaMemsetAsync(b->res, 0, sizeof(double) * N * nt * ny, st));
    if (b->sres && npl) AB_CUDA(cudaMemsetAsync(b->sres, 0, sizeof(double) * N * nt * ny * npl, st));
    if (b->FIM && npl) AB_CUDA(cudaMemsetAsync(b->FIM, 0, sizeof(double) * N * npl * npl, st));
    return e->launch(*opt, *b, work_counter, st);
}

int ab_reduce_objective(int N, int nplist, const int* status, const double* llh, const double* sllh, double* out, void* stream) {
    if (N < 0 || nplist < 0 || !llh || !out || (nplist > 0 && !sllh)) { g_err = "bad argument"; return AB_E_BAD_ARGUMENT; }
    reduce_objective_kernel<<<nplist + 2, 256, 0, (cudaStream_t)stream>>>(N, nplist, status, llh, sllh, out);
    g_launches++;
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

int ab_measure_fp64_peak(double* tflops) {
    if (!tflops) return AB_E_BAD_ARGUMENT;
    int dev = 0, sms = 0;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    AB_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, 0);
        fp64_peak_kernel<<<blocks, threads>>>(sink, iters, 0.999999, 1e-9);
        g_launches++;
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    AB_CUDA(cudaGetLastError());
    *tflops = best;
    return AB_OK;
}

long ab_kernel_launch_count(void) { return g_launches.load(); }

const char* ab_last_error(void) { return last_error().c_str(); }

// ---- host-buffer entry point ------------------------------------------------------------------
namespace {
struct DevBuf {
    void* d = nullptr;
    ~DevBuf() { if (d) cudaFree(d); }
};
}  // namespace

int ab_run_batch(const char* model, const ab_solver_options* opt, const ab_batch* hb, float* kernel_ms) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    int rc = check_args(opt, hb);
    if (rc != AB_OK) return rc;
    if (hb->N == 0) return AB_OK;
    const ab_dims& d = e->dims;
    const size_t N = hb->N, nt = hb->nt, q = (opt->sensi >= 1) ? hb->nplist : 0;
    const size_t nx = d.nx, ny = d.ny;
    ab_batch db = *hb;
    constexpr int MAXB = 40;
    DevBuf bufs[MAXB];
    int nb = 0;
    cudaStream_t st;
    AB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct Out { void* host; void* dev; size_t bytes; } outs[MAXB];
    int nout = 0;
    bool fail = false;
    auto in = [&](const void* h, size_t bytes) -> void* {
        if (!h || bytes == 0 || fail) return nullptr;
        if (!cuda_ok(cudaMalloc(&bufs[nb].d, bytes), "cudaMalloc")) { fail = true; return nullptr; }
        if (!cuda_ok(cudaMemcpyAsync(bufs[nb].d, h, bytes, cudaMemcpyHostToDevice, st), "H2D")) fail = true;
        return bufs[nb++].d;
    };
    auto out = [&](void* h, size_t bytes) -> void* {
        if (!h || bytes == 0 || fail) return nullptr;
        if (!cuda_ok(cudaMalloc(&bufs[nb].d, bytes), "cudaMalloc")) { fail = true; return nullptr; }
        outs[nout++] = Out{h, bufs[nb].d, bytes};
        return bufs[nb++].d;
    };
    const size_t D = sizeof(double), I = sizeof(int);
    db.ts = (const double*)in(hb->ts, D * nt);
    db.plist = (const int*)in(hb->plist, I * hb->nplist);
    db.pscale = (const int*)in(hb->pscale, I * d.np);
    db.p = (const double*)in(hb->p, D * N * d.np);
    db.k = (const double*)in(hb->k, D * N * d.nk);
    db.my = (const double*)in(hb->my, D * N * nt * ny);
    db.sigmay_in = (const double*)in(hb->sigmay_in, D * N * nt * ny);
    db.status = (int*)out(hb->status, I * N);
    db.llh = (double*)out(hb->llh, D * N);
    db.chi2 = (double*)out(hb->chi2, D * N);
    db.sllh = (double*)out(hb->sllh, D * N * q);
    db.x = (double*)out(hb->x, D * N * nt * nx);
    db.sx = (double*)out(hb->sx, D * N * nt * q * nx);
    db.y = (double*)out(hb->y, D * N * nt * ny);
    db.sy = (double*)out(hb->sy, D * N * nt * q * ny);
    db.sigmay = (double*)out(hb->sigmay, D * N * nt * ny);
    db.ssigmay = (double*)out(hb->ssigmay, D * N * nt * q * ny);
    db.res = (double*)out(hb->res, D * N * nt * ny);
    db.sres = (double*)out(hb->sres, D * N * nt * ny * q);
    db.FIM = (double*)out(hb->FIM, D * N * q * q);
    db.x0 = (double*)out(hb->x0, D * N * nx);
    db.sx0 = (double*)out(hb->sx0, D * N * q * nx);
    db.xdot = (double*)out(hb->xdot, D * N * nx);
    db.J = (double*)out(hb->J, D * N * nx * nx);
    db.numsteps = (int*)out(hb->numsteps, I * N * nt);
    db.numrhsevals = (int*)out(hb->numrhsevals, I * N * nt);
    db.numerrtestfails = (int*)out(hb->numerrtestfails, I * N * nt);
    db.numnonlinsolvconvfails = (int*)out(hb->numnonlinsolvconvfails, I * N * nt);
    db.order = (int*)out(hb->order, I * N * nt);
    DevBuf counter;
    if (!fail && !cuda_ok(cudaMalloc(&counter.d, sizeof(int)), "cudaMalloc")) fail = true;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!fail) {
        // outputs the kernel may leave untouched must not be garbage
        for (int i = 0; i < nout && !fail; ++i)
            if (!cuda_ok(cudaMemsetAsync(outs[i].dev, 0, outs[i].bytes, st), "memset")) fail = true;
    }
    if (!fail) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        rc = ab_solve_device(model, opt, &db, (int*)counter.d, st);
        cudaEventRecord(e1, st);
        if (rc == AB_OK) {
            for (int i = 0; i < nout && !fail; ++i)
                if (!cuda_ok(cudaMemcpyAsync(outs[i].host, outs[i].dev, outs[i].bytes, cudaMemcpyDeviceToHost, st), "D2H"))
                    fail = true;
        }
        if (!cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize")) fail = true;
        if (kernel_ms && !fail) cudaEventElapsedTime(kernel_ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    cudaStreamDestroy(st);
    if (fail) return AB_E_CUDA;
    return rc;
}

}  // extern "C"
