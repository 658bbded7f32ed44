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
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <vector>
#include <string>
#include <thread>

#include "model_common.cuh"
#include "registry.h"

namespace amici_b200 {

std::atomic<long> g_launches{0};
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
std::string& last_kernel() {
    static thread_local std::string k;
    return k;
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
        entry_events(),
        entry_neuron(),
        entry_massaction(),
        entry_massaction_small(),
        entry_spline_modes(),
        entry_ev_heavisides(),
        entry_ev_nested(),
        entry_ev_pw_simple(),
        entry_ev_ddeltax(),
        entry_ev_pw_semi(),
        entry_ev_pw_state(),
        entry_dirac(),
        entry_nested_events(),
        entry_robertson(),
        entry_calvetti(),
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

static int check_args(const ab_solver_options* o, const ab_batch* b, const int np = 1) {
    if (!o || !b) { g_err = "null argument"; return AB_E_BAD_ARGUMENT; }
    if (b->N < 0 || b->nt < 0 || b->nplist < 0) { g_err = "negative size"; return AB_E_BAD_ARGUMENT; }
    if (!b->ts && b->nt > 0) { g_err = "ts is NULL"; return AB_E_BAD_ARGUMENT; }
    if ((!b->p && np > 0) || !b->status) { g_err = "p/status is NULL"; return AB_E_BAD_ARGUMENT; }   // a model without free parameters has no p
    if (!std::isfinite(o->t0)) { g_err = "t0 must be finite"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi >= 1 && b->nplist > 0 && !b->plist) { g_err = "plist is NULL"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi >= 1 && o->ism != 1 && o->ism != 2) {
        g_err = "InternalSensitivityMethod must be simultaneous (1) or staggered (2); staggered1 is not implemented";
        return AB_E_UNSUPPORTED;
    }
    if (!(o->rtol >= 0) || !(o->atol >= 0)) { g_err = "negative tolerance"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi_meth != 0 && o->sensi_meth != 1 && o->sensi_meth != 2) { g_err = "unknown SensitivityMethod"; return AB_E_BAD_ARGUMENT; }
    if (o->sensi >= 1 && o->sensi_meth == 2 && (b->k_preeq || b->t_presim || o->posteq || o->reinit_states_sim || o->reinit_states_presim ||
                                                  o->x_nonneg != 0ull || o->maxtime > 0.0)) {
        g_err = "adjoint sensitivity analysis: plain simulations only (no pre-/post-equilibration, presimulation, re-initialisation, non-negativity, time limit)";
        return AB_E_UNSUPPORTED;
    }
    if (o->nmaxevent < 0) { g_err = "nmaxevent must be non-negative"; return AB_E_BAD_ARGUMENT; }
    if (b->k_presim && !b->t_presim) { g_err = "k_presim without t_presim"; return AB_E_BAD_ARGUMENT; }
    if ((b->k_preeq || o->posteq) && (!(o->rtol_ss >= 0) || !(o->atol_ss >= 0) || !(o->rtol_ss_sensi >= 0) || !(o->atol_ss_sensi >= 0))) {
        g_err = "steady-state tolerances must be resolved (non-negative) when k_preeq is given";
        return AB_E_BAD_ARGUMENT;
    }
    if ((b->k_preeq || o->posteq) && (o->ss_mode < 0 || o->ss_mode > 2 || o->ss_sensi_mode < 0 || o->ss_sensi_mode > 2)) {
        g_err = "invalid steady-state mode";
        return AB_E_BAD_ARGUMENT;
    }
    if ((b->k_preeq || o->posteq) && o->sensi >= 1 && o->ss_mode == 0 && o->ss_sensi_mode == 1) {
        // SteadyStateProblem constructor (reference src/forwardproblem.cpp:792-806)
        g_err = "steady-state computation mode 'newtonOnly' and sensitivity mode 'integrationOnly' are not compatible";
        return AB_E_BAD_ARGUMENT;
    }
    return AB_OK;
}

// forward sensitivities (or none): the model's own kernels; SensitivityMethod::adjoint: adjoint.cuh
static int launch_solve(const ModelEntry* e, const ab_solver_options& o, const ab_batch& b, int* queue, const Progress& pg, cudaStream_t st) {
    if (o.sensi >= 1 && o.sensi_meth == 2 && b.nplist > 0) return launch_adjoint(e->name, o, b, queue, pg, st);
    return e->launch(o, b, queue, pg, st);
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
    std::vector<double> pt(e->dims.np + 1), kt(e->dims.nk + 1);
    e->nominal(pt.data(), kt.data());
    if (p) std::memcpy(p, pt.data(), sizeof(double) * e->dims.np);
    if (k) std::memcpy(k, kt.data(), sizeof(double) * e->dims.nk);
    return AB_OK;
}

int ab_solve_device(const char* model, const ab_solver_options* opt, const ab_batch* b, int* work_counter, void* stream) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    int rc = check_args(opt, b, e->dims.np);
    if (rc != AB_OK) return rc;
    if (!work_counter) { g_err = "work_counter is NULL"; return AB_E_BAD_ARGUMENT; }
    if (b->N == 0) return AB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int npl = (opt->sensi >= 1) ? b->nplist : 0;
    const size_t N = b->N, nt = b->nt, ny = e->dims.ny;
    if (b->res) AB_CUDA(cudaMemsetAsync(b->res, 0, sizeof(double) * N * nt * ny, st));
    if (b->sres && npl) AB_CUDA(cudaMemsetAsync(b->sres, 0, sizeof(double) * N * nt * ny * npl, st));
    if (b->FIM && npl) AB_CUDA(cudaMemsetAsync(b->FIM, 0, sizeof(double) * N * npl * npl, st));
    // outputs that a failing condition does not write from the failing timepoint on keep the reference's default 0
    // (ReturnData::invalidate touches only the fields it lists, src/rdata.cpp:684-735); callers re-use their output tensors
    if (b->sigmay) AB_CUDA(cudaMemsetAsync(b->sigmay, 0, sizeof(double) * N * nt * ny, st));
    if (b->ssigmay && npl) AB_CUDA(cudaMemsetAsync(b->ssigmay, 0, sizeof(double) * N * nt * ny * npl, st));
    for (int* cnt : {b->numsteps, b->numrhsevals, b->numerrtestfails, b->numnonlinsolvconvfails, b->order})
        if (cnt) AB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * N * nt, st));
    for (int* st3 : {b->posteq_status, b->posteq_numsteps, b->preeq_status, b->preeq_numsteps})
        if (st3) AB_CUDA(cudaMemsetAsync(st3, 0, sizeof(int) * N * 3, st));
    // event-observable slots that are never reached stay 0 as in the reference (src/rdata.cpp:160-165, :203-205)
    const size_t nzs = (size_t)e->dims.nz * (opt->nmaxevent > 0 ? opt->nmaxevent : 0) * N;
    if (nzs) {
        if (b->z) AB_CUDA(cudaMemsetAsync(b->z, 0, sizeof(double) * nzs, st));
        if (b->rz) AB_CUDA(cudaMemsetAsync(b->rz, 0, sizeof(double) * nzs, st));
        if (b->sigmaz) AB_CUDA(cudaMemsetAsync(b->sigmaz, 0, sizeof(double) * nzs, st));
        if (b->sz && npl) AB_CUDA(cudaMemsetAsync(b->sz, 0, sizeof(double) * nzs * npl, st));
        if (b->srz && npl) AB_CUDA(cudaMemsetAsync(b->srz, 0, sizeof(double) * nzs * npl, st));
    }
    return launch_solve(e, *opt, *b, work_counter, Progress{}, st);
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

const char* ab_last_kernel_name(void) { return last_kernel().c_str(); }

// ---- host-buffer entry point ------------------------------------------------------------------
namespace {
// Field table of ab_batch: per-condition element counts, so that host and device views can be built generically.
struct FieldDesc {
    size_t offset;      // offsetof(ab_batch, field)
    size_t elem;        // sizeof element
    size_t per_cond;    // elements per condition
    bool input;
};

// Device workspace of ab_run_batch, kept between calls: allocating and freeing ~70 buffers (2.6 GB for the JAK-STAT
// bench batch) per call costs more than the copies.  Released by ab_release_workspace() or at process exit.
constexpr int WS_SLOTS = 64;
constexpr int WS_MAX_CHUNKS = 4096;
struct Workspace {
    std::mutex mu;
    void* dev[WS_SLOTS] = {};
    size_t cap[WS_SLOTS] = {};
    int* host_flags = nullptr;      // mapped pinned memory, WS_MAX_CHUNKS ints
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int device = -1;

    bool reserve(int slot, size_t bytes) {
        if (bytes <= cap[slot]) return true;
        if (dev[slot]) cudaFree(dev[slot]);
        dev[slot] = nullptr;
        cap[slot] = 0;
        if (!cuda_ok(cudaMalloc(&dev[slot], bytes), "cudaMalloc")) return false;
        cap[slot] = bytes;
        return true;
    }
    void release() {
        for (int i = 0; i < WS_SLOTS; ++i) {
            if (dev[i]) cudaFree(dev[i]);
            dev[i] = nullptr;
            cap[i] = 0;
        }
        if (host_flags) cudaFreeHost(host_flags);
        host_flags = nullptr;
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (compute) cudaStreamDestroy(compute);
        if (copy) cudaStreamDestroy(copy);
        ev0 = ev1 = nullptr;
        compute = copy = nullptr;
        device = -1;
    }
    bool init() {
        int d = 0;
        if (!cuda_ok(cudaGetDevice(&d), "cudaGetDevice")) return false;
        if (device == d) return true;
        release();   // first use on this device
        if (!cuda_ok(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
        if (!cuda_ok(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
        if (!cuda_ok(cudaEventCreate(&ev0), "cudaEventCreate") || !cuda_ok(cudaEventCreate(&ev1), "cudaEventCreate")) return false;
        if (!cuda_ok(cudaHostAlloc((void**)&host_flags, sizeof(int) * WS_MAX_CHUNKS, cudaHostAllocMapped), "cudaHostAlloc")) return false;
        device = d;
        return true;
    }
};
// One workspace per device: a process that drives several GPUs (one host thread each) neither serialises on a global
// lock nor re-allocates when the current device changes.
constexpr int WS_MAX_DEVICES = 64;
Workspace* workspaces() {
    static Workspace w[WS_MAX_DEVICES];
    return w;
}
Workspace* workspace_of_current_device() {
    int d = 0;
    if (!cuda_ok(cudaGetDevice(&d), "cudaGetDevice")) return nullptr;
    if (d < 0 || d >= WS_MAX_DEVICES) { last_error() = "device index out of range"; return nullptr; }
    return workspaces() + d;
}
}  // namespace

void ab_release_workspace(void) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    for (int d = 0; d < WS_MAX_DEVICES; ++d) {
        Workspace& w = workspaces()[d];
        std::lock_guard<std::mutex> lock(w.mu);
        if (w.device < 0) continue;
        cudaSetDevice(w.device);
        w.release();
    }
    cudaSetDevice(cur);
}

// One launch per super-batch (all inputs resident first), outputs streamed back while the kernel runs: the kernel raises
// a flag in mapped host memory whenever a sub-batch of `chunk` consecutive conditions is complete (Progress in
// registry.h), and this thread then queues that sub-batch's D2H copies on a second stream.  Compared with one launch
// per sub-batch there is no drain of the persistent grid between sub-batches, and the copies still overlap the solve.
int ab_run_batch(const char* model, const ab_solver_options* opt, const ab_batch* hb, float* kernel_ms) {
    const ModelEntry* e = find_model(model);
    if (!e) return AB_E_UNKNOWN_MODEL;
    int rc = check_args(opt, hb, e->dims.np);
    if (rc != AB_OK) return rc;
    if (kernel_ms) *kernel_ms = 0.f;
    if (hb->N == 0) return AB_OK;
    // t = inf output points (post-equilibration, reference src/forwardproblem.cpp:376-387) must be the last ones
    ab_solver_options opt_local = *opt;
    opt_local.posteq = 0;
    for (int it = 0; it < hb->nt; ++it) {
        const double t = hb->ts[it];
        if (t != t || t == -INFINITY || (opt_local.posteq && std::isfinite(t))) {
            g_err = "timepoints must be finite, followed by t = inf entries only";
            return AB_E_BAD_ARGUMENT;
        }
        if (t == INFINITY) opt_local.posteq = 1;
    }
    opt = &opt_local;
    const ab_dims& d = e->dims;
    const size_t N = hb->N, nt = hb->nt, q = (opt->sensi >= 1) ? hb->nplist : 0;
    const size_t nx = d.nx, ny = d.ny, nz = d.nz, D = sizeof(double), I = sizeof(int);
    const size_t nme = opt->nmaxevent > 0 ? opt->nmaxevent : 0;
#define AB_F(f, el, cnt, in) FieldDesc{offsetof(ab_batch, f), el, (size_t)(cnt), in}
    const FieldDesc fields[] = {
        AB_F(p, D, d.np, true), AB_F(k, D, d.nk, true), AB_F(my, D, nt * ny, true), AB_F(sigmay_in, D, nt * ny, true),
        AB_F(k_preeq, D, d.nk, true), AB_F(mz, D, nme * nz, true), AB_F(sigmaz_in, D, nme * nz, true),
        AB_F(x0_in, D, nx, true), AB_F(sx0_in, D, q * nx, true), AB_F(k_presim, D, d.nk, true), AB_F(t_presim, D, 1, true),
        AB_F(status, I, 1, false), AB_F(llh, D, 1, false), AB_F(chi2, D, 1, false), AB_F(sllh, D, q, false),
        AB_F(x, D, nt * nx, false), AB_F(sx, D, nt * q * nx, false), AB_F(y, D, nt * ny, false),
        AB_F(sy, D, nt * q * ny, false), AB_F(sigmay, D, nt * ny, false), AB_F(ssigmay, D, nt * q * ny, false),
        AB_F(res, D, nt * ny, false), AB_F(sres, D, nt * ny * q, false), AB_F(FIM, D, q * q, false),
        AB_F(x0, D, nx, false), AB_F(sx0, D, q * nx, false), AB_F(xdot, D, nx, false), AB_F(J, D, nx * nx, false),
        AB_F(numsteps, I, nt, false), AB_F(numrhsevals, I, nt, false), AB_F(numerrtestfails, I, nt, false),
        AB_F(numnonlinsolvconvfails, I, nt, false), AB_F(order, I, nt, false),
        AB_F(x_ss, D, nx, false), AB_F(sx_ss, D, q * nx, false), AB_F(preeq_status, I, 3, false),
        AB_F(preeq_numsteps, I, 3, false), AB_F(preeq_wrms, D, 1, false), AB_F(preeq_t, D, 1, false),
        AB_F(posteq_status, I, 3, false), AB_F(posteq_numsteps, I, 3, false), AB_F(posteq_wrms, D, 1, false), AB_F(posteq_t, D, 1, false),
        AB_F(z, D, nme * nz, false), AB_F(sz, D, nme * q * nz, false), AB_F(rz, D, nme * nz, false),
        AB_F(srz, D, nme * q * nz, false), AB_F(sigmaz, D, nme * nz, false)};
#undef AB_F
    constexpr int NF = (int)(sizeof(fields) / sizeof(fields[0]));
    static_assert(NF + 8 <= WS_SLOTS, "workspace slots");
    enum { S_TS = NF, S_PLIST, S_PSCALE, S_QUEUE, S_DONE };
    auto host_ptr = [&](const FieldDesc& f) -> char* { return *(char* const*)((const char*)hb + f.offset); };

    // sizes: sub-batch (granularity of the streamed copies) and super-batch (one launch; bounds the device footprint)
    // 64 sub-batches per batch, between 512 and 8192 conditions each: what is still on the device when the kernel ends (the last
    // sub-batch) is the un-overlapped tail of the call, and with eight ranks on one host that tail shares the host's memory
    // system (8 GPUs, 65,536 conditions each: 8192 -> 2048 conditions per sub-batch took the end-to-end step from 469 to 449 ms)
    size_t chunk = (size_t)hb->N / 64;
    chunk = chunk < 512 ? 512 : chunk > 8192 ? 8192 : chunk;
    if (const char* env = getenv("AMICI_B200_CHUNK")) chunk = (size_t)atol(env) > 0 ? (size_t)atol(env) : chunk;
    size_t bytes_per_cond = 0;
    for (int f = 0; f < NF; ++f)
        if (host_ptr(fields[f])) bytes_per_cond += fields[f].elem * fields[f].per_cond;
    size_t budget = (size_t)16 << 30;   // device bytes per launch
    if (const char* env = getenv("AMICI_B200_LAUNCH_BYTES")) budget = (size_t)atoll(env) > 0 ? (size_t)atoll(env) : budget;
    size_t super = bytes_per_cond ? budget / bytes_per_cond : N;
    if (super > chunk * (size_t)WS_MAX_CHUNKS) super = chunk * (size_t)WS_MAX_CHUNKS;
    super = super / chunk * chunk;
    if (super < chunk) super = chunk;
    if (super > N) super = N;

    Workspace* wp = workspace_of_current_device();
    if (!wp) return AB_E_CUDA;
    Workspace& w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    if (!w.init()) return AB_E_CUDA;
    if (!w.reserve(S_TS, D * (nt ? nt : 1)) || !w.reserve(S_PLIST, I * (hb->nplist ? hb->nplist : 1)) ||
        !w.reserve(S_PSCALE, I * (d.np ? d.np : 1)) || !w.reserve(S_QUEUE, I) || !w.reserve(S_DONE, I * WS_MAX_CHUNKS))
        return AB_E_CUDA;
    for (int f = 0; f < NF; ++f) {
        const size_t bytes = fields[f].elem * fields[f].per_cond * super;
        if (host_ptr(fields[f]) && bytes && !w.reserve(f, bytes)) return AB_E_CUDA;
    }
    bool fail = false;
    auto ok = [&](cudaError_t err, const char* what) { if (!cuda_ok(err, what)) fail = true; return !fail; };
    ab_batch base = *hb;
    if (nt) ok(cudaMemcpyAsync(w.dev[S_TS], hb->ts, D * nt, cudaMemcpyHostToDevice, w.compute), "H2D ts");
    if (hb->plist && hb->nplist) ok(cudaMemcpyAsync(w.dev[S_PLIST], hb->plist, I * hb->nplist, cudaMemcpyHostToDevice, w.compute), "H2D plist");
    if (hb->pscale && d.np) ok(cudaMemcpyAsync(w.dev[S_PSCALE], hb->pscale, I * d.np, cudaMemcpyHostToDevice, w.compute), "H2D pscale");
    base.ts = (const double*)w.dev[S_TS];
    base.plist = hb->plist ? (const int*)w.dev[S_PLIST] : nullptr;
    base.pscale = hb->pscale ? (const int*)w.dev[S_PSCALE] : nullptr;
    int* dev_flags = nullptr;
    ok(cudaHostGetDevicePointer((void**)&dev_flags, w.host_flags, 0), "cudaHostGetDevicePointer");

    float kms_total = 0.f;
    for (size_t lo = 0; lo < N && !fail && rc == AB_OK; lo += super) {
        const size_t cnt = (lo + super <= N) ? super : N - lo;
        const int nchunks = (int)((cnt + chunk - 1) / chunk);
        // the previous super-batch's copies read the same device buffers
        ok(cudaStreamSynchronize(w.copy), "cudaStreamSynchronize");
        ab_batch db = base;
        db.N = (int)cnt;
        for (int f = 0; f < NF && !fail; ++f) {
            char* h = host_ptr(fields[f]);
            const size_t stride = fields[f].elem * fields[f].per_cond;
            *(void**)((char*)&db + fields[f].offset) = (h && stride) ? w.dev[f] : nullptr;
            if (!h || !stride) continue;
            if (fields[f].input) ok(cudaMemcpyAsync(w.dev[f], h + lo * stride, stride * cnt, cudaMemcpyHostToDevice, w.compute), "H2D");
            else ok(cudaMemsetAsync(w.dev[f], 0, stride * cnt, w.compute), "memset");
        }
        ok(cudaMemsetAsync(w.dev[S_DONE], 0, I * nchunks, w.compute), "memset");
        for (int c = 0; c < nchunks; ++c) w.host_flags[c] = 0;
        if (fail) break;
        Progress pg;
        pg.done = (int*)w.dev[S_DONE];
        pg.host_flags = dev_flags;
        pg.chunk = (int)chunk;
        cudaEventRecord(w.ev0, w.compute);
        rc = launch_solve(e, *opt, db, (int*)w.dev[S_QUEUE], pg, w.compute);
        cudaEventRecord(w.ev1, w.compute);
        if (rc != AB_OK) break;
        volatile int* flags = w.host_flags;
        for (int c = 0; c < nchunks && !fail; ++c) {
            unsigned spins = 0;
            bool finished = false;
            // A sub-batch takes tens of milliseconds: after a short spin the thread sleeps between polls instead of burning a
            // host core per rank for the whole solve (the reference arm and other ranks' copies need those cores).
            while (!flags[c]) {
                ++spins;
                if ((spins & 0x3fu) == 0) {
                    // the kernel ended (or died) without raising the flag: stop waiting; a clean end has set every flag
                    if (cudaEventQuery(w.ev1) != cudaErrorNotReady) { finished = true; break; }
                }
                if (spins < 64) std::this_thread::yield();
                else std::this_thread::sleep_for(std::chrono::microseconds(spins < 1024 ? 20 : 100));
            }
            if (finished && !flags[c]) {
                if (!cuda_ok(cudaStreamSynchronize(w.compute), "solve kernel")) fail = true;
                if (!flags[c] && !fail) { g_err = "solve kernel ended without completing a sub-batch"; fail = true; }
                if (fail) break;
            }
            const size_t clo = (size_t)c * chunk, ccnt = (clo + chunk <= cnt) ? chunk : cnt - clo;
            for (int f = 0; f < NF && !fail; ++f) {
                char* h = host_ptr(fields[f]);
                const size_t stride = fields[f].elem * fields[f].per_cond;
                if (!h || fields[f].input || stride == 0) continue;
                ok(cudaMemcpyAsync(h + (lo + clo) * stride, (char*)w.dev[f] + clo * stride, stride * ccnt, cudaMemcpyDeviceToHost, w.copy), "D2H");
            }
        }
        if (!cuda_ok(cudaStreamSynchronize(w.compute), "cudaStreamSynchronize")) fail = true;
        if (!fail) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, w.ev0, w.ev1);
            kms_total += ms;
        }
    }
    if (!cuda_ok(cudaStreamSynchronize(w.copy), "cudaStreamSynchronize")) fail = true;
    if (kernel_ms) *kernel_ms = kms_total;
    if (fail) return AB_E_CUDA;
    return rc;
}

}  // extern "C"
