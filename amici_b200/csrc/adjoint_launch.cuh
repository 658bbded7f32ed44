// adjoint_launch.cuh — kernel and history workspace of the adjoint method (adjoint.cuh); one translation unit per model
// (tu/adjoint_<model>.cu), dispatch by model name in tu/adjoint.cu.
#ifndef AMICI_B200_ADJOINT_LAUNCH_CUH
#define AMICI_B200_ADJOINT_LAUNCH_CUH
#include <cstdlib>

#include "adjoint.cuh"
#include "block_launch.cuh"   // pooled, event-guarded device workspaces

namespace amici_b200 {

constexpr int ADJ_THREADS = 64;

// workspace layout, all arrays interleaved over the `slots` threads of the grid (element stride = slots):
//   double t[cap], double y[cap][NX], double dJydx[nt][NX], int ord[cap]
template <class M>
__global__ void __launch_bounds__(ADJ_THREADS) adjoint_kernel(const ab_solver_options o, const ab_batch b, int* queue_head, const Progress pg,
                                                              double* ws, const int cap) {
    const size_t slots = (size_t)gridDim.x * blockDim.x, slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    FwdHistory<M::NX> hist;
    hist.t = ws + slot;
    hist.y = ws + slots * cap + slot;
    double* dj = ws + slots * cap * (1 + M::NX) + slot;
    hist.ord = reinterpret_cast<int*>(ws + slots * ((size_t)cap * (1 + M::NX) + (size_t)(b.nt > 0 ? b.nt : 1) * M::NX)) + slot;
    hist.stride = slots;
    hist.cap = cap;
    for (;;) {
        const int n = atomicAdd(queue_head, 1);
        if (n >= b.N) break;
        AdjointSimulation<M> sim(o, b, hist, dj, n);
        sim.run();
        if (pg.done) {
            __threadfence();
            const int c = n / pg.chunk;
            const int size = (b.N - c * pg.chunk < pg.chunk) ? b.N - c * pg.chunk : pg.chunk;
            if (atomicAdd(pg.done + c, 1) == size - 1) {
                __threadfence_system();
                *reinterpret_cast<volatile int*>(pg.host_flags + c) = 1;
            }
        }
    }
}

template <class M>
int launch_adjoint_model(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    auto kern = adjoint_kernel<M>;
    int dev = 0, sms = 0, per_sm = 0;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ADJ_THREADS, 0));
    if (per_sm < 1) {
        last_error() = "adjoint kernel does not fit on an SM";
        return AB_E_CUDA;
    }
    if (const char* capenv = std::getenv("AMICI_B200_ADJ_BLOCKS_PER_SM"))   // experiments (profiles/r02_adjoint.md)
        if (std::atoi(capenv) > 0 && std::atoi(capenv) < per_sm) per_sm = std::atoi(capenv);
    // one history slot per thread: every internal step of the forward pass is a data point, so a slot holds maxsteps + 2 points
    // (AMICI calls CVodeAdjInit with steps = maxsteps: one check-point interval, no re-integration)
    const int cap = o.maxsteps + 2;
    const size_t per_slot = sizeof(double) * ((size_t)cap * (1 + M::NX) + (size_t)(b.nt > 0 ? b.nt : 1) * M::NX) + sizeof(int) * (size_t)cap;
    size_t free_b = 0, total_b = 0;
    AB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    long grid = (long)sms * per_sm;
    const long want = ((long)b.N + ADJ_THREADS - 1) / ADJ_THREADS;
    if (want < grid) grid = want;
    const long fit = (long)((free_b / 2) / (per_slot * ADJ_THREADS));   // at most half of the free device memory
    if (fit < grid) grid = fit;
    if (grid < 1) {
        last_error() = "not enough device memory for the forward history of the adjoint method (maxsteps too large)";
        return AB_E_CUDA;
    }
    const size_t need = per_slot * (size_t)grid * ADJ_THREADS + 64;
    const int slot = block_workspace_acquire(dev, need);
    if (slot < 0) return AB_E_CUDA;
    double* ws_dev;
    cudaEvent_t ws_busy;
    {
        std::lock_guard<std::mutex> lock(block_pool().mu);
        ws_dev = block_pool().items[slot].dev;
        ws_busy = block_pool().items[slot].busy;
    }
    AB_CUDA(cudaMemsetAsync(queue_head, 0, sizeof(int), st));
    kern<<<(unsigned)grid, ADJ_THREADS, 0, st>>>(o, b, queue_head, pg, ws_dev, cap);
    AB_CUDA(cudaEventRecord(ws_busy, st));
    g_launches++;
    last_kernel() = std::string("adjoint_kernel<") + M::name() + ">";
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

}  // namespace amici_b200
#endif
