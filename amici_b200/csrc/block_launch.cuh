// block_launch.cuh — kernel, workspace and registry entry of the block solver (bdf_block.cuh); one translation unit per model.
#ifndef AMICI_B200_BLOCK_LAUNCH_CUH
#define AMICI_B200_BLOCK_LAUNCH_CUH
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "bdf_block.cuh"
#include "registry.h"

#ifndef AB_BLOCK_MINB
#define AB_BLOCK_MINB 2
#endif

namespace amici_b200 {

template <class M, int NCP>
__global__ void __launch_bounds__(NCP, AB_BLOCK_MINB) block_solve_kernel(const ab_solver_options o, const ab_batch b, int* queue_head,
                                                             const Progress pg, double* workspace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlockShared<M>& S = *reinterpret_cast<BlockShared<M>*>(smem_raw);
    __shared__ int next;
    double* ws = workspace + (size_t)blockIdx.x * BV_COUNT * M::NX * NCP;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next = atomicAdd(queue_head, 1);
        __syncthreads();
        const int n = next;
        if (n >= b.N) break;
        BlockSimulation<M, NCP> sim(S, o, b, ws, n);
        sim.run();
        if (pg.done) {   // completion tracking for the streaming host path (see bdf_solve_kernel)
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                const int c = n / pg.chunk;
                const int size = (b.N - c * pg.chunk < pg.chunk) ? b.N - c * pg.chunk : pg.chunk;
                if (atomicAdd(pg.done + c, 1) == size - 1) {
                    __threadfence_system();
                    *reinterpret_cast<volatile int*>(pg.host_flags + c) = 1;
                }
            }
        }
    }
}

// Workspaces of the block solver (one slice of BV_COUNT vectors per resident block).  ab_solve_device is asynchronous on the
// caller's stream, so two launches may be in flight at once (two DeviceBatch objects, two models): every launch takes its OWN
// buffer from a pool and marks it busy with an event recorded behind its kernel; a buffer is handed out again only when that
// event has completed, and nothing is freed while a kernel may still use it.
struct BlockWorkspace {
    double* dev = nullptr;
    size_t cap = 0;
    int device = -1;
    cudaEvent_t busy = nullptr;
};
struct BlockWorkspacePool {
    std::mutex mu;
    std::vector<BlockWorkspace> items;
};
inline BlockWorkspacePool& block_pool() {
    static BlockWorkspacePool p;
    return p;
}
// returns the index of a free buffer of at least `need` bytes on device `dev` (allocating one if necessary), -1 on error
inline int block_workspace_acquire(int dev, size_t need) {
    BlockWorkspacePool& pool = block_pool();
    std::lock_guard<std::mutex> lock(pool.mu);
    for (size_t i = 0; i < pool.items.size(); ++i) {
        BlockWorkspace& w = pool.items[i];
        if (w.device == dev && w.cap >= need && cudaEventQuery(w.busy) == cudaSuccess) return (int)i;
    }
    BlockWorkspace w;
    if (!cuda_ok(cudaMalloc((void**)&w.dev, need), "cudaMalloc (block workspace)")) return -1;
    if (!cuda_ok(cudaEventCreateWithFlags(&w.busy, cudaEventDisableTiming), "cudaEventCreate")) return -1;
    w.cap = need;
    w.device = dev;
    pool.items.push_back(w);
    return (int)pool.items.size() - 1;
}

template <class M, int NCP>
int launch_block_variant(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    auto kern = block_solve_kernel<M, NCP>;
    const size_t smem = sizeof(BlockShared<M>);
    int dev = 0, sms = 0, per_sm = 0;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NCP, smem));
    if (per_sm < 1) {
        last_error() = "block kernel does not fit on an SM";
        return AB_E_CUDA;
    }
    long grid = (long)sms * per_sm;
    if (b.N < grid) grid = b.N;
    if (grid < 1) grid = 1;
    const size_t need = (size_t)grid * BV_COUNT * M::NX * NCP * sizeof(double);
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
    kern<<<(unsigned)grid, NCP, smem, st>>>(o, b, queue_head, pg, ws_dev);
    AB_CUDA(cudaEventRecord(ws_busy, st));   // the buffer is taken until this launch has finished
    g_launches++;
    last_kernel() = std::string("block_solve_kernel<") + M::name() + "," + std::to_string(NCP) + ">";
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

template <class M>
int launch_block_model(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if (b.k_preeq || o.posteq || b.t_presim || o.reinit_states_sim || o.reinit_states_presim) {
        last_error() = "pre-/post-equilibration and presimulation are not implemented for block-solver models";
        return AB_E_UNSUPPORTED;
    }
    if (o.x_nonneg != 0ull || o.maxtime > 0.0) {
        last_error() = "state non-negativity and the wall-clock guard are not implemented for block-solver models";
        return AB_E_UNSUPPORTED;
    }
    const int npl = (o.sensi >= 1) ? b.nplist : 0;
    if (npl > 0 && o.ism != 1) {
        last_error() = "block-solver models implement the simultaneous corrector only";
        return AB_E_UNSUPPORTED;
    }
    if (npl > M::NP) {
        last_error() = "nplist exceeds the number of model parameters";
        return AB_E_UNSUPPORTED;
    }
    // one thread per column, padded to a compiled block size
    constexpr int FULL = ((M::NP + 1 + 31) / 32) * 32;
    static_assert(FULL <= 320, "block solver supports at most 319 sensitivity columns");
    const int ncols = 1 + npl;
    if (ncols <= 32) return launch_block_variant<M, 32>(o, b, queue_head, pg, st);
    if (FULL > 96 && ncols <= 96) return launch_block_variant<M, 96>(o, b, queue_head, pg, st);
    return launch_block_variant<M, FULL>(o, b, queue_head, pg, st);
}

template <class M>
ModelEntry make_block_entry() {
    ModelEntry e;
    e.name = M::name();
    e.dims = ab_dims{M::NX, M::NP, M::NK, M::NY, M::NZ, M::NE, M::NSPL, 0, 0};   // group_lanes = 0: one block per condition
    e.nominal = &M::nominal;
    e.launch = &launch_block_model<M>;
    return e;
}

}  // namespace amici_b200
#endif
