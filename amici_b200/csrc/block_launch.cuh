// block_launch.cuh — kernel, workspace and registry entry of the block solver (bdf_block.cuh); one translation unit per model.
#ifndef AMICI_B200_BLOCK_LAUNCH_CUH
#define AMICI_B200_BLOCK_LAUNCH_CUH
#include <cuda_runtime.h>

#include <mutex>

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

// Per-process workspace of the block solver (one slice of BV_COUNT vectors per resident block), kept between launches.
struct BlockWorkspace {
    std::mutex mu;
    double* dev = nullptr;
    size_t cap = 0;
    int device = -1;
};
inline BlockWorkspace& block_workspace() {
    static BlockWorkspace w;
    return w;
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
    BlockWorkspace& w = block_workspace();
    std::lock_guard<std::mutex> lock(w.mu);
    const size_t need = (size_t)grid * BV_COUNT * M::NX * NCP * sizeof(double);
    if (w.device != dev || need > w.cap) {
        // a previous launch may still be using the old buffer: cudaFree synchronises with it
        if (w.dev) cudaFree(w.dev);
        w.dev = nullptr;
        w.cap = 0;
        AB_CUDA(cudaMalloc((void**)&w.dev, need));
        w.cap = need;
        w.device = dev;
    }
    AB_CUDA(cudaMemsetAsync(queue_head, 0, sizeof(int), st));
    kern<<<(unsigned)grid, NCP, smem, st>>>(o, b, queue_head, pg, w.dev);
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
    const int npl = (o.sensi >= 1) ? b.nplist : 0;
    if (npl > M::NP) {
        last_error() = "nplist exceeds the number of model parameters";
        return AB_E_UNSUPPORTED;
    }
    if (b.FIM && !b.sres && npl > 0) {
        last_error() = "block-solver models compute FIM from sres: request sres together with FIM";
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
