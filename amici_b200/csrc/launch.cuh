// launch.cuh — solve kernel, per-model launch configuration and registry entry (one translation unit per model).
#ifndef AMICI_B200_LAUNCH_CUH
#define AMICI_B200_LAUNCH_CUH
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "registry.h"
#include "simulate.cuh"

#ifndef AB_MINBLOCKS
#define AB_MINBLOCKS 4
#endif
#ifndef AB_STEP_SYNC
#define AB_STEP_SYNC 1
#endif
#ifndef AB_THREADS_G32
#define AB_THREADS_G32 256
#endif
#ifndef AB_THREADS_SMALL
#define AB_THREADS_SMALL 64
#endif

namespace amici_b200 {

constexpr int SYNC_BYTES = 128;

template <class M, int G, int C, int THREADS, int MINBLOCKS, bool PREEQ>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) bdf_solve_kernel(const ab_solver_options o, const ab_batch b, int* queue_head, const Progress pg) {
    using Sim = Simulation<M, G, C, PREEQ>;
    using Shared = typename Sim::Shared;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [0, SYNC_BYTES): step-alignment generations of the block's warps (Bdf::step_sync_*), then one slice per group
    int* sync_pos = reinterpret_cast<int*>(smem_raw);
    Shared* slices = reinterpret_cast<Shared*>(smem_raw + SYNC_BYTES);
    Shared& S = slices[threadIdx.x / G];
    Group<G> g;
    constexpr bool STEP_SYNC = AB_STEP_SYNC && G == 32 && THREADS > 32;
    if (STEP_SYNC) {
        if (threadIdx.x < THREADS / 32) sync_pos[threadIdx.x] = 0x7fffffff;
        __syncthreads();
    }
    for (;;) {
        int n = 0;
        if (g.lane == 0) n = atomicAdd(queue_head, 1);
        n = g.bcast(n, 0);
        if (n >= b.N) break;
        Sim sim(S, o, b, n);
        if (STEP_SYNC) {
            sim.bdf.sync_pos = sync_pos;
            sim.bdf.sync_n = THREADS / 32;
            sim.bdf.sync_w = threadIdx.x / 32;
        }
        sim.run();
        g.sync();
        if (pg.done) {
            // Completion tracking for the streaming host path (ab_run_batch): conditions are handed out in index order, so
            // the sub-batches [c*chunk, (c+1)*chunk) complete roughly in order; the group that finishes the last condition
            // of a sub-batch raises its flag in mapped host memory and the host starts that sub-batch's D2H copies.
            __threadfence();   // every lane: its ReturnData stores are visible device-wide before the count below
            g.sync();
            if (g.lane == 0) {
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

// ---------------------------------------------------------------------------------------------
// per-model launch configuration
// ---------------------------------------------------------------------------------------------
constexpr int pow2_at_least(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

template <class M>
struct Config {
    static constexpr int COLS = M::NP + 1;
#ifdef AB_FORCE_G   // experiment only (tools/build_variant.py)
    static constexpr int G = AB_FORCE_G;
    static constexpr int C = 1;
#else
    static constexpr int G = pow2_at_least(COLS);
    static constexpr int C = (COLS + G - 1) / G;
#endif
    // One warp per condition (G == 32): all resident warps of an SM in ONE block, so that they can align their steps
    // (Bdf::step_sync_*).  Otherwise small blocks: shared-memory granularity decides occupancy.
    static constexpr int THREADS = (G == 32) ? AB_THREADS_G32 : AB_THREADS_SMALL;
    static constexpr int MINBLOCKS = (G == 32) ? 1 : AB_MINBLOCKS;  // register cap = 65536 / (MINBLOCKS * THREADS)
};

template <class M, int G, int C, int THREADS, int MINBLOCKS, bool PREEQ>
int launch_variant(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    using Sim = Simulation<M, G, C, PREEQ>;
    auto kern = bdf_solve_kernel<M, G, C, THREADS, MINBLOCKS, PREEQ>;
    constexpr int groups = THREADS / G;
    const size_t smem = SYNC_BYTES + sizeof(typename Sim::Shared) * groups;
    int dev = 0, sms = 0;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) {
        last_error() = "kernel does not fit on an SM";
        return AB_E_CUDA;
    }
    long want = ((long)b.N + groups - 1) / groups;
    long grid = (long)sms * per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    AB_CUDA(cudaMemsetAsync(queue_head, 0, sizeof(int), st));
    kern<<<(unsigned)grid, THREADS, smem, st>>>(o, b, queue_head, pg);
    g_launches++;
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

// PREEQ = true variants (conditions with ExpData::fixed_parameters_pre_equilibration) are instantiated in their own
// translation unit (tu/<model>_preeq.cu) to keep the build parallel.
template <class M, bool PREEQ>
int launch_model_impl(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if (o.sensi >= 1 && b.nplist > 0) {
        if (b.nplist + 1 > Config<M>::COLS || b.nplist + 1 > Config<M>::G * Config<M>::C) {
            last_error() = "nplist exceeds the number of model parameters";
            return AB_E_UNSUPPORTED;
        }
        return launch_variant<M, Config<M>::G, Config<M>::C, Config<M>::THREADS, Config<M>::MINBLOCKS, PREEQ>(o, b, queue_head, pg, st);
    }
    // no sensitivities: one thread per condition
    return launch_variant<M, 1, 1, 32, 1, PREEQ>(o, b, queue_head, pg, st);
}

template <class M>
int launch_model_preeq(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st);

template <class M>
int launch_model(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if ((o.reinit_states_sim | o.reinit_states_presim) && !M::X0_FIXED_REINIT_ALLOWED) {
        // Model::set_reinitialize_fixed_parameter_initial_states (reference src/model.cpp:1074-1081)
        last_error() = "State reinitialization cannot be enabled for this model: initial states depending on fixed parameters also depend on parameters";
        return AB_E_UNSUPPORTED;
    }
    if (b.k_preeq || o.posteq || b.t_presim) return launch_model_preeq<M>(o, b, queue_head, pg, st);   // the variant with the steady-state code
    return launch_model_impl<M, false>(o, b, queue_head, pg, st);
}

template <class M>
ModelEntry make_entry() {
    ModelEntry e;
    e.name = M::name();
    e.dims = ab_dims{M::NX, M::NP, M::NK, M::NY, M::NZ, M::NE, M::NSPL, Config<M>::G, Config<M>::C};
    e.nominal = &M::nominal;
    e.launch = &launch_model<M>;
    return e;
}

}  // namespace amici_b200
#endif
