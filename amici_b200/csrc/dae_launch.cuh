// dae_launch.cuh — kernel, launch and registry entry of the DAE path (dae.cuh); one translation unit per model (tu/<model>.cu).
#ifndef AMICI_B200_DAE_LAUNCH_CUH
#define AMICI_B200_DAE_LAUNCH_CUH
#include <string>

#include "dae.cuh"
#include "registry.h"

namespace amici_b200 {

constexpr int DAE_THREADS = 32;   // BASELINE configs[0] is 4,096 conditions: 128 one-warp blocks cover the 148 SMs better than 64 two-warp ones

template <class M>
__global__ void __launch_bounds__(DAE_THREADS) dae_kernel(const ab_solver_options o, const ab_batch b, int* queue_head, const Progress pg) {
    if ((int)(threadIdx.x & 31) >= pg.lanes) return;   // small batches: see Progress::lanes
    for (;;) {
        const int n = atomicAdd(queue_head, 1);
        if (n >= b.N) break;
        DaeSimulation<M> sim(o, b, n);
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
int launch_dae_model(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if (o.sensi >= 1 && b.nplist > 0 && (o.sensi > 1 || o.sensi_meth == 2 || M::NE > 0)) {
        last_error() = "DAE models: first-order forward sensitivities of event-free models only";
        return AB_E_UNSUPPORTED;
    }
    if (b.k_preeq || o.posteq || b.t_presim || o.x_nonneg != 0ull || o.maxtime > 0.0 || o.reinit_states_sim || o.reinit_states_presim) {
        last_error() = "pre-/post-equilibration, presimulation, state re-initialisation, non-negativity and the wall-clock guard are not implemented for DAE models";
        return AB_E_UNSUPPORTED;
    }
    auto kern = dae_kernel<M>;
    int dev = 0, sms = 0, per_sm = 0;
    AB_CUDA(cudaGetDevice(&dev));
    AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DAE_THREADS, 0));
    if (per_sm < 1) {
        last_error() = "DAE kernel does not fit on an SM";
        return AB_E_CUDA;
    }
    long grid = (long)sms * per_sm;
    Progress pgl = pg;
    pgl.lanes = lanes_for_batch(b.N, grid * (DAE_THREADS / 32));
    const long per_block = (long)pgl.lanes * (DAE_THREADS / 32);
    const long want = ((long)b.N + per_block - 1) / per_block;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    AB_CUDA(cudaMemsetAsync(queue_head, 0, sizeof(int), st));
    kern<<<(unsigned)grid, DAE_THREADS, 0, st>>>(o, b, queue_head, pgl);
    g_launches++;
    last_kernel() = std::string("dae_kernel<") + M::name() + ">";
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

template <class M>
ModelEntry make_dae_entry() {
    static_assert(M::NE_SOLVER == 0 && M::NZ == 0 && M::NSPL == 0 && alg_mask_of<M>::value != 0ull,
                  "DAE path: spline-free models with algebraic states whose only events are time-triggered Heaviside switches");
    ModelEntry e;
    e.name = M::name();
    e.dims = ab_dims{M::NX, M::NP, M::NK, M::NY, M::NZ, M::NE, M::NSPL, 1, 1};
    e.nominal = &M::nominal;
    e.launch = &launch_dae_model<M>;
    return e;
}

}  // namespace amici_b200
#endif
