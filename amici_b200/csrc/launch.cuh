// launch.cuh — solve kernel, per-model launch configuration and registry entry (one translation unit per model).
#ifndef AMICI_B200_LAUNCH_CUH
#define AMICI_B200_LAUNCH_CUH
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <string>

#include "registry.h"
#include "flat.cuh"

#ifndef AB_MINBLOCKS
#define AB_MINBLOCKS 4
#endif
#ifndef AB_STEP_SYNC
#define AB_STEP_SYNC 1
#endif
#ifndef AB_THREADS_G32
#define AB_THREADS_G32 256
#endif
#ifndef AB_STEP_SYNC_G16
#define AB_STEP_SYNC_G16 0   // opportunistic alignment of warps that hold two conditions (Bdf::step_sync_arrive): measured neutral
#endif
#ifndef AB_THREADS_SMALL
#define AB_THREADS_SMALL 64
#endif
#ifndef AB_FLAT
// 1: event-free models without pre-equilibration run the flat driver (flat.cuh: one attempt per iteration of a warp-uniform
// loop, so that the step alignment also works for two conditions per warp).  Measured on B200 (profiles/r02_variants.md) the
// nested driver is faster at the current code size, so it stays the default; both are covered by the tests.
#define AB_FLAT 0
#endif
#ifndef AB_COMPACT
#define AB_COMPACT 1     // leave identically-zero sensitivity columns out of the integration when that halves the lane group
#endif
#ifndef AB_THREADS_G16
#define AB_THREADS_G16 256
#endif

namespace amici_b200 {

constexpr int SYNC_BYTES = 128;

// Completion tracking for the streaming host path (ab_run_batch): conditions are handed out in index order, so the
// sub-batches [c*chunk, (c+1)*chunk) complete roughly in order; the group that finishes the last condition of a sub-batch
// raises its flag in mapped host memory and the host starts that sub-batch's D2H copies.
template <int G>
__device__ __forceinline__ void report_done(const Progress& pg, const Group<G>& g, const int N, const int n) {
    if (!pg.done) return;
    __threadfence();   // every lane: its ReturnData stores are visible device-wide before the count below
    g.sync();
    if (g.lane == 0) {
        const int c = n / pg.chunk;
        const int size = (N - c * pg.chunk < pg.chunk) ? N - c * pg.chunk : pg.chunk;
        if (atomicAdd(pg.done + c, 1) == size - 1) {
            __threadfence_system();
            *reinterpret_cast<volatile int*>(pg.host_flags + c) = 1;
        }
    }
}

#if AB_SLDET_SVC
// cvSLdet service warps (AB_SLDET_SVC; the last warpgroup of a block of 16-lane groups): four lanes serve one lane group, the
// three orders in three of them (bdf_sldet_xchg).  A group posts (order, ring head) in its slice's mailbox (Bdf::sldet_remote),
// the ring buffer itself is read from the group's slice, the return code goes back through the same word.
template <class Shared, int NG>
__device__ void sldet_service(Shared* slices, volatile int* done, const int st) {
    constexpr int NGW = (NG + 3) / 4;            // lane groups served by one of the four warps
    const int w = st >> 5, wl = st & 31, sub = wl >> 2;
    const int gi = w * NGW + sub;
    const bool mine = sub < NGW && gi < NG;
#if defined(AB_SLDET_SVC_LOCAL) && AB_SLDET_SVC_LOCAL == 2
    return;
#endif
    Group<4> g4;
    Shared& S = slices[mine ? gi : 0];
    volatile int* svc = &S.svc;
    for (;;) {
        int v = 0;
        if (mine && g4.lane == 0) v = *svc;
        v = __shfl_sync(0xffffffffu, v, wl & ~3);
        const bool req = (v & SVC_REQ) != 0;
        if (__any_sync(0xffffffffu, req)) {
            // the groups of this warp that have a request pending evaluate the routine together
            if (req) {
                __threadfence_block();
                const int rc = bdf_sldet_xchg<4>(g4, S.ssdat, S.sld, v & 0xff, (v >> 8) & 0xff);
                g4.sync();
                if (g4.lane == 0) *svc = SVC_ANS | (rc + 16);
            }
            __syncwarp();
        } else {
            if (*done >= NG) break;
            if (AB_SLDET_SVC_SLEEP > 0) __nanosleep(AB_SLDET_SVC_SLEEP);
        }
    }
}

#endif

template <int G>
struct SldetSvc { static constexpr bool value = AB_SLDET_SVC && G == 16; };

template <class M, int G, int C, int THREADS, int MINBLOCKS, bool PREEQ>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) bdf_solve_kernel(const ab_solver_options o, const ab_batch b, int* queue_head, const Progress pg) {
    using Sim = Simulation<M, G, C, PREEQ>;
    using Shared = typename Sim::Shared;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [0, SYNC_BYTES): step-alignment generations of the block's warps (Bdf::step_sync_*), then one slice per group
    int* sync_pos = reinterpret_cast<int*>(smem_raw);
    Shared* slices = reinterpret_cast<Shared*>(smem_raw + SYNC_BYTES);
    constexpr bool SVC = SldetSvc<G>::value;
    constexpr int CT = SVC ? THREADS - 128 : THREADS;   // threads that integrate
    Shared& S = slices[(SVC && threadIdx.x >= CT) ? 0 : threadIdx.x / G];
    Group<G> g;
    constexpr bool STEP_SYNC = AB_STEP_SYNC && (G == 32 || (G == 16 && AB_STEP_SYNC_G16)) && THREADS > 32;
    static_assert(!(SVC && STEP_SYNC), "the service warp does not take part in the step alignment");
    if (G == 1 && (int)(threadIdx.x & 31) >= pg.lanes) return;   // small batches: see Progress::lanes
    if (STEP_SYNC) {
        if (threadIdx.x < THREADS / 32) sync_pos[threadIdx.x] = 0x7fffffff;
        __syncthreads();
    }
#if AB_SLDET_SVC
    if (SVC) {
        if (threadIdx.x < CT / G) slices[threadIdx.x].svc = 0;
        if (threadIdx.x == 0) sync_pos[31] = 0;
        __syncthreads();
        // register file per scheduler: two integrating warps + one service warp; the service warpgroup hands most of its
        // share to the integrating ones (launch bound 168 -> 232 + 232 + 40)
        if (threadIdx.x >= CT) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            sldet_service<Shared, CT / G>(slices, sync_pos + 31, (int)threadIdx.x - CT);
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    }
#endif
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
        report_done<G>(pg, g, b.N, n);
    }
    if (SVC) {
        g.sync();
        if (g.lane == 0) atomicAdd(sync_pos + 31, 1);
    }
}

// Event-free models without pre-equilibration: the flat driver (flat.cuh), one step attempt per iteration of a warp-uniform loop.
template <class M, int G, int C, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) bdf_flat_kernel(const ab_solver_options o, const ab_batch b, int* queue_head, const Progress pg) {
    using Shared = typename FlatDriver<M, G, C>::Shared;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* sync_pos = reinterpret_cast<int*>(smem_raw);
    Shared* slices = reinterpret_cast<Shared*>(smem_raw + SYNC_BYTES);
    Shared& S = slices[threadIdx.x / G];
    constexpr bool STEP_SYNC = AB_STEP_SYNC && G >= 16 && THREADS > 32;
    WarpAlign align;
    if (STEP_SYNC) {
        if (threadIdx.x < THREADS / 32) sync_pos[threadIdx.x] = 0x7fffffff;
        __syncthreads();
        align.pos = sync_pos;
        align.n = THREADS / 32;
        align.w = threadIdx.x / 32;
    }
    Group<G> g;
    flat_run<M, G, C>(S, o, b, queue_head, align, [&](const int n) { report_done<G>(pg, g, b.N, n); });
}

// ---------------------------------------------------------------------------------------------
// per-model launch configuration
// ---------------------------------------------------------------------------------------------
constexpr int pow2_at_least(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

// COMPACT: one column per parameter whose state sensitivity is not identically zero (bdf_core.cuh Layout)
template <class M, bool COMPACT = false>
struct Config {
    static constexpr int COLS = COMPACT ? M::NP_DYN + 1 : M::NP + 1;
#ifdef AB_FORCE_G   // experiment only (tools/build_variant.py)
    static constexpr int G = AB_FORCE_G;
    static constexpr int C = 1;
#else
    static constexpr int G = pow2_at_least(COLS);
    static constexpr int C = (COLS + G - 1) / G;
#endif
    // One warp per condition (G == 32), or two (G == 16): all resident warps of an SM in ONE block, so that they can align
    // their steps (Bdf::step_sync_*, WarpAlign).  Otherwise small blocks: shared-memory granularity decides occupancy.
    static constexpr int THREADS = (G == 32) ? AB_THREADS_G32 : (G == 16) ? AB_THREADS_G16 + (AB_SLDET_SVC ? 128 : 0) : AB_THREADS_SMALL;
    static constexpr int MINBLOCKS = (G >= 16) ? 1 : AB_MINBLOCKS;  // register cap = 65536 / (MINBLOCKS * THREADS)
};
// compaction pays when it halves the lane group (JAK-STAT: 18 columns in 32 lanes -> 11 in 16, two conditions per warp)
template <class M>
constexpr bool compact_pays() {
    return AB_COMPACT && M::SX_STATIC_MASK != 0ull && Config<M, true>::G * Config<M, true>::C < Config<M, false>::G * Config<M, false>::C &&
           Config<M, true>::G * Config<M, true>::C < M::NP + 1;
}

template <class M, int G, int C, int THREADS, int MINBLOCKS, bool PREEQ>
auto select_kernel() {
    if constexpr (AB_FLAT && !PREEQ && M::NE == 0) return bdf_flat_kernel<M, G, C, THREADS, MINBLOCKS>;
    else return bdf_solve_kernel<M, G, C, THREADS, MINBLOCKS, PREEQ>;
}

template <class M, int G, int C, int THREADS, int MINBLOCKS, bool PREEQ>
int launch_variant(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    using Sim = Simulation<M, G, C, PREEQ>;
    auto kern = select_kernel<M, G, C, THREADS, MINBLOCKS, PREEQ>();
    constexpr int groups = (SldetSvc<G>::value ? THREADS - 128 : THREADS) / G;
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
    if (const char* cap = std::getenv("AMICI_B200_BLOCKS_PER_SM"))   // experiments only (profiles/r02_variants.md)
        if (std::atoi(cap) > 0 && std::atoi(cap) < per_sm) per_sm = std::atoi(cap);
    long want = ((long)b.N + groups - 1) / groups;
    long grid = (long)sms * per_sm;
    Progress pgl = pg;
    if (G == 1) {
        pgl.lanes = lanes_for_batch(b.N, grid * (THREADS / 32));
        want = ((long)b.N + pgl.lanes * (THREADS / 32) - 1) / (pgl.lanes * (THREADS / 32));
    }
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    AB_CUDA(cudaMemsetAsync(queue_head, 0, sizeof(int), st));
    kern<<<(unsigned)grid, THREADS, smem, st>>>(o, b, queue_head, pgl);
    g_launches++;
    last_kernel() = std::string((AB_FLAT && !PREEQ && M::NE == 0) ? "bdf_flat_kernel<" : "bdf_solve_kernel<") + M::name() + "," + std::to_string(G) + "," +
                    std::to_string(C) + "," + std::to_string(THREADS) + "," + std::to_string(MINBLOCKS) + "," + (PREEQ ? "preeq" : "plain") + ">";
    AB_CUDA(cudaGetLastError());
    return AB_OK;
}

// PREEQ = true variants (conditions with ExpData::fixed_parameters_pre_equilibration) are instantiated in their own
// translation unit (tu/<model>_preeq.cu) to keep the build parallel.
template <class M, bool PREEQ>
int launch_model_impl(const ab_solver_options& o, const ab_batch& b, int* queue_head, const Progress& pg, cudaStream_t st) {
    if (o.sensi >= 1 && b.nplist > 0) {
        if (b.nplist + 1 > Config<M>::COLS) {
            last_error() = "nplist exceeds the number of model parameters";
            return AB_E_UNSUPPORTED;
        }
        if constexpr (!PREEQ && compact_pays<M>()) {
            // given initial sensitivities may be non-zero for any parameter: every column is integrated then
            using CC = Config<M, true>;
            if (!b.sx0_in) return launch_variant<M, CC::G, CC::C, CC::THREADS, CC::MINBLOCKS, PREEQ>(o, b, queue_head, pg, st);
        }
        if (b.nplist + 1 > Config<M>::G * Config<M>::C) {
            last_error() = "nplist exceeds the lane group of this build";
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
    if (M::NE > 0 && o.posteq) {
        // the reference has no working path here either: SteadyStateProblem(is_preeq = false) of a model with events ends in
        // std::out_of_range ("vector::_M_range_check") for every condition, status AMICI_ERROR (observed with the compiled
        // reference on model_events_py, profiles/r02_events_phases.md)
        last_error() = "post-equilibration (t = inf output points) of models with events is not implemented";
        return AB_E_UNSUPPORTED;
    }
    if (b.k_preeq || o.posteq || b.t_presim) return launch_model_preeq<M>(o, b, queue_head, pg, st);   // the variant with the steady-state code
    // state non-negativity / wall-clock guard (AB_FULL translation units) and the
    // staggered corrector: compiled into that variant too (event models carry them in their only kernel)
    if ((o.x_nonneg != 0ull || o.maxtime > 0.0) && M::NE == 0) return launch_model_preeq<M>(o, b, queue_head, pg, st);
    if (o.sensi >= 1 && o.ism == 2 && b.nplist > 0 && M::NE == 0) return launch_model_preeq<M>(o, b, queue_head, pg, st);
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
