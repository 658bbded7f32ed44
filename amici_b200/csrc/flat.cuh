// flat.cuh — the condition driver as ONE warp-uniform loop: every iteration is one step ATTEMPT of every lane group of the
// warp that is integrating, followed by whatever else a group has to do (write a timepoint, finish a condition, load the
// next one, enter the next solver call).
//
// Why (profiles/r02_*): the step loop is bound by instruction supply and dependent-issue latency with two warps per
// scheduler, and the sensitivity columns of one condition fill 11-18 of 32 lanes.  With the nested driver
// (Simulation::run -> Bdf::solve_to -> step -> attempt) two conditions that share a warp (G = 16) diverge at the first
// data-dependent loop exit and run as two serialised instruction streams.  Here the only loop is the outer one, so the
// groups of a warp re-converge at its top before every attempt and execute the attempt — > 95 % of all instructions — as
// one stream: one instruction fetch and issue serves two conditions.  Inside the attempt the groups still take their own
// branches (order q, number of Newton iterations, a linear-solver setup); those re-converge at the end of the branch.
//
// It is the same solver: the pieces (BdfControl::solve_begin / step_pre / step_begin / attempt / step_finish / step_post,
// Simulation::load_condition / datapoint / finish) are shared with the nested driver, which stays in use for models with
// events and for the pre-equilibration / presimulation variants.  tests/cpu_emul runs both drivers on the host and the
// CPU suite checks that they agree bit for bit.
#ifndef AMICI_B200_FLAT_CUH
#define AMICI_B200_FLAT_CUH
#include "simulate.cuh"

namespace amici_b200 {

// votes over the whole warp (all 32 lanes take part: they are converged at the top of the flat loop)
AB_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}
AB_HD bool warp_all(bool p) {
#if defined(__CUDA_ARCH__)
    return __all_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}

// Step alignment of the warps of a block (same protocol as Bdf::step_sync_*, kept at warp level here): sync_pos[w] is the
// generation warp w has arrived at, INT_MAX while it is not integrating.
struct WarpAlign {
    volatile int* pos = nullptr;
    int n = 0, w = 0, gen = 0;
    bool member = false;
    AB_HD void join() {
#if defined(__CUDA_ARCH__)
        if (!pos || member) return;
        int m = 0x7fffffff;
        for (int v = 0; v < n; ++v) {
            const int pv = pos[v];
            if (v != w && pv < m) m = pv;
        }
        if (m != 0x7fffffff) gen = m;
        if ((threadIdx.x & 31) == 0) pos[w] = gen;
        __syncwarp();
        member = true;
#endif
    }
    AB_HD void leave() {
#if defined(__CUDA_ARCH__)
        if (!pos || !member) return;
        if ((threadIdx.x & 31) == 0) pos[w] = 0x7fffffff;
        __syncwarp();
        member = false;
#endif
    }
    AB_HD void arrive() {
#if defined(__CUDA_ARCH__)
        if (!pos) return;
        ++gen;
        if ((threadIdx.x & 31) == 0) pos[w] = gen;
        __syncwarp();
        for (;;) {
            int m = 0x7fffffff;
            for (int v = 0; v < n; ++v) {
                const int pv = pos[v];
                if (pv < m) m = pv;
            }
            if (m >= gen) break;
            __nanosleep(40);
        }
#endif
    }
};

template <class M, int G, int C>
struct FlatDriver {
    using Sim = Simulation<M, G, C, false>;
    using Shared = typename Sim::Shared;
    static_assert(M::NE == 0, "the flat driver is for event-free models");
    enum : int { ST_LOAD = 0, ST_NEXT = 1, ST_RUN = 2, ST_RETURNED = 3, ST_DONE = 4 };
    static constexpr int CONTINUE = Sim::Solver::CONTINUE;

    Sim sim;
    int state = ST_LOAD, it = 0, status = AB_SUCCESS, rc = 0;
    bool in_step = false;
    double t = 0.0;

    AB_HD FlatDriver(Shared& S, const SolverOptions& o, const BatchView& b) : sim(S, o, b, 0) {}

    // next condition index of this group, or -1 (device: one atomic on the queue head; host: the caller's counter)
    AB_HD int next_condition(int* queue_head) {
        int n = 0;
#if defined(__CUDA_ARCH__)
        if (sim.bdf.g.lane == 0) n = atomicAdd(queue_head, 1);
        n = sim.bdf.g.bcast(n, 0);
#else
        n = (*queue_head)++;
#endif
        return n < sim.b.N ? n : -1;
    }

    // One attempt of the step in progress (or the first attempt of a new step).  The groups of a warp run this converged.
    AB_HD void attempt() {
        auto& bdf = sim.bdf;
        int r = CONTINUE;
        if (!in_step) {
            r = bdf.step_pre();
            if (r == CONTINUE) {
                bdf.step_begin();
                in_step = true;
            }
        }
        if (r == CONTINUE) {
            const int a = bdf.attempt();
            if (a == AB_SUCCESS) {
                bdf.step_finish();
                in_step = false;
                r = bdf.step_post();
            } else if (a != CONTINUE) {
                bdf.tretlast = bdf.tn;
                in_step = false;
                r = a;
            }
        }
        if (r != CONTINUE) {
            bdf.flush_pending();
            rc = r;
            state = ST_RETURNED;
        }
    }

    // Everything between two solver calls (the body of Simulation::run's loop for a model without events or
    // pre-equilibration); returns when the group is integrating again or has no more work.
    template <class Done>
    AB_HD void tail(int* queue_head, Done&& condition_done) {
        auto& bdf = sim.bdf;
        auto& g = bdf.g;
        const int nt = sim.b.nt;
        for (;;) {
            if (state == ST_LOAD) {
                const int n = next_condition(queue_head);
                if (n < 0) { state = ST_DONE; return; }
                sim.n = n;
                sim.reset_objective();
                sim.load_condition(sim.b.k);
                t = sim.o.t0;
                status = AB_SUCCESS;
                it = 0;
                in_step = false;
                sim.write_initial_state();
                bdf.init(sim.o.t0);
                state = ST_NEXT;
            }
            if (state == ST_RETURNED) {
                // the solver call is over (Solver::run, reference src/solver.cpp:141-160); on failure the reference keeps
                // the last returned solution
                if (rc >= 0) t = bdf.tretlast;
                if (rc < 0) {
                    status = rc;
                    it = finish_condition(condition_done);
                    continue;
                }
                g.sync();
                if (g.lane == 0) {
                    AB_UNROLL
                    for (int r = 0; r < M::NX; ++r) sim.S.xfin[r] = sim.S.zn[0][r][0];
                }
                state = ST_NEXT;
            }
            if (state == ST_NEXT) {
                if (it >= nt) {
                    it = finish_condition(condition_done);
                    continue;
                }
                const double tout = sim.b.ts[it];
                if (!(tout > sim.o.t0) || !(t < tout) || next_t_too_close(t, tout)) {
                    sim.datapoint(it);
                    ++it;
                    continue;
                }
                const int r = bdf.solve_begin(tout, false);
                if (r == CONTINUE) {
                    state = ST_RUN;
                    return;
                }
                rc = r;
                state = ST_RETURNED;
            }
        }
    }

    template <class Done>
    AB_HD int finish_condition(Done&& condition_done) {
        sim.write_steady_state(false);
        if (status != AB_SUCCESS) sim.invalidate(it, ab_nan());
        sim.finish(status, t);
        sim.bdf.g.sync();
        condition_done(sim.n);
        state = ST_LOAD;
        return 0;
    }
};

// The flat loop of one lane group (all groups of a warp call this together).
template <class M, int G, int C, class Done>
AB_HD void flat_run(typename FlatDriver<M, G, C>::Shared& S, const SolverOptions& o, const BatchView& b, int* queue_head,
                    WarpAlign& align, Done&& condition_done) {
    using D = FlatDriver<M, G, C>;
    D d(S, o, b);
    for (;;) {
        // top of the loop: the warp is converged here
        const bool running = warp_any(d.state == D::ST_RUN);
        if (running) {
            align.join();
            align.arrive();
            if (d.state == D::ST_RUN) d.attempt();
        }
        if (warp_any(d.state != D::ST_RUN && d.state != D::ST_DONE)) {
            align.leave();   // the other warps of the block do not wait for outputs / loads
            if (d.state != D::ST_RUN && d.state != D::ST_DONE) d.tail(queue_head, condition_done);
        }
        if (warp_all(d.state == D::ST_DONE)) break;
    }
    align.leave();
}

}  // namespace amici_b200
#endif
