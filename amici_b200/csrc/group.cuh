// Lane-group primitives.
//
// One trajectory is integrated by a group of G lanes of one warp (G = 1, 2, 4, 8, 16, 32);
// lane g owns the columns g, g+G, ... of the stacked (state | sensitivities) system.
// All scalar step-control logic is executed redundantly and identically by every lane
// of a group, so the only communication is: broadcast of the state column, max-reduction
// of the per-column WRMS norms, and visibility of the group's shared-memory slice.
// Different groups inside one warp may diverge freely (independent thread scheduling);
// every collective below names the group's own lane mask.
//
// With G == 1 all collectives are the identity, which is also how the same solver source is
// compiled for the host in tests/cpu_emul (an algorithm check that runs without a GPU).
#ifndef AMICI_B200_GROUP_CUH
#define AMICI_B200_GROUP_CUH
#include "model_common.cuh"

namespace amici_b200 {

template <int G>
struct Group {
    static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    int lane;        // lane index inside the group
    unsigned mask;   // warp mask of the group's lanes

    AB_HD Group() : lane(0), mask(1u) {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            const int wl = threadIdx.x & 31;
            lane = wl & (G - 1);
            mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
        }
#endif
    }

    // G == 16: two conditions share a warp.  When both halves execute a collective together (the common case: they follow
    // the same instruction stream until a data-dependent branch separates them), a collective with the half's own mask is
    // compiled into a divergence check plus a serialised WARPSYNC.COLLECTIVE region per half (~20 out-of-line instructions
    // per call site).  If the whole warp is here (__activemask() is full: all 32 lanes execute this very instruction, so all
    // 32 also reach the full-mask collective that follows), ONE full-warp instruction serves both halves instead: shuffles
    // address the own half, reductions are done once per half with the other half's input neutralised.
    AB_HD bool whole_warp() const {
#if defined(__CUDA_ARCH__)
        if (G == 16) return __activemask() == 0xffffffffu;
#endif
        return false;
    }
    AB_HD bool upper() const { return (mask & 1u) == 0u; }   // G == 16: this group is the upper half of its warp

    AB_HD void sync() const {
#if defined(__CUDA_ARCH__)
        if (G > 1) __syncwarp(mask);
#endif
    }

    // Maximum of a NON-NEGATIVE value (a norm) over the group.  Non-negative doubles order like their bit patterns, and
    // a NaN pattern is larger than every finite one and +inf, so two 32-bit integer warp reductions (REDUX) give a
    // NaN-propagating maximum that is identical on every lane — instead of a five-stage 64-bit shuffle tree.
    AB_HD double max(double v) const {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu, lo = (unsigned)__double2loint(v);
            unsigned mhi, mlo;
            if (whole_warp()) {
                const bool up = upper();
                const unsigned a = __reduce_max_sync(0xffffffffu, up ? 0u : hi), b = __reduce_max_sync(0xffffffffu, up ? hi : 0u);
                mhi = up ? b : a;
                const unsigned l2 = hi == mhi ? lo : 0u;
                const unsigned c = __reduce_max_sync(0xffffffffu, up ? 0u : l2), d = __reduce_max_sync(0xffffffffu, up ? l2 : 0u);
                mlo = up ? d : c;
            } else {
                mhi = __reduce_max_sync(mask, hi);
                mlo = __reduce_max_sync(mask, hi == mhi ? lo : 0u);
            }
            v = __hiloint2double((int)mhi, (int)mlo);
        }
#endif
        return v;
    }

    // Minimum of a non-negative value over the group (same bit-pattern argument as max)
    AB_HD double min(double v) const {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu, lo = (unsigned)__double2loint(v);
            unsigned mhi, mlo;
            if (whole_warp()) {
                const bool up = upper();
                const unsigned a = __reduce_min_sync(0xffffffffu, up ? 0xffffffffu : hi), b = __reduce_min_sync(0xffffffffu, up ? hi : 0xffffffffu);
                mhi = up ? b : a;
                const unsigned l2 = hi == mhi ? lo : 0xffffffffu;
                const unsigned c = __reduce_min_sync(0xffffffffu, up ? 0xffffffffu : l2), d = __reduce_min_sync(0xffffffffu, up ? l2 : 0xffffffffu);
                mlo = up ? d : c;
            } else {
                mhi = __reduce_min_sync(mask, hi);
                mlo = __reduce_min_sync(mask, hi == mhi ? lo : 0xffffffffu);
            }
            v = __hiloint2double((int)mhi, (int)mlo);
        }
#endif
        return v;
    }

    AB_HD double bcast(double v, int src) const {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            if (whole_warp()) v = __shfl_sync(0xffffffffu, v, src, G);   // width G: every half reads its own lane src
            else v = __shfl_sync(mask, v, src, G);
        }
#else
        (void)src;
#endif
        return v;
    }

    AB_HD int bcast(int v, int src) const {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            if (whole_warp()) v = __shfl_sync(0xffffffffu, v, src, G);
            else v = __shfl_sync(mask, v, src, G);
        }
#else
        (void)src;
#endif
        return v;
    }

    AB_HD bool any(bool p) const {
#if defined(__CUDA_ARCH__)
        if (G > 1) {
            if (whole_warp()) return (__ballot_sync(0xffffffffu, p) & mask) != 0u;
            return (__ballot_sync(mask, p) & mask) != 0u;
        }
#endif
        return p;
    }
};

}  // namespace amici_b200
#endif
