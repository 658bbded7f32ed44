// conditions.cpp — host side of the drop-in for amici::run_simulations when the caller holds a LIST of conditions
// (std::vector<ExpData*>): ab_run_conditions (include/amici_b200.h).
//
// What the reference does per condition on the host — ConditionContext (reference src/edata.cpp:409-581: every ExpData field
// that is set overrides the Model's, incl. Model::set_parameter_scale's re-scaling of the current parameters, src/model.cpp:513-527),
// the fan-out loop with `failfast` (src/amici.cpp:220-264) — is done here for the whole list at once: effective settings per
// condition, grouping of the conditions that can share one launch (same timepoints / plist / pscale / t0 / phases /
// re-initialisation sets), struct-of-arrays assembly, ab_run_batch per group, results scattered back to per-condition
// storage (the std::vector buffers of the caller's ReturnData objects).  Plain C++ on top of the C ABI: no CUDA here.
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/amici_b200.h"
#include "registry.h"

namespace {

using amici_b200::last_error;

struct Sizes {
    int nx, np, nk, ny, nz, nt, q, nme;
};

// per-condition element counts of the output fields of ab_batch (layouts: include/amici_b200.h)
struct DField { double* ab_batch::*ptr; size_t (*count)(const Sizes&); };
struct IField { int* ab_batch::*ptr; size_t (*count)(const Sizes&); };
#define CNT(expr) [](const Sizes& s) -> size_t { (void)s; return (size_t)(expr); }
const DField D_FIELDS[] = {
    {&ab_batch::llh, CNT(1)}, {&ab_batch::chi2, CNT(1)}, {&ab_batch::sllh, CNT(s.q)}, {&ab_batch::x, CNT(s.nt * s.nx)},
    {&ab_batch::sx, CNT(s.nt * s.q * s.nx)}, {&ab_batch::y, CNT(s.nt * s.ny)}, {&ab_batch::sy, CNT(s.nt * s.q * s.ny)},
    {&ab_batch::sigmay, CNT(s.nt * s.ny)}, {&ab_batch::ssigmay, CNT(s.nt * s.q * s.ny)}, {&ab_batch::res, CNT(s.nt * s.ny)},
    {&ab_batch::sres, CNT(s.nt * s.ny * s.q)}, {&ab_batch::FIM, CNT(s.q * s.q)}, {&ab_batch::x0, CNT(s.nx)}, {&ab_batch::sx0, CNT(s.q * s.nx)},
    {&ab_batch::xdot, CNT(s.nx)}, {&ab_batch::J, CNT(s.nx * s.nx)}, {&ab_batch::x_ss, CNT(s.nx)}, {&ab_batch::sx_ss, CNT(s.q * s.nx)},
    {&ab_batch::preeq_wrms, CNT(1)}, {&ab_batch::preeq_t, CNT(1)}, {&ab_batch::posteq_wrms, CNT(1)}, {&ab_batch::posteq_t, CNT(1)},
    {&ab_batch::z, CNT(s.nme * s.nz)}, {&ab_batch::sz, CNT(s.nme * s.q * s.nz)}, {&ab_batch::rz, CNT(s.nme * s.nz)},
    {&ab_batch::srz, CNT(s.nme * s.q * s.nz)}, {&ab_batch::sigmaz, CNT(s.nme * s.nz)},
};
const IField I_FIELDS[] = {
    {&ab_batch::status, CNT(1)}, {&ab_batch::numsteps, CNT(s.nt)}, {&ab_batch::numrhsevals, CNT(s.nt)}, {&ab_batch::numerrtestfails, CNT(s.nt)},
    {&ab_batch::numnonlinsolvconvfails, CNT(s.nt)}, {&ab_batch::order, CNT(s.nt)}, {&ab_batch::preeq_status, CNT(3)},
    {&ab_batch::preeq_numsteps, CNT(3)}, {&ab_batch::posteq_status, CNT(3)}, {&ab_batch::posteq_numsteps, CNT(3)},
};
#undef CNT

double unscale(double v, int s) { return s == 1 ? std::exp(v) : s == 2 ? std::pow(10.0, v) : v; }   // reference include/amici/misc.h:164-195
double scale(double v, int s) { return s == 1 ? std::log(v) : s == 2 ? std::log10(v) : v; }

// the settings of one condition after ConditionContext
struct Effective {
    std::vector<double> ts, p, k;
    std::vector<int> plist, pscale;
    double t0 = 0.0;
    bool preeq = false, presim = false, has_x0 = false, has_sx0 = false;
    unsigned long long reinit_sim = 0, reinit_presim = 0;
};

bool mask_of(const int* idx, int n, int nx, unsigned long long* out) {
    unsigned long long m = 0;
    for (int i = 0; i < n; ++i) {
        if (idx[i] < 0 || idx[i] >= nx) { last_error() = "Invalid state index given: " + std::to_string(idx[i]); return false; }
        if (idx[i] >= 64) { last_error() = "state re-initialisation is limited to models with nx <= 64"; return false; }
        m |= 1ull << idx[i];
    }
    *out = m;
    return true;
}

template <class T>
void append(std::string& key, const T* v, size_t n) {
    key.append(reinterpret_cast<const char*>(&n), sizeof n);
    key.append(reinterpret_cast<const char*>(v), n * sizeof(T));
}

}  // namespace

extern "C" int ab_run_conditions(const char* model, const ab_solver_options* opt, const ab_model_state* ms, int n, const ab_condition* conds,
                                 const ab_batch* outs, int failfast, int* n_batches) {
    if (!model || !opt || !ms || n < 0 || (n > 0 && (!conds || !outs))) return AB_E_BAD_ARGUMENT;
    ab_dims d;
    const int rc = ab_get_dims(model, &d);
    if (rc != AB_OK) return rc;
    if (n_batches) *n_batches = 0;
    if (!ms->p || (d.nk > 0 && !ms->k) || !ms->pscale || (ms->nplist > 0 && !ms->plist) || (ms->nt > 0 && !ms->ts)) return AB_E_BAD_ARGUMENT;
    const unsigned long long ALL = ~0ull;

    // ---- effective settings per condition (ConditionContext, reference src/edata.cpp:438-556) ----
    std::vector<Effective> eff(n);
    std::map<std::string, std::vector<int>> groups;
    std::vector<std::string> order;
    for (int i = 0; i < n; ++i) {
        const ab_condition& c = conds[i];
        Effective& e = eff[i];
        const bool none = c.none != 0;   // ExpData == nullptr: the model as it is
        const bool own_plist = !none && c.nplist > 0 && c.plist, own_pscale = !none && c.pscale, own_p = !none && c.p;
        e.plist.assign(own_plist ? c.plist : ms->plist, (own_plist ? c.plist : ms->plist) + (own_plist ? c.nplist : ms->nplist));
        for (int ip : e.plist)
            if (ip < 0 || ip >= d.np) { last_error() = "plist entry " + std::to_string(ip) + " outside [0, np)"; return AB_E_BAD_ARGUMENT; }
        e.pscale.assign(own_pscale ? c.pscale : ms->pscale, (own_pscale ? c.pscale : ms->pscale) + d.np);
        e.p.resize(d.np);
        for (int j = 0; j < d.np; ++j) {
            if (own_p) e.p[j] = c.p[j];
            // a new scale without new values: Model::set_parameter_scale keeps the UNSCALED parameters (reference src/model.cpp:513-527)
            else e.p[j] = own_pscale ? scale(unscale(ms->p[j], ms->pscale[j]), c.pscale[j]) : ms->p[j];
        }
        const bool own_k = !none && c.k;
        e.k.assign(own_k ? c.k : ms->k, (own_k ? c.k : ms->k) + d.nk);
        const bool own_ts = !none && c.nt > 0 && c.ts;
        e.ts.assign(own_ts ? c.ts : ms->ts, (own_ts ? c.ts : ms->ts) + (own_ts ? c.nt : ms->nt));
        e.t0 = none ? ms->t0 : c.t_start;   // ConditionContext always applies ExpData::t_start
        e.preeq = !none && c.k_preeq;
        e.presim = !none && c.t_presim > 0.0;
        e.has_x0 = !none && c.x0;
        e.has_sx0 = !none && c.sx0 && opt->sensi;
        // state re-initialisation (src/edata.cpp:492-548, src/model.cpp:1074-1092): the ExpData flag always replaces the model's
        // setting (flag off clears the model's index list); the ExpData's index lists apply where it carries the fixed
        // parameters of that phase and the flag is off
        unsigned long long m_model = 0;
        if (!mask_of(ms->reinit_idxs, ms->reinit_idxs ? ms->n_reinit : 0, d.nx, &m_model)) return AB_E_BAD_ARGUMENT;
        const bool model_all = ms->reinit_all && ms->n_reinit == 0;
        if (none) {
            e.reinit_sim = model_all ? ALL : m_model;
            e.reinit_presim = model_all ? ALL : 0;
        } else if (c.reinit_all) {
            e.reinit_sim = e.reinit_presim = ALL;
        } else {
            e.reinit_sim = 0;
            e.reinit_presim = 0;
            if (own_k && !mask_of(c.reinit_sim, c.reinit_sim ? c.n_reinit_sim : 0, d.nx, &e.reinit_sim)) return AB_E_BAD_ARGUMENT;
            if (c.k_presim && !mask_of(c.reinit_presim, c.reinit_presim ? c.n_reinit_presim : 0, d.nx, &e.reinit_presim)) return AB_E_BAD_ARGUMENT;
        }
        // conditions with the same key can share one launch
        std::string key;
        append(key, e.ts.data(), e.ts.size());
        append(key, e.plist.data(), e.plist.size());
        append(key, e.pscale.data(), e.pscale.size());
        append(key, &e.t0, 1);
        const unsigned long long flags[3] = {(unsigned long long)(e.preeq | e.presim << 1 | e.has_x0 << 2 | e.has_sx0 << 3), e.reinit_sim, e.reinit_presim};
        append(key, flags, 3);
        auto it = groups.find(key);
        if (it == groups.end()) {
            order.push_back(key);
            groups[key] = {i};
        } else {
            it->second.push_back(i);
        }
    }

    // ---- one ab_run_batch per group ----
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    struct GroupRun {
        std::vector<int> idx;
        Sizes s;
        std::vector<std::vector<double>> dbuf;
        std::vector<std::vector<int>> ibuf;
    };
    std::vector<GroupRun> runs;
    std::vector<int> status(n, 0);
    for (const std::string& key : order) {
        const std::vector<int>& idx = groups[key];
        const int N = (int)idx.size();
        const Effective& e0 = eff[idx[0]];
        Sizes s{d.nx, d.np, d.nk, d.ny, d.nz, (int)e0.ts.size(), opt->sensi ? (int)e0.plist.size() : 0, opt->nmaxevent};
        const size_t nty = (size_t)s.nt * s.ny, nmz = (size_t)s.nme * s.nz;
        std::vector<double> p((size_t)N * d.np), k((size_t)N * d.nk), my(N * nty, NaN), sg(N * nty, NaN), kpe, kps, tps, x0, sx0, mz, sz;
        bool any_mz = false;
        for (int j = 0; j < N; ++j) any_mz = any_mz || (!conds[idx[j]].none && (conds[idx[j]].mz || conds[idx[j]].sigmaz));
        if (e0.preeq) kpe.resize((size_t)N * d.nk);
        if (e0.presim) { kps.resize((size_t)N * d.nk); tps.resize(N); }
        if (e0.has_x0) x0.resize((size_t)N * d.nx);
        if (e0.has_sx0) sx0.resize((size_t)N * s.q * d.nx);
        if (any_mz && nmz) { mz.assign(N * nmz, NaN); sz.assign(N * nmz, NaN); }
        for (int j = 0; j < N; ++j) {
            const ab_condition& c = conds[idx[j]];
            const Effective& e = eff[idx[j]];
            std::copy(e.p.begin(), e.p.end(), p.begin() + (size_t)j * d.np);
            std::copy(e.k.begin(), e.k.end(), k.begin() + (size_t)j * d.nk);
            if (c.none) continue;
            if (c.my) std::copy(c.my, c.my + nty, my.begin() + j * nty);
            if (c.sigmay) std::copy(c.sigmay, c.sigmay + nty, sg.begin() + j * nty);
            if (e0.preeq) std::copy(c.k_preeq, c.k_preeq + d.nk, kpe.begin() + (size_t)j * d.nk);
            if (e0.presim) {
                tps[j] = c.t_presim;
                const double* src = c.k_presim ? c.k_presim : e.k.data();   // no dedicated parameters: the simulation's (src/edata.cpp:531-548)
                std::copy(src, src + d.nk, kps.begin() + (size_t)j * d.nk);
            }
            if (e0.has_x0) std::copy(c.x0, c.x0 + d.nx, x0.begin() + (size_t)j * d.nx);
            if (e0.has_sx0) std::copy(c.sx0, c.sx0 + (size_t)s.q * d.nx, sx0.begin() + (size_t)j * s.q * d.nx);
            if (!mz.empty() && c.mz) std::copy(c.mz, c.mz + nmz, mz.begin() + j * nmz);
            if (!sz.empty() && c.sigmaz) std::copy(c.sigmaz, c.sigmaz + nmz, sz.begin() + j * nmz);
        }
        ab_batch b;
        std::memset(&b, 0, sizeof b);
        b.N = N; b.nt = s.nt; b.nplist = (int)e0.plist.size();
        b.ts = e0.ts.data(); b.plist = e0.plist.data(); b.pscale = e0.pscale.data();
        b.p = p.data(); b.k = k.data(); b.my = my.data(); b.sigmay_in = sg.data();
        if (e0.preeq) b.k_preeq = kpe.data();
        if (e0.presim) { b.k_presim = kps.data(); b.t_presim = tps.data(); }
        if (e0.has_x0) b.x0_in = x0.data();
        if (e0.has_sx0) b.sx0_in = sx0.data();
        if (!mz.empty()) { b.mz = mz.data(); b.sigmaz_in = sz.data(); }
        // outputs: every field any condition of the group asks for
        runs.emplace_back();
        GroupRun& run = runs.back();
        run.idx = idx;
        run.s = s;
        std::vector<std::vector<double>>& dbuf = run.dbuf;
        std::vector<std::vector<int>>& ibuf = run.ibuf;
        dbuf.resize(sizeof D_FIELDS / sizeof D_FIELDS[0]);
        ibuf.resize(sizeof I_FIELDS / sizeof I_FIELDS[0]);
        for (size_t f = 0; f < dbuf.size(); ++f) {
            bool want = false;
            for (int j = 0; j < N; ++j) want = want || outs[idx[j]].*(D_FIELDS[f].ptr);
            const size_t cnt = D_FIELDS[f].count(s);
            if (want && cnt) { dbuf[f].assign(N * cnt, 0.0); b.*(D_FIELDS[f].ptr) = dbuf[f].data(); }
        }
        for (size_t f = 0; f < ibuf.size(); ++f) {
            bool want = f == 0;   // status always
            for (int j = 0; j < N; ++j) want = want || outs[idx[j]].*(I_FIELDS[f].ptr);
            const size_t cnt = I_FIELDS[f].count(s);
            if (want && cnt) { ibuf[f].assign(N * cnt, 0); b.*(I_FIELDS[f].ptr) = ibuf[f].data(); }
        }
        ab_solver_options o = *opt;
        o.t0 = e0.t0;
        o.reinit_states_sim = e0.reinit_sim;
        o.reinit_states_presim = e0.reinit_presim;
        const int brc = ab_run_batch(model, &o, &b, nullptr);
        if (brc != AB_OK) return brc;
        if (n_batches) ++*n_batches;
        for (int j = 0; j < N; ++j) status[idx[j]] = ibuf[0][j];
    }
    // ---- failfast (reference src/amici.cpp:246-260 with one thread): everything after the first failed condition is "not run".
    // The batch has been integrated as a whole; the buffers of the skipped conditions are left as the caller initialised them (a
    // freshly constructed ReturnData in the reference) and only their status is written.
    int first_failed = n;
    if (failfast)
        for (int i = n - 1; i >= 0; --i)
            if (status[i] < 0) first_failed = i;
    for (const GroupRun& run : runs) {
        for (size_t j = 0; j < run.idx.size(); ++j) {
            const int i = run.idx[j];
            const ab_batch& out = outs[i];
            if (i > first_failed) {
                if (out.status) *out.status = -1001;   // AMICI_NOT_RUN (reference include/amici/rdata.h:498)
                continue;
            }
            for (size_t f = 0; f < run.dbuf.size(); ++f) {
                const size_t cnt = D_FIELDS[f].count(run.s);
                if (out.*(D_FIELDS[f].ptr) && !run.dbuf[f].empty())
                    std::copy(run.dbuf[f].begin() + j * cnt, run.dbuf[f].begin() + (j + 1) * cnt, out.*(D_FIELDS[f].ptr));
            }
            for (size_t f = 0; f < run.ibuf.size(); ++f) {
                const size_t cnt = I_FIELDS[f].count(run.s);
                if (out.*(I_FIELDS[f].ptr) && !run.ibuf[f].empty())
                    std::copy(run.ibuf[f].begin() + j * cnt, run.ibuf[f].begin() + (j + 1) * cnt, out.*(I_FIELDS[f].ptr));
            }
        }
    }
    return AB_OK;
}
