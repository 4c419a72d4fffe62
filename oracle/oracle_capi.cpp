// ORACLE — TEST INFRASTRUCTURE ONLY (see dlwgd_oracle.hpp).  C entry points for ctypes.
// Mirrors the reference's user-level calls so that parity tests read like test/tests.jl:
//   DLWGD(nw, df, λ, μ, η, nt)           model.jl:22-28      -> orc_new
//   logpdf!(d, p) / logpdf!(d[i], p)     profile.jl:56-60    -> orc_logpdf_full / orc_logpdf_from
//   logpdfroot(d[1], p)                  profile.jl:62-63    -> orc_logpdfroot
//   gradient / gradient_cr               profile.jl:71-75    -> orc_gradient / orc_gradient_cr
//   set!/rev!/extend!/shrink!            profile.jl:79-121   -> orc_set / orc_rev / orc_extend / orc_shrink
//   update!, addwgd!, addwgt!, removewgd!, setrates!          node.jl:99-116, model.jl:69-183
#include "dlwgd_oracle.hpp"

#include <chrono>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

struct OrcHandle {
    NwTree tree;
    Model<double>* d = nullptr;
    PArray p;
    std::string err;
    ~OrcHandle() { delete d; }
};

#define ORC_TRY(h, failret, ...)                         \
    try { __VA_ARGS__ }                                  \
    catch (const std::exception& e) { (h)->err = e.what(); return failret; }

static double weighted_sum(const std::vector<double>& l, const PArray& p) {
    double s = 0.0;
    bool first = true;
    for (size_t f = 0; f < l.size(); ++f) {                 // mapreduce(x -> x.n * logpdf!(...), +, p)
        double v = (double)p.p[f].n * l[f];
        s = first ? v : s + v;
        first = false;
    }
    return s;
}

extern "C" {

// counts: nfam × ntaxa row-major, names: ntaxa C strings (column names of the DataFrame).
// nfam == 0 gives the data-less constructor (model.jl:30-36, m = 3, mock profile).
OrcHandle* orc_new(const char* nw, const char* const* names, int ntaxa, const int64_t* counts, int nfam,
                   double lam, double mu, double eta, int nodetype) {
    auto* h = new OrcHandle();
    try {
        readnw(nw, h->tree);
        int m = 3;
        if (nfam > 0) {
            std::vector<std::string> nm(names, names + ntaxa);
            std::vector<int64_t> c(counts, counts + (size_t)nfam * ntaxa);
            int mm = build_profiles(h->tree, nm, c, nfam, h->p);
            m = std::max(3, mm);
        }
        h->d = initmodel<double>(h->tree, eta, lam, mu, m, (NodeType)nodetype);
        h->d->setall();
    } catch (const std::exception& e) { h->err = e.what(); }
    return h;
}
void orc_free(OrcHandle* h) { delete h; }
const char* orc_last_error(OrcHandle* h) { return h->err.c_str(); }
int orc_ok(OrcHandle* h) { return h->d != nullptr && h->err.empty(); }

int orc_nnodes(OrcHandle* h) { return (int)h->d->nodes.size(); }
int orc_maxid(OrcHandle* h) { return h->d->maxid(); }
int orc_npatterns(OrcHandle* h) { return h->p.empty() ? 0 : (int)h->p.p.size(); }
int orc_m(OrcHandle* h) { return h->d->m; }
int orc_ncols(OrcHandle* h) { return h->p.empty() ? 0 : (int)h->p.p[0].xp.size(); }
int orc_nwgd(OrcHandle* h) { return h->d->nwgd(); }
int orc_ne(OrcHandle* h) { return h->d->ne(); }

// Export the model graph, indexed by id-1 up to maxid (absent ids: kind = -1).
// child_pos[i] = position of node i in its parent's child vector (iteration order of n.c).
int orc_export(OrcHandle* h, int cap, int* parent, int* kind, int* child_pos, double* t, double* lam, double* mu,
               double* q, double* eta, double* eps0, double* eps1) {
    int mx = h->d->maxid();
    if (cap < mx) return -1;
    for (int i = 0; i < mx; ++i) { parent[i] = 0; kind[i] = -1; child_pos[i] = 0; t[i] = lam[i] = mu[i] = q[i] = eps0[i] = eps1[i] = kNaN; }
    for (auto& kv : h->d->nodes) {
        auto* n = kv.second;
        int i = kv.first - 1;
        parent[i] = n->p ? n->p->i : 0;
        kind[i] = (int)n->kind;
        if (n->p) child_pos[i] = (int)(std::find(n->p->c.begin(), n->p->c.end(), n) - n->p->c.begin());
        t[i] = n->t; lam[i] = n->lam; mu[i] = n->mu; q[i] = n->q;
        eps0[i] = n->eps[0]; eps1[i] = n->eps[1];
    }
    *eta = h->d->root()->eta;
    return mx;
}
// W of node i (column-major m×m, as the reference stores it) into out[m*m].
int orc_get_W(OrcHandle* h, int i, double* out) {
    ORC_TRY(h, -1, { auto* n = h->d->at(i); std::copy(n->W.begin(), n->W.end(), out); return h->d->m; })
}
// pattern multiplicities and the (proposal) extended counts xp, ncols × npatterns (node fastest,
// i.e. permutedims(M) of model.jl:48)
int orc_get_profile(OrcHandle* h, int64_t* weight, int64_t* xp) {
    if (h->p.empty()) return 0;
    int nc = (int)h->p.p[0].xp.size();
    for (size_t f = 0; f < h->p.p.size(); ++f) {
        weight[f] = h->p.p[f].n;
        for (int c = 0; c < nc; ++c) xp[f * nc + c] = h->p.p[f].xp[c];
    }
    return (int)h->p.p.size();
}
// Lp column of one family/node (log space), rows = max(x)+1 of that family.  which: 0 = L, 1 = Lp.
int orc_get_column(OrcHandle* h, int fam, int node, int which, double* out, int cap) {
    auto& pr = h->p.p[fam];
    auto& L = which ? pr.Lp : pr.L;
    if (node < 1 || node > L.cols) return -1;
    int r = std::min(cap, L.rows);
    for (int k = 0; k < r; ++k) out[k] = L(k, node - 1);
    return L.rows;
}

// ---- likelihood -------------------------------------------------------------------------
static double run_families(OrcHandle* h, int mode, int node, int nthreads, double* per_family) {
    // mode 0: full sweep, 1: from node, 2: root only
    if (h->p.empty()) return 0.0;                                          // profile.jl:56
    size_t N = h->p.p.size();
    std::vector<double> l(N);
    MNode<double>* start = mode == 1 ? h->d->at(node) : h->d->root();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long f = 0; f < (long)N; ++f) {
        auto& pr = h->p.p[f];
        if (mode == 0) l[f] = logpdf_full(pr.Lp, *h->d, pr.xp);
        else if (mode == 1) l[f] = logpdf_from(pr.Lp, start, pr.xp);
        else l[f] = logpdfroot(pr.Lp, start);
    }
    if (per_family) std::copy(l.begin(), l.end(), per_family);
    return weighted_sum(l, h->p);
}
double orc_logpdf_full(OrcHandle* h, int nthreads, double* per_family) {
    ORC_TRY(h, kNaN, { return run_families(h, 0, 1, nthreads, per_family); })
}
double orc_logpdf_from(OrcHandle* h, int node, int nthreads, double* per_family) {
    ORC_TRY(h, kNaN, { return run_families(h, 1, node, nthreads, per_family); })
}
double orc_logpdfroot(OrcHandle* h, int nthreads, double* per_family) {
    ORC_TRY(h, kNaN, { return run_families(h, 2, 1, nthreads, per_family); })
}
// logpdf(d, x) for one explicit extended count vector (model.jl:347-352; test/tests.jl:109)
double orc_logpdf_x(OrcHandle* h, const int64_t* x, int len) {
    ORC_TRY(h, kNaN, { std::vector<int64_t> xv(x, x + len); return logpdf_fresh(*h->d, xv); })
}
// csuros_miklos(d[1], x)[:, node] (model.jl:394-400; test/tests.jl:9-12)
int orc_csuros_miklos_x(OrcHandle* h, const int64_t* x, int len, int node, double* out, int cap) {
    ORC_TRY(h, -1, {
        std::vector<int64_t> xv(x, x + len);
        int mx = (int)*std::max_element(xv.begin(), xv.end());
        LMat<double> L(mx + 1, len);
        for (auto* n : h->d->postorder()) csuros_miklos_node(L, n, xv);
        for (int k = 0; k < std::min(cap, mx + 1); ++k) out[k] = L(k, node - 1);
        return mx + 1;
    })
}

// ---- profile state ----------------------------------------------------------------------
void orc_set(OrcHandle* h) { if (!h->p.empty()) pa_set(h->p); }
void orc_rev(OrcHandle* h) { if (!h->p.empty()) pa_rev(h->p); }
void orc_extend(OrcHandle* h, int i) { if (!h->p.empty()) pa_extend(h->p, i); }
void orc_shrink(OrcHandle* h, int i) { if (!h->p.empty()) pa_shrink(h->p, i); }

// ---- model edits ------------------------------------------------------------------------
// update!(n, (λ=…, μ=…)) / update!(n, :η, v): NaN leaves a field unchanged.
int orc_update(OrcHandle* h, int i, double lam, double mu, double eta) {
    ORC_TRY(h, -1, {
        auto* n = h->d->at(i);
        if (!std::isnan(lam)) n->lam = lam;
        if (!std::isnan(mu)) n->mu = mu;
        if (!std::isnan(eta)) n->eta = eta;
        h->d->update(n);
        return 0;
    })
}
// raw θ assignment without update! (rjmcmc.jl:499-506: n[:q] = …; n[:t] = …), then orc_update_node.
int orc_setfield(OrcHandle* h, int i, int field, double v) {   // 0 t, 1 λ, 2 μ, 3 q, 4 η
    ORC_TRY(h, -1, {
        auto* n = h->d->at(i);
        switch (field) { case 0: n->t = v; break; case 1: n->lam = v; break; case 2: n->mu = v; break;
                         case 3: n->q = v; break; case 4: n->eta = v; break; default: throw std::runtime_error("bad field"); }
        return 0;
    })
}
int orc_update_node(OrcHandle* h, int i) { ORC_TRY(h, -1, { h->d->update(h->d->at(i)); return 0; }) }
int orc_setall(OrcHandle* h) { ORC_TRY(h, -1, { h->d->setall(); return 0; }) }
// setrates!(model, X) with X 2 × n column-major (model.jl:69-76)
int orc_setrates(OrcHandle* h, const double* X, int n) {
    ORC_TRY(h, -1, {
        for (auto& kv : h->d->nodes) {
            if (isawgd(kv.second)) break;
            if (kv.first > n) throw std::runtime_error("setrates!: BoundsError");
            kv.second->lam = X[2 * (kv.first - 1)];
            kv.second->mu = X[2 * (kv.first - 1) + 1];
        }
        h->d->setall();
        return 0;
    })
}
int orc_addwgd(OrcHandle* h, int i, double t, double q) { ORC_TRY(h, -1, { return h->d->addwgm(h->d->at(i), t, q, WGD)->i; }) }
int orc_addwgt(OrcHandle* h, int i, double t, double q) { ORC_TRY(h, -1, { return h->d->addwgm(h->d->at(i), t, q, WGT)->i; }) }
int orc_removewgd(OrcHandle* h, int i, int reindex, int set) {
    ORC_TRY(h, -1, { return h->d->removewgd(h->d->at(i), reindex != 0, set != 0)->i; })
}
int orc_nonwgdchild(OrcHandle* h, int i) { ORC_TRY(h, -1, { return nonwgdchild(h->d->at(i))->i; }) }
int orc_getwgds(OrcHandle* h, int* out, int cap) {
    auto w = h->d->getwgds();
    for (int k = 0; k < (int)w.size() && k < cap; ++k) out[k] = w[k];
    return (int)w.size();
}
int orc_asvector(OrcHandle* h, double* out, int cap) {
    auto v = h->d->asvector();
    for (int k = 0; k < (int)v.size() && k < cap; ++k) out[k] = v[k];
    return (int)v.size();
}
// model = model(θ) (README.md:44-46): replace the handle's model by the rebuilt one.
int orc_set_from_vector(OrcHandle* h, const double* th, int len) {
    ORC_TRY(h, -1, {
        std::vector<double> v(th, th + len);
        Model<double>* nd = rebuild<double, double>(*h->d, v);
        delete h->d;
        h->d = nd;
        return 0;
    })
}

// ---- gradients (model.jl:517-533, profile.jl:71-75) -----------------------------------------
// literal = 1: seed with asvector(d) exactly as the reference does (q in getwgds order fed to a
//              constructor that reads q by ascending id — differs for >= 2 WGDs).
// literal = 0: self-consistent: q block ordered by ascending WGD id, evaluated at the model itself.
static int gradient_impl(OrcHandle* h, double* g, int cap, int literal, int cr, int nthreads) {
    std::vector<double> v = h->d->asvector();
    int n = h->d->ne() + 1;
    int k = (int)v.size() - 2 * n - 1;
    if (!literal && k > 1) std::reverse(v.begin() + 2 * n, v.begin() + 2 * n + k);
    int P = cr ? 2 : (int)v.size();
    if (cap < P) return -1;
    for (int j = 0; j < P; ++j) g[j] = 0.0;
    if (h->p.empty()) return P;
    size_t N = h->p.p.size();
    std::vector<std::vector<double>> gf(N, std::vector<double>(P, 0.0));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        Dual::ntan() = P;
        std::vector<Dual> th(v.size());
        if (!cr) {
            for (size_t j = 0; j < v.size(); ++j) { th[j] = Dual(v[j]); th[j].d[j] = 1.0; }
        } else {                                        // fullvec(x) = [repeat(x[1], n); repeat(x[2], n); v[end]]
            if (k != 0) { /* the reference's fullvec drops q: model(θ) would then fail for k > 0 */ }
            th.assign(2 * n + 1, Dual(0.0));
            for (int j = 0; j < n; ++j) { th[j] = Dual(v[0]); th[j].d[0] = 1.0; th[n + j] = Dual(v[n]); th[n + j].d[1] = 1.0; }
            th[2 * n] = Dual(v.back());
        }
        Model<Dual>* dd = nullptr;
        std::string err;
        try { dd = rebuild<Dual, double>(*h->d, th); } catch (const std::exception& e) { err = e.what(); }
        if (dd) {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
            for (long f = 0; f < (long)N; ++f) {
                Dual l = logpdf_fresh(*dd, h->p.p[f].xp);
                for (int j = 0; j < P; ++j) gf[f][j] = l.d[j];
            }
            delete dd;
        } else {
#ifdef _OPENMP
#pragma omp critical
#endif
            h->err = err;
        }
    }
    if (!h->err.empty()) return -1;
    for (size_t f = 0; f < N; ++f)
        for (int j = 0; j < P; ++j) g[j] += (double)h->p.p[f].n * gf[f][j];
    return P;
}
int orc_gradient(OrcHandle* h, double* g, int cap, int literal, int nthreads) { return gradient_impl(h, g, cap, literal, 0, nthreads); }
int orc_gradient_cr(OrcHandle* h, double* g, int nthreads) { return gradient_impl(h, g, 2, 1, 1, nthreads); }

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
