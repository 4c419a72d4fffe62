// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, fp64, log space) of the gene-family likelihood hot path of
// arzwa/Beluga.jl, written to be *literal*: same loop structure, same operation order,
// same corner-case behaviour (NaN -> -Inf, isapprox switches, clamps) as the reference.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this.  The CUDA product never links or calls it.
//
// Parity status: PINNED for the DL / DL+WGD likelihood by the reference's own golden
// vectors (test/runtests.jl:15-22, test/tests.jl:109, test/gradient.jl:2-7), checked in
// tests/test_oracle_golden.py.  UNPINNED (no reference test exists): WGT nodes
// (wstar_wgt! incl. the never-written w[3,3]), gradient with WGDs present, gradient_cr,
// logpdfroot, polytomies, Branch-model log-likelihood values.  For those this file is
// the definition and is kept literal.
//
// Every function cites the reference file:line (under /root/reference) it follows.
// Third-party arithmetic restated here (not vendored in the reference):
//   Base.isapprox default rtol = sqrt(eps(Float64))            (Julia Base)
//   StatsFuns 0.9.4 logaddexp / log1mexp                       (Manifest.toml:273-283)
//   Distributions 0.23.2 logpdf(Binomial(n,p),s) -> Rmath dbinom (Manifest.toml:108-112)
//   ForwardDiff 0.10.10 forward-mode duals                     (Manifest.toml:126-130)
//   NewickTree#beluga readnw: ids pre-order, root = 1          (Manifest.toml:189-195)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kNaN = std::numeric_limits<double>::quiet_NaN();

// ---------------------------------------------------------------------------------------
// Forward-mode dual number (stands in for ForwardDiff.Dual; all tangents carried at once).
// Comparisons act on the value part only, like ForwardDiff.
// ---------------------------------------------------------------------------------------
struct Dual {
    double v;
    std::vector<double> d;
    static int& ntan() { static thread_local int n = 0; return n; }
    Dual() : v(0.0), d(ntan(), 0.0) {}
    Dual(double x) : v(x), d(ntan(), 0.0) {}
    Dual(int x) : v((double)x), d(ntan(), 0.0) {}
};
inline double val(double x) { return x; }
inline double val(const Dual& x) { return x.v; }

inline Dual operator+(const Dual& a, const Dual& b) { Dual r(a.v + b.v); for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r(a.v - b.v); for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
inline Dual operator-(const Dual& a) { Dual r(-a.v); for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = -a.d[k]; return r; }
// NaN-safe products (README.md:97-102 asks for ForwardDiff's NANSAFE mode: a zero tangent
// times an infinite partial is 0, not NaN).
inline double nsmul(double partial, double tangent) { return tangent == 0.0 ? 0.0 : partial * tangent; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r(a.v * b.v); for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = nsmul(b.v, a.d[k]) + nsmul(a.v, b.d[k]); return r; }
inline Dual operator/(const Dual& a, const Dual& b) {
    Dual r(a.v / b.v);
    double inv = 1.0 / b.v, q = -a.v / (b.v * b.v);
    for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = nsmul(inv, a.d[k]) + nsmul(q, b.d[k]);
    return r;
}
inline Dual& operator+=(Dual& a, const Dual& b) { a = a + b; return a; }
inline Dual& operator-=(Dual& a, const Dual& b) { a = a - b; return a; }
inline Dual& operator*=(Dual& a, const Dual& b) { a = a * b; return a; }
inline bool operator<(const Dual& a, const Dual& b) { return a.v < b.v; }
inline bool operator>(const Dual& a, const Dual& b) { return a.v > b.v; }
inline bool operator<=(const Dual& a, const Dual& b) { return a.v <= b.v; }
inline bool operator>=(const Dual& a, const Dual& b) { return a.v >= b.v; }
inline bool operator==(const Dual& a, const Dual& b) { return a.v == b.v; }
inline Dual unary(const Dual& a, double f, double df) { Dual r(f); for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = nsmul(df, a.d[k]); return r; }
inline Dual exp(const Dual& a) { double e = std::exp(a.v); return unary(a, e, e); }
inline Dual log(const Dual& a) { return unary(a, std::log(a.v), 1.0 / a.v); }
inline Dual log1p(const Dual& a) { return unary(a, std::log1p(a.v), 1.0 / (1.0 + a.v)); }
inline Dual expm1(const Dual& a) { return unary(a, std::expm1(a.v), std::exp(a.v)); }
inline double exp(double a) { return std::exp(a); }
inline double log(double a) { return std::log(a); }
inline double log1p(double a) { return std::log1p(a); }
inline double expm1(double a) { return std::expm1(a); }
inline bool isfinite_(double a) { return std::isfinite(a); }
inline bool isfinite_(const Dual& a) { return std::isfinite(a.v); }

// Base.isapprox(x, y) with default tolerances: x == y || (finite && |x-y| <= rtol*max(|x|,|y|)),
// rtol = sqrt(eps(Float64)).  Used by utils.jl:8,9,17,19.
template <class T> inline bool isapprox(const T& x, const T& y) {
    const double rtol = 1.4901161193847656e-8;
    double a = val(x), b = val(y);
    if (a == b) return true;
    if (!std::isfinite(a) || !std::isfinite(b)) return false;
    return std::fabs(a - b) <= rtol * std::max(std::fabs(a), std::fabs(b));
}
// utils.jl:19   approx1(x) = x ≈ one(x) ? one(x) : x
template <class T> inline T approx1(const T& x) { return isapprox(x, T(1.0)) ? T(1.0) : x; }

// StatsFuns.logaddexp (0.9.x): max(x,y) + log1pexp(-|x-y|), with Δ = 0 when x == y (±Inf).
template <class T> inline T logaddexp(const T& x, const T& y) {
    double a = val(x), b = val(y);
    if (std::isnan(a) || std::isnan(b)) return x + y;  // NaN propagates
    if (a == b) return x + T(0.6931471805599453);      // includes -Inf == -Inf -> -Inf
    if (a > b) { if (b == -kInf) return x; return x + log1p(exp(y - x)); }
    if (a == -kInf) return y;
    return y + log1p(exp(x - y));
}
// StatsFuns.log1mexp(x) = x < log(1/2) ? log1p(-exp(x)) : log(-expm1(x))
template <class T> inline T log1mexp(const T& x) {
    return val(x) < -0.6931471805599453 ? log1p(-exp(x)) : log(-expm1(x));
}

// log n! in extended precision (exact products summed in long double), used for the
// binomial coefficient of the Binomial log-pmf.  The table is filled once (thread-safe static
// init) so OpenMP workers only ever read it.
struct LFactTable {
    std::vector<long double> tab;
    LFactTable() : tab(1, 0.0L) { for (int k = 1; k <= 65536; ++k) tab.push_back(tab.back() + std::log((long double)k)); }
};
inline long double lfact(int n) {
    static const LFactTable T;
    if (n < (int)T.tab.size()) return T.tab[n];
    return std::lgamma((long double)n + 1.0L);
}
inline double lchoose(int n, int s) { return (double)(lfact(n) - lfact(s) - lfact(n - s)); }
// logpdf(Binomial(n, p), s)  (model.jl:448) -> Rmath dbinom(s, n, p, log=true).
// Rmath's special cases (p == 0, q == 0, n == 0) are kept; the general case uses the exact
// log binomial coefficient instead of Loader's saddle-point form (agrees to ~1 ulp).
template <class T> inline T binomlogpdf(int n, const T& p, int s) {
    if (s < 0 || s > n) return T(-kInf);
    double pv = val(p);
    if (pv == 0.0) return s == 0 ? T(0.0) : T(-kInf);
    if (pv == 1.0) return s == n ? T(0.0) : T(-kInf);
    if (n == 0) return T(0.0);
    T r = T(lchoose(n, s));
    if (s > 0) r = r + T((double)s) * log(p);
    if (n - s > 0) r = r + T((double)(n - s)) * log1p(-p);
    return r;
}

// ---------------------------------------------------------------------------------------
// Model graph (node.jl:12-77).  kind mirrors the :root/:sp/:wgd/:wgt/:wgdafter symbols.
// ---------------------------------------------------------------------------------------
enum Kind { ROOT = 0, SP = 1, WGD = 2, WGT = 3, WGDAFTER = 4 };
enum NodeType { NODE_MODEL = 0, BRANCH_MODEL = 1 };  // node.jl:12-26 (Node vs Branch)

template <class T> struct MNode {
    int i = 0;
    Kind kind = SP;
    T t = T(0.0), lam = T(0.0), mu = T(0.0), q = T(0.0), eta = T(0.0);  // θ Dict entries
    T eps[2] = {T(1.0), T(1.0)};                                        // ϵ = ones(T, 2)
    std::vector<T> W;                                                   // m×m, column-major
    int m = 0;
    MNode* p = nullptr;
    std::vector<MNode*> c;
    T& w(int r, int col) { return W[(size_t)col * m + r]; }             // 0-based
    bool isleaf() const { return c.empty(); }
    bool isroot() const { return p == nullptr; }
};

template <class T> inline bool iswgd(const MNode<T>* n) { return n->kind == WGD; }
template <class T> inline bool iswgt(const MNode<T>* n) { return n->kind == WGT; }
template <class T> inline bool iswgm(const MNode<T>* n) { return iswgd(n) || iswgt(n); }
template <class T> inline bool iswgdafter(const MNode<T>* n) { return n->kind == WGDAFTER && n->p->kind == WGD; }
template <class T> inline bool iswgtafter(const MNode<T>* n) { return n->kind == WGDAFTER && n->p->kind == WGT; }
template <class T> inline bool iswgmafter(const MNode<T>* n) { return iswgdafter(n) || iswgtafter(n); }
template <class T> inline bool isawgd(const MNode<T>* n) { return iswgd(n) || iswgdafter(n) || iswgt(n) || iswgtafter(n); }
template <class T> inline bool issp(const MNode<T>* n) { return n->kind == SP; }
// node.jl:79-87
template <class T> inline MNode<T>* nonwgdparent(MNode<T>* n) { while (isawgd(n)) n = n->p; return n; }
template <class T> inline MNode<T>* nonwgdchild(MNode<T>* n) { while (isawgd(n)) n = n->c.front(); return n; }

// utils.jl:8-9
template <class T> inline T getphi(const T& t, const T& l, const T& m) {
    if (isapprox(l, m)) return l * t / (T(1.0) + l * t);
    return m * (exp(t * (l - m)) - T(1.0)) / (l * exp(t * (l - m)) - m);
}
template <class T> inline T getpsi(const T& t, const T& l, const T& m) {
    if (isapprox(l, m)) return l * t / (T(1.0) + l * t);
    return (l / m) * getphi(t, l, m);
}
// utils.jl:17-18
template <class T> inline T ep(const T& l, const T& m, const T& t, const T& e) {
    if (isapprox(l, m)) return T(1.0) + (T(1.0) - e) / (m * (e - T(1.0)) * t - T(1.0));
    return approx1((m + (l - m) / (T(1.0) + exp((l - m) * t) * l * (e - T(1.0)) / (m - l * e))) / l);
}

template <class T> struct Model {
    std::map<int, MNode<T>*> nodes;       // DLWGD.nodes (model.jl:8-11)
    std::map<int, std::string> leaves;    // DLWGD.leaves
    NodeType nt = BRANCH_MODEL;
    int m = 3;
    std::vector<MNode<T>*> owned;         // everything ever allocated (freed in dtor)
    ~Model() { for (auto* n : owned) delete n; }
    MNode<T>* at(int i) const { auto it = nodes.find(i); if (it == nodes.end()) throw std::runtime_error("no such node"); return it->second; }
    MNode<T>* root() const { return at(1); }
    int maxid() const { return nodes.rbegin()->first; }
    MNode<T>* mk(int i, Kind k) { auto* n = new MNode<T>(); n->i = i; n->kind = k; n->m = m; n->W.assign((size_t)m * m, T(0.0)); owned.push_back(n); return n; }

    // node.jl:220-230  getλμ(parent, child)
    void getlm(MNode<T>* a, MNode<T>* b, T& l, T& mu) const {
        if (nt == NODE_MODEL) {
            a = isawgd(a) ? nonwgdparent(a) : a;
            b = isawgd(b) ? nonwgdchild(b) : b;
            l = (a->lam + b->lam) / T(2.0);
            mu = (a->mu + b->mu) / T(2.0);
        } else {
            b = isawgd(b) ? nonwgdchild(b) : b;
            l = b->lam;
            mu = b->mu;
        }
    }
    // node.jl:136-162  setϵ!
    void seteps(MNode<T>* n) const {
        if (n->isleaf()) {
            n->eps[1] = T(0.0);
        } else if (iswgd(n)) {
            T q = n->q;
            MNode<T>* c = n->c.front();
            T ec = c->eps[1];
            c->eps[0] = ec;
            n->eps[1] = c->eps[0] = q * (ec * ec) + (T(1.0) - q) * ec;
        } else if (iswgt(n)) {
            T q = n->q;
            MNode<T>* c = n->c.front();
            T ec = c->eps[1];
            c->eps[0] = ec;
            n->eps[1] = c->eps[0] = q * (ec * ec * ec) + T(2.0) * q * (T(1.0) - q) * (ec * ec) + (T(1.0) - q) * (T(1.0) - q) * ec;
        } else {
            n->eps[1] = T(1.0);
            for (MNode<T>* c : n->c) {
                c->eps[0] = T(1.0);
                T l, mu;
                getlm(n, c, l, mu);
                c->eps[0] = approx1(ep(l, mu, c->t, c->eps[1]));
                n->eps[1] = n->eps[1] * c->eps[0];
            }
            n->eps[1] = approx1(n->eps[1]);
        }
    }
    // node.jl:181-192  wstar!
    static void wstar(MNode<T>* n, const T& t, const T& l, const T& mu, const T& e) {
        T phi = getphi(t, l, mu), psi = getpsi(t, l, mu);
        T nn = T(1.0) - psi * e;
        T phip = approx1((phi * (T(1.0) - e) + (T(1.0) - psi) * e) / nn);
        T psip = approx1(psi * (T(1.0) - e) / nn);
        n->w(0, 0) = T(1.0);
        for (int mm = 1; mm <= n->m - 1; ++mm)
            for (int k = 1; k <= mm; ++k)
                n->w(k, mm) = psip * n->w(k, mm - 1) + (T(1.0) - phip) * (T(1.0) - psip) * n->w(k - 1, mm - 1);
    }
    // node.jl:194-203  wstar_wgd!   (1-based w[i+1,j+1] = w[2,2]*w[i,j] + w[2,3]*w[i,j-1])
    static void wstar_wgd(MNode<T>* n, const T& q, const T& e) {
        n->w(0, 0) = T(1.0);
        n->w(1, 1) = ((T(1.0) - q) + T(2.0) * q * e) * (T(1.0) - e);
        n->w(1, 2) = q * ((T(1.0) - e) * (T(1.0) - e));
        int mmax = n->m - 1;
        for (int i = 1; i <= mmax; ++i)
            for (int j = 2; j <= mmax; ++j)
                n->w(i, j) = n->w(1, 1) * n->w(i - 1, j - 1) + n->w(1, 2) * n->w(i - 1, j - 2);
    }
    // node.jl:205-218  wstar_wgt!   (the j loop starts at 3, so 1-based w[i+1,3] is never
    // written for i >= 2: w[3,3] stays 0.  Kept literally; untested in the reference.)
    static void wstar_wgt(MNode<T>* n, const T& q, const T& e) {
        T q1 = q, q2 = T(2.0) * q * (T(1.0) - q), q3 = (T(1.0) - q) * (T(1.0) - q);
        T ome = T(1.0) - e;
        n->w(0, 0) = T(1.0);
        n->w(1, 1) = q1 * ome + T(2.0) * q2 * e * ome + T(3.0) * q3 * (e * e) * ome;
        n->w(1, 2) = q2 * (ome * ome) + T(3.0) * q3 * e * (ome * ome);
        n->w(1, 3) = q3 * (ome * ome * ome);
        int mmax = n->m - 1;
        for (int i = 1; i <= mmax; ++i)
            for (int j = 3; j <= mmax; ++j)
                n->w(i, j) = n->w(1, 1) * n->w(i - 1, j - 1) + n->w(1, 2) * n->w(i - 1, j - 2) + n->w(1, 3) * n->w(i - 1, j - 3);
    }
    // node.jl:164-179  setW!
    void setW(MNode<T>* n) const {
        if (n->isroot()) return;
        T l, mu;
        getlm(n->p, n, l, mu);
        T e = n->eps[1];
        if (iswgdafter(n)) wstar_wgd(n, n->p->q, e);
        else if (iswgtafter(n)) { if (n->m < 4) throw std::runtime_error("wgt needs m >= 4"); wstar_wgt(n, n->p->q, e); }
        else wstar(n, n->t, l, mu, e);
    }
    void set(MNode<T>* n) const { seteps(n); setW(n); }                       // node.jl:118-121
    void setabove(MNode<T>* n) const { while (n) { set(n); n = n->p; } }        // node.jl:123-128
    void setbelow(MNode<T>* n) const { for (auto* c : n->c) set(c); }           // node.jl:130-134
    void update(MNode<T>* n) const {                                            // node.jl:99-104
        if (nt == BRANCH_MODEL) setabove(n); else { setbelow(n); setabove(n); }
    }
    void postwalk(MNode<T>* n, std::vector<MNode<T>*>& out) const { for (auto* c : n->c) postwalk(c, out); out.push_back(n); }
    std::vector<MNode<T>*> postorder() const { std::vector<MNode<T>*> v; postwalk(root(), v); return v; }
    void setall() const { for (auto* n : postorder()) set(n); }                 // model.jl:56-60

    // model.jl:211-225  getwgds: walks ids downwards from the maximum -> DESCENDING ids.
    std::vector<int> getwgds() const {
        std::vector<int> w;
        int i = maxid();
        bool cont = true;
        while (cont && i >= 1) {
            auto it = nodes.find(i);
            if (it == nodes.end() || iswgmafter(it->second)) {}
            else if (iswgm(it->second)) w.push_back(i);
            else cont = false;
            --i;
        }
        return w;
    }
    int nwgd() const { return (int)getwgds().size(); }
    int ne() const { return (int)nodes.size() - 2 * nwgd() - 1; }               // model.jl:20

    // model.jl:106-138  addwgd!/addwgt!
    MNode<T>* insertwgm(MNode<T>* n, MNode<T>* w, MNode<T>* a) {
        MNode<T>* par = n->p;
        par->c.erase(std::remove(par->c.begin(), par->c.end(), n), par->c.end());
        par->c.push_back(w); w->p = par;
        w->c.push_back(a); a->p = w;
        a->c.push_back(n);
        n->p = a;
        n->t = n->t - w->t;
        nodes[w->i] = w;
        nodes[a->i] = a;
        setabove(nonwgdchild(n));
        return w;
    }
    MNode<T>* addwgm(MNode<T>* n, const T& t, const T& q, Kind k) {
        if (n->isroot()) throw std::runtime_error("Cannot add WGD above root node");
        if (!(val(n->t - t) > 0.0)) throw std::runtime_error("Invalid WGD time");
        int i = maxid() + 1;
        MNode<T>* w = mk(i, k); w->q = q; w->t = n->t - t;
        MNode<T>* a = mk(i + 1, WGDAFTER); a->t = T(0.0);
        return insertwgm(n, w, a);
    }
    // model.jl:158-183  removewgd! / reindex!
    MNode<T>* removewgd(MNode<T>* n, bool reindex = true, bool set = true) {
        if (!iswgm(n)) throw std::runtime_error("Not a WGD/T node");
        MNode<T>* parent = n->p;
        MNode<T>* after = n->c.front();
        MNode<T>* child = after->c.front();
        after->c.clear();
        n->c.clear();
        parent->c.erase(std::remove(parent->c.begin(), parent->c.end(), n), parent->c.end());
        parent->c.push_back(child);
        child->p = parent;
        child->t = child->t + n->t;
        nodes.erase(n->i);
        nodes.erase(n->i + 1);
        if (reindex) reindexfrom(n->i + 2);
        if (set) setabove(nonwgdchild(child));
        return child;
    }
    void reindexfrom(int i) {
        if (nodes.empty()) return;
        int mx = maxid();
        for (int j = i; j <= mx; ++j) {
            auto it = nodes.find(j);
            if (it == nodes.end()) throw std::runtime_error("reindex!: KeyError");
            MNode<T>* nd = it->second;
            nodes[j - 2] = nd;
            nd->i = j - 2;
            nodes.erase(j);
        }
    }
    // model.jl:541-546 asvector = [λ1..λn, μ1..μn, q (getwgds order, i.e. descending id), η]
    std::vector<double> asvector() const {
        int n = ne() + 1;
        std::vector<double> v(2 * n, 0.0);
        for (auto& kv : nodes) {                       // getrates, model.jl:83-91 (sorted ids, break at first wgd)
            if (isawgd(kv.second)) break;
            v[kv.first - 1] = val(kv.second->lam);
            v[n + kv.first - 1] = val(kv.second->mu);
        }
        for (int i : getwgds()) v.push_back(val(at(i)->q));
        v.push_back(val(root()->eta));
        return v;
    }
};

// model.jl:305-339  (model::DLWGD)(θ): rebuild a model of scalar type Y from a flat vector.
// q index: j = x.i % n; j -= (j-1)÷2  (ASCENDING wgd id — differs from asvector's order when
// there are >= 2 WGDs; literal).  `q_by_ascending_id=false` keeps that; the flag exists so
// tests can also ask for the self-consistent ordering.
template <class Y, class T>
Model<Y>* rebuild(const Model<T>& src, const std::vector<Y>& th) {
    int n = src.ne() + 1;
    auto* dst = new Model<Y>();
    dst->nt = src.nt; dst->m = src.m; dst->leaves = src.leaves;
    int nq = (int)th.size() - 2 * n - 1;
    struct Rec {
        static void go(const MNode<T>* x, MNode<Y>* y, Model<Y>* dst, const std::vector<Y>& th, int n, int nq) {
            MNode<Y>* x_ = nullptr;
            if (x->isroot()) {
                x_ = dst->mk(1, ROOT);
                x_->eta = th.back(); x_->lam = th[0]; x_->mu = th[n]; x_->t = Y(0.0);
                dst->nodes[1] = x_;
            } else if (x->kind == SP) {
                x_ = dst->mk(x->i, SP);
                x_->lam = th[x->i - 1]; x_->mu = th[n + x->i - 1]; x_->t = Y(val(x->t));
                dst->nodes[x->i] = x_; x_->p = y; y->c.push_back(x_);
            } else if (x->kind == WGD || x->kind == WGT) {
                int j = x->i % n;
                j -= (j - 1) / 2;
                if (j < 1 || j > nq) throw std::runtime_error("model(θ): q index out of range");
                x_ = dst->mk(x->i, x->kind);
                x_->q = th[2 * n + j - 1]; x_->t = Y(val(x->t));
                dst->nodes[x->i] = x_; x_->p = y; y->c.push_back(x_);
            } else {
                x_ = dst->mk(x->i, WGDAFTER);
                x_->t = Y(0.0);
                dst->nodes[x->i] = x_; x_->p = y; y->c.push_back(x_);
            }
            for (const MNode<T>* c : x->c) go(c, x_, dst, th, n, nq);
        }
    };
    Rec::go(src.root(), nullptr, dst, th, n, nq);
    dst->setall();
    return dst;
}

// ---------------------------------------------------------------------------------------
// Likelihood (model.jl:341-501).  L is (mx+1) × ncols column-major, log space, -Inf init.
// ---------------------------------------------------------------------------------------
template <class T> struct LMat {
    int rows = 0, cols = 0;
    std::vector<T> a;
    LMat() {}
    LMat(int r, int c) : rows(r), cols(c), a((size_t)r * c, T(-kInf)) {}   // minfs, utils.jl:2-4
    T& operator()(int r, int c) { return a[(size_t)c * rows + r]; }
    const T& operator()(int r, int c) const { return a[(size_t)c * rows + r]; }
};

// model.jl:402-463  csuros_miklos!(L, node, x)      (x and L columns are indexed by node id-1)
template <class T>
void csuros_miklos_node(LMat<T>& L, MNode<T>* node, const std::vector<int64_t>& x) {
    int mx = (int)*std::max_element(x.begin(), x.end());
    if (node->isleaf()) {
        L((int)x[node->i - 1], node->i - 1) = T(0.0);
        return;
    }
    int k = (int)node->c.size();
    std::vector<int> Mc(k), cM(k + 1, 0);
    for (int i = 0; i < k; ++i) { Mc[i] = (int)x[node->c[i]->i - 1]; cM[i + 1] = cM[i] + Mc[i]; }
    std::vector<T> ce(k + 1, T(1.0));                         // _ϵ = cumprod([1; ϵ_c[1]...])
    for (int i = 0; i < k; ++i) ce[i + 1] = ce[i] * node->c[i]->eps[0];
    for (int i = 0; i <= k; ++i) if (val(ce[i]) > 1.0) ce[i] = T(1.0);
    int S = cM[k];
    // B[i][t][s], A[i][n]
    std::vector<std::vector<T>> A(k, std::vector<T>(S + 1, T(-kInf)));
    std::vector<T> expL(mx + 1);
    for (int i = 0; i < k; ++i) {
        MNode<T>* c = node->c[i];
        int Mi = Mc[i];
        std::vector<std::vector<T>> B(cM[i] + 1 > 1 ? cM[i] + 1 : 1, std::vector<T>(mx + 2, T(-kInf)));
        // B[i,1,:] = log.(Wc * exp.(L[1:mx+1, c.i]))   — dense (mx+1)² mat-vec, model.jl:425-426
        for (int j = 0; j <= mx; ++j) expL[j] = exp(L(j, c->i - 1));
        for (int s = 0; s <= mx; ++s) {
            T acc = T(0.0);
            for (int j = 0; j <= mx; ++j) acc = acc + c->w(s, j) * expL[j];
            B[0][s] = log(acc);
        }
        T le = log(c->eps[0]);
        for (int t = 1; t <= cM[i]; ++t)                         // model.jl:428-431
            for (int s = 0; s <= Mi; ++s)
                B[t][s] = (s == Mi) ? B[t - 1][s] + le : logaddexp(B[t - 1][s + 1], le + B[t - 1][s]);
        if (i == 0) {
            T l1 = log(T(1.0) - ce[1]);
            for (int n = 0; n <= cM[1]; ++n) A[0][n] = B[0][n] - T((double)n) * l1;   // model.jl:432-435
        } else {
            for (int n = 0; n <= cM[i + 1]; ++n)                 // model.jl:438-452
                for (int t = 0; t <= cM[i]; ++t) {
                    int s = n - t;
                    if (s < 0 || s > Mi) continue;
                    T p = ce[i];
                    if (!(val(p) >= 0.0 && val(p) <= 1.0)) p = T(1.0);
                    T lp = binomlogpdf(n, p, s) + A[i - 1][t] + B[t][s];
                    A[i][n] = logaddexp(A[i][n], lp);
                }
            T l1 = log(T(1.0) - ce[i + 1]);
            for (int n = 0; n <= cM[i + 1]; ++n) A[i][n] = A[i][n] - T((double)n) * l1;   // model.jl:453-455
        }
    }
    for (int n = 0; n <= (int)x[node->i - 1]; ++n) L(n, node->i - 1) = A[k - 1][n];       // model.jl:458-461
}

// model.jl:465-475 integrate_root
template <class T> T integrate_root(const LMat<T>& L, const MNode<T>* n) {
    T eta = n->eta;
    T e = log(n->eps[1]);
    T p = T(-kInf);
    for (int i = 2; i <= L.rows; ++i) {
        T f = T((double)(i - 1)) * log1mexp(e) + log(eta) + T((double)(i - 2)) * log(T(1.0) - eta);
        f = f - T((double)i) * log1mexp(log(T(1.0) - eta) + e);
        p = logaddexp(p, L(i - 1, 0) + f);
    }
    return p;
}
// utils.jl:22-23
template <class T> T geometric_extinctionp(const T& e, const T& eta) { return eta + e - log1mexp(log1mexp(eta) + e); }
// model.jl:492-501 condition_oib
template <class T> T condition_oib(const MNode<T>* n) {
    T leta = log(n->eta);
    if (n->c.size() < 2) throw std::runtime_error("condition_oib: root needs two children");
    T lr1 = geometric_extinctionp(log(n->c[0]->eps[0]), leta);
    T lr2 = geometric_extinctionp(log(n->c[1]->eps[0]), leta);
    if (val(lr1) > 0.0 || val(lr2) > 0.0) return T(-kInf);
    return log1mexp(lr1) + log1mexp(lr2);
}
template <class T> T finite_or_minf(const T& l) { return isfinite_(l) ? l : T(-kInf); }

// model.jl:360-367 logpdf!(L, d, x)
template <class T> T logpdf_full(LMat<T>& L, const Model<T>& d, const std::vector<int64_t>& x) {
    for (auto* n : d.postorder()) csuros_miklos_node(L, n, x);
    T l = integrate_root(L, d.root());
    l = l - condition_oib(d.root());
    return finite_or_minf(l);
}
// model.jl:347-352 logpdf(d, x) (fresh L, model.jl:394-400)
template <class T> T logpdf_fresh(const Model<T>& d, const std::vector<int64_t>& x) {
    int mx = (int)*std::max_element(x.begin(), x.end());
    LMat<T> L(mx + 1, (int)x.size());
    return logpdf_full(L, d, x);
}
// model.jl:376-385 logpdf!(L, n, x): path n -> root only
template <class T> T logpdf_from(LMat<T>& L, MNode<T>* n, const std::vector<int64_t>& x) {
    while (!n->isroot()) { csuros_miklos_node(L, n, x); n = n->p; }
    csuros_miklos_node(L, n, x);
    T l = integrate_root(L, n);
    l = l - condition_oib(n);
    return finite_or_minf(l);
}
// model.jl:388-392 logpdfroot
template <class T> T logpdfroot(const LMat<T>& L, const MNode<T>* n) {
    T l = integrate_root(L, n);
    l = l - condition_oib(n);
    return finite_or_minf(l);
}

// ---------------------------------------------------------------------------------------
// Profiles (profile.jl:14-121)
// ---------------------------------------------------------------------------------------
struct Profile {
    int64_t n = 1;                 // multiplicity of this site pattern
    std::vector<int64_t> x, xp;
    LMat<double> L, Lp;
};
struct PArray {
    std::vector<Profile> p;
    bool empty() const { return p.empty() || p[0].x.empty(); }   // the mock profile, profile.jl:28,31
};
inline void pa_set(PArray& a) { for (auto& f : a.p) { f.x = f.xp; f.L = f.Lp; } }            // profile.jl:79,82-86
inline void pa_rev(PArray& a) { for (auto& f : a.p) { f.xp = f.x; f.Lp = f.L; } }            // profile.jl:80,88-92
inline void pa_extend(PArray& a, int i) {                                                      // profile.jl:111-115
    for (auto& f : a.p) {
        int64_t v = f.xp[i - 1];
        f.xp.push_back(v); f.xp.push_back(v);
        LMat<double> nl(f.Lp.rows, f.Lp.cols + 2);
        std::copy(f.Lp.a.begin(), f.Lp.a.end(), nl.a.begin());
        f.Lp = nl;
    }
}
inline void pa_shrink(PArray& a, int i) {                                                      // profile.jl:117-121
    for (auto& f : a.p) {
        f.xp.erase(f.xp.begin() + (i - 1), f.xp.begin() + (i + 1));
        LMat<double> nl(f.Lp.rows, f.Lp.cols - 2);
        int cc = 0;
        for (int c = 0; c < f.Lp.cols; ++c) {
            if (c == i - 1 || c == i) continue;
            for (int r = 0; r < f.Lp.rows; ++r) nl(r, cc) = f.Lp(r, c);
            ++cc;
        }
        f.Lp = nl;
    }
}

// ---------------------------------------------------------------------------------------
// Newick reader (NewickTree#beluga readnw): ids pre-order, root = 1, children in file order.
// ---------------------------------------------------------------------------------------
struct NwNode { int id = 0; std::string name; double dist = 0.0; bool hasdist = false; NwNode* p = nullptr; std::vector<NwNode*> c; };
struct NwTree {
    std::vector<NwNode*> all;
    NwNode* root = nullptr;
    ~NwTree() { for (auto* n : all) delete n; }
    NwNode* mk() { all.push_back(new NwNode()); return all.back(); }
};
inline void nw_skipws(const std::string& s, size_t& i) { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
inline NwNode* nw_parse(NwTree& t, const std::string& s, size_t& i) {
    NwNode* n = t.mk();
    nw_skipws(s, i);
    if (i < s.size() && s[i] == '(') {
        ++i;
        for (;;) {
            NwNode* c = nw_parse(t, s, i);
            c->p = n; n->c.push_back(c);
            nw_skipws(s, i);
            if (i < s.size() && s[i] == ',') { ++i; continue; }
            if (i < s.size() && s[i] == ')') { ++i; break; }
            throw std::runtime_error("newick: expected , or )");
        }
    }
    nw_skipws(s, i);
    size_t j = i;
    while (j < s.size() && s[j] != ':' && s[j] != ',' && s[j] != ')' && s[j] != ';' && s[j] != '(') ++j;
    n->name = s.substr(i, j - i);
    while (!n->name.empty() && std::isspace((unsigned char)n->name.back())) n->name.pop_back();
    i = j;
    if (i < s.size() && s[i] == ':') {
        ++i;
        size_t k = i;
        while (k < s.size() && s[k] != ',' && s[k] != ')' && s[k] != ';') ++k;
        n->dist = std::stod(s.substr(i, k - i));
        n->hasdist = true;
        i = k;
    }
    return n;
}
inline void nw_number(NwNode* n, int& ctr) { n->id = ++ctr; for (auto* c : n->c) nw_number(c, ctr); }
inline void readnw(const std::string& s, NwTree& t) {
    size_t i = 0;
    t.root = nw_parse(t, s, i);
    int ctr = 0;
    nw_number(t.root, ctr);
}

// model.jl:227-244 initmodel
template <class T>
Model<T>* initmodel(const NwTree& t, double eta, double lam, double mu, int m, NodeType nt) {
    auto* d = new Model<T>();
    d->nt = nt; d->m = m;
    struct Rec {
        static void go(const NwNode* x, MNode<T>* y, Model<T>* d, double eta, double lam, double mu) {
            MNode<T>* x_;
            if (!x->p) {
                x_ = d->mk(1, ROOT); x_->eta = T(eta); x_->lam = T(lam); x_->mu = T(mu); x_->t = T(0.0);
                d->nodes[1] = x_;
            } else {
                x_ = d->mk(x->id, SP); x_->lam = T(lam); x_->mu = T(mu); x_->t = T(x->dist);
                d->nodes[x->id] = x_; x_->p = y; y->c.push_back(x_);
            }
            if (x->c.empty()) d->leaves[x->id] = x->name;
            else for (auto* c : x->c) go(c, x_, d, eta, lam, mu);
        }
    };
    Rec::go(t.root, nullptr, d, eta, lam, mu);
    return d;
}

// model.jl:40-49 profile(t, l, df): site-pattern compression (first-appearance order) and
// extended counts (leaf = observed, internal = Σ children).  `counts` is nfam × ntaxa row-major,
// column j belongs to leaf name names[j].
inline int build_profiles(const NwTree& t, const std::vector<std::string>& names,
                          const std::vector<int64_t>& counts, int nfam, PArray& out) {
    int ntaxa = (int)names.size();
    std::map<std::vector<int64_t>, int> seen;
    std::vector<std::vector<int64_t>> pats;
    std::vector<int64_t> mult;
    for (int f = 0; f < nfam; ++f) {
        std::vector<int64_t> row(counts.begin() + (size_t)f * ntaxa, counts.begin() + (size_t)(f + 1) * ntaxa);
        auto it = seen.find(row);
        if (it == seen.end()) { seen[row] = (int)pats.size(); pats.push_back(row); mult.push_back(1); }
        else mult[it->second] += 1;
    }
    int nn = (int)t.all.size();
    std::map<std::string, int> col;
    for (int j = 0; j < ntaxa; ++j) col[names[j]] = j;
    // post-order over the newick tree
    std::vector<const NwNode*> post;
    struct PW { static void go(const NwNode* n, std::vector<const NwNode*>& o) { for (auto* c : n->c) go(c, o); o.push_back(n); } };
    PW::go(t.root, post);
    int64_t gmax = 0;
    out.p.clear();
    for (size_t k = 0; k < pats.size(); ++k) {
        Profile pr;
        pr.n = mult[k];
        pr.x.assign(nn, 0);
        for (const NwNode* n : post) {
            if (n->c.empty()) {
                auto it = col.find(n->name);
                if (it == col.end()) throw std::runtime_error("leaf " + n->name + " not in data");
                pr.x[n->id - 1] = pats[k][it->second];
            } else {
                int64_t s = 0;
                for (auto* c : n->c) s += pr.x[c->id - 1];
                pr.x[n->id - 1] = s;
            }
        }
        pr.xp = pr.x;
        int64_t mx = *std::max_element(pr.x.begin(), pr.x.end());
        gmax = std::max(gmax, mx);
        pr.L = LMat<double>((int)mx + 1, nn);                                 // profile.jl:33
        pr.Lp = pr.L;
        out.p.push_back(pr);
    }
    return (int)gmax + 1;
}

}  // namespace orc
