// gradient.cuh — forward-mode, path-sparse tangent kernels (gradient / gradient_cr of src/model.jl:517-533,
// src/profile.jl:71-75).  Included by beluga_b200.cu after the primal kernels.
//
// A parameter of branch b only perturbs ε, W and the partial-likelihood columns on the path b→root, so a
// tangent pass visits just those nodes.  Every quantity of the primal general path (setϵ!, setW!, the merge
// tables, the pruning step, the root integration) is re-evaluated here in dual arithmetic Du<T> — value plus T
// tangent lanes — which is what ForwardDiff does to the reference (src/model.jl:517-522), restricted to the
// perturbed nodes.  Value-based switches (λ ≈ μ, approx1, clamps) look at the value part only, so, like the
// reference, the λ ≈ μ branch differentiates the limit formula.
#pragma once

template <int T> struct Du {
    double v;
    double d[T];
    __device__ __forceinline__ Du() {}
    __device__ __forceinline__ Du(double x) : v(x) {
#pragma unroll
        for (int t = 0; t < T; ++t) d[t] = 0.0;
    }
};
__device__ __forceinline__ double val(double x) { return x; }
template <int T> __device__ __forceinline__ double val(const Du<T>& x) { return x.v; }

#define DU_BIN(op, vexpr, dexpr)                                                                          \
    template <int T> __device__ __forceinline__ Du<T> operator op(const Du<T>& a, const Du<T>& b) {      \
        Du<T> r;                                                                                           \
        r.v = vexpr;                                                                                       \
        _Pragma("unroll") for (int t = 0; t < T; ++t) r.d[t] = dexpr;                                    \
        return r;                                                                                          \
    }
DU_BIN(+, a.v + b.v, a.d[t] + b.d[t])
DU_BIN(-, a.v - b.v, a.d[t] - b.d[t])
DU_BIN(*, a.v * b.v, a.d[t] * b.v + a.v * b.d[t])
template <int T> __device__ __forceinline__ Du<T> operator/(const Du<T>& a, const Du<T>& b) {
    Du<T> r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int t = 0; t < T; ++t) r.d[t] = (a.d[t] - r.v * b.d[t]) * inv;
    return r;
}
template <int T> __device__ __forceinline__ Du<T> operator+(const Du<T>& a, double b) { Du<T> r = a; r.v += b; return r; }
template <int T> __device__ __forceinline__ Du<T> operator+(double b, const Du<T>& a) { Du<T> r = a; r.v += b; return r; }
template <int T> __device__ __forceinline__ Du<T> operator-(const Du<T>& a, double b) { Du<T> r = a; r.v -= b; return r; }
template <int T> __device__ __forceinline__ Du<T> operator-(double b, const Du<T>& a) {
    Du<T> r;
    r.v = b - a.v;
#pragma unroll
    for (int t = 0; t < T; ++t) r.d[t] = -a.d[t];
    return r;
}
template <int T> __device__ __forceinline__ Du<T> operator-(const Du<T>& a) { return 0.0 - a; }
template <int T> __device__ __forceinline__ Du<T> operator*(const Du<T>& a, double b) {
    Du<T> r;
    r.v = a.v * b;
#pragma unroll
    for (int t = 0; t < T; ++t) r.d[t] = a.d[t] * b;
    return r;
}
template <int T> __device__ __forceinline__ Du<T> operator*(double b, const Du<T>& a) { return a * b; }
template <int T> __device__ __forceinline__ Du<T> operator/(const Du<T>& a, double b) { return a * (1.0 / b); }
template <int T> __device__ __forceinline__ Du<T> operator/(double a, const Du<T>& b) { return Du<T>(a) / b; }
using ::exp;      // keep the double overloads visible next to the Du ones below
using ::log;
using ::log1p;
using ::expm1;
template <int T> __device__ __forceinline__ Du<T> du_unary(const Du<T>& a, double f, double df) {
    Du<T> r;
    r.v = f;
#pragma unroll
    for (int t = 0; t < T; ++t) r.d[t] = a.d[t] == 0.0 ? 0.0 : df * a.d[t];     // NaN-safe like README.md:97-102 asks
    return r;
}
template <int T> __device__ __forceinline__ Du<T> exp(const Du<T>& a) { double e = exp(a.v); return du_unary(a, e, e); }
template <int T> __device__ __forceinline__ Du<T> log(const Du<T>& a) { return du_unary(a, log(a.v), 1.0 / a.v); }
template <int T> __device__ __forceinline__ Du<T> log1p(const Du<T>& a) { return du_unary(a, log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int T> __device__ __forceinline__ Du<T> expm1(const Du<T>& a) { return du_unary(a, expm1(a.v), exp(a.v)); }

// generic (double or Du) versions of the BDP scalars: utils.jl:8-19
template <class S> __device__ __forceinline__ S g_approx1(const S& x) { return d_isapprox(val(x), 1.0) ? S(1.0) : x; }
template <class S> __device__ __forceinline__ S g_getphi(const S& t, const S& l, const S& m) {
    if (d_isapprox(val(l), val(m))) return l * t / (1.0 + l * t);
    S e = exp(t * (l - m));
    return m * (e - 1.0) / (l * e - m);
}
template <class S> __device__ __forceinline__ S g_getpsi(const S& t, const S& l, const S& m) {
    if (d_isapprox(val(l), val(m))) return l * t / (1.0 + l * t);
    return (l / m) * g_getphi(t, l, m);
}
template <class S> __device__ __forceinline__ S g_ep(const S& l, const S& m, const S& t, const S& e) {
    if (d_isapprox(val(l), val(m))) return 1.0 + (1.0 - e) / (m * (e - 1.0) * t - 1.0);
    return g_approx1((m + (l - m) / (1.0 + exp((l - m) * t) * l * (e - 1.0) / (m - l * e))) / l);
}
template <class S> __device__ __forceinline__ S g_powi(S b, int n) {
    S r(1.0);
    while (n > 0) { if (n & 1) r = r * b; b = b * b; n >>= 1; }
    return r;
}
template <class S> __device__ __forceinline__ S g_log1mexp(const S& x) { return val(x) < -0.6931471805599453 ? log1p(-exp(x)) : log(-expm1(x)); }

// which parameter each tangent lane differentiates
enum { SEED_LAM = 0, SEED_MU = 1, SEED_Q = 2, SEED_ETA = 3, SEED_LAM_ALL = 4, SEED_MU_ALL = 5, SEED_NONE = 6 };
template <int T> struct Seeds { int kind[T]; int node[T]; };

template <int T> __device__ __forceinline__ Du<T> seeded(double v, int what, int node, const Seeds<T>& sd) {
    Du<T> r(v);
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int k = sd.kind[t];
        if ((k == what && sd.node[t] == node) || (what == SEED_LAM && k == SEED_LAM_ALL) || (what == SEED_MU && k == SEED_MU_ALL)) r.d[t] = 1.0;
    }
    return r;
}
template <int T> __device__ __forceinline__ Du<T> load_du(const double* v, const double* dv, size_t stride, size_t i) {
    Du<T> r;
    r.v = v[i];
#pragma unroll
    for (int t = 0; t < T; ++t) r.d[t] = dv ? dv[t * stride + i] : 0.0;
    return r;
}
template <int T> __device__ __forceinline__ void store_tan(double* dv, size_t stride, size_t i, const Du<T>& x) {
#pragma unroll
    for (int t = 0; t < T; ++t) dv[t * stride + i] = x.d[t];
}

template <int T> __device__ __forceinline__ void g_getlm(const ModelDev& m, const Seeds<T>& sd, int a, int b, Du<T>& l, Du<T>& mu) {
    b = d_isawgd(m, b) ? d_nonwgdchild(m, b) : b;
    if (m.model_kind == BLG_NODE_MODEL) {
        a = d_isawgd(m, a) ? d_nonwgdparent(m, a) : a;
        l = (seeded<T>(m.lam[a], SEED_LAM, a, sd) + seeded<T>(m.lam[b], SEED_LAM, b, sd)) / 2.0;
        mu = (seeded<T>(m.mu[a], SEED_MU, a, sd) + seeded<T>(m.mu[b], SEED_MU, b, sd)) / 2.0;
    } else {
        l = seeded<T>(m.lam[b], SEED_LAM, b, sd);
        mu = seeded<T>(m.mu[b], SEED_MU, b, sd);
    }
}

// tangents of ϵ[1], ϵ[2] for the listed nodes (setϵ!, node.jl:136-162, in dual arithmetic).
// de0/de1: [T][ncols], zero for nodes outside the list.
template <int T>
__global__ void eps_dual_kernel(ModelDev m, Seeds<T> sd, const int* __restrict__ list, int nlist, double* de0, double* de1) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const size_t st = (size_t)m.n;
    for (int li = 0; li < nlist; ++li) {
        int n = list[li];
        int c0 = m.child_off[n], c1 = m.child_off[n + 1];
        uint8_t kd = m.kind[n];
        if (c0 == c1) continue;                                  // leaf: ϵ[2] = 0, no tangent
        if (kd == BLG_WGD || kd == BLG_WGT) {
            Du<T> q = seeded<T>(m.q[n], SEED_Q, n, sd);
            int c = m.child_idx[c0];
            Du<T> ec = load_du<T>(m.eps1, de1, st, c);
            Du<T> v = (kd == BLG_WGD) ? q * (ec * ec) + (1.0 - q) * ec
                                      : q * (ec * ec * ec) + 2.0 * q * (1.0 - q) * (ec * ec) + (1.0 - q) * (1.0 - q) * ec;
            store_tan<T>(de0, st, c, v);
            store_tan<T>(de1, st, n, v);
        } else {
            Du<T> e(1.0);
            for (int ci = c0; ci < c1; ++ci) {
                int c = m.child_idx[ci];
                Du<T> l, mu;
                g_getlm<T>(m, sd, n, c, l, mu);
                Du<T> ec = g_approx1(g_ep(l, mu, Du<T>(m.t[c]), load_du<T>(m.eps1, de1, st, c)));
                store_tan<T>(de0, st, c, ec);
                e = e * ec;
            }
            e = g_approx1(e);
            store_tan<T>(de1, st, n, e);
        }
    }
}

// tangent tables of one node: dWt [T][wdim*wdp], dipow [T][nrow], dcoefD[i] [T][nrow*cdp[i]]
struct DualTab {
    double* dWt;
    double* dipow;
    double* dcoefD[BLG_MAXK];
};

template <int T>
__global__ void __launch_bounds__(128) tables_dual_kernel(ModelDev m, Seeds<T> sd, const int* __restrict__ list, const DualTab* __restrict__ dtabs,
                                                          const double* __restrict__ de0, const double* __restrict__ de1) {
    extern __shared__ double sm[];        // 2 * maxdim * (1+T) doubles
    const int li = blockIdx.x;
    const int n = list[li];
    const NodeTab tb = m.tab[n];
    const DualTab dt = dtabs[li];
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t st = (size_t)m.n;
    uint8_t kd = m.kind[n];
    if (kd != BLG_ROOT && tb.Wt != nullptr && dt.dWt != nullptr) {
        const int d = tb.wdim, dp = tb.wdp;
        const size_t wsz = (size_t)d * dp;
        double* W = dt.dWt;
#define DWT(lane_, n_, j_) W[(size_t)(lane_) * wsz + (size_t)(j_) * dp + (n_)]
        for (size_t i = tid; i < wsz * T; i += nt) W[i] = 0.0;
        __syncthreads();
        int par = m.parent[n];
        Du<T> e = load_du<T>(m.eps1, de1, st, n);
        if (kd == BLG_WGDAFTER) {
            Du<T> q = seeded<T>(m.q[par], SEED_Q, par, sd);
            Du<T> a, b, c(0.0);
            const bool wgt = m.kind[par] == BLG_WGT;
            if (!wgt) {
                a = ((1.0 - q) + 2.0 * q * e) * (1.0 - e);
                b = q * ((1.0 - e) * (1.0 - e));
            } else {
                Du<T> q1 = q, q2 = 2.0 * q * (1.0 - q), q3 = (1.0 - q) * (1.0 - q), ome = 1.0 - e;
                a = q1 * ome + 2.0 * q2 * e * ome + 3.0 * q3 * (e * e) * ome;
                b = q2 * (ome * ome) + 3.0 * q3 * e * (ome * ome);
                c = q3 * (ome * ome * ome);
            }
            // rows are polynomial powers: row i from row i-1; kept in shared memory as Du (value needed for the product rule)
            Du<T>* prev = reinterpret_cast<Du<T>*>(sm);
            Du<T>* cur = prev + d;
            for (int j = tid; j < d; j += nt) prev[j] = Du<T>(j == 0 ? 1.0 : 0.0);          // row 0
            __syncthreads();
            for (int i = 1; i < d; ++i) {
                for (int j = tid; j < d; j += nt) {
                    Du<T> v(0.0);
                    if (i == 1) {
                        if (j == 1) v = a; else if (j == 2) v = b; else if (j == 3 && wgt) v = c;
                    } else if (!wgt) {
                        if (j >= 2) v = a * prev[j - 1] + b * prev[j - 2];
                    } else {
                        if (j >= 3) v = a * prev[j - 1] + b * prev[j - 2] + c * prev[j - 3];      // literal: column 2 of rows >= 2 stays 0
                    }
                    cur[j] = v;
#pragma unroll
                    for (int t = 0; t < T; ++t) DWT(t, i, j) = v.d[t];
                }
                __syncthreads();
                Du<T>* tmp = prev; prev = cur; cur = tmp;
            }
        } else {
            Du<T> l, mu;
            g_getlm<T>(m, sd, par, n, l, mu);
            Du<T> t(m.t[n]);
            Du<T> phi = g_getphi(t, l, mu), psi = g_getpsi(t, l, mu);
            Du<T> nn = 1.0 - psi * e;
            Du<T> phip = g_approx1((phi * (1.0 - e) + (1.0 - psi) * e) / nn);
            Du<T> psip = g_approx1(psi * (1.0 - e) / nn);
            Du<T> cc = (1.0 - phip) * (1.0 - psip);
            Du<T>* prev = reinterpret_cast<Du<T>*>(sm);
            Du<T>* cur = prev + d;
            for (int i = tid; i < d; i += nt) prev[i] = Du<T>(i == 0 ? 1.0 : 0.0);          // column 0
            __syncthreads();
            for (int j = 1; j < d; ++j) {
                for (int r = tid; r < d; r += nt) {
                    Du<T> v(0.0);
                    if (r >= 1 && r <= j) {
                        v = psip * prev[r] + cc * prev[r - 1];
#pragma unroll
                        for (int t2 = 0; t2 < T; ++t2) DWT(t2, r, j) = v.d[t2];
                    }
                    cur[r] = v;
                }
                __syncthreads();
                Du<T>* tmp = prev; prev = cur; cur = tmp;
            }
        }
#undef DWT
    }
    int c0 = m.child_off[n], c1 = m.child_off[n + 1];
    int k = c1 - c0;
    if (k > 0 && tb.ipow != nullptr && dt.dipow != nullptr) {
        Du<T> E[BLG_MAXK + 1];
        Du<T> raw(1.0);
        E[0] = Du<T>(1.0);
        for (int i = 0; i < k; ++i) {
            raw = raw * load_du<T>(m.eps0, de0, st, tb.mo[i]);
            E[i + 1] = raw.v > 1.0 ? Du<T>(1.0) : raw;
        }
        Du<T> inv = 1.0 / (1.0 - E[k]);
        for (int r = tid; r < tb.nrow; r += nt) {
            Du<T> p = g_powi(inv, r);
#pragma unroll
            for (int t = 0; t < T; ++t) dt.dipow[(size_t)t * tb.nrow + r] = p.d[t];
        }
        for (int i = 1; i < k; ++i) {
            double* cf = dt.dcoefD[i];
            int cs = tb.cdp[i];
            int smax = m.tab[tb.mo[i]].wdim - 1;
            const size_t csz = (size_t)tb.nrow * cs;
            for (int idx = tid; idx < tb.nrow * cs; idx += nt) {
                int t_ = idx / cs, s = idx - t_ * cs;
                int r = t_ + s;
                Du<T> v(0.0);
                if (r < tb.nrow && s <= smax) v = m.pascal[(size_t)r * m.pdim + s] * g_powi(E[i], s);
#pragma unroll
                for (int t = 0; t < T; ++t) cf[(size_t)t * csz + idx] = v.d[t];
            }
        }
    }
}

// dual pruning step (general kernel in dual arithmetic).  Writes only the tangent column of the node,
// expressed in the scale 2^scl of the node's stored primal column.
template <int T> struct DualChild {
    const double* dell;     // child's tangent column [T][rows][ld] or nullptr (child not perturbed)
    size_t dell_stride;     // rows*ld
    const double* dWt;      // [T][wdim*wdp] or nullptr
    size_t dWt_stride;
    const double* dcoefD;   // [T][nrow*cdp] or nullptr
    size_t dcoef_stride;
};
template <int T> struct DualStep {
    DualChild<T> child[BLG_MAXK];
    const double* dipow;    // [T][nrow] or nullptr
    int nrow;
    const double* de0;      // [T][ncols]
    size_t de_stride;
    double* dout;           // [T][rows][ld]
    size_t dout_stride;
    const int* out_scl;     // exponent of the node's stored primal column
};

template <int T, int CAP>
__global__ void __launch_bounds__(64) prune_dual_kernel(const NodeStep st, const DualStep<T> ds, const double* __restrict__ eps0, int N, size_t ld) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= N) return;
    Du<T> A0[CAP], A1[CAP], Bc[CAP], lc[CAP];
    Du<T>* Acur = A0;
    Du<T>* Anew = A1;
    int S = 0;
    int sclsum = 0;
    for (int i = 0; i < st.k; ++i) {
        const ChildRef& c = st.child[i];
        const DualChild<T>& dc = ds.child[i];
        const int Mi = c.cnt[f];
        const double* __restrict__ Wt = c.Wt;
        const int wd = c.wdp;
        if (c.ell == nullptr) {
            for (int s = 0; s <= Mi; ++s) Bc[s] = load_du<T>(Wt, dc.dWt, dc.dWt_stride, (size_t)Mi * wd + s);
        } else {
            for (int j = 0; j <= Mi; ++j) {
                Du<T> x;
                x.v = c.ell[(size_t)j * ld + f];
#pragma unroll
                for (int t = 0; t < T; ++t) x.d[t] = dc.dell ? dc.dell[t * dc.dell_stride + (size_t)j * ld + f] : 0.0;
                lc[j] = x;
            }
            sclsum += c.scl[f];
            for (int s = 0; s <= Mi; ++s) {
                Du<T> acc(0.0);
                for (int j = s; j <= Mi; ++j) acc = acc + load_du<T>(Wt, dc.dWt, dc.dWt_stride, (size_t)j * wd + s) * lc[j];
                Bc[s] = acc;
            }
        }
        if (i == 0) {
            for (int n = 0; n <= Mi; ++n) Acur[n] = Bc[n];
            S = Mi;
        } else {
            Du<T> e = load_du<T>(eps0, ds.de0, ds.de_stride, (size_t)c.id);
            const double* __restrict__ cf = c.coefD;
            const int cs = c.cdp;
            for (int n = 0; n <= S + Mi; ++n) Anew[n] = Du<T>(0.0);
            for (int t = 0; t <= S; ++t) {
                const Du<T> a = Acur[t];
                for (int s = 0; s <= Mi; ++s)
                    Anew[t + s] = Anew[t + s] + load_du<T>(cf, dc.dcoefD, dc.dcoef_stride, (size_t)t * cs + s) * (a * Bc[s]);
                for (int s = 0; s < Mi; ++s) Bc[s] = e * Bc[s] + Bc[s + 1];
                Bc[Mi] = e * Bc[Mi];
            }
            Du<T>* tmp = Acur; Acur = Anew; Anew = tmp;
            S += Mi;
        }
    }
    const int X = st.cnt[f];
    // tangent of ℓ_u[n] = Ã[n]·ipow[n], moved from scale 2^sclsum to the stored column's scale 2^out_scl
    const int shift = sclsum - ds.out_scl[f];
    for (int n = 0; n <= X; ++n) {
        Du<T> v = (n <= S ? Acur[n] : Du<T>(0.0)) * load_du<T>(st.ipow, ds.dipow, (size_t)ds.nrow, (size_t)n);
#pragma unroll
        for (int t = 0; t < T; ++t) ds.dout[t * ds.dout_stride + (size_t)n * ld + f] = scalbn(v.d[t], shift);
    }
}

// root: weights of integrate_root (model.jl:465-475) and condition_oib (model.jl:492-501) in dual arithmetic.
// rootwD: [(1+T)][nrow]: lane 0 = values, lanes 1..T = tangents; entry 0 of each lane = conditioning constant.
template <int T>
__global__ void root_tables_dual_kernel(ModelDev m, Seeds<T> sd, const double* __restrict__ de0, const double* __restrict__ de1,
                                        double* rootwD, int nrow) {
    const int root = 0;
    const size_t st = (size_t)m.n;
    Du<T> eta = seeded<T>(m.eta, SEED_ETA, 0, sd);
    Du<T> le = log(load_du<T>(m.eps1, de1, st, root));
    Du<T> a = g_log1mexp(le), b = log(eta), c = log(1.0 - eta), d = g_log1mexp(log(1.0 - eta) + le);
    for (int n = 1 + threadIdx.x; n < nrow; n += blockDim.x) {
        Du<T> f = (double)n * a + b + (double)(n - 1) * c;
        f = f - (double)(n + 1) * d;
        Du<T> w = exp(f);
        rootwD[n] = w.v;
#pragma unroll
        for (int t = 0; t < T; ++t) rootwD[(size_t)(1 + t) * nrow + n] = w.d[t];
    }
    if (threadIdx.x == 0) {
        int c0 = m.child_off[root];
        Du<T> leta = log(eta);
        Du<T> cond(0.0);
        for (int i = 0; i < 2; ++i) {
            Du<T> e = log(load_du<T>(m.eps0, de0, st, m.child_idx[c0 + i]));
            Du<T> lr = leta + e - g_log1mexp(g_log1mexp(leta) + e);
            cond = cond + g_log1mexp(lr);
        }
        rootwD[0] = cond.v;
#pragma unroll
        for (int t = 0; t < T; ++t) rootwD[(size_t)(1 + t) * nrow] = cond.d[t];
    }
}

// per family: ∂logL = ∂P/P - ∂cond; weighted block sums per lane into partial[lane][block]
template <int T>
__global__ void __launch_bounds__(256) root_dual_kernel(const double* __restrict__ ell, const double* __restrict__ dell, size_t dell_stride,
                                                        const uint16_t* __restrict__ cnt, const double* __restrict__ rootwD, int nrow,
                                                        const double* __restrict__ weight, double* __restrict__ partial, int nblocks,
                                                        int N, size_t ld) {
    __shared__ double red[256];
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double g[T];
#pragma unroll
    for (int t = 0; t < T; ++t) g[t] = 0.0;
    if (f < N) {
        int X = cnt[f];
        double p = 0.0;
        double dp[T];
#pragma unroll
        for (int t = 0; t < T; ++t) dp[t] = 0.0;
        for (int n = 1; n <= X; ++n) {
            const double l = ell[(size_t)n * ld + f];
            p = fma(l, rootwD[n], p);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const double dl = dell ? dell[t * dell_stride + (size_t)n * ld + f] : 0.0;
                dp[t] += dl * rootwD[n] + l * rootwD[(size_t)(1 + t) * nrow + n];
            }
        }
        const double w = weight[f];
#pragma unroll
        for (int t = 0; t < T; ++t) g[t] = w * (dp[t] / p - rootwD[(size_t)(1 + t) * nrow]);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
        red[threadIdx.x] = g[t];
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[(size_t)t * nblocks + blockIdx.x] = red[0];
        __syncthreads();
    }
}
// out[lane] += Σ_b partial[lane][b]   (fixed order)
__global__ void __launch_bounds__(256) grad_reduce_kernel(const double* __restrict__ partial, int nblocks, int nlanes, double* __restrict__ out) {
    __shared__ double red[256];
    for (int lane = 0; lane < nlanes; ++lane) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[(size_t)lane * nblocks + i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int k = 128; k > 0; k >>= 1) {
            if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[lane] = red[0];
        __syncthreads();
    }
}
