// adjoint.cuh — reverse-mode gradient (gradient / gradient_cr of src/model.jl:517-533, src/profile.jl:71-75).
// Included by beluga_b200.cu after gradient.cuh.
//
// The forward-mode path (gradient.cuh) costs one tangent pass per branch; this file computes the same vector from ONE
// adjoint sweep root→leaves plus one small per-parameter chain-rule kernel, for trees whose nodes have at most two
// children (binary species trees with WGD/WGT chains).  The reference differentiates with ForwardDiff
// (model.jl:517-522); what is reproduced is its result, not its method.
//
// With ℓ̂_u the stored (power-of-two scaled) column of node u and L = Σ_f w_f·log-likelihood_f:
//   stage 1  ô_u[n] = ∂L/∂ℓ̂_u[n] for every internal node, root first.  At a node with children (0, 1) in merge order
//            the forward step is  b_i = W_i ℓ̂_i,  A[n] = Σ_{t+s=n} C(n,s) e_0^s b_0[t] B_1^(t)[s]  (B_1^(0) = b_1,
//            B^(t)[s] = e_1 B^(t-1)[s] + B^(t-1)[s+1]),  ℓ̂_u[n] = 2^(-ex) A[n] (1-e_0 e_1)^(-n).  A is symmetric in
//            the two children, so each child's adjoint comes from a forward-style pass with the OTHER child's B rows:
//              b̄_0[t] = Σ_s C(t+s,s) e_0^s Ā[t+s] B_1^(t)[s],      b̄_1[t] = Σ_s C(t+s,s) e_1^s Ā[t+s] B_0^(t)[s]
//            and the same passes give ∂L/∂e_i (the explicit power s·e_i^(s-1)); no B row is ever stored.
//            Branch scalars: for an ordinary branch w*(n|j) = C(j-1,n-1) ψ'^(j-n) c^n gives ∂b[n]/∂c = n b[n]/c and
//            ∂b[n]/∂ψ' = n b[n+1]/c, so ψ̄' and c̄ are dot products of b̄ with b; the branch below a WGD/WGT contracts
//            b̄ and ℓ̂ with the dual tables ∂W/∂q, ∂W/∂ε.  Per-family values are reduced in a fixed order (per block, then over
//            the blocks), so the gradient is bit-reproducible.
//   stage 2  per parameter (same seeds as the forward-mode path): ε tangents for the whole tree in dual arithmetic,
//            then  g_p = Σ_c [ψ̄'_c dψ'_c + c̄_c dc_c + ē_c de_c] + root terms.
#pragma once

struct AdjChild {
    const double* ell;       // child's stored column (nullptr: leaf)
    const int* scl;
    const uint16_t* cnt;
    const double* Wt;        // transposed w* of the child's branch
    int wdp;
    int id;
    double* adj;             // child's adjoint column to write (internal children), [n][f]
    int* adj_scl;
    const double* dW;        // branch below a WGD/WGT: [2][wdim*wdp] = ∂W/∂q, ∂W/∂ε(child); else nullptr
    size_t dW_stride;
    double e;                // ϵ[1] of the child
    double cc;               // c = (1-ϕ')(1-ψ') of the child's branch (ordinary branches)
};
struct AdjStep {
    int k;
    AdjChild child[2];       // merge order
    const double* ell;       // the node's own stored column, exponent, counts
    const int* scl;
    const uint16_t* cnt;
    const double* adj;       // the node's adjoint column and its exponent
    const int* adj_scl;
    const double* ipow;      // (1-E_k)^(-n)
    double inv;              // 1/(1-E_k)
    double* partial;         // per-block partial sums of the six branch scalars, [6][nblocks]: child i → {ψ̄' | Q̄, c̄ | Ē(W), ē}
};

// root: ô_root[n] = w_f·rootw[n] / Σ_m rootw[m] ℓ̂_root[m];  r̄w[n] = Σ_f w_f ℓ̂_root[n] / (…) as per-block partial sums
// partial[n][block] (fixed-order reductions throughout: the gradient is bit-reproducible)
__global__ void __launch_bounds__(256) root_adjoint_kernel(const double* __restrict__ ell, const uint16_t* __restrict__ cnt,
                                                           const double* __restrict__ rootw, const double* __restrict__ weight,
                                                           double* __restrict__ adj, int* __restrict__ adj_scl,
                                                           double* __restrict__ partial, int nblocks, int nrow, int N, size_t ld) {
    __shared__ double red[8];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int X = -1;
    double wr = 0.0;
    if (f < N) {
        X = cnt[f];
        double p = 0.0;
        for (int n = 1; n <= X; ++n) p = fma(ell[(size_t)n * ld + f], rootw[n], p);
        const bool ok = p > 0.0 && isfinite(p);
        wr = ok ? weight[f] / p : 0.0;
        double mx = 0.0;
        for (int n = 1; n <= X; ++n) mx = fmax(mx, wr * rootw[n]);
        int m = 0;
        if (mx > 0.0 && isfinite(mx)) m = ilogb(mx);
        adj[f] = 0.0;
        for (int n = 1; n <= X; ++n) adj[(size_t)n * ld + f] = scalbn(wr * rootw[n], -m);
        adj_scl[f] = m;
    }
    for (int n = 0; n < nrow; ++n) {
        double v = (n >= 1 && n <= X) ? wr * ell[(size_t)n * ld + f] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0)
            partial[(size_t)n * nblocks + blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
        __syncthreads();
    }
}

// one adjoint step at an internal node with one or two children, every family
template <int CAP>
__global__ void __launch_bounds__(128) adjoint_node_kernel(const AdjStep st, const double* __restrict__ pascal, int pdim, int N, size_t ld) {
    __shared__ double red[4][6];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    double sc[2][3];                   // per child: {ψ̄' | Q̄, c̄ | Ē(W), ē}
#pragma unroll
    for (int i = 0; i < 2; ++i) { sc[i][0] = 0.0; sc[i][1] = 0.0; sc[i][2] = 0.0; }
    if (f < N) {
        const int X = st.cnt[f];
        double Ab[CAP], b0[CAP], b1[CAP], Br[CAP], bb[CAP], lc[CAP];
        int sclc = 0;
        for (int i = 0; i < st.k; ++i) if (st.child[i].scl) sclc += st.child[i].scl[f];
        const int au = st.adj_scl[f];
        const int kappa = au - (st.scl[f] - sclc);
        // Ā[n] = ô_u[n]·2^(-ex)·(1-E_k)^(-n) (in units of 2^kappa); adjoint of the normalisation
        double Ekb = 0.0;
        for (int n = 0; n <= X; ++n) {
            const double o = st.adj[(size_t)n * ld + f];
            Ab[n] = o * st.ipow[n];
            Ekb = fma((double)n * o, st.ell[(size_t)n * ld + f], Ekb);
        }
        Ekb = scalbn(Ekb * st.inv, au);
        // b_i = W_i ℓ̂_i
        int M[2] = {0, 0};
        for (int i = 0; i < st.k; ++i) {
            const AdjChild& c = st.child[i];
            const int Mi = c.cnt[f];
            M[i] = Mi;
            double* b = i == 0 ? b0 : b1;
            if (c.ell == nullptr) {
                for (int s = 0; s <= Mi; ++s) b[s] = c.Wt[(size_t)Mi * c.wdp + s];
            } else {
                for (int j = 0; j <= Mi; ++j) lc[j] = c.ell[(size_t)j * ld + f];
                for (int s = 0; s <= Mi; ++s) {
                    double acc = 0.0;
                    for (int j = s; j <= Mi; ++j) acc = fma(c.Wt[(size_t)j * c.wdp + s], lc[j], acc);
                    b[s] = acc;
                }
            }
        }
        const double e0 = st.child[0].e, e1 = st.k == 2 ? st.child[1].e : 1.0;
        if (st.k == 2) {
            const double prod = e0 * e1;
            if (prod < 1.0) { sc[0][2] += Ekb * e1; sc[1][2] += Ekb * e0; }
        } else if (e0 < 1.0) {
            sc[0][2] += Ekb;
        }
        for (int i = 0; i < st.k; ++i) {
            const AdjChild& c = st.child[i];
            const int Mi = M[i];
            const double* bi = i == 0 ? b0 : b1;
            // ---- b̄_i ----
            if (st.k == 1) {
                for (int t = 0; t <= Mi; ++t) bb[t] = t <= X ? Ab[t] : 0.0;
            } else {
                const double* bo = i == 0 ? b1 : b0;         // the other child: its B rows
                const int Mo = M[1 - i];
                const double ei = i == 0 ? e0 : e1, eo = i == 0 ? e1 : e0;
                const double Ei = ei > 1.0 ? 1.0 : ei;
                for (int s = 0; s <= Mo; ++s) Br[s] = bo[s];
                double eacc = 0.0;
                for (int t = 0; t <= Mi; ++t) {
                    double acc = 0.0, accs = 0.0, pw = 1.0;
                    for (int s = 0; s <= Mo; ++s) {
                        const int n = t + s;
                        const double cf = n <= X ? pascal[(size_t)n * pdim + s] * pw * Ab[n] * Br[s] : 0.0;
                        acc += cf;
                        accs = fma((double)s, cf, accs);
                        pw *= Ei;
                    }
                    bb[t] = acc;
                    eacc = fma(bi[t], accs, eacc);
                    for (int s = 0; s < Mo; ++s) Br[s] = fma(eo, Br[s], Br[s + 1]);
                    Br[Mo] = eo * Br[Mo];
                }
                if (ei < 1.0 && ei > 0.0) sc[i][2] += scalbn(eacc / ei, kappa);
            }
            // ---- branch scalars ----
            if (c.dW == nullptr) {                         // ordinary branch: closed form in (ψ', c)
                double ps = 0.0, cs = 0.0;
                for (int n = 1; n <= Mi; ++n) {
                    const double w = (double)n * bb[n];
                    cs = fma(w, bi[n], cs);
                    if (n < Mi) ps = fma(w, bi[n + 1], ps);
                }
                sc[i][0] += scalbn(ps / c.cc, kappa);
                sc[i][1] += scalbn(cs / c.cc, kappa);
            } else {                                       // branch below a WGD/WGT: dual tables
                for (int j = 0; j <= Mi; ++j) lc[j] = c.ell ? c.ell[(size_t)j * ld + f] : (j == Mi ? 1.0 : 0.0);
                double q = 0.0, ee = 0.0;
                for (int s = 0; s <= Mi; ++s) {
                    double aq = 0.0, ae = 0.0;
                    for (int j = s; j <= Mi; ++j) {
                        aq = fma(c.dW[(size_t)j * c.wdp + s], lc[j], aq);
                        ae = fma(c.dW[c.dW_stride + (size_t)j * c.wdp + s], lc[j], ae);
                    }
                    q = fma(bb[s], aq, q);
                    ee = fma(bb[s], ae, ee);
                }
                sc[i][0] += scalbn(q, kappa);
                sc[i][1] += scalbn(ee, kappa);
            }
            // ---- the child's adjoint column ô_i = W_iᵀ b̄_i ----
            if (c.adj != nullptr) {
                double mx = 0.0;
                for (int j = 0; j <= Mi; ++j) {
                    double acc = 0.0;
                    for (int s = 0; s <= j; ++s) acc = fma(c.Wt[(size_t)j * c.wdp + s], bb[s], acc);
                    lc[j] = acc;
                    mx = fmax(mx, fabs(acc));
                }
                int m = 0;
                if (mx > 0.0 && isfinite(mx)) m = ilogb(mx);
                for (int j = 0; j <= Mi; ++j) c.adj[(size_t)j * ld + f] = scalbn(lc[j], -m);
                c.adj_scl[f] = kappa + m;
            }
        }
    }
    // block reduction of the six scalars, fixed order; the blocks are combined by adjoint_reduce_kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double v = sc[i][j];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][i * 3 + j] = v;
        }
    __syncthreads();
    if (threadIdx.x < 6)
        st.partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] =
            (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// sadj[id_i][j] = Σ_b partial[3i + j][b]   (every node is the child of exactly one step: plain stores)
__global__ void __launch_bounds__(256) adjoint_reduce_kernel(const double* __restrict__ partial, int nblocks, int k, int id0, int id1,
                                                             double* __restrict__ sadj) {
    __shared__ double red[256];
    for (int slot = 0; slot < 3 * k; ++slot) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[(size_t)slot * nblocks + i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) sadj[(size_t)(slot < 3 ? id0 : id1) * 4 + slot % 3] = red[0];
        __syncthreads();
    }
}

// stage 2: one block per parameter group (two tangent lanes).  ε tangents of the whole tree, then the chain rule.
// rwbar[0] is unused; wtot = Σ_f w_f (weight of the conditioning term)
__global__ void __launch_bounds__(32) adjoint_params_kernel(ModelDev m, const Seeds<2>* __restrict__ groups, const int* __restrict__ post,
                                                            int npost, const double* __restrict__ sadj, const double* __restrict__ rwbar,
                                                            int nrow, double wtot, double* __restrict__ out) {
    extern __shared__ double dsm[];            // de0[2][n] | de1[2][n]
    const int n = m.n;
    double* de0 = dsm;
    double* de1 = dsm + 2 * (size_t)n;
    const Seeds<2> sd = groups[blockIdx.x];
    const size_t st = (size_t)n;
    for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) dsm[i] = 0.0;
    __syncwarp();
    if (threadIdx.x == 0) {                    // setϵ! in dual arithmetic, leaves → root (node.jl:136-162)
        for (int li = 0; li < npost; ++li) {
            const int u = post[li];
            const int c0 = m.child_off[u], c1 = m.child_off[u + 1];
            const uint8_t kd = m.kind[u];
            if (c0 == c1) continue;
            if (kd == BLG_WGD || kd == BLG_WGT) {
                Du<2> q = seeded<2>(m.q[u], SEED_Q, u, sd);
                const int c = m.child_idx[c0];
                Du<2> ec = load_du<2>(m.eps1, de1, st, c);
                Du<2> v = (kd == BLG_WGD) ? q * (ec * ec) + (1.0 - q) * ec
                                          : q * (ec * ec * ec) + 2.0 * q * (1.0 - q) * (ec * ec) + (1.0 - q) * (1.0 - q) * ec;
                store_tan<2>(de0, st, c, v);
                store_tan<2>(de1, st, u, v);
            } else {
                Du<2> e(1.0);
                for (int ci = c0; ci < c1; ++ci) {
                    const int c = m.child_idx[ci];
                    Du<2> l, mu;
                    g_getlm<2>(m, sd, u, c, l, mu);
                    Du<2> ec = g_approx1(g_ep(l, mu, Du<2>(m.t[c]), load_du<2>(m.eps1, de1, st, c)));
                    store_tan<2>(de0, st, c, ec);
                    e = e * ec;
                }
                e = g_approx1(e);
                store_tan<2>(de1, st, u, e);
            }
        }
    }
    __syncwarp();
    double g[2] = {0.0, 0.0};
    for (int li = threadIdx.x; li < npost; li += blockDim.x) {
        const int c = post[li];
        const uint8_t kd = m.kind[c];
        if (kd == BLG_ROOT) continue;
        const double* a = sadj + (size_t)c * 4;
        const int par = m.parent[c];
#pragma unroll
        for (int t = 0; t < 2; ++t) g[t] = fma(a[2], de0[t * st + c], g[t]);
        if (kd == BLG_WGDAFTER) {
            Du<2> q = seeded<2>(m.q[par], SEED_Q, par, sd);
#pragma unroll
            for (int t = 0; t < 2; ++t) g[t] += a[0] * q.d[t] + a[1] * de1[t * st + c];
        } else {
            Du<2> l, mu;
            g_getlm<2>(m, sd, par, c, l, mu);
            Du<2> tt(m.t[c]);
            Du<2> e = load_du<2>(m.eps1, de1, st, c);
            Du<2> phi = g_getphi(tt, l, mu), psi = g_getpsi(tt, l, mu);
            Du<2> nn = 1.0 - psi * e;
            Du<2> phip = g_approx1((phi * (1.0 - e) + (1.0 - psi) * e) / nn);
            Du<2> psip = g_approx1(psi * (1.0 - e) / nn);
            Du<2> cc = (1.0 - phip) * (1.0 - psip);
#pragma unroll
            for (int t = 0; t < 2; ++t) g[t] += a[0] * psip.d[t] + a[1] * cc.d[t];
        }
    }
    {   // root: integrate_root weights and the conditioning constant (model.jl:465-501) in dual arithmetic
        const int root = 0;
        Du<2> eta = seeded<2>(m.eta, SEED_ETA, 0, sd);
        Du<2> le = log(load_du<2>(m.eps1, de1, st, root));
        Du<2> a = g_log1mexp(le), b = log(eta), c = log(1.0 - eta), d = g_log1mexp(log(1.0 - eta) + le);
        for (int r = 1 + threadIdx.x; r < nrow; r += blockDim.x) {
            Du<2> f = (double)r * a + b + (double)(r - 1) * c;
            f = f - (double)(r + 1) * d;
            Du<2> w = exp(f);
#pragma unroll
            for (int t = 0; t < 2; ++t) g[t] = fma(rwbar[r], w.d[t], g[t]);
        }
        if (threadIdx.x == 0) {
            const int c0 = m.child_off[root];
            Du<2> leta = log(eta);
            Du<2> cond(0.0);
            for (int i = 0; i < 2; ++i) {
                Du<2> e = log(load_du<2>(m.eps0, de0, st, m.child_idx[c0 + i]));
                Du<2> lr = leta + e - g_log1mexp(g_log1mexp(leta) + e);
                cond = cond + g_log1mexp(lr);
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) g[t] -= wtot * cond.d[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        double v = g[t];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[(size_t)blockIdx.x * 2 + t] = v;
    }
}
