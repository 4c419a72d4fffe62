// padjoint.cuh — the reverse-mode gradient sweep (adjoint.cuh) over NODE PATTERNS (pattern.cuh) instead of families.
// Included by beluga_b200.cu after adjoint.cuh.  Backs gradient / gradient_cr (src/model.jl:517-533, src/profile.jl:71-75).
//
// The forward sweep stores one column per distinct sub-pattern of every node; many patterns of a node share one pattern
// of a child.  The adjoint of a child's pattern column is therefore the SUM of what the node's patterns send down:
//   phase A (padj_parent_kernel, every node of a tree depth in one launch; one thread per pattern of the node, or one warp
//            per pattern where a depth holds few patterns): Ā from ô_u, b_i = W_i ℓ̂_i from the children's pattern columns,
//            both b̄_i from one set of rows of the child with fewer counts (padj_merge2_reg), the branch scalars ψ̄', c̄
//            (or Q̄, Ē below a WGD/WGT) and ē_i — b̄_i of an internal child is stored per pattern of the NODE instead of
//            being pushed through W_iᵀ;
//   phase B (padj_child_kernel, every internal child of a tree depth in one launch; one thread, warp or block per pattern of
//            the child): sums b̄_i over the node patterns that use the child pattern (stored next to each other: the
//            inverse of the link's pattern map; fixed order: the gradient stays bit-reproducible) and applies W_iᵀ once
//            per child pattern: ô_i = W_iᵀ Σ b̄_i.
// Single-child nodes (WGD-type nodes) share their child's patterns one to one: phase A writes ô_i directly.
//
// Scaling.  L = Σ_f w_f log L_f and L_f is linear in every node column, so Σ_n ô_u[n] ℓ̂_u[n] = (weight of the families
// with this pattern) for the STORED (power-of-two scaled) column ℓ̂_u: adjoints relative to the stored columns need no
// exponent of their own.  Entries that would leave the double range (ℓ̂_u[n] below 1e-300 of the column's top) are clamped.
#pragma once

#define BLG_PB_HIT 48                 // phase B: child patterns with more node patterns than this (or, at depths with few patterns,
#define BLG_PB_MID 16                 //   with at least BLG_PB_MID counts) are summed by a warp,
#define BLG_PB_BLOCK 2048             //   with more node patterns than this by a whole block

struct PAChild {
    const double* ell;       // child's stored pattern columns (nullptr: leaf), [n][ld]
    const int* scl;
    const uint16_t* cnt;     // child's count per CHILD pattern
    const int* pidx;         // node pattern → child pattern (nullptr: identity)
    const uint16_t* xcnt;    // child's count per NODE pattern (nullptr with an identity link)
    const double* Wt;        // transposed w* of the child's branch
    const double* dW;        // branch below a WGD/WGT: [2][wdim*wdp] = ∂W/∂q, ∂W/∂ε(child); else nullptr
    double* bbar;            // internal child of a two-child node: b̄ per NODE pattern, [t][ld of the node], the node patterns
    const int* pos;          //   in the order of the child patterns they use (pos[p]): phase B sums contiguous runs
    double* adj;             // identity link: the child's adjoint columns, written directly
    size_t dW_stride;
    int wdp, id, ld, pad_;
};
struct PAStep {
    int k, node, U, ld;      // the node's patterns and their stride
    int blk0;                // first block of the step inside its launch
    int pblk0;               // first slot of the step's per-block partial sums
    int nblk;
    int T;                   // patterns with at least T counts whose children both have at least msmin go to the warp path
    int wblk0, nwblk;        // the step's blocks inside the warp kernel's launch (4 patterns each)
    int wpblk0, nbig;        // first slot of their partial sums; number of such patterns
    int msmin, pad_;
    const int* big;          // their indices
    PAChild child[2];        // merge order
    const double* ell;       // the node's stored columns, exponents, counts
    const int* scl;
    const uint16_t* cnt;
    const double* adj;       // the node's adjoint columns
    const double* ipow;      // (1-E_k)^(-n)
};
struct PBStep {
    int Uc, ldc, ldp, blk0;  // child patterns, their stride, the node's stride, first block inside the launch
    int hblk0, nhit, wdp;
    int midM;                // child patterns with at least this many counts go to the warp kernel (small tree depths only)
    int mblk0, nmid;         // blocks of the step inside the warp kernel's launch (4 child patterns each)
    const int* mid;          // the child patterns of the warp kernel
    const uint16_t* cnt;     // child's counts
    const int* off;          // per child pattern: its run of node patterns in b̄ (PAChild::pos)
    const int* hit;          // the child patterns of the block kernel
    const double* bbar;
    const double* Wt;
    double* adj;
};

__device__ __forceinline__ double padj_clamp(double v) { return fmin(v, 1e300); }
__constant__ double c_rinv[257];      // 1/s, s = 1..256 (filled by the host at the first gradient)

// which patterns of a step the warp path takes (the host builds PAStep::big with the same rule)
__host__ __device__ __forceinline__ bool padj_is_big(int k, int T, int msmin, int X, int M0, int M1) {
    return X >= T && (k == 1 || (M0 < M1 ? M0 : M1) >= msmin);
}

template <class Step>
__device__ __forceinline__ int padj_find(const Step* __restrict__ steps, int nsteps, int block) {
    int lo = 0, hi = nsteps - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (steps[mid].blk0 <= block) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// root: ô_root[n] = w_p·rootw[n] / Σ_m rootw[m] ℓ̂_root[m] per root pattern p (w_p: weight of the families with the
// pattern);  r̄w[n] = Σ_p w_p ℓ̂_root[n] / (…) as per-block partial sums partial[n][block]
__global__ void __launch_bounds__(256) root_padjoint_kernel(const double* __restrict__ ell, const uint16_t* __restrict__ cnt,
                                                            const double* __restrict__ rootw, const double* __restrict__ pweight,
                                                            double* __restrict__ adj, double* __restrict__ partial, int nblocks, int nrow,
                                                            int U, size_t ld) {
    __shared__ double red[8];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int X = -1;
    double wr = 0.0;
    if (p < U) {
        X = cnt[p];
        double s = 0.0;
        for (int n = 1; n <= X; ++n) s = fma(ell[(size_t)n * ld + p], rootw[n], s);
        const bool ok = s > 0.0 && isfinite(s);
        wr = ok ? pweight[p] / s : 0.0;
        adj[p] = 0.0;
        for (int n = 1; n <= X; ++n) adj[(size_t)n * ld + p] = padj_clamp(wr * rootw[n]);
    }
    for (int n = 0; n < nrow; ++n) {
        double v = (n >= 1 && n <= X) ? wr * ell[(size_t)n * ld + p] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0)
            partial[(size_t)n * nblocks + blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
        __syncthreads();
    }
}

// fixed-order sum over the lanes of a warp, result in every lane
__device__ __forceinline__ double padj_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// phase A for a pattern with many counts: one WARP per pattern, the columns in shared memory, lanes over the inner index of
// every loop of the thread path (a thread there walks O(M0·M1 + M0² + M1²) dependent steps, which for the heaviest
// patterns of a tree depth takes longer than everything else at that depth)
__device__ __forceinline__ void padj_warp_pattern(const PAStep& st, int item, const NodeK* __restrict__ nk, const double* __restrict__ pascal,
                                                  int pdim, double* __restrict__ Ab, int cap, double* __restrict__ P, double (&sc)[2][3]) {
    const int lane = threadIdx.x & 31;
    const size_t ld = (size_t)st.ld;
    double* bq[2] = {Ab + cap, Ab + 2 * cap};
    double* Brq[2] = {Ab + 3 * cap, Ab + 4 * cap};
    double* bb = Ab + 5 * cap;
    double* lc = Ab + 6 * cap;
    double* pws = Ab + 7 * cap;
    if (item < st.nbig) {
        const int p = st.big[item];
        const int X = st.cnt[p];
        int cidx[2] = {p, p}, M[2] = {0, 0};
        int ex = st.scl[p];
        for (int i = 0; i < st.k; ++i) {
            const PAChild& c = st.child[i];
            if (c.pidx) cidx[i] = c.pidx[p];
            M[i] = c.xcnt ? (int)c.xcnt[p] : (int)c.cnt[cidx[i]];
            if (c.scl) ex -= c.scl[cidx[i]];
        }
        double Ekb = 0.0;
        for (int n = lane; n <= X; n += 32) {
            const double o = st.adj[(size_t)n * ld + p];
            Ab[n] = padj_clamp(scalbn(o * st.ipow[n], -ex));
            Ekb = fma((double)n * o, st.ell[(size_t)n * ld + p], Ekb);
        }
        const double SAA = padj_warp_sum(Ekb);             // Σ_n n Ā[n] A[n]
        Ekb = SAA * nk[st.node].inv;
        for (int i = 0; i < st.k; ++i) {
            const PAChild& c = st.child[i];
            const int Mi = M[i];
            double* b = bq[i];
            if (c.ell == nullptr) {
                for (int s = lane; s <= Mi; s += 32) b[s] = c.Wt[(size_t)Mi * c.wdp + s];
            } else {
                __syncwarp();
                for (int j = lane; j <= Mi; j += 32) lc[j] = c.ell[(size_t)j * c.ld + cidx[i]];
                __syncwarp();
                for (int s = lane; s <= Mi; s += 32) {
                    double acc = 0.0;
                    for (int j = s; j <= Mi; ++j) acc = fma(c.Wt[(size_t)j * c.wdp + s], lc[j], acc);
                    b[s] = acc;
                }
            }
        }
        __syncwarp();
        const double e0 = nk[st.child[0].id].e, e1 = st.k == 2 ? nk[st.child[1].id].e : 1.0;
        if (lane == 0) {
            if (st.k == 2) {
                if (e0 * e1 < 1.0) { sc[0][2] += Ekb * e1; sc[1][2] += Ekb * e0; }
            } else if (e0 < 1.0) {
                sc[0][2] += Ekb;
            }
        }
        const bool two = st.k == 2 && st.child[0].dW == nullptr && st.child[1].dW == nullptr && (M[0] < M[1] ? M[0] : M[1]) <= 31;
        if (two) {
            // both b̄ from the rows of the child with fewer counts (padj_merge2_reg explains the recursions): lane s holds
            // B_S^(t)[s], then R^(t)[s]; the products K(t,s)·B_S^(t)[s] of 32 steps t go to P, whose rows are then summed by
            // one lane each (no reduction across lanes anywhere)
            const int L = M[0] >= M[1] ? 0 : 1, S = 1 - L;
            const int ML = M[L], MS = M[S];
            const PAChild& cL = st.child[L];
            const PAChild& cS = st.child[S];
            const double* bL = bq[L];
            const double* bS = bq[S];
            const double eL = L == 0 ? e0 : e1, eS = L == 0 ? e1 : e0;
            const double EL = eL > 1.0 ? 1.0 : eL;
            double* outL = (cL.bbar != nullptr && ML > 0) ? cL.bbar + cL.pos[p] : nullptr;
            double* outS = (cS.bbar != nullptr && MS > 0) ? cS.bbar + cS.pos[p] : nullptr;
            double* rinv = pws;
            for (int q = lane; q < cap; q += 32) rinv[q] = c_rinv[q < 256 ? q : 256];
            __syncwarp();
            const bool on = lane <= MS;
            double row = on ? bS[lane] : 0.0;
            double cb = 0.0, nd = (double)lane;
            if (on) {
                double pw = 1.0, base = EL;                    // C(s, s)·E_L^s
                for (int q = lane; q > 0; q >>= 1) { if (q & 1) pw *= base; base *= base; }
                cb = pw;
            }
            double csL = 0.0, psL = 0.0;
            for (int t0 = 0; t0 <= ML; t0 += 32) {
                const int tn = ML + 1 - t0 < 32 ? ML + 1 - t0 : 32;
                for (int r = 0; r < tn; ++r) {
                    const int t = t0 + r;
                    const double K = on ? cb * Ab[t + lane] : 0.0;
                    P[r * 33 + lane] = K * row;
                    double up = __shfl_down_sync(0xffffffffu, row, 1);
                    if (lane == 31) up = 0.0;
                    row = fma(eS, row, up);
                    nd += 1.0;
                    cb *= nd * rinv[t + 1];                    // C(t+1+s, s) = C(t+s, s)·(t+s+1)/(t+1)
                }
                __syncwarp();
                if (lane < tn) {
                    const int t = t0 + lane;
                    double acc = 0.0;
                    for (int q = 0; q <= MS; ++q) acc += P[lane * 33 + q];
                    acc = padj_clamp(acc);
                    const double w = (double)t * acc;
                    csL = fma(w, bL[t], csL);
                    if (t < ML) psL = fma(w, bL[t + 1], psL);
                    if (outL) outL[(size_t)t * ld] = acc;
                }
                __syncwarp();
            }
            double R = 0.0;
            for (int t = ML; t >= 0; --t) {
                if (on) cb *= (double)(t + 1) * rinv[t + 1 + lane];   // C(t+s, s) = C(t+1+s, s)·(t+1)/(t+1+s)
                const double K = on ? cb * Ab[t + lane] : 0.0;
                double dn = __shfl_up_sync(0xffffffffu, R, 1);
                if (lane == 0) dn = 0.0;
                R = on ? fma(K, bL[t], fma(eS, R, dn)) : 0.0;
            }
            double csS = 0.0, psS = 0.0;
            if (on) {
                const double r = padj_clamp(R);
                const double w = (double)lane * r;
                csS = w * bS[lane];
                if (lane < MS) psS = w * bS[lane + 1];
                if (outS) outS[(size_t)lane * ld] = r;
            }
            const double ccL = nk[cL.id].cc, ccS = nk[cS.id].cc;
            sc[L][0] += psL / ccL; sc[L][1] += csL / ccL;
            sc[S][0] += psS / ccS; sc[S][1] += csS / ccS;
            const double saa = lane == 0 ? SAA : 0.0;          // (per-lane shares: the block reduction adds them up)
            if (eL < 1.0 && eL > 0.0) sc[L][2] += (saa - csL) / eL;
            if (eS < 1.0 && eS > 0.0) sc[S][2] += (saa - csS) / eS;
            __syncwarp();
        }
        for (int i = 0; i < (two ? 0 : st.k); ++i) {
            const PAChild& c = st.child[i];
            const int Mi = M[i];
            const double* bi = bq[i];
            if (st.k == 1) {
                for (int t = lane; t <= Mi; t += 32) bb[t] = t <= X ? Ab[t] : 0.0;
            } else {
                const double* bo = bq[1 - i];
                const int Mo = M[1 - i];
                const double ei = i == 0 ? e0 : e1, eo = i == 0 ? e1 : e0;
                const double Ei = ei > 1.0 ? 1.0 : ei;
                for (int s = lane; s <= Mo; s += 32) {
                    Brq[0][s] = bo[s];
                    double pw = 1.0, base = Ei;                  // Ei^s
                    for (int q = s; q > 0; q >>= 1) { if (q & 1) pw *= base; base *= base; }
                    pws[s] = pw;
                }
                __syncwarp();
                double eacc = 0.0;
                int cur = 0;
                for (int t = 0; t <= Mi; ++t) {
                    const double* Br = Brq[cur];
                    double* Bn = Brq[cur ^ 1];
                    double acc = 0.0, accs = 0.0;
                    for (int s = lane; s <= Mo; s += 32) {
                        const int n = t + s;
                        const double br = Br[s];
                        const double cf = n <= X ? pascal[(size_t)n * pdim + s] * pws[s] * Ab[n] * br : 0.0;
                        acc += cf;
                        accs = fma((double)s, cf, accs);
                        Bn[s] = s < Mo ? fma(eo, br, Br[s + 1]) : eo * br;
                    }
                    eacc = fma(bi[t], accs, eacc);
                    acc = padj_warp_sum(acc);
                    if (lane == 0) bb[t] = padj_clamp(acc);
                    cur ^= 1;
                    __syncwarp();
                }
                if (ei < 1.0 && ei > 0.0) sc[i][2] += eacc / ei;
            }
            __syncwarp();
            if (c.dW == nullptr) {
                double ps = 0.0, cs = 0.0;
                for (int n = 1 + lane; n <= Mi; n += 32) {
                    const double w = (double)n * bb[n];
                    cs = fma(w, bi[n], cs);
                    if (n < Mi) ps = fma(w, bi[n + 1], ps);
                }
                const double cc = nk[c.id].cc;
                sc[i][0] += ps / cc;
                sc[i][1] += cs / cc;
            } else {
                for (int j = lane; j <= Mi; j += 32) lc[j] = c.ell ? c.ell[(size_t)j * c.ld + cidx[i]] : (j == Mi ? 1.0 : 0.0);
                __syncwarp();
                double q = 0.0, ee = 0.0;
                for (int s = lane; s <= Mi; s += 32) {
                    double aq = 0.0, ae = 0.0;
                    for (int j = s; j <= Mi; ++j) {
                        aq = fma(c.dW[(size_t)j * c.wdp + s], lc[j], aq);
                        ae = fma(c.dW[c.dW_stride + (size_t)j * c.wdp + s], lc[j], ae);
                    }
                    q = fma(bb[s], aq, q);
                    ee = fma(bb[s], ae, ee);
                }
                sc[i][0] += q;
                sc[i][1] += ee;
            }
            if (c.bbar != nullptr && Mi > 0) {
                const int pp = c.pos[p];
                for (int t = lane; t <= Mi; t += 32) c.bbar[(size_t)t * ld + pp] = bb[t];
            } else if (c.adj != nullptr && Mi > 0) {
                for (int j = lane; j <= Mi; j += 32) {
                    double acc = 0.0;
                    for (int s = 0; s <= j; ++s) acc = fma(c.Wt[(size_t)j * c.wdp + s], bb[s], acc);
                    c.adj[(size_t)j * c.ld + cidx[i]] = padj_clamp(acc);
                }
            }
            __syncwarp();
        }
    }
}

// b̄ of both children of a two-child node from ONE set of rows.  With L the child with more counts and S the other,
//   A[n] = Σ_{t+s=n} K(t,s)/Ā[n] · b_L[t] · B_S^(t)[s],   K(t,s) = C(t+s,s)·E_L^s·Ā[t+s],   B_S^(t)[s] = e_S B_S^(t-1)[s] + B_S^(t-1)[s+1]:
//   b̄_L[t] = Σ_s K(t,s) B_S^(t)[s]                                   (t ascending, the row B_S^(t) updated in place)
//   b̄_S    = R^(0),  R^(t)[s] = K(t,s) b_L[t] + e_S R^(t+1)[s] + R^(t+1)[s-1]   (t descending: the adjoint of the row recursion)
// so the only vectors that change in the loops have the SHORTER child's length: in registers when it is below RS.
// Also returns cs_i = Σ_n n b̄_i[n] b_i[n] and ps_i = Σ_n n b̄_i[n] b_i[n+1] (the branch scalars, and through
// Σ_n n Ā[n] A[n] - cs_i the derivative with respect to e_i: every term of A[n] carries e_i to the power n - (index of b_i)).
// (K is built along s from C(t+s+1, s+1) = C(t+s, s)·(t+s+1)/(s+1) with the reciprocals from constant memory: no table
// look-up per term)
template <int RS>
__device__ __forceinline__ void padj_merge2_reg(const double* __restrict__ Ab, const double* __restrict__ bL, const double* __restrict__ bS,
                                                int ML, int MS, double EL, double eS, double* __restrict__ outL, double* __restrict__ outS,
                                                size_t ld, double& csL, double& psL, double& csS, double& psS) {
    double row[RS];
#pragma unroll
    for (int s = 0; s < RS; ++s) row[s] = s <= MS ? bS[s] : 0.0;
    double cs = 0.0, ps = 0.0;
    for (int t = 0; t <= ML; ++t) {
        double acc = 0.0, cb = 1.0, nd = (double)t;
#pragma unroll
        for (int s = 0; s < RS; ++s)
            if (s <= MS) {
                acc = fma(cb * Ab[t + s], row[s], acc);
                nd += 1.0;
                cb *= nd * (EL * c_rinv[s + 1]);
            }
        acc = padj_clamp(acc);
        const double w = (double)t * acc;
        cs = fma(w, bL[t], cs);
        if (t < ML) ps = fma(w, bL[t + 1], ps);
        if (outL) outL[(size_t)t * ld] = acc;
#pragma unroll
        for (int s = 0; s < RS - 1; ++s) row[s] = fma(eS, row[s], row[s + 1]);
        row[RS - 1] *= eS;
    }
    csL = cs; psL = ps;
#pragma unroll
    for (int s = 0; s < RS; ++s) row[s] = 0.0;                 // R
    for (int t = ML; t >= 0; --t) {
        const double bl = bL[t];
        double cb = bl, nd = (double)t, prev = 0.0;
#pragma unroll
        for (int s = 0; s < RS; ++s)
            if (s <= MS) {
                const double old = row[s];
                row[s] = fma(cb, Ab[t + s], fma(eS, old, prev));
                prev = old;
                nd += 1.0;
                cb *= nd * (EL * c_rinv[s + 1]);
            }
    }
    cs = 0.0; ps = 0.0;
#pragma unroll
    for (int s = 0; s < RS; ++s)
        if (s <= MS) {
            const double r = padj_clamp(row[s]);
            const double w = (double)s * r;
            cs = fma(w, bS[s], cs);
            if (s < MS) ps = fma(w, bS[s + 1], ps);
            if (outS) outS[(size_t)s * ld] = r;
        }
    csS = cs; psS = ps;
}
// the same with the rows in (local) memory: both children have many counts
__device__ __forceinline__ void padj_merge2_mem(const double* __restrict__ Ab, const double* __restrict__ bL, const double* __restrict__ bS,
                                                int ML, int MS, double EL, double eS, double* __restrict__ outL, double* __restrict__ outS,
                                                size_t ld, double& csL, double& psL, double& csS, double& psS, double* __restrict__ row) {
    for (int s = 0; s <= MS; ++s) row[s] = bS[s];
    double cs = 0.0, ps = 0.0;
    for (int t = 0; t <= ML; ++t) {
        double acc = 0.0, cb = 1.0, nd = (double)t, cur = row[0];
        for (int s = 0; s <= MS; ++s) {
            const double nxt = s < MS ? row[s + 1] : 0.0;
            acc = fma(cb * Ab[t + s], cur, acc);
            row[s] = fma(eS, cur, nxt);
            cur = nxt;
            nd += 1.0;
            cb *= nd * (EL * c_rinv[s + 1]);
        }
        acc = padj_clamp(acc);
        const double w = (double)t * acc;
        cs = fma(w, bL[t], cs);
        if (t < ML) ps = fma(w, bL[t + 1], ps);
        if (outL) outL[(size_t)t * ld] = acc;
    }
    csL = cs; psL = ps;
    for (int s = 0; s <= MS; ++s) row[s] = 0.0;
    for (int t = ML; t >= 0; --t) {
        double cb = bL[t], nd = (double)t, prev = 0.0;
        for (int s = 0; s <= MS; ++s) {
            const double old = row[s];
            row[s] = fma(cb, Ab[t + s], fma(eS, old, prev));
            prev = old;
            nd += 1.0;
            cb *= nd * (EL * c_rinv[s + 1]);
        }
    }
    cs = 0.0; ps = 0.0;
    for (int s = 0; s <= MS; ++s) {
        const double r = padj_clamp(row[s]);
        const double w = (double)s * r;
        cs = fma(w, bS[s], cs);
        if (s < MS) ps = fma(w, bS[s + 1], ps);
        if (outS) outS[(size_t)s * ld] = r;
    }
    csS = cs; psS = ps;
}

// phase A: every node of one tree depth.  Blocks [0, nwblocks): one warp per pattern with at least PAStep::T counts
// (padj_warp_pattern); the other blocks: one thread per pattern
#ifndef BLG_PADJ_MINB
#define BLG_PADJ_MINB 6
#endif
template <int CAP>
__global__ void __launch_bounds__(128, BLG_PADJ_MINB) padj_parent_kernel(const PAStep* __restrict__ steps, int nsteps, const NodeK* __restrict__ nk,
                                                          const double* __restrict__ pascal, int pdim, double* __restrict__ partial,
                                                          int nwblocks, int wcap) {
    extern __shared__ double blg_smem[];
    __shared__ PAStep st;
    __shared__ double red[4][6];
    const bool wpath = (int)blockIdx.x < nwblocks;
    const int vb = wpath ? (int)blockIdx.x : (int)blockIdx.x - nwblocks;
    {
        int lo = 0, hi = nsteps - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((wpath ? steps[mid].wblk0 : steps[mid].blk0) <= vb) lo = mid; else hi = mid - 1;
        }
        static_assert(sizeof(PAStep) % 4 == 0 && sizeof(PAStep) / 4 <= 128, "PAStep is staged by one pass of the block");
        if (threadIdx.x < sizeof(PAStep) / 4) ((int*)&st)[threadIdx.x] = ((const int*)&steps[lo])[threadIdx.x];
        __syncthreads();
    }
    const int p = wpath ? -1 : (vb - st.blk0) * 128 + (int)threadIdx.x;
    const size_t ld = (size_t)st.ld;
    double sc[2][3];                   // per child: {ψ̄' | Q̄, c̄ | Ē(W), ē}
#pragma unroll
    for (int i = 0; i < 2; ++i) { sc[i][0] = 0.0; sc[i][1] = 0.0; sc[i][2] = 0.0; }
    int X = (p >= 0 && p < st.U) ? (int)st.cnt[p] : 0;
    if (wpath) {
        const int item = (vb - st.wblk0) * 4 + (int)(threadIdx.x >> 5);
        double* wsm = blg_smem + (size_t)(threadIdx.x >> 5) * (8 * (size_t)wcap + 32 * 33);
        if (item < st.nbig) padj_warp_pattern(st, item, nk, pascal, pdim, wsm, wcap, wsm + 8 * (size_t)wcap, sc);
    }
    double Ab[CAP], b0[CAP], b1[CAP], Br[CAP], bb[CAP], lc[CAP];
    int cidx[2] = {p, p}, M[2] = {0, 0};
    int ex = 0;
    if (X > 0) {
        ex = st.scl[p];
        for (int i = 0; i < st.k; ++i) {
            const PAChild& c = st.child[i];
            if (c.pidx) cidx[i] = c.pidx[p];
            M[i] = c.xcnt ? (int)c.xcnt[p] : (int)c.cnt[cidx[i]];
            if (c.scl) ex -= c.scl[cidx[i]];
        }
        if (padj_is_big(st.k, st.T, st.msmin, X, M[0], M[1])) X = 0;      // (the warp path has it)
    }
    if (X > 0) {
        // Ā[n] = ô_u[n]·2^(-ex)·(1-E_k)^(-n); adjoint of the normalisation
        double Ekb = 0.0;
        for (int n = 0; n <= X; ++n) {
            const double o = st.adj[(size_t)n * ld + p];
            Ab[n] = padj_clamp(scalbn(o * st.ipow[n], -ex));
            Ekb = fma((double)n * o, st.ell[(size_t)n * ld + p], Ekb);
        }
        const double SAA = Ekb;            // Σ_n n Ā[n] A[n]
        Ekb *= nk[st.node].inv;
        // b_i = W_i ℓ̂_i
        for (int i = 0; i < st.k; ++i) {
            const PAChild& c = st.child[i];
            const int Mi = M[i];
            double* b = i == 0 ? b0 : b1;
            if (c.ell == nullptr) {
                for (int s = 0; s <= Mi; ++s) b[s] = c.Wt[(size_t)Mi * c.wdp + s];
            } else {
                for (int j = 0; j <= Mi; ++j) lc[j] = c.ell[(size_t)j * c.ld + cidx[i]];
                for (int s = 0; s <= Mi; ++s) {
                    double acc = 0.0;
                    for (int j = s; j <= Mi; ++j) acc = fma(c.Wt[(size_t)j * c.wdp + s], lc[j], acc);
                    b[s] = acc;
                }
            }
        }
        const double e0 = nk[st.child[0].id].e, e1 = st.k == 2 ? nk[st.child[1].id].e : 1.0;
        if (st.k == 2) {
            const double prod = e0 * e1;
            if (prod < 1.0) { sc[0][2] += Ekb * e1; sc[1][2] += Ekb * e0; }
        } else if (e0 < 1.0) {
            sc[0][2] += Ekb;
        }
        const bool two = st.k == 2 && st.child[0].dW == nullptr && st.child[1].dW == nullptr;
        if (two) {                                           // both b̄ from the rows of the child with fewer counts
            const int L = M[0] >= M[1] ? 0 : 1, S = 1 - L;
            const PAChild& cL = st.child[L];
            const PAChild& cS = st.child[S];
            const double* bL = L == 0 ? b0 : b1;
            const double* bS = L == 0 ? b1 : b0;
            const double eL = L == 0 ? e0 : e1, eS = L == 0 ? e1 : e0;
            double* outL = (cL.bbar != nullptr && M[L] > 0) ? cL.bbar + cL.pos[p] : nullptr;
            double* outS = (cS.bbar != nullptr && M[S] > 0) ? cS.bbar + cS.pos[p] : nullptr;
            double csL, psL, csS, psS;
            if (M[S] < 8) padj_merge2_reg<8>(Ab, bL, bS, M[L], M[S], eL > 1.0 ? 1.0 : eL, eS, outL, outS, ld, csL, psL, csS, psS);
            else padj_merge2_mem(Ab, bL, bS, M[L], M[S], eL > 1.0 ? 1.0 : eL, eS, outL, outS, ld, csL, psL, csS, psS, Br);
            const double ccL = nk[cL.id].cc, ccS = nk[cS.id].cc;
            sc[L][0] += psL / ccL; sc[L][1] += csL / ccL;
            sc[S][0] += psS / ccS; sc[S][1] += csS / ccS;
            if (eL < 1.0 && eL > 0.0) sc[L][2] += (SAA - csL) / eL;
            if (eS < 1.0 && eS > 0.0) sc[S][2] += (SAA - csS) / eS;
        }
        for (int i = 0; i < (two ? 0 : st.k); ++i) {
            const PAChild& c = st.child[i];
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
                    bb[t] = padj_clamp(acc);
                    eacc = fma(bi[t], accs, eacc);
                    for (int s = 0; s < Mo; ++s) Br[s] = fma(eo, Br[s], Br[s + 1]);
                    Br[Mo] = eo * Br[Mo];
                }
                if (ei < 1.0 && ei > 0.0) sc[i][2] += eacc / ei;
            }
            // ---- branch scalars ----
            if (c.dW == nullptr) {                         // ordinary branch: closed form in (ψ', c)
                double ps = 0.0, cs = 0.0;
                for (int n = 1; n <= Mi; ++n) {
                    const double w = (double)n * bb[n];
                    cs = fma(w, bi[n], cs);
                    if (n < Mi) ps = fma(w, bi[n + 1], ps);
                }
                const double cc = nk[c.id].cc;
                sc[i][0] += ps / cc;
                sc[i][1] += cs / cc;
            } else {                                       // branch below a WGD/WGT: dual tables
                for (int j = 0; j <= Mi; ++j) lc[j] = c.ell ? c.ell[(size_t)j * c.ld + cidx[i]] : (j == Mi ? 1.0 : 0.0);
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
                sc[i][0] += q;
                sc[i][1] += ee;
            }
            // ---- what goes down to the child ----
            if (c.bbar != nullptr && Mi > 0) {             // summed over the node's patterns by phase B
                const int pp = c.pos[p];
                for (int t = 0; t <= Mi; ++t) c.bbar[(size_t)t * ld + pp] = bb[t];
            } else if (c.adj != nullptr && Mi > 0) {       // one to one: ô_i = W_iᵀ b̄_i
                for (int j = 0; j <= Mi; ++j) {
                    double acc = 0.0;
                    for (int s = 0; s <= j; ++s) acc = fma(c.Wt[(size_t)j * c.wdp + s], bb[s], acc);
                    c.adj[(size_t)j * c.ld + cidx[i]] = padj_clamp(acc);
                }
            }
        }
    }
    // block reduction of the six scalars, fixed order; the blocks are combined by padj_reduce_kernel
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
        partial[(size_t)(wpath ? st.wpblk0 + vb - st.wblk0 : st.pblk0 + vb - st.blk0) * 6 + threadIdx.x] =
            (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// phase B: ô_c = W_cᵀ Σ_{node patterns p using c} b̄[p] for every internal child of one tree depth, in one launch.
// Blocks [0, nhblocks): one BLOCK per child pattern that very many node patterns use; blocks [nhblocks, nhblocks + nmblocks):
// one WARP per child pattern with many node patterns (lanes over them) or, at depths with few patterns, many counts (lanes
// over the counts); the other blocks: one THREAD per child pattern.  All sums in a fixed order.
template <int CAP>
__global__ void __launch_bounds__(128) padj_child_kernel(const PBStep* __restrict__ steps, int nsteps, int nhblocks, int nmblocks, int cap) {
    extern __shared__ double blg_smem[];
    __shared__ double red[4];
    const int path = (int)blockIdx.x < nhblocks ? 0 : (int)blockIdx.x < nhblocks + nmblocks ? 1 : 2;
    const int vb = (int)blockIdx.x - (path == 0 ? 0 : path == 1 ? nhblocks : nhblocks + nmblocks);
    int lo = 0, hi = nsteps - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const int b0 = path == 0 ? steps[mid].hblk0 : path == 1 ? steps[mid].mblk0 : steps[mid].blk0;
        if (b0 <= vb) lo = mid; else hi = mid - 1;
    }
    const PBStep& st = steps[lo];
    const size_t ldp = (size_t)st.ldp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (path == 2) {
        const int c = (vb - st.blk0) * 128 + (int)threadIdx.x;
        if (c >= st.Uc) return;
        const int M = st.cnt[c];
        if (M == 0) return;                                    // a constant column: nothing depends on it
        const int k0 = st.off[c], k1 = st.off[c + 1];
        if (k1 - k0 > BLG_PB_HIT || M >= st.midM) return;      // (the other paths)
        double bs[CAP];
        for (int t = 0; t <= M; ++t) bs[t] = 0.0;
        for (int k = k0; k < k1; ++k)
            for (int t = 0; t <= M; ++t) bs[t] += st.bbar[(size_t)t * ldp + k];
        for (int j = 0; j <= M; ++j) {
            double acc = 0.0;
            for (int s = 0; s <= j; ++s) acc = fma(st.Wt[(size_t)j * st.wdp + s], bs[s], acc);
            st.adj[(size_t)j * st.ldc + c] = padj_clamp(acc);
        }
    } else if (path == 1) {
        const int item = (vb - st.mblk0) * 4 + warp;
        if (item >= st.nmid) return;
        double* bs = blg_smem + (size_t)warp * cap;
        const int c = st.mid[item];
        const int M = st.cnt[c];
        const int k0 = st.off[c], k1 = st.off[c + 1];
        if (k1 - k0 <= BLG_PB_HIT && M < st.midM) return;      // (listed for a lower bound: the thread path has it)
        if (k1 - k0 >= 32) {
            for (int t = 0; t <= M; ++t) {
                const double* row = st.bbar + (size_t)t * ldp;
                double v = 0.0;
                for (int k = k0 + lane; k < k1; k += 32) v += row[k];
                v = padj_warp_sum(v);
                if (lane == 0) bs[t] = v;
            }
        } else {
            for (int t = lane; t <= M; t += 32) {
                const double* row = st.bbar + (size_t)t * ldp;
                double v = 0.0;
                for (int k = k0; k < k1; ++k) v += row[k];
                bs[t] = v;
            }
        }
        __syncwarp();
        for (int j = lane; j <= M; j += 32) {
            double acc = 0.0;
            for (int s = 0; s <= j; ++s) acc = fma(st.Wt[(size_t)j * st.wdp + s], bs[s], acc);
            st.adj[(size_t)j * st.ldc + c] = padj_clamp(acc);
        }
    } else {
        double* bs = blg_smem;
        const int c = st.hit[vb - st.hblk0];
        const int M = st.cnt[c];
        const int k0 = st.off[c], k1 = st.off[c + 1];
        for (int t = 0; t <= M; ++t) {
            const double* row = st.bbar + (size_t)t * ldp;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;             // (four loads in flight per thread)
            int k = k0 + (int)threadIdx.x;
            for (; k + 384 < k1; k += 512) { v0 += row[k]; v1 += row[k + 128]; v2 += row[k + 256]; v3 += row[k + 384]; }
            for (; k < k1; k += 128) v0 += row[k];
            double v = (v0 + v1) + (v2 + v3);
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp] = v;
            __syncthreads();
            if (threadIdx.x == 0) bs[t] = (red[0] + red[1]) + (red[2] + red[3]);
            __syncthreads();
        }
        for (int j = threadIdx.x; j <= M; j += 128) {
            double acc = 0.0;
            for (int s = 0; s <= j; ++s) acc = fma(st.Wt[(size_t)j * st.wdp + s], bs[s], acc);
            st.adj[(size_t)j * st.ldc + c] = padj_clamp(acc);
        }
    }
}

// sadj[child id][j] = Σ_b partial[b][3i + j] over the blocks of the child's parent step, one block per step
__global__ void __launch_bounds__(256) padj_reduce_kernel(const double* __restrict__ partial, const PAStep* __restrict__ steps,
                                                          double* __restrict__ sadj) {
    __shared__ double red[256];
    const PAStep& st = steps[blockIdx.x];
    for (int slot = 0; slot < 3 * st.k; ++slot) {
        double s = 0.0;
        for (int i = threadIdx.x; i < st.nblk; i += 256) s += partial[(size_t)(st.pblk0 + i) * 6 + slot];
        for (int i = threadIdx.x; i < st.nwblk; i += 256) s += partial[(size_t)(st.wpblk0 + i) * 6 + slot];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) sadj[(size_t)st.child[slot < 3 ? 0 : 1].id * 4 + slot % 3] = red[0];
        __syncthreads();
    }
}
