// general.cuh — the GENERAL pruning kernel: one warp per family, any count bound up to BLG_MAXCOUNT, any parameters.
// Included by beluga_b200.cu.  Same recursion as sweep_kernel (csuros_miklos!, src/model.jl:402-463), evaluated in a
// form whose every factor is a probability, so that nothing over- or underflows before it is negligible:
//
//   b̂_c[s]   = b_c[s]/(1-e_c)^s,  b_c = W_c ℓ_c                      (model.jl:425-426; row s of W_c/(1-e_c)^s sums to 1)
//        ordinary branch: Horner on wstar!'s recurrence (node.jl:181-192), r[n] ← ψ'·r[n] + ĉ·r[n-1], ĉ = c/(1-e_c)
//        below a WGD:     w*(n|j) = [z^j](a·z + b·z²)^n (node.jl:194-203)  ⇒  v_n[i] = â·v_{n-1}[i+1] + b̂·v_{n-1}[i+2],
//                         b̂[n] = v_n[0],  â = a/(1-e_c), b̂ = b/(1-e_c), â + b̂ = 1
//        below a WGT:     the (literal, quirky) table of node.jl:205-218
//   B̂[0,s]   = b̂[s];   B̂[t,s] = (1-e)·B̂[t-1,s+1] + e·B̂[t-1,s]        (model.jl:428-431 divided by (1-e)^s: a convex combination)
//   Â[n]     = Σ_{t+s=n} Binom(n; p_s)(s) · Â_t[t] · B̂_s[t,s]          (model.jl:438-455)
//        with E_t, E_s the (clamped) extinction probabilities of the two groups being merged,
//        p_s = E_t(1-E_s)/(1-E_tE_s), p_t = (1-E_t)/(1-E_tE_s), p_s + p_t = 1.
//        The reference's A[i,·] (its (1-_ϵ)^(-n) normalisation applied after every child) IS Â; the formula is symmetric in the
//        two groups, so the side that is spread over the lanes is chosen for speed.
//   the binomial weight is folded into the row:  B̃[t,s] = Binom(t+s; p_s)(s)·B̂[t,s],
//        B̃[t,s] = (p_t/t)·[ (1-e)/p_s·(s+1)·B̃[t-1,s+1] + e·(t+s)·B̃[t-1,s] ],   Â[n] = Σ_{t+s=n} Â_t[t]·B̃[t,s]
//        — five fp64 operations per (t,s), no table, no factorial, no exp/log.
//
// Work decomposition: a WARP owns ONE family for the whole node list (dynamic hand-out, heaviest first); the lanes hold
// contiguous chunks of the count index in registers (Horner tiles, the B̃ row and the sliding anti-diagonal
// accumulators), neighbours exchange one value per stage by shuffle.  Columns live in per-warp shared memory; the stored
// copy is normalised to a power-of-two exponent per (family, node) like everywhere else in this library.
// Storage is addressed as base[off(f) + n·stride]: stride = ld, off = f for the node-major SoA slabs of the fast shard;
// stride = 1, off = per-family prefix offsets for the ragged slabs of the heavy shard (coalesced along n).
#pragma once

#define BLG_GH 900          // headroom exponent of the weighted row: B̃ is carried as B̃·2^900 (p_s^s may be far below 2^-1074)

struct GChild {
    const double* ell;      // nullptr: leaf (one-hot at its count)
    const int* scl;
    const uint16_t* cnt;
    const uint32_t* off;    // ragged layout: per-family offset into ell (nullptr: SoA with stride ld)
    const int* pidx;        // pattern layout: item of the step → column of the child (nullptr: the same index)
    const double* Wt;       // transposed w* table (branch below a WGT only)
    int wdp;
    int node;               // node index of the child (NodeK)
    int mode;               // 0 ordinary branch, 1 below a WGD (polynomial form), 2 table
    int ld;                 // SoA stride of the child's columns
    // heavy family set: where the family's entry at this child is below the heavy bound, the column is the node pattern
    // fpat[f] of the fast set (SoA pattern columns pell/pscl, stride pld), not a ragged per-family column
    const int* fpat;
    const double* pell;
    const int* pscl;
    int pld, pad_;
};
struct GStep {
    int k, first, resident, node;
    int item0, nchain, is_chain, pad_;   // item mode (pattern layout): first work item of this step in the launch; single-child steps
                                         // that follow on the same pattern index (their columns stay in shared memory)
    double* out;
    int* out_scl;
    const uint16_t* cnt;
    const uint32_t* out_off;
    size_t out_ld;          // SoA stride of the output columns
    const int* fpat;        // heavy family set: fpat[f] >= 0 — the entry is below the heavy bound, the fast set's pattern path
                            // computes it: skip the step for this family
};
#define BLG_G_MAXSTEPS 128
#define BLG_G_MAXCHILD 200
struct GList {
    int nsteps;
    int pad_[3];
    GStep step[BLG_G_MAXSTEPS];
    GChild child[BLG_G_MAXCHILD];
};
static_assert(sizeof(GList) <= 32000, "kernel parameter space");

__device__ __forceinline__ double g_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// power-of-two exponent of a positive finite maximum (0 otherwise), and the exact scale 2^-ex
__device__ __forceinline__ int g_exponent(double mx) {
    int ex = 0;
    if (mx > 0.0) {
        const int be = (__double2hiint(mx) >> 20) & 0x7ff;
        if (be != 0 && be != 0x7ff) ex = be - 1023;
    }
    return ex;
}
__device__ __forceinline__ double g_pow2(int ex) {       // 2^ex for |ex| <= 1022
    return __hiloint2double((1023 + ex) << 20, 0);
}
// rescale col[0..M] by the power of two that brings its maximum into [1,2); returns the exponent
__device__ __forceinline__ int g_normalise(double* col, int M, int lane) {
    double mx = 0.0;
    for (int n = lane; n <= M; n += 32) mx = fmax(mx, col[n]);
    mx = g_warp_max(mx);
    const int ex = g_exponent(mx);
    if (ex != 0) {
        // two steps so that |ex| up to 2044 stays exact and inside the range of g_pow2
        const int e1 = ex / 2, e2 = ex - e1;
        const double s1 = g_pow2(-e1), s2 = g_pow2(-e2);
        for (int n = lane; n <= M; n += 32) col[n] = (col[n] * s1) * s2;
    }
    __syncwarp();
    return ex;
}

// in-place Horner contraction over an ordinary branch: col[j] = ℓ[j] → col[n] = b̂[n], n = 0..M (b̂[0] = ℓ[0]).
// Lane L of a tile holds outputs n = base + 1 + L·C + k; a tile hands the old value of its top element to the tile above
// through the column itself (slot base+32C+j, already consumed), exactly as sweep_contract_horner does per thread.
template <int C>
__device__ __noinline__ void g_horner(double* col, int M, double psi, double chat, int lane) {
    constexpr int W = 32 * C;
#pragma unroll 1
    for (int base = 0; base < M; base += W) {
        double r[C];
#pragma unroll
        for (int k = 0; k < C; ++k) r[k] = 0.0;
        const int lim = M - base - W;
#pragma unroll 1
        for (int j = M - base; j >= 1; --j) {
            const double h = col[base + j];
            const double top = r[C - 1];
            const double up = __shfl_up_sync(0xffffffffu, top, 1);
            const double in = (lane == 0) ? h : up;
#pragma unroll
            for (int k = C - 1; k >= 1; --k) r[k] = fma(psi, r[k], chat * r[k - 1]);
            r[0] = fma(psi, r[0], chat * in);
            if (j <= lim && lane == 31) col[base + W + j] = top;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int n = base + 1 + lane * C + k;
            if (n <= M) col[n] = r[k];
        }
        __syncwarp();
    }
}
__device__ __forceinline__ void g_horner_any(double* col, int M, double psi, double chat, int lane) {
    if (M <= 32) g_horner<1>(col, M, psi, chat, lane);
    else if (M <= 64) g_horner<2>(col, M, psi, chat, lane);
    else if (M <= 128) g_horner<4>(col, M, psi, chat, lane);
    else g_horner<8>(col, M, psi, chat, lane);
}

// contraction below a WGD in polynomial form: v (= ℓ, in place in `v`) is advanced M times, out[n] = v_n[0].
__device__ __noinline__ void g_wgd_poly_iter(double* v, double* out, int M, double ahat, double bhat, int lane) {
    if (lane == 0) out[0] = v[0];
    // v[M+1], v[M+2] are read as zero
    if (lane == 0) { v[M + 1] = 0.0; v[M + 2] = 0.0; }
    __syncwarp();
#pragma unroll 1
    for (int n = 1; n <= M; ++n) {
        const int top = M - n;                      // v_n[i] is non-zero for i <= M - n only
#pragma unroll 1
        for (int i0 = 0; i0 <= top; i0 += 32) {
            const int i = i0 + lane;
            double nv = 0.0;
            if (i <= top) nv = fma(ahat, v[i + 1], bhat * v[i + 2]);
            __syncwarp();
            if (i <= top) v[i] = nv;
            __syncwarp();
        }
        if (lane == 0) { out[n] = v[0]; v[top + 1] = 0.0; }      // (v_n[M-n+1] = 0 for the next stage)
        __syncwarp();
    }
}

// The same contraction in closed form.  With S the shift (Sℓ)[i] = ℓ[i+1] the stages above are v_n = (â S + b̂ S²)ⁿ ℓ, so
//   out[n] = v_n[0] = Σ_k C(n,k) âⁿ⁻ᵏ b̂ᵏ ℓ[n+k],   k = 0 … min(n, M−n):
// every output is an independent sum — lane L takes n = n0 + L, all lanes walk k together (the reciprocal 1/(k+1) is the
// same for the warp), no shared-memory stores and no warp barriers between the M stages.  âⁿ is factored out and applied
// at the end through its exponent (it underflows for n in the hundreds while the sum does not); the running term
// C(n,k)(b̂/â)ᵏ and the sum are rescaled together by a power of two whenever they grow large.
__device__ __noinline__ void g_wgd_poly(double* v, double* out, int M, double ahat, double bhat, int lane) {
    if (!(ahat > 1e-12) || !(bhat >= 0.0) || bhat > 1e3 * ahat) { g_wgd_poly_iter(v, out, M, ahat, bhat, lane); return; }
    const double ba = bhat / ahat;
    const double l2a = log2(ahat);
#pragma unroll 1
    for (int n0 = 0; n0 <= M; n0 += 32) {
        const int n = n0 + lane;
        const int kmax = n <= M ? (n < M - n ? n : M - n) : -1;
        const int kw = __reduce_max_sync(0xffffffffu, kmax);
        const double nd = (double)n;
        double t = 1.0, acc = 0.0;
        int e = 0;
#pragma unroll 1
        for (int k0 = 0; k0 <= kw; k0 += 16) {
            const int k1 = k0 + 15 < kw ? k0 + 15 : kw;
#pragma unroll 4
            for (int k = k0; k <= k1; ++k) {
                if (k <= kmax) {
                    acc = fma(t, v[n + k], acc);
                    t *= (nd - (double)k) * (c_inv[k + 1] * ba);
                }
            }
            if (t > 0x1p+300 || acc > 0x1p+300) { t *= 0x1p-300; acc *= 0x1p-300; e += 300; }
        }
        if (n <= M) {
            const double x = nd * l2a, xi = floor(x);
            out[n] = scalbn(acc * exp2(x - xi), (int)xi + e);
        }
    }
    __syncwarp();
}

// table contraction (branch below a WGT; literal table of node.jl:205-218): out[s] = (1-e)^(-s)·Σ_j Wt[j·wdp+s]·in[j]
__device__ __noinline__ void g_table(const double* in, double* out, int M, const double* __restrict__ Wt, int wdp, double ome, int lane) {
    const double iome = 1.0 / ome;
#pragma unroll 1
    for (int s = lane; s <= M; s += 32) {
        double acc = 0.0;
        const int jhi = min(M, 3 * s);
#pragma unroll 1
        for (int j = s; j <= jhi; ++j) acc = fma(Wt[(size_t)j * wdp + s], in[j], acc);
        out[s] = acc * d_powi(iome, s);
    }
    __syncwarp();
}

// register merge: the "s" side (column S[0..Ms], extinction Es) is spread over the lanes, C consecutive s per lane; the "t"
// side (column A[0..Mt]) is walked.  O (may alias S) receives 2^BLG_GH·Â[n], n = 0..Mt+Ms.
template <int C>
__device__ __noinline__ void g_merge_reg(const double* A, int Mt, const double* S, int Ms, double* O, double Es, double ps, double pt, int lane) {
    double B[C], acc[C], c1[C], es[C];
    const double u = (1.0 - Es) / ps;
    {
        double pw = d_powi(ps, lane * C) * g_pow2(BLG_GH / 2);
        const double h2 = g_pow2(BLG_GH - BLG_GH / 2);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int s = lane * C + k;
            B[k] = (s <= Ms) ? (S[s] * pw) * h2 : 0.0;
            pw *= ps;
            c1[k] = u * (double)(s + 1);
            es[k] = Es * (double)s;
        }
    }
    __syncwarp();                                   // every lane has read S before O (which may alias it) is written
    {
        const double a0 = A[0];
#pragma unroll
        for (int k = 0; k < C; ++k) acc[k] = a0 * B[k];
    }
#pragma unroll 1
    for (int t = 0; t < Mt; ++t) {
        if (lane == 0) O[t] = acc[0];               // output t is complete
        double nbB = __shfl_down_sync(0xffffffffu, B[0], 1);
        double nbA = __shfl_down_sync(0xffffffffu, acc[0], 1);
        if (lane == 31) { nbB = 0.0; nbA = 0.0; }
        const double g = pt * c_inv[t + 1];
        const double et = Es * (double)(t + 1);
        const double a = A[t + 1];
#pragma unroll
        for (int k = 0; k < C - 1; ++k) {
            B[k] = g * fma(c1[k], B[k + 1], (es[k] + et) * B[k]);
            acc[k] = fma(a, B[k], acc[k + 1]);
        }
        B[C - 1] = g * fma(c1[C - 1], nbB, (es[C - 1] + et) * B[C - 1]);
        acc[C - 1] = fma(a, B[C - 1], nbA);
    }
    __syncwarp();
    // the accumulators now hold outputs n = Mt + s
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int s = lane * C + k;
        if (s <= Ms) O[Mt + s] = acc[k];
    }
    __syncwarp();
}
// shared-memory merge for an "s" side too wide for the register tiles: the weighted row lives in S (in place),
// O (a third buffer, zero-filled here) accumulates 2^BLG_GH·Â[n].
__device__ __noinline__ void g_merge_smem(const double* A, int Mt, double* S, int Ms, double* O, double Es, double ps, double pt, int lane) {
    const double u = (1.0 - Es) / ps;
    {
        double pw = d_powi(ps, lane) * g_pow2(BLG_GH / 2);
        const double step = d_powi(ps, 32), h2 = g_pow2(BLG_GH - BLG_GH / 2);
        for (int s = lane; s <= Ms; s += 32) { S[s] = (S[s] * pw) * h2; pw *= step; }
        for (int n = lane; n <= Mt + Ms; n += 32) O[n] = 0.0;
        if (lane == 0) S[Ms + 1] = 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int t = 0; t <= Mt; ++t) {
        const double a = A[t];
        for (int s = lane; s <= Ms; s += 32) O[t + s] = fma(a, S[s], O[t + s]);
        __syncwarp();
        if (t == Mt) break;
        const double g = pt * c_inv[t + 1];
        const double td = (double)(t + 1);
#pragma unroll 1
        for (int s0 = 0; s0 <= Ms; s0 += 32) {
            const int s = s0 + lane;
            double nv = 0.0;
            if (s <= Ms) nv = g * fma(u * (double)(s + 1), S[s + 1], (Es * (td + (double)s)) * S[s]);
            __syncwarp();
            if (s <= Ms) S[s] = nv;
            __syncwarp();
        }
    }
}
// no extinction on either side (E_t = E_s = 0): every lineage survives in both groups, Â[n] = Â_t[n]·Â_s[n]
__device__ __forceinline__ void g_merge_product(const double* A, int Mt, const double* S, int Ms, double* O, int lane) {
    const double h2 = g_pow2(BLG_GH / 2), h1 = g_pow2(BLG_GH - BLG_GH / 2);
    for (int n = lane; n <= Mt + Ms; n += 32) {
        const double v = (n <= Mt && n <= Ms) ? A[n] * S[n] : 0.0;
        O[n] = (v * h1) * h2;
    }
    __syncwarp();
}

__device__ __forceinline__ int g_cpad(int M) {           // lanes × C register slots needed for s = 0..M
    const int c = (M + 32) / 32;
    return c <= 1 ? 1 : c <= 2 ? 2 : c <= 4 ? 4 : c <= 8 ? 8 : c <= 16 ? 16 : 0;
}

// flags[0]: may the table-free sweep kernel be used with the current parameters?  `when`: 0 = always run, 1 = run only if
// flags[0] == 0 (the fallback of the fast shard).  ran[1] counts the launches of this kernel that did the work.
// Families f0..f1-1 are handed out one per warp by an atomic counter (heaviest first in the heavy shard).
// Shared memory: nbuf (2 or 3) buffers of `cap` doubles per warp.
template <int CMAX>
__global__ void __launch_bounds__(128, (CMAX <= 4) ? 3 : 2)
gsweep_kernel(const __grid_constant__ GList L, const NodeK* __restrict__ nk, const int* __restrict__ flags, int when,
              unsigned int* __restrict__ ran, int f0, int f1, int count_run, int cap, int nbuf, unsigned int* __restrict__ counter,
              int item_mode) {
    extern __shared__ double gsm[];
    if (when == 1 && flags[0] != 0) return;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (count_run && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ran[1], 1u);
    double* buf0 = gsm + (size_t)warp * nbuf * cap;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 1
    for (;;) {
        int f = 0;
        if (lane == 0) f = f0 + (int)atomicAdd(counter, 1u);
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= f1) break;
        // three roles rotate over the buffers: pA = the column being built / kept, pB and pC scratch
        double* pA = buf0;
        double* pB = buf0 + cap;
        double* pC = (nbuf >= 3) ? buf0 + 2 * (size_t)cap : nullptr;
        int prevNode = -1, prevX = 0, prevScl = 0;   // column left in pA by the previous step (normalised, exponent prevScl)
        int s_lo = 0, s_hi = L.nsteps - 1;
        if (item_mode) {
            // item → (step, pattern): the last step whose first item is not beyond it, then back to the head of its chain
            int lo = 0, hi = L.nsteps - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (L.step[mid].item0 <= f) lo = mid; else hi = mid - 1;
            }
            while (L.step[lo].is_chain) --lo;
            f -= L.step[lo].item0;
            s_lo = lo;
            s_hi = lo + L.step[lo].nchain;
        }
#pragma unroll 1
        for (int si = s_lo; si <= s_hi; ++si) {
            const GStep& st = L.step[si];
            const int k = st.k;
            const GChild* ch = &L.child[st.first];
            if (st.fpat != nullptr && st.fpat[f] >= 0) { prevNode = -1; continue; }   // light entry: the pattern path has it
            const int X = st.cnt[f];
            const double ip0 = nk[st.node].ip0;                    // 1, or NaN where the reference's column is NaN
            double* outp = st.out + (st.out_off ? (size_t)st.out_off[f] : (size_t)f);
            const size_t ostride = st.out_off ? 1 : st.out_ld;
            if (X == 0) {
                // no gene below the node: ℓ = [1] (n = 0 lineages leave no descendants with probability 1), or the NaN
                if (lane == 0) { outp[0] = ip0; st.out_scl[f] = 0; pA[0] = ip0; }
                prevNode = st.node; prevX = 0; prevScl = 0;
                __syncwarp();
                continue;
            }
            int sclsum = 0;
            int S = 0;                    // Σ counts merged so far; pA[0..S] holds the normalised Â of the group
            double Eg = 1.0;              // clamped extinction probability of the group
#pragma unroll 1
            for (int i = 0; i < k; ++i) {
                const GChild& c = ch[i];
                const NodeK ck = nk[c.node];
                const int cf = c.pidx ? c.pidx[f] : f;            // the child's column of this item
                const int M = c.cnt[cf];
                // ---- the child's column → b̂ (normalised), into `dst` -------------------------------------------
                double* dst = (i == 0) ? pA : pB;
                const bool resident = (i == 0) && st.resident && prevNode == c.node;
                if (c.ell == nullptr) {
                    for (int n = lane; n <= M; n += 32) dst[n] = (n == M) ? 1.0 : 0.0;
                } else if (resident) {
                    sclsum += prevScl;
                } else {
                    const int pc = c.fpat ? c.fpat[f] : -1;
                    if (pc >= 0) {                                 // the child's entry is light: its node pattern's column
                        const double* src = c.pell + pc;
                        for (int n = lane; n <= M; n += 32) dst[n] = src[(size_t)n * c.pld];
                        sclsum += c.pscl[pc];
                    } else {
                        const double* src = c.ell + (c.off ? (size_t)c.off[cf] : (size_t)cf);
                        const size_t sstride = c.off ? 1 : (size_t)c.ld;
                        for (int n = lane; n <= M; n += 32) dst[n] = src[(size_t)n * sstride];
                        sclsum += c.scl[cf];
                    }
                }
                __syncwarp();
                double* bcol = dst;
                if (c.mode == 0) {
                    g_horner_any(dst, M, ck.psi, ck.chat, lane);
                } else if (c.mode == 1) {
                    double* o = (i == 0) ? pB : pC;            // (k == 1 always for these branches: i == 0)
                    g_wgd_poly(dst, o, M, ck.chat, ck.bhat, lane);
                    bcol = o;
                } else {
                    double* o = (i == 0) ? pB : pC;
                    g_table(dst, o, M, c.Wt, c.wdp, 1.0 - ck.e, lane);
                    bcol = o;
                }
                if (bcol != dst) {                             // result landed in the other buffer: swap roles
                    if (i == 0) { double* t_ = pA; pA = pB; pB = t_; }
                }
                sclsum += g_normalise((i == 0) ? pA : bcol, M, lane);
                const double ei = ck.e;
                if (i == 0) { S = M; Eg = fmin(ei, 1.0); continue; }
                // ---- merge the group (pA[0..S], Eg) with this child (pB[0..M], ei) ------------------------------------
                const double Es_c = fmin(ei, 1.0);
                const double E2 = fmin(Eg * Es_c, 1.0);
                const double den = 1.0 / (1.0 - E2);
                // orientation 0: s side = child; orientation 1: s side = group
                const double ps0 = Eg * (1.0 - Es_c) * den, pt0 = (1.0 - Eg) * den;
                const double ps1 = Es_c * (1.0 - Eg) * den, pt1 = (1.0 - Es_c) * den;
                const bool ok0 = ps0 > 1e-280 && ps0 <= 1.0, ok1 = ps1 > 1e-280 && ps1 <= 1.0;
                const int c0 = g_cpad(M), c1 = g_cpad(S);
                // cost of a register merge: (Mt+1)·(5·C + 6) issue slots per lane
                const long long cost0 = (c0 && c0 <= CMAX) ? (long long)(S + 1) * (5 * c0 + 6) : (1LL << 60);
                const long long cost1 = (c1 && c1 <= CMAX) ? (long long)(M + 1) * (5 * c1 + 6) : (1LL << 60);
                int orient = -1;                               // -1: no register orientation
                if (ok0 && ok1) orient = (cost0 <= cost1) ? (cost0 < (1LL << 60) ? 0 : -1) : 1;
                else if (ok0) orient = cost0 < (1LL << 60) ? 0 : -1;
                else if (ok1) orient = cost1 < (1LL << 60) ? 1 : -1;
                double* O;
                if (!(E2 < 1.0) || !(Eg >= 0.0) || !(Es_c >= 0.0)) {
                    // a clamped prefix of 1 (or an invalid ε): the reference's column is NaN here (the node is poisoned)
                    O = pB;
                    for (int n = lane; n <= S + M; n += 32) O[n] = NaN;
                    __syncwarp();
                } else if (!ok0 && !ok1) {
                    O = pB;
                    g_merge_product(pA, S, pB, M, O, lane);
                } else if (orient >= 0) {
                    const double* A = orient == 0 ? pA : pB;
                    double* Sd = orient == 0 ? pB : pA;
                    const int Mt = orient == 0 ? S : M, Ms = orient == 0 ? M : S;
                    const double Es = orient == 0 ? Es_c : Eg, ps = orient == 0 ? ps0 : ps1, pt = orient == 0 ? pt0 : pt1;
                    const int cc = orient == 0 ? c0 : c1;
                    O = Sd;
                    if (cc == 1) g_merge_reg<1>(A, Mt, Sd, Ms, O, Es, ps, pt, lane);
                    else if (cc == 2) g_merge_reg<2>(A, Mt, Sd, Ms, O, Es, ps, pt, lane);
                    else if (cc == 4) g_merge_reg<4>(A, Mt, Sd, Ms, O, Es, ps, pt, lane);
                    else if (CMAX >= 8 && cc == 8) g_merge_reg<(CMAX >= 8 ? 8 : 1)>(A, Mt, Sd, Ms, O, Es, ps, pt, lane);
                    else g_merge_reg<(CMAX >= 16 ? 16 : 1)>(A, Mt, Sd, Ms, O, Es, ps, pt, lane);
                } else {
                    // too wide for the register tiles on every usable side: shared-memory row, third buffer
                    const bool o0 = ok0 && (!ok1 || M <= S);
                    const double* A = o0 ? pA : pB;
                    double* Sd = o0 ? pB : pA;
                    const int Mt = o0 ? S : M, Ms = o0 ? M : S;
                    O = pC;
                    g_merge_smem(A, Mt, Sd, Ms, O, o0 ? Es_c : Eg, o0 ? ps0 : ps1, o0 ? pt0 : pt1, lane);
                }
                // the merged column becomes the group: rotate the buffers so that it sits in pA
                if (O == pB) { double* t_ = pA; pA = pB; pB = t_; }
                else if (O == pC) { double* t_ = pA; pA = pC; pC = t_; }
                S += M;
                Eg = E2;
                sclsum += g_normalise(pA, S, lane) - BLG_GH;
            }
            // ---- ℓ_u[n] = Â_k[n], n = 0..X; stored normalised (pA already is), NaN-poisoned where the reference's is ----
            for (int n = lane; n <= X; n += 32) {
                const double v = (n <= S ? pA[n] : 0.0) * ip0;
                outp[(size_t)n * ostride] = v;
                if (n > S || ip0 != 1.0) pA[n] = v;
            }
            if (lane == 0) st.out_scl[f] = sclsum;
            prevNode = st.node; prevX = X; prevScl = sclsum;
            __syncwarp();
        }
    }
    (void)NaN;
}

// root integration for the ragged heavy shard, one warp per family: log P = log Σ_n ℓ[n]·exp(lw[n]) (integrate_root,
// model.jl:465-475) minus the conditioning constant (condition_oib, model.jl:492-501).  The weights come in LOG form: with
// root counts in the hundreds exp(lw[n]) underflows long before the product does, so the sum is taken relative to the
// largest term (found from the exponent bits of ℓ[n], no log in the loop).  Weighted block sums; the last block of ALL root
// launches of the call (ticket counts to nblocks_total) adds the partial sums in a fixed order.
__global__ void __launch_bounds__(256) root_ragged_kernel(const double* __restrict__ ell, const int* __restrict__ scl,
                                                          const uint16_t* __restrict__ cnt, const uint32_t* __restrict__ off,
                                                          const double* __restrict__ rootlw,
                                                          const double* __restrict__ rootw0, const double* __restrict__ weight,
                                                          double* __restrict__ fam_ll, double* __restrict__ partial, int part0,
                                                          int N, unsigned int* __restrict__ ticket, unsigned int nblocks_total,
                                                          const int* __restrict__ flags, double* __restrict__ marker, double* __restrict__ out) {
    __shared__ double red[8];
    __shared__ bool last_block;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 8 + warp;
    double contrib = 0.0;
    if (f < N) {
        const int X = cnt[f];
        const double* col = ell + off[f];
        const double LN2 = 0.6931471805599453;
        // pass 1: the largest log-term, to within one binary order of magnitude
        double top = -INFINITY;
        bool bad = false;
        for (int n = 1 + lane; n <= X; n += 32) {
            const double v = col[n];
            if (v != v) bad = true;
            if (v > 0.0) top = fmax(top, (double)g_exponent(v) * LN2 + rootlw[n]);
        }
        top = g_warp_max(top);
        bad = __any_sync(0xffffffffu, bad);
        double p = 0.0;
        if (top > -INFINITY) {
            for (int n = 1 + lane; n <= X; n += 32) p = fma(col[n], exp(rootlw[n] - top), p);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        double l = (top > -INFINITY && !bad) ? log(p) + top + (double)scl[f] * LN2 - rootw0[0] : -INFINITY;
        if (!isfinite(l)) l = -INFINITY;                                        // model.jl:366
        if (lane == 0) fam_ll[f] = l;
        contrib = weight[f] * l;
    }
    if (lane == 0) red[warp] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
        partial[part0 + blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
        __threadfence();
        last_block = atomicAdd(ticket, 1u) == nblocks_total - 1;
    }
    __syncthreads();
    if (last_block) {
        __shared__ double red2[256];
        __threadfence();
        const volatile double* vp = partial;
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)nblocks_total; i += 256) s += vp[i];
        red2[threadIdx.x] = s;
        __syncthreads();
        for (int k = 128; k > 0; k >>= 1) {
            if ((int)threadIdx.x < k) red2[threadIdx.x] += red2[threadIdx.x + k];
            __syncthreads();
        }
        if (threadIdx.x == 0) { out[0] = red2[0]; marker[0] = flags[0] ? 0.0 : 1.0; *ticket = 0u; }
    }
}
