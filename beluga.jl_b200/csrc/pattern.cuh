// pattern.cuh — the pruning step over NODE PATTERNS.  Included by beluga_b200.cu (uses the sweep kernel's helpers).
//
// The column ℓ_u of a family depends only on the family's counts in the subtree below u.  The reference compresses
// identical families once, over the whole profile (site patterns, src/model.jl:40-49); here the same idea is applied at
// EVERY node: the families of a handle are deduplicated per node by the leaf counts below it, and csuros_miklos!
// (src/model.jl:402-463) is evaluated once per distinct sub-pattern.  A cherry of a 100-taxon tree sees a few dozen
// distinct count pairs however many families there are; only the nodes next to the root see (nearly) every family.
// Every family still owns exactly the value the per-family recursion gives it — a pattern's column is computed from the
// columns of its children's patterns with the same arithmetic — so the results are those of the reference.
//
// Storage per node u: U_u pattern columns, node-major SoA  ell[n][p], p < U_u (stride ld_u), one power-of-two exponent
// per pattern, the pattern's count cnt_u[p], and per child j the map pidx_j[p] → pattern of the child (its count, for the
// family → pattern map fam2pat only the root needs on the device).  Patterns of a node are sorted by their counts, so the
// 32 patterns of a warp run (nearly) identical loop shapes — an order the per-family layout, where one family order must
// serve all nodes, cannot have.
//
// One launch handles the nodes of one LEVEL of the tree (all of a node's children are in earlier launches); a warp takes
// one tile of 32 patterns of one node and runs the sweep kernel's step on it (same table-free contraction and merge).
#pragma once

struct PChild {
    const double* ell;     // child's pattern columns (nullptr: leaf — one-hot at its count)
    const int* scl;
    const uint16_t* cnt;   // child's count per pattern
    const int* pidx;       // node pattern → child pattern (nullptr: identity, the single child of a WGD-type node)
    const uint16_t* xcnt;  // the child's count per NODE pattern (nullptr: read cnt[child pattern])
    const double* Wt;      // transposed w* of the child's branch (leaf gather, WGD/WGT-after contraction)
    int wdp, table;
    int node, ld;          // node index of the child (NodeK); stride of its pattern columns
};
struct PStep {
    int k, first, node, U;
    int tile0, ld;         // first tile of this step in the launch; stride of the node's pattern columns
    int p0;                // first pattern the tiles cover (the patterns before it run on the general kernel, one warp each)
    const int* tstart;     // host-cut tiles: tile t of this step covers patterns tstart[t] .. tstart[t+1]-1 (at most 32, and few
                           // enough that the tile fits the shared slice); nullptr: fixed tiles of 32 from p0
    int chain_first, nchain, pad_;   // single-child steps on the same patterns (WGD-type nodes above this one): they follow in
                                     // the same warp, with the column still in shared memory; stored at step[nsteps + chain_first ...]
    double* out;
    int* out_scl;
    const uint16_t* cnt;
    // the root's step integrates its column on the spot (integrate_root / condition_oib, model.jl:465-501):
    // pat_ll[p] = log Σ_n ℓ[n]·rootw[n] - rootw[0], the value every family of the pattern gets
    const double* rootw;
    double* pat_ll;
};
#define BLG_P_MAXSTEPS 110
#define BLG_P_MAXCHILD 270
struct PList {
    int nsteps, ntiles, first_of_call, pad_;
    PStep step[BLG_P_MAXSTEPS];
    PChild child[BLG_P_MAXCHILD];
};
static_assert(sizeof(PList) <= 32000, "kernel parameter space");
__device__ __forceinline__ ChildArgs child_args(const PChild& c, double psi, double cc) { return ChildArgs{c.ell, c.Wt, c.wdp, c.table, psi, cc}; }

// WIDE: the instantiation for patterns whose smaller child has 20..35 counts — register-tiled merges 28 and 36 wide, which
// need more than 128 registers (2 blocks per SM instead of 4); such patterns sit at the front of a node's order.
// SH: every tile of the launch fits its shared slice (the host cut the tiles that way): the working column is addressed as
// an offset into shared memory (SCol: LDS/STS).  !SH: the column may live in global scratch (GCol, generic accesses) —
// polytomies and anything else the host could not bound.
template <bool WIDE, bool SH>
__global__ void __launch_bounds__(128, WIDE ? 2 : 4) ptile_kernel(const __grid_constant__ PList L, const NodeK* __restrict__ nk,
                                                       const int* __restrict__ flags, unsigned int* __restrict__ ran,
                                                       int slice_cap, double* gscratch, size_t gs_stride) {
    extern __shared__ double smem[];
    typedef typename std::conditional<SH, SCol, GCol>::type C;
    if (flags[0] == 0) return;        // some node is outside the table-free forms' range: the caller re-runs the general kernel
    const unsigned lane = threadIdx.x & 31u;
    const int warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * (int)(blockDim.x >> 5) + warp;
    if (tile >= L.ntiles) return;
    int lo = 0, hi = L.nsteps - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (L.step[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    const PStep& st = L.step[lo];
    const int k = st.k;
    const PChild* ch = &L.child[st.first];
    int p, pend;
    if (st.tstart != nullptr) { p = st.tstart[tile - st.tile0] + (int)lane; pend = st.tstart[tile - st.tile0 + 1]; }
    else { p = st.p0 + (tile - st.tile0) * 32 + (int)lane; pend = st.U; }
    const bool live = p < pend;
    if (!live) p = pend - 1;          // idle lanes shadow the tile's last pattern (loads stay in bounds; they never store)
    const size_t ldo = (size_t)st.ld;
    const int sloff = warp * slice_cap * 32 + (int)lane;
    const int X = st.cnt[p];
    const int Xw = warp_max(X);
    const double ip0 = nk[st.node].ip0;          // 1, or NaN where the reference's column is NaN (here or below)
    if (Xw == 0) {
        // no pattern of this tile has a gene below the node: ℓ = [1] (n = 0 lineages leave no descendants with probability 1)
        if (live) {
            st.out[p] = ip0; st.out_scl[p] = 0;
#pragma unroll 1
            for (int cj = 0; cj < st.nchain; ++cj) {
                const PStep& sv = L.step[L.nsteps + st.chain_first + cj];
                sv.out[p] = nk[sv.node].ip0; sv.out_scl[p] = 0;
            }
        }
        return;
    }
    const int c0 = ch[0].pidx ? ch[0].pidx[p] : p;
    const int M0 = ch[0].xcnt ? ch[0].xcnt[p] : ch[0].cnt[c0];
    int c1 = 0, M1 = 0;
    if (k >= 2) { c1 = ch[1].pidx ? ch[1].pidx[p] : p; M1 = ch[1].xcnt ? ch[1].xcnt[p] : ch[1].cnt[c1]; }
    int sclsum = (ch[0].scl ? ch[0].scl[c0] : 0) + ((k >= 2 && ch[1].scl) ? ch[1].scl[c1] : 0);
    const int M0w = warp_max(M0);
    const int M1w = warp_max(M1);
    int sumw = M0w + M1w;
    // which side of a binary merge is register-tiled: the child with the smaller bound in THIS tile
    const bool swap = (k == 2) && (M0w < M1w);
    constexpr int TMAX = WIDE ? 36 : 20;      // widest register-tiled merge of this instantiation
    // (host-cut tiles never need the chained-tile merge: the smaller maximum stays below TMAX by construction)
    bool tiled = SH ? false : ((k == 2) ? (min(M0w, M1w) >= TMAX) : (M1w >= TMAX));
#pragma unroll 1
    for (int i = 2; i < k; ++i) {
        const int ci = ch[i].pidx ? ch[i].pidx[p] : p;
        const int Miw = warp_max((int)ch[i].cnt[ci]);
        sumw += Miw;
        if (!SH && Miw >= TMAX) tiled = true;
        if (ch[i].scl) sclsum += ch[i].scl[ci];
    }
    const int need = tiled ? 3 * (sumw + 1) : sumw + k + 1;
    C R;
    bool insl = true;
    if constexpr (SH) { R = SCol{sloff}; (void)need; }
    else { insl = need <= slice_cap; R = GCol{insl ? smem + sloff : gscratch + (size_t)tile * gs_stride + lane}; }
    const NodeK k0 = nk[ch[0].node];
    // the input columns of the first two children are gathers (every lane reads another pattern's column): all requests
    // are issued at once as asynchronous copies — a loop of dependent loads would pay one L2 round trip per element
    const bool async0 = insl;
    const bool async1 = (k >= 2) && !tiled && insl;
    if (async0) {
        if (ch[0].ell == nullptr) {
            const double* row = ch[0].Wt + (size_t)M0 * ch[0].wdp;
#pragma unroll 1
            for (int s = 0; s <= M0w; ++s) cp_async_f64(R.ptr(s * 32), (s <= M0) ? row + s : row, s <= M0);
        } else {
            const double* src = ch[0].ell + c0;
            const size_t ld0 = (size_t)ch[0].ld;
#pragma unroll 1
            for (int j = 0; j <= M0w; ++j) cp_async_f64(R.ptr(j * 32), (j <= M0) ? src + (size_t)j * ld0 : src, j <= M0);
        }
        cp_async_commit();
    }
    if (async1) {
        const C dst = R + (M0w + 1) * 32;
        if (ch[1].ell == nullptr) {
            const double* row = ch[1].Wt + (size_t)M1 * ch[1].wdp;
#pragma unroll 1
            for (int s = 0; s <= M1w; ++s) cp_async_f64(dst.ptr(s * 32), (s <= M1) ? row + s : row, s <= M1);
        } else {
            const double* src = ch[1].ell + c1;
            const size_t ld1 = (size_t)ch[1].ld;
#pragma unroll 1
            for (int j = 0; j <= M1w; ++j) cp_async_f64(dst.ptr(j * 32), (j <= M1) ? src + (size_t)j * ld1 : src, j <= M1);
        }
        cp_async_commit();
    }
    // child 0's column is waited for now, child 1's only after child 0 has been contracted
    if (async0) { if (async1) cp_async_wait_but_one(); else cp_async_wait_all(); }
    sweep_child(child_args(ch[0], k0.psi, k0.cc), c0, (size_t)ch[0].ld, M0, M0w, R, async0 ? 2 : 0, 1.0);
    if (async1) cp_async_wait_all();
    int Sw = M0w;
    double mxv = 0.0;
    bool scaled = false;        // has the (1-E_k)^(-n) normalisation already been applied (and mxv taken)?
    double Eraw = k0.e;         // product of the e's of the children merged so far (polytomies: clamped when used)
    const double inv = nk[st.node].inv;
    if (!tiled) {
#pragma unroll 1
        for (int i = 1; i < k; ++i) {
            const int ci = (i == 1) ? c1 : (ch[i].pidx ? ch[i].pidx[p] : p);
            const int Mi = (i == 1) ? M1 : (int)ch[i].cnt[ci];
            const int Miw = (i == 1) ? M1w : warp_max(Mi);
            const C colB = R + (Sw + 1) * 32;
            const NodeK ki = nk[ch[i].node];
            sweep_child(child_args(ch[i], ki.psi, ki.cc), ci, (size_t)ch[i].ld, Mi, Miw, colB, (i == 1 && async1) ? 2 : 0, 1.0);
            const bool last = (i == k - 1);
            const double em0 = last ? ip0 : 1.0, emg = last ? inv : 1.0;
            const int Xm = last ? X : 0x7fffffff;
            const C tA = swap ? colB : R;
            const C sB = swap ? R : colB;
            const int tW = swap ? Miw : Sw, sW = swap ? Sw : Miw;
            const double Epre = fmin(Eraw, 1.0);              // clamped prefix of the children merged before this one
            const double E = swap ? fmin(ki.e, 1.0) : Epre;
            const double e = swap ? k0.e : ki.e;
            const double invE = swap ? ki.inve : ((k == 2) ? k0.inve : 1.0 / E);
            const double inve = swap ? k0.inve : ki.inve;
            Eraw *= ki.e;
            double mx;
            if (WIDE) {
                if (sW < 20) mx = sweep_merge_fast<20>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
                else if (sW < 28) mx = sweep_merge_fast<28>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
                else mx = sweep_merge_fast<36>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
            } else {
                if (sW < 4) mx = sweep_merge_fast<4>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
                else if (sW < 8) mx = sweep_merge_fast<8>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
                else if (sW < 12) mx = sweep_merge_fast<12>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
                else mx = sweep_merge_fast<20>(R, tA, tW, sB, sW, E, invE, e, inve, em0, emg, Xm);
            }
            Sw += Miw;
            if (last) { mxv = mx; scaled = true; }
        }
    } else if constexpr (!SH) {
        const int cap = sumw + 1;
        const C rA = R;
        const C rH = R + cap * 32;
        const C rO = R + 2 * cap * 32;
#pragma unroll 1
        for (int i = 1; i < k; ++i) {
            const int ci = (i == 1) ? c1 : (ch[i].pidx ? ch[i].pidx[p] : p);
            const int Mi = (i == 1) ? M1 : (int)ch[i].cnt[ci];
            const int Miw = (i == 1) ? M1w : warp_max(Mi);
            const NodeK ki = nk[ch[i].node];
            sweep_child(child_args(ch[i], ki.psi, ki.cc), ci, (size_t)ch[i].ld, Mi, Miw, rO, 0, 1.0);
#pragma unroll 1
            for (int n = Miw + 1; n <= Sw + Miw; ++n) rO[n * 32] = 0.0;
            sweep_merge_tiled(rA, Sw, rO, Miw, rH, fmin(Eraw, 1.0), ki.e);
            Eraw *= ki.e;
            Sw += Miw;
#pragma unroll 1
            for (int n = 0; n <= Sw; ++n) rA[n * 32] = rO[n * 32];     // result back to the front of R
        }
    }
    // ℓ_u[n] = Ã[n]·(1-E_k)^(-n), rescale to a power-of-two exponent, store
    if (!scaled) {
        double ip = ip0;
#pragma unroll 1
        for (int n = 0; n <= Xw; ++n) {
            double v = 0.0;
            if (n <= X) {
                v = (n <= Sw ? R[n * 32] : 0.0) * ip;
                mxv = fmax(mxv, v);
            }
            R[n * 32] = v;
            ip *= inv;
        }
    }
    int ex = 0;
    if (mxv > 0.0) {                               // (fmax skips NaNs; an infinite maximum keeps exponent 0)
        const int be = (__double2hiint(mxv) >> 20) & 0x7ff;
        if (be != 0 && be != 0x7ff) ex = be - 1023;
    }
    const double sc = __hiloint2double((1023 - ex) << 20, 0);         // 2^-ex, exact (|ex| <= 1022)
    const bool keep = st.nchain > 0;               // a chain follows: the resident copy is rescaled too (same bits as the stored one)
    {
        double* o = st.out + p;
        int n = 0;
#pragma unroll 1
        for (; n < X; n += 2) {                    // two loads in flight
            const double v0 = R[n * 32] * sc, v1 = R[(n + 1) * 32] * sc;
            if (live) { o[(size_t)n * ldo] = v0; o[(size_t)(n + 1) * ldo] = v1; }
            if (keep) { R[n * 32] = v0; R[(n + 1) * 32] = v1; }
        }
        if (n == X) {
            const double v0 = R[n * 32] * sc;
            if (live) o[(size_t)n * ldo] = v0;
            if (keep) R[n * 32] = v0;
        }
        if (live) st.out_scl[p] = sclsum + ex;
    }
    if (st.pat_ll != nullptr) {                    // same arithmetic, in the same order, as root_pattern_kernel on the stored column
        const double* __restrict__ rw = st.rootw;
        // four interleaved partial sums (n ≡ 1, 2, 3, 0 mod 4), added pairwise: four loads in flight per trip
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int n = 1;
#pragma unroll 1
        for (; n + 3 <= Xw; n += 4) {
            const double w0 = rw[n], w1 = rw[n + 1], w2 = rw[n + 2], w3 = rw[n + 3];
            const double r0 = R[n * 32], r1 = R[(n + 1) * 32], r2 = R[(n + 2) * 32], r3 = R[(n + 3) * 32];
            if (n <= X) a0 = fma(r0 * sc, w0, a0);
            if (n + 1 <= X) a1 = fma(r1 * sc, w1, a1);
            if (n + 2 <= X) a2 = fma(r2 * sc, w2, a2);
            if (n + 3 <= X) a3 = fma(r3 * sc, w3, a3);
        }
#pragma unroll 1
        for (; n <= Xw; ++n) {
            const double w = rw[n];
            const int k = (n - 1) & 3;
            if (n <= X) {
                const double v = R[n * 32] * sc;
                if (k == 0) a0 = fma(v, w, a0); else if (k == 1) a1 = fma(v, w, a1); else if (k == 2) a2 = fma(v, w, a2); else a3 = fma(v, w, a3);
            }
        }
        const double acc = (a0 + a1) + (a2 + a3);
        double l = log(acc) + (double)(sclsum + ex) * 0.6931471805599453 - rw[0];
        if (!isfinite(l)) l = -INFINITY;                                        // model.jl:366
        if (live) st.pat_ll[p] = l;
    }
    // ---- the chain: single-child nodes above this one, on the same patterns.  The column stays in R, rescaled like its
    // stored copy (so a partial recompute that starts from the stored copy sees the same bits) ------------------------------
    int prevScl = sclsum + ex;
#pragma unroll 1
    for (int cj = 0; cj < st.nchain; ++cj) {
        const PStep& sv = L.step[L.nsteps + st.chain_first + cj];
        const PChild& cv = L.child[sv.first];
        const NodeK kv = nk[cv.node];
        __syncwarp();
        sweep_child(child_args(cv, kv.psi, kv.cc), 0, 0, X, Xw, R, 2, 1.0);
        // ℓ_v[n] = b[n]·(1-e)^(-n): first pass the values and their maximum, second pass the rescaled column out
        double ip = nk[sv.node].ip0;
        const double invv = nk[sv.node].inv;
        double mv = 0.0;
        {
            int n = 0;
#pragma unroll 1
            for (; n < X; n += 2) {
                const double v0 = R[n * 32] * ip, v1 = R[(n + 1) * 32] * (ip * invv);
                mv = fmax(mv, fmax(v0, v1));
                R[n * 32] = v0; R[(n + 1) * 32] = v1;
                ip *= invv * invv;
            }
            if (n == X) { const double v0 = R[n * 32] * ip; mv = fmax(mv, v0); R[n * 32] = v0; }
        }
        int exv = 0;
        if (mv > 0.0) {
            const int be = (__double2hiint(mv) >> 20) & 0x7ff;
            if (be != 0 && be != 0x7ff) exv = be - 1023;
        }
        const double scv = __hiloint2double((1023 - exv) << 20, 0);
        {
            double* o = sv.out + p;
            int n = 0;
#pragma unroll 1
            for (; n < X; n += 2) {
                const double v0 = R[n * 32] * scv, v1 = R[(n + 1) * 32] * scv;
                if (live) { o[(size_t)n * ldo] = v0; o[(size_t)(n + 1) * ldo] = v1; }
                R[n * 32] = v0; R[(n + 1) * 32] = v1;
            }
            if (n == X) {
                const double v0 = R[n * 32] * scv;
                if (live) o[(size_t)n * ldo] = v0;
                R[n * 32] = v0;
            }
            if (live) sv.out_scl[p] = prevScl + exv;
        }
        prevScl += exv;
    }
}

// root integration over pattern columns: family f reads the column of its root pattern fam2pat[f]
// (integrate_root model.jl:465-475, condition_oib model.jl:492-501; same reduction as root_kernel)
__global__ void __launch_bounds__(256) root_pattern_kernel(const double* __restrict__ ell, const int* __restrict__ scl,
                                                           const uint16_t* __restrict__ pcnt, const int* __restrict__ fam2pat,
                                                           const double* __restrict__ rootw, const double* __restrict__ weight,
                                                           double* __restrict__ fam_ll, double* __restrict__ partial, int N, size_t ld,
                                                           unsigned int* __restrict__ ticket, unsigned int nblocks_total,
                                                           const int* __restrict__ flags, unsigned int* __restrict__ ran,
                                                           const double* __restrict__ pat_ll, int pat_lo,
                                                           double* __restrict__ marker, double* __restrict__ out) {
    __shared__ double red[256];
    __shared__ bool last_block;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double contrib = 0.0;
    if (f == 0 && flags[0]) atomicAdd(&ran[0], 1u);          // the table-free pattern path did this call's work
    if (f < N) {
        const int p = fam2pat[f];
        double l;
        if (p < 0) {
            l = 0.0;                             // a family of the heavy set: its value comes from root_ragged_kernel
        } else if (pat_ll != nullptr && p >= pat_lo) {
            l = pat_ll[p];                       // integrated by the tile that computed the root's column
        } else {
            const int X = pcnt[p];
            // (the same four interleaved partial sums as the root's tiles: both routes give the same bits)
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            for (int n = 1; n <= X; ++n) a[(n - 1) & 3] = fma(ell[(size_t)n * ld + p], rootw[n], a[(n - 1) & 3]);
            const double acc = (a[0] + a[1]) + (a[2] + a[3]);
            l = log(acc) + (double)scl[p] * 0.6931471805599453 - rootw[0];
            if (!isfinite(l)) l = -INFINITY;                                        // model.jl:366
        }
        fam_ll[f] = l;
        contrib = p < 0 ? 0.0 : weight[f] * l;
    }
    red[threadIdx.x] = contrib;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {                                         // fixed-order tree: deterministic
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = red[0];
        __threadfence();
        last_block = atomicAdd(ticket, 1u) == nblocks_total - 1;
    }
    __syncthreads();
    if (last_block) {
        __threadfence();
        const volatile double* vp = partial;
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)nblocks_total; i += 256) s += vp[i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int k = 128; k > 0; k >>= 1) {
            if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            out[0] = red[0];
            marker[0] = flags[0] ? 0.0 : 1.0;
            *ticket = 0u;
        }
    }
}
