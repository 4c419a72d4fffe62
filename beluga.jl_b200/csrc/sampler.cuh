// sampler.cuh — rand(model, N) at the count level, on the device.  Included by beluga_b200.cu.
//
// The reference simulates gene trees lineage by lineage (src/sim.jl:93-145: exponential waiting times, birth with
// probability λ/(λ+μ), a WGD keeps the duplicate with probability q) and only then counts the leaves (`profile`, :72-80).
// The counts have a closed-form law: one lineage entering a branch of length t leaves 0 descendants with probability ϕ(t)
// and otherwise 1 + Geometric(1-ψ(t)) (utils.jl:8-9 are exactly these ϕ, ψ), independently per lineage; the root holds
// 1 + Geometric(η) lineages (sim.jl:83); a WGD turns n lineages into n + Binomial(n, q) (sim.jl:112-115).  One thread
// draws one family with its own Philox stream (seed, family index): the result does not depend on the launch shape.
// A WGT (src/sim.jl does not simulate them; the reference's own ϵ and w* formulas for WGTs, node.jl:144-149,205-218, do not
// agree on a retention law — their weights do not even sum to one) keeps each of the two extra copies independently with
// probability q: 1, 2 or 3 copies per lineage with probabilities (1-q)², 2q(1-q), q².
// `rand(model, N; condition)` (sim.jl:17-27) is a rejection sampler: draws are made in chunks, the accepted ones are
// compacted in draw order (a scan, not atomics: the output is reproducible).
#pragma once
// (needs <curand_kernel.h>, included at global scope by beluga_b200.cu)

struct SimNode {          // one node in pre-order (parents before children)
    int parent;           // position of the parent in the pre-order list (-1: root)
    int ret;              // retention event at the TOP of the branch: 0 none, 2 below a WGD, 3 below a WGT
    int leaf;             // leaf column (ascending leaf id), -1 for internal nodes
    int clade;            // which child of the root this node descends from (-1: the root itself)
    double q, phi, psi;   // retention probability; extinction / geometric parameter of the branch (phi = psi = 0: t = 0)
};

__device__ __forceinline__ int sim_geometric(curandStatePhilox4_32_10_t* st, double p_fail) {   // failures before the first success
    if (!(p_fail > 0.0)) return 0;
    const double u = curand_uniform_double(st);                      // (0, 1]
    const double k = floor(log(u) / log(p_fail));
    return k > 2.0e9 ? 2000000000 : (int)k;
}

// cnt: scratch [nnodes][chunk] (column per node, coalesced over the families of the chunk)
__global__ void __launch_bounds__(256) rand_counts_kernel(const SimNode* __restrict__ nodes, int nnodes, int nleaves, int nclades, double eta,
                                                          unsigned long long seed, unsigned long long draw0, int chunk, int condition,
                                                          int max_leaf, int max_rootsum, int* __restrict__ cnt, uint16_t* __restrict__ leaves,
                                                          unsigned int* __restrict__ keep) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= chunk) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, draw0 + (unsigned long long)f, 0ULL, &st);
    long long total = 0;
    int maxleaf = 0;
    unsigned int clade_seen = 0u;
    bool overflow = false;
    for (int i = 0; i < nnodes; ++i) {
        const SimNode nd = nodes[i];
        int y;
        if (nd.parent < 0) {
            y = 1 + sim_geometric(&st, 1.0 - eta);                   // sim.jl:83
        } else {
            int n = cnt[(size_t)nd.parent * chunk + f];
            if (nd.ret == 2) {                                       // sim.jl:112-115
                int extra = 0;
                for (int l = 0; l < n; ++l) extra += curand_uniform_double(&st) <= nd.q ? 1 : 0;
                n += extra;
            } else if (nd.ret == 3) {
                const double p1 = (1.0 - nd.q) * (1.0 - nd.q), p2 = p1 + 2.0 * nd.q * (1.0 - nd.q);
                int extra = 0;
                for (int l = 0; l < n; ++l) { const double u = curand_uniform_double(&st); extra += u <= p1 ? 0 : (u <= p2 ? 1 : 2); }
                n += extra;
            }
            if (nd.phi == 0.0 && nd.psi == 0.0) {
                y = n;                                               // a branch of length 0
            } else {
                y = 0;
                for (int l = 0; l < n; ++l)
                    if (curand_uniform_double(&st) > nd.phi) y += 1 + sim_geometric(&st, nd.psi);
            }
            if (y > 60000) { overflow = true; y = 60000; }
        }
        cnt[(size_t)i * chunk + f] = y;
        if (nd.leaf >= 0) {
            leaves[(size_t)f * nleaves + nd.leaf] = (uint16_t)y;
            total += y;
            maxleaf = max(maxleaf, y);
            if (y > 0 && nd.clade >= 0) clade_seen |= 1u << nd.clade;
        }
    }
    bool ok = !overflow;
    if (condition == 1) ok = ok && total > 0;                                         // non-extinct (sim.jl:17 default)
    if (condition == 2) ok = ok && clade_seen == ((1u << nclades) - 1u);              // one gene in every root clade (sim.jl:51)
    if (max_leaf >= 0) ok = ok && maxleaf <= max_leaf;
    if (max_rootsum >= 0) ok = ok && total <= (long long)max_rootsum;
    keep[f] = ok ? 1u : 0u;
}

// ordered compaction of the accepted draws: exclusive scan of keep[] in three small launches, then the copy
__global__ void __launch_bounds__(256) sim_block_sums(const unsigned int* __restrict__ keep, int n, unsigned int* __restrict__ bsum) {
    __shared__ unsigned int red[256];
    const int base = blockIdx.x * 1024;
    unsigned int v = 0;
    for (int i = threadIdx.x; i < 1024; i += 256) if (base + i < n) v += keep[base + i];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) bsum[blockIdx.x] = red[0];
}
__global__ void sim_scan_sums(unsigned int* bsum, int nb, unsigned int* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned int run = 0;
        for (int i = 0; i < nb; ++i) { const unsigned int v = bsum[i]; bsum[i] = run; run += v; }
        *total = run;
    }
}
__global__ void __launch_bounds__(256) sim_compact(const unsigned int* __restrict__ keep, int n, const unsigned int* __restrict__ bsum,
                                                   const uint16_t* __restrict__ leaves, int nleaves, uint16_t* __restrict__ out,
                                                   long long out0, long long out_cap) {
    __shared__ unsigned int part[256];
    const int base = blockIdx.x * 1024 + threadIdx.x * 4;            // four consecutive draws per thread
    unsigned int v[4], tot = 0;
    for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? keep[base + i] : 0u; tot += v[i]; }
    part[threadIdx.x] = tot;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int run = bsum[blockIdx.x]; for (int i = 0; i < 256; ++i) { const unsigned int t_ = part[i]; part[i] = run; run += t_; } }
    __syncthreads();
    unsigned int run = part[threadIdx.x];
    for (int i = 0; i < 4; ++i) {
        if (v[i]) {
            const long long dst = out0 + (long long)run;
            if (dst < out_cap) {
                const uint16_t* src = leaves + (size_t)(base + i) * nleaves;
                uint16_t* d = out + (size_t)dst * nleaves;
                for (int j = 0; j < nleaves; ++j) d[j] = src[j];
            }
            ++run;
        }
    }
}
