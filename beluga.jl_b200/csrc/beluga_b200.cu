// beluga_b200.cu — B200 (sm_100a) implementation of the DLWGD gene-family likelihood hot path.
//
// What this file replaces in arzwa/Beluga.jl (all Julia, file:line under the reference tree):
//   per-branch numerics  setϵ!/setW!/wstar*!     src/node.jl:136-218, src/utils.jl:8-19
//   pruning recursion    csuros_miklos!          src/model.jl:402-463
//   root + conditioning  integrate_root, condition_oib   src/model.jl:465-501
//   profile storage/DP   Profile/PArray, logpdf!, set!/rev!/extend!/shrink!   src/profile.jl:14-121
//
// Design: counts and partial likelihoods live on the device node-major and
// family-contiguous; ℓ_u[n][f] is kept in LINEAR fp64 with one power-of-two exponent per
// (family, node).  ε, W and the per-node merge tables are built on the device per parameter
// update.  One pruning launch per internal node applies the edge contraction W·ℓ and the
// conditional-survival merge to every family at once (one thread = one family, all global
// accesses coalesced along the family axis).  No CPU fallback: every entry fails loudly when
// CUDA is unavailable.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <limits>
#include <cstdint>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "beluga_b200.h"

#define BLG_MAXK 8          // max children of one node (polytomies)
#define BLG_MAXCOUNT 4095   // largest extended count supported (general kernel: probability-form recursion, no tables)
#define BLG_MMA_MIN 12         // count bound of a warp from which a contraction runs as a DMMA table product
#define BLG_FAST_MAXCOUNT 100   // largest count bound the table-free sweep kernel handles (families above it: heavy shard)

namespace {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                            std::to_string(__LINE__) + ")");                                       \
    } while (0)

thread_local std::string g_create_error;

// Every device allocation and release of the library goes through these two: a cached CUDA graph of a likelihood call
// (blg_handle::logpdf) stays valid only while no pointer it could hold has changed, and nothing may be allocated while a
// call is being captured (the capture is dropped and the call runs launch by launch instead).
std::atomic<unsigned long long> g_mem_epoch{0};
thread_local bool g_capturing = false;
template <class T> cudaError_t blgMalloc(T** p, size_t bytes) {
    if (g_capturing) throw std::runtime_error("internal: device allocation inside a graph capture");
    g_mem_epoch.fetch_add(1, std::memory_order_relaxed);
    return cudaMalloc((void**)p, bytes);
}
inline cudaError_t blgFree(void* p) {
    g_mem_epoch.fetch_add(1, std::memory_order_relaxed);
    return cudaFree(p);
}

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen): the library owns the communicator of a multi-GPU run (one process per GPU, one
// handle per process) and every likelihood / gradient entry returns the all-reduced value (src/profile.jl:56-75 is a
// mapreduce over worker processes).  Binding at run time keeps a single copy of NCCL in a process that has already loaded
// one (a PyTorch host), and lets single-GPU users run without NCCL installed.
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void* Comm;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        const char* names[] = {getenv("BLG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
            why = dlerror();
        }
        if (!lib) return false;
        GetUniqueId = (int (*)(UniqueId*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(Comm*, int, UniqueId, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        CommDestroy = (int (*)(Comm))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { why = "NCCL symbols missing"; lib = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        int r_ = (call);                                                                                   \
        if (r_ != 0)                                                                                       \
            throw std::runtime_error(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error")); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side model description
// ------------------------------------------------------------------------------------------------
#define BLG_TS 8            // register tile (counts per thread) of the contraction and the merge
struct NodeTab {        // per node id, device-resident tables
    double* Wt;         // TRANSPOSED w*: Wt[j*wdp + n] = w*(n|j) for the branch above the node, n <= j < wdim
    int wdim;           // xmax + 1
    int wdp;            // row stride, wdim rounded up to BLG_TS (padding is zero)
    double* ipow;       // (1-E_k)^(-n), n = 0..nrow-1  (NaN-poisoned when the reference would be)
    int nrow;           // xmax_u + 1 (fast shard; 0 when it holds no family)
    int xfast;          // count bound of this column in the fast shard
    // merge coefficients for the child in merge position i >= 1, diagonal-major:
    // coefD[i][t*cdp[i] + s] = C(t+s, s)·E_{i-1}^s  (0 where t+s > xmax_u or s > xmax of that child)
    double* coefD[BLG_MAXK];
    int cdp[BLG_MAXK];
    int mo[BLG_MAXK];   // merge order: node indices of the children in the order they are merged
};

// per node id, DEVICE-RESIDENT: written by tables_kernel at every parameter update, read by the pruning kernels by node
// index (the host never needs these numbers, so a parameter update involves no read-back).  80 bytes, 16-byte aligned:
// the sweep kernel stages {e, inve, psi, cc} and {inv, ip0} with 16-byte cp.async.
struct NodeK {
    // the node as a CHILD (the branch above it)
    double e;                 // ϵ[1]: extinction probability of one lineage at the top of the branch
    double inve;              // 1/min(e, 1)
    double psi, cc;           // ψ', c = (1-ϕ')(1-ψ') of an ordinary branch
    double chat;              // c/(1-e) (ordinary branch);  a/(1-e) below a WGD  (general kernel)
    double bhat;              // b/(1-e) below a WGD
    // the node as a PARENT
    double inv;               // 1/(1-E_k)
    double ip0;               // 1, or NaN when the reference's column is NaN here or anywhere below (model.jl:433-435)
    double ip0raw;            // the same for this node alone
    double pad_;
};
static_assert(sizeof(NodeK) == 80, "NodeK layout");

struct ModelDev {
    int n;                     // ncols (max id)
    int model_kind;
    const uint8_t* kind;
    const int* parent;         // 0-based, -1 root/absent
    const int* child_off;      // CSR over ordered children
    const int* child_idx;
    const double* t;
    const double* lam;
    const double* mu;
    const double* q;
    double eta;
    double* eps0;              // ϵ[1]: extinction prob at the top of the branch above the node
    double* eps1;              // ϵ[2]: at the node
    NodeTab* tab;
    const double* pascal;      // C(n,s), row stride pdim
    int pdim;
    NodeK* nk;                 // per-node scalars of the pruning kernels
    int* fast_ok;              // per node: 1 if the factorial-scaled forms of the sweep kernel stay inside fp64 here
    double* rootw;             // root integration weights, rebuilt with the root's tables (nullptr: not now)
    double* rootlw;            // ... in log form (heavy shard; nullptr: not needed)
    int rootw_n;
    // summary written by the last block of tables_kernel: flags[0] = every present node is fast_ok
    int* flags;
    unsigned int* ticket;
    const int* post;           // all present nodes in post-order
    int npost;
};

__device__ __forceinline__ bool d_isapprox(double a, double b) {
    if (a == b) return true;
    if (!isfinite(a) || !isfinite(b)) return false;
    return fabs(a - b) <= 1.4901161193847656e-8 * fmax(fabs(a), fabs(b));
}
__device__ __forceinline__ double d_approx1(double x) { return d_isapprox(x, 1.0) ? 1.0 : x; }   // utils.jl:19
__device__ __forceinline__ double d_getphi(double t, double l, double m) {                       // utils.jl:8
    if (d_isapprox(l, m)) return l * t / (1.0 + l * t);
    double e = exp(t * (l - m));
    return m * (e - 1.0) / (l * e - m);
}
__device__ __forceinline__ double d_getpsi(double t, double l, double m) {                       // utils.jl:9
    if (d_isapprox(l, m)) return l * t / (1.0 + l * t);
    return (l / m) * d_getphi(t, l, m);
}
__device__ __forceinline__ double d_ep(double l, double m, double t, double e) {                 // utils.jl:17-18
    if (d_isapprox(l, m)) return 1.0 + (1.0 - e) / (m * (e - 1.0) * t - 1.0);
    return d_approx1((m + (l - m) / (1.0 + exp((l - m) * t) * l * (e - 1.0) / (m - l * e))) / l);
}
__device__ __forceinline__ bool d_isawgd(const ModelDev& m, int i) {
    uint8_t k = m.kind[i];
    return k == BLG_WGD || k == BLG_WGT || k == BLG_WGDAFTER;
}
__device__ __forceinline__ int d_nonwgdparent(const ModelDev& m, int i) { while (d_isawgd(m, i)) i = m.parent[i]; return i; }
__device__ __forceinline__ int d_nonwgdchild(const ModelDev& m, int i) { while (d_isawgd(m, i)) i = m.child_idx[m.child_off[i]]; return i; }
// node.jl:220-230 getλμ(parent a, child b)
__device__ __forceinline__ void d_getlm(const ModelDev& m, int a, int b, double& l, double& mu) {
    b = d_isawgd(m, b) ? d_nonwgdchild(m, b) : b;
    if (m.model_kind == BLG_NODE_MODEL) {
        a = d_isawgd(m, a) ? d_nonwgdparent(m, a) : a;
        l = (m.lam[a] + m.lam[b]) / 2.0;
        mu = (m.mu[a] + m.mu[b]) / 2.0;
    } else {
        l = m.lam[b];
        mu = m.mu[b];
    }
}
__device__ __noinline__ double d_powi(double b, int n) {   // b^n, n >= 0, by squaring (not inlined: it is called from two dozen places)
    double r = 1.0;
    while (n > 0) { if (n & 1) r *= b; b *= b; n >>= 1; }
    return r;
}

// setϵ! (node.jl:136-162) for the listed nodes.  A node needs its children's ϵ[2], so the list is cut into
// segments (seg[0..nseg]) such that no node depends on another node of its own segment: the host passes either
// one node per segment (a literal replay of the reference's call order) or the nodes grouped by height above
// the leaves (set!(d): every level in parallel).  One block; threads of a segment work concurrently.
// stage != 0: the topology, the parameters, ε and the two lists are first copied to shared memory (one parallel pass): the
// levels are a chain of dependent reads, one L2 round trip each if left in global memory (61·n + 4·(nlist + nseg) bytes).
__global__ void __launch_bounds__(128) eps_kernel(ModelDev mg, const int* __restrict__ glist, const int* __restrict__ gseg, int nseg, int nlist,
                                                  int stage) {
    extern __shared__ double esm[];
    ModelDev m = mg;
    const int* list = glist;
    const int* seg = gseg;
    if (stage) {
        const int n = mg.n;
        double* s_t = esm; double* s_lam = s_t + n; double* s_mu = s_lam + n; double* s_q = s_mu + n;
        double* s_e0 = s_q + n; double* s_e1 = s_e0 + n;
        int* s_parent = reinterpret_cast<int*>(s_e1 + n);
        int* s_off = s_parent + n;
        int* s_idx = s_off + n + 1;
        int* s_list = s_idx + n;
        int* s_seg = s_list + nlist;
        uint8_t* s_kind = reinterpret_cast<uint8_t*>(s_seg + nseg + 1);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            s_t[i] = mg.t[i]; s_lam[i] = mg.lam[i]; s_mu[i] = mg.mu[i]; s_q[i] = mg.q[i];
            s_e0[i] = mg.eps0[i]; s_e1[i] = mg.eps1[i];
            s_parent[i] = mg.parent[i]; s_off[i] = mg.child_off[i]; s_idx[i] = mg.child_idx[i]; s_kind[i] = mg.kind[i];
        }
        if (threadIdx.x == 0) s_off[n] = mg.child_off[n];
        for (int i = threadIdx.x; i < nlist; i += blockDim.x) s_list[i] = glist[i];
        for (int i = threadIdx.x; i <= nseg; i += blockDim.x) s_seg[i] = gseg[i];
        m.t = s_t; m.lam = s_lam; m.mu = s_mu; m.q = s_q; m.eps0 = s_e0; m.eps1 = s_e1;
        m.parent = s_parent; m.child_off = s_off; m.child_idx = s_idx; m.kind = s_kind;
        list = s_list; seg = s_seg;
        __syncthreads();
    }
    for (int sg = 0; sg < nseg; ++sg) {
      for (int li = seg[sg] + (int)threadIdx.x; li < seg[sg + 1]; li += (int)blockDim.x) {
        int n = list[li];
        int c0 = m.child_off[n], c1 = m.child_off[n + 1];
        uint8_t kd = m.kind[n];
        if (c0 == c1) {
            m.eps1[n] = 0.0;
        } else if (kd == BLG_WGD) {
            double q = m.q[n];
            int c = m.child_idx[c0];
            double ec = m.eps1[c];
            double v = q * (ec * ec) + (1.0 - q) * ec;
            m.eps0[c] = v;
            m.eps1[n] = v;
        } else if (kd == BLG_WGT) {
            double q = m.q[n];
            int c = m.child_idx[c0];
            double ec = m.eps1[c];
            double v = q * (ec * ec * ec) + 2.0 * q * (1.0 - q) * (ec * ec) + (1.0 - q) * (1.0 - q) * ec;
            m.eps0[c] = v;
            m.eps1[n] = v;
        } else {
            double e = 1.0;
            for (int ci = c0; ci < c1; ++ci) {
                int c = m.child_idx[ci];
                double l, mu;
                d_getlm(m, n, c, l, mu);
                double ec = d_approx1(d_ep(l, mu, m.t[c], m.eps1[c]));
                m.eps0[c] = ec;
                e *= ec;
            }
            m.eps1[n] = d_approx1(e);
        }
      }
      __syncthreads();
    }
    if (stage)      // (entries the lists did not touch go back as they came)
        for (int i = threadIdx.x; i < mg.n; i += blockDim.x) { mg.eps0[i] = m.eps0[i]; mg.eps1[i] = m.eps1[i]; }
}

__device__ __forceinline__ double d_log1mexp(double x) { return x < -0.6931471805599453 ? log1p(-exp(x)) : log(-expm1(x)); }

// rootw[n] = exp((n)·log1mexp(lε) + log η + (n-1)·log(1-η) - (n+1)·log1mexp(log(1-η)+lε)), n >= 1;
// rootw[0] holds the conditioning constant log(1-r_1) + log(1-r_2) (+Inf marks "invalid").
// (integrate_root model.jl:465-475, condition_oib model.jl:492-501; one block, tid of nt threads)
// rootlw (optional): the same weights in log form, n >= 1 — the heavy shard's root kernel needs them: with root counts in the
// hundreds the linear weights underflow long before the products ℓ[n]·weight[n] do.
__device__ void d_root_tables(const ModelDev& m, double* rootw, double* rootlw, int nrow, int tid, int nt) {
    int root = 0;
    double eta = m.eta;
    double le = log(m.eps1[root]);
    double a = d_log1mexp(le), b = log(eta), c = log(1.0 - eta), d = d_log1mexp(log(1.0 - eta) + le);
    for (int n = 1 + tid; n < nrow; n += nt) {
        double f = (double)n * a + b + (double)(n - 1) * c;
        f -= (double)(n + 1) * d;
        rootw[n] = exp(f);
        if (rootlw) rootlw[n] = f;
    }
    if (tid == 0) {
        int c0 = m.child_off[root], c1 = m.child_off[root + 1];
        double cond = __longlong_as_double(0x7ff8000000000000LL);
        if (c1 - c0 >= 2) {
            double leta = log(eta);
            double lr[2];
            for (int i = 0; i < 2; ++i) {
                double e = log(m.eps0[m.child_idx[c0 + i]]);
                lr[i] = leta + e - d_log1mexp(d_log1mexp(leta) + e);           // utils.jl:22-23
            }
            if (lr[0] > 0.0 || lr[1] > 0.0) cond = -INFINITY;                   // model.jl:495-497
            else cond = d_log1mexp(lr[0]) + d_log1mexp(lr[1]);
        }
        rootw[0] = cond;
    }
}

// setW! (node.jl:164-218), the per-node scalars of the pruning kernels (NodeK) and the merge tables of the gradient
// kernels, one block per listed node (the list holds every node once).  The W table is only filled where one is
// allocated (tb.Wt != nullptr: dimensions of the FAST shard, i.e. counts <= BLG_FAST_MAXCOUNT; the branch below a WGT keeps
// a table for both shards); the general kernel needs no table for ordinary branches and branches below a WGD.
// The last block to finish folds the NaN rule over the tree (a NaN column makes every column above it NaN) and publishes
// flags[0] = "every node is inside the table-free sweep kernel's range" — no host round trip.
__global__ void __launch_bounds__(128) tables_kernel(ModelDev m, const int* __restrict__ list, int nlist) {
    extern __shared__ double sm[];   // max(2 * maxdim doubles, ncols ints)
    __shared__ bool last_block;
    int n = list[blockIdx.x];
    const NodeTab tb = m.tab[n];
    const int tid = threadIdx.x, nt = blockDim.x;
    uint8_t kd = m.kind[n];
    bool ok = true;                 // may the table-free sweep kernel be used at this node?
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    // ---- the branch above n ---------------------------------------------------------------------
    if (kd != BLG_ROOT) {
        const int d = tb.Wt ? tb.wdim : 0;
        const int dp = tb.wdp;
        double* W = tb.Wt;            // element (row n, column j) of w* lives at W[j*dp + n]
#define WT(n_, j_) W[(size_t)(j_) * dp + (n_)]
        for (int i = tid; i < d * dp; i += nt) W[i] = 0.0;
        __syncthreads();
        int par = m.parent[n];
        double e = m.eps1[n];
        const double etop = m.eps0[n];           // ϵ[1] of n: final, eps_kernel has finished
        if (tid == 0) {
            NodeK& k_ = m.nk[n];
            k_.e = etop;
            k_.inve = 1.0 / (etop > 1.0 ? 1.0 : etop);
            k_.psi = 0.0; k_.cc = 0.0; k_.chat = 0.0; k_.bhat = 0.0;
        }
        if (kd == BLG_WGDAFTER && m.kind[par] == BLG_WGD) {                         // wstar_wgd!
            double q = m.q[par];
            double a = ((1.0 - q) + 2.0 * q * e) * (1.0 - e);
            double b = q * ((1.0 - e) * (1.0 - e));
            if (tid == 0) { m.nk[n].chat = a / (1.0 - etop); m.nk[n].bhat = b / (1.0 - etop); }
            if (d > 0) {
                if (tid == 0) { WT(0, 0) = 1.0; if (d > 1) WT(1, 1) = a; if (d > 2) WT(1, 2) = b; }
                __syncthreads();
                for (int i = 1; i < d; ++i) {
                    for (int j = 2 + tid; j < d; j += nt)
                        WT(i, j) = a * WT(i - 1, j - 1) + b * WT(i - 1, j - 2);
                    __syncthreads();
                }
            }
        } else if (kd == BLG_WGDAFTER && m.kind[par] == BLG_WGT) {                  // wstar_wgt! (literal, j from 3)
            double q = m.q[par];
            double q1 = q, q2 = 2.0 * q * (1.0 - q), q3 = (1.0 - q) * (1.0 - q), ome = 1.0 - e;
            double a = q1 * ome + 2.0 * q2 * e * ome + 3.0 * q3 * (e * e) * ome;
            double b = q2 * (ome * ome) + 3.0 * q3 * e * (ome * ome);
            double c = q3 * (ome * ome * ome);
            if (d > 0) {
                if (tid == 0) { WT(0, 0) = 1.0; if (d > 1) WT(1, 1) = a; if (d > 2) WT(1, 2) = b; if (d > 3) WT(1, 3) = c; }
                __syncthreads();
                for (int i = 1; i < d; ++i) {
                    for (int j = 3 + tid; j < d; j += nt)
                        WT(i, j) = a * WT(i - 1, j - 1) + b * WT(i - 1, j - 2) + c * WT(i - 1, j - 3);
                    __syncthreads();
                }
            }
        } else {                                                                     // wstar!
            double l, mu;
            d_getlm(m, par, n, l, mu);
            double t = m.t[n];
            double phi = d_getphi(t, l, mu), psi = d_getpsi(t, l, mu);
            double nn = 1.0 - psi * e;
            double phip = d_approx1((phi * (1.0 - e) + (1.0 - psi) * e) / nn);
            double psip = d_approx1(psi * (1.0 - e) / nn);
            double cc = (1.0 - phip) * (1.0 - psip);
            if (tid == 0) { m.nk[n].psi = psip; m.nk[n].cc = cc; m.nk[n].chat = cc / (1.0 - etop); }
            {   // (the contractions — Horner on the recurrence, or the w* table on the tensor pipe — need no range rule of their
                // own: every factor is a probability; only the count bound and the sanity of ψ', c are checked here)
                const int xm = tb.xfast;
                if (xm > BLG_FAST_MAXCOUNT) ok = false;
                if (!(psip >= 0.0 && psip <= 1.0 && cc >= 0.0 && cc <= 1.0)) ok = false;
            }
            if (d > 0) {
                double* prev = sm;
                double* cur = sm + d;
                for (int i = tid; i < d; i += nt) prev[i] = (i == 0) ? 1.0 : 0.0;   // column 0
                if (tid == 0) WT(0, 0) = 1.0;
                __syncthreads();
                for (int j = 1; j < d; ++j) {
                    for (int r = tid; r < d; r += nt) {
                        double v = 0.0;
                        if (r >= 1 && r <= j) v = psip * prev[r] + cc * prev[r - 1];
                        cur[r] = v;
                        if (r >= 1 && r <= j) WT(r, j) = v;
                    }
                    __syncthreads();
                    double* tmp = prev; prev = cur; cur = tmp;
                }
            }
        }
#undef WT
    }
    // ---- n as a parent: normalisation, NaN rule, range rule, merge tables of the gradient kernels -----------------------
    int c0 = m.child_off[n], c1 = m.child_off[n + 1];
    int k = c1 - c0;
    if (k > 0) {
        // _ϵ = cumprod([1; ϵ_c[1]...]) then clamp >1 to 1 (model.jl:415-419)
        double E[BLG_MAXK + 1];
        double raw = 1.0;
        bool poison = false;
        E[0] = 1.0;
        const bool have_mo = tb.ipow != nullptr;
        for (int i = 0; i < k; ++i) {          // prefixes in MERGE order: what the merge coefficients use
            raw *= m.eps0[have_mo ? tb.mo[i] : m.child_idx[c0 + i]];
            E[i + 1] = raw > 1.0 ? 1.0 : raw;
        }
        raw = 1.0;
        for (int i = 0; i < k; ++i) {          // prefixes in the REFERENCE's child order decide the NaN rule:
            raw *= m.eps0[m.child_idx[c0 + i]];
            double v = raw > 1.0 ? 1.0 : raw;
            // the reference divides by (1-_ϵ)^n in log space after every child: _ϵ == 1 gives 0·(-Inf) = NaN
            // at n = 0 (model.jl:433-435,453-455); a prefix outside [0,1) is treated the same way here.
            if (!(v >= 0.0 && v < 1.0)) poison = true;
        }
        double inv = 1.0 / (1.0 - E[k]);
        if (tid == 0) {
            m.nk[n].inv = inv;
            m.nk[n].ip0raw = poison ? NaN : 1.0;
        }
        const int nrow = tb.nrow;              // fast-shard bound of this column + 1 (0: no fast families)
        if (nrow - 1 > BLG_FAST_MAXCOUNT) ok = false;
        // the fast merge scales the "t" side by E^-t (E: prefix product; in a binary merge the e of whichever child is
        // the "t" side in a warp) and the register-tiled side by e^-n (that child's own e): both must stay inside fp64
        // over the count bound of this node
        for (int i = 1; i < k; ++i)
            if (!poison && !(E[i] > 0.0 && -(double)nrow * log(E[i]) < 600.0)) ok = false;
        if (k >= 2) {
            for (int i = 0; i < k; ++i) {
                const double ei = m.eps0[m.child_idx[c0 + i]];
                if (!poison && !(ei > 0.0 && ei <= 1.0 && -(double)nrow * log(ei) < 600.0)) ok = false;
            }
        }
        if (tb.ipow != nullptr) {
            for (int r = tid; r < nrow; r += nt) tb.ipow[r] = poison ? NaN : d_powi(inv, r);
            for (int i = 1; i < k; ++i) {
                double* cf = tb.coefD[i];
                if (cf == nullptr) continue;
                int cs = tb.cdp[i];
                int smax = m.tab[tb.mo[i]].xfast;      // fast-shard bound of that child
                double p = E[i];
                for (int idx = tid; idx < nrow * cs; idx += nt) {
                    int t = idx / cs, s = idx - t * cs;
                    int r = t + s;
                    cf[idx] = (r < nrow && s <= smax) ? m.pascal[(size_t)r * m.pdim + s] * d_powi(p, s) : 0.0;
                }
            }
        }
    } else if (tid == 0) {
        m.nk[n].inv = 1.0;
        m.nk[n].ip0raw = 1.0;
    }
    if (tid == 0) m.fast_ok[n] = ok ? 1 : 0;
    if (kd == BLG_ROOT && m.rootw != nullptr) d_root_tables(m, m.rootw, m.rootlw, m.rootw_n, tid, nt);   // ϵ of the root's children is final here
    // ---- last block: NaN rule over the tree, fast-path summary ------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last_block = atomicAdd(m.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_block) {
        __threadfence();
        // everything the serial fold below touches goes to shared memory first (in parallel): the fold itself is a chain
        // of dependent reads, one L2 round trip each if left in global memory
        int* pz = reinterpret_cast<int*>(sm);          // [n] NaN flag of the subtree
        int* sok = pz + m.n;                           // [n] range flag
        int* soff = sok + m.n;                         // [n + 1] CSR of the children
        int* sidx = soff + m.n + 1;                    // [n]
        int* spost = sidx + m.n;                       // [npost]
        for (int i = tid; i < m.n; i += nt) {
            pz[i] = isnan(__ldcg(&m.nk[i].ip0raw)) ? 1 : 0;
            sok[i] = __ldcg(&m.fast_ok[i]);
            sidx[i] = m.child_idx[i];                  // (a tree has n - 1 child entries; the array holds at least n)
        }
        for (int i = tid; i <= m.n; i += nt) soff[i] = m.child_off[i];
        for (int i = tid; i < m.npost; i += nt) spost[i] = m.post[i];
        __syncthreads();
        if (tid == 0) {
            int all_ok = 1;
            for (int li = 0; li < m.npost; ++li) {
                const int u = spost[li];
                int p = (soff[u + 1] > soff[u]) ? pz[u] : 0;
                for (int ci = soff[u]; ci < soff[u + 1]; ++ci) p |= pz[sidx[ci]];
                pz[u] = p;
                if (sok[u] == 0) all_ok = 0;
            }
            m.flags[0] = all_ok;
            *m.ticket = 0u;
        }
        __syncthreads();
        for (int li = tid; li < m.npost; li += nt) { const int u = spost[li]; m.nk[u].ip0 = pz[u] ? NaN : 1.0; }
    }
}

// Pascal triangle C(n,s), n,s < dim, by the addition recurrence (exact while < 2^53).
__global__ void pascal_kernel(double* P, int dim) {
    extern __shared__ double sm[];
    double* prev = sm;
    double* cur = sm + dim;
    int tid = threadIdx.x, nt = blockDim.x;
    for (int s = tid; s < dim; s += nt) { prev[s] = (s == 0) ? 1.0 : 0.0; P[s] = prev[s]; }
    __syncthreads();
    for (int n = 1; n < dim; ++n) {
        for (int s = tid; s < dim; s += nt) {
            double v = (s == 0) ? 1.0 : (s <= n ? prev[s - 1] + prev[s] : 0.0);
            cur[s] = v;
            P[(size_t)n * dim + s] = v;
        }
        __syncthreads();
        double* t = prev; prev = cur; cur = t;
    }
}

// ------------------------------------------------------------------------------------------------
// pruning step at one internal node for every family (csuros_miklos!, model.jl:402-463)
// ------------------------------------------------------------------------------------------------
struct ChildRef {
    const double* ell;     // child's slab [n][f] (nullptr: leaf, one-hot at its count)
    const int* scl;        // child's per-family exponent (nullptr for leaves)
    const uint16_t* cnt;   // child's extended counts
    const double* Wt;      // child's transposed w*: Wt[j*wdp + s]
    int wdp;
    int id;                // child's node index (for ϵ[1])
    const double* coefD;   // merge coefficients for this merge position (i >= 1), [t*cdp + s]
    int cdp;
};
struct NodeStep {
    int k;
    int nrow;              // xmax_u + 1
    ChildRef child[BLG_MAXK];   // in MERGE order
    const double* ipow;
    double* out;           // node slab [n][f]
    int* out_scl;
    const uint16_t* cnt;   // node's own extended counts
};

// Linear-space form of the recursion.  With ℓ_c the child's column, e_i = ϵ_{c_i}[1],
// E_i = min(1, e_1⋯e_i):
//   b_i[s]   = Σ_{j>=s} W_{c_i}[s,j] ℓ_{c_i}[j]                                   (model.jl:425-426)
//   B_i[0,s] = b_i[s];  B_i[t,s] = B_i[t-1,s+1] + e_i B_i[t-1,s]                   (model.jl:428-431)
//   Ã_1[n]   = b_1[n];  Ã_i[n] = Σ_{t+s=n} C(n,s) E_{i-1}^s Ã_{i-1}[t] B_i[t,s]   (model.jl:438-452)
//   ℓ_u[n]   = Ã_k[n] / (1-E_k)^n                                                  (model.jl:453-461)
// The reference normalises by (1-E_i)^n after every child and multiplies the next merge by the
// Binomial(n, E_i) pmf; the (1-E_i) powers cancel exactly, so only the last division is kept.
// The children of a node may be merged in any order (ℓ_u does not depend on it); the host puts
// the child with the smaller counts first.

// ---- the sweep kernel.  ONE launch walks a list of nodes (a full post-order sweep, or the path
// node→root of a partial recompute).  A warp owns 32 consecutive families for the whole list and never
// synchronises with another warp: a parent only reads what the same thread wrote.
//   * step descriptors travel as a __grid_constant__ kernel parameter (constant bank: uniform,
//     cache-resident reads that streaming data cannot evict);
//   * the per-family working column lives in a per-warp shared-memory slice laid out [index][lane]
//     (conflict-free 64-bit accesses); a node needs S+3 doubles per family because both inner loops
//     run IN PLACE;
//   * NO coefficient tables are read in the inner loops and no register is moved there:
//       contraction over an ordinary branch: Horner's rule on the recurrence of wstar! (node.jl:181-192),
//           r[n] ← ψ'·r[n] + r[n-1]  for j = M..1 (r[0] := ℓ[j]),   b[n] = c^n·r[n],   c = (1-ϕ')(1-ψ')
//         — one DFMA per element, TS accumulators in registers, tiles chained through the column itself
//         (sweep_contract_horner);
//       merge: C(n,s)·E^s = n!·E^n · 1/(t!·E^t) · 1/s!  (t+s = n), hence with a'[t] = A[t]/(t!E^t) and
//         B‡[t,s] = B[t,s]·e^-(t+s)/s!:   B‡[t,s] = (s+1)·B‡[t-1,s+1] + B‡[t-1,s]      (one DFMA, immediate multiplier)
//                                          acc[k] ← acc[k+1] + a'[t+1]·B‡[t+1,k]         (the slide is the FMA's addend)
//                                          Ã[n]   = n!E^n · e^n · acc[0] at step n
//         with the B‡ row and the anti-diagonal accumulators of the smaller child in registers (sweep_merge_fast);
//     the branch below a WGD/WGT keeps its (banded) table, read from global memory;
//   * the scalings (1/t!E^t, e^-n) limit this kernel to count bounds <= BLG_FAST_MAXCOUNT and to parameter
//     ranges where they stay inside fp64 (checked per node on the device when ε/W are built);
//     otherwise the per-node general kernels above are used;
//   * the code is kept SMALL on purpose (rolled loops, noinline helpers, one merge dispatch): with 16 warps per SM
//     at different places of the sweep, the kernel's speed follows its instruction-cache footprint more closely than
//     anything else (80 KB of SASS: 12.0 ms per benchmark sweep; 54 KB: 2.7 ms; 41 KB: 2.3 ms — same arithmetic);
//   * no helper may use so many registers that the kernel's loop invariants get spilled around the calls (the widest
//     merge is 20 for that reason): a spill reload here is an L2 round trip.
__constant__ double c_invfact[BLG_FAST_MAXCOUNT + 72];   // 1/n!
__constant__ double c_inv[BLG_MAXCOUNT + 9];             // 1/n (c_inv[0] = 0)

struct SweepChild {
    const double* ell;     // nullptr: leaf
    const int* scl;
    const uint16_t* cnt;
    const double* Wt;      // transposed w* of the child's branch (leaf gather, WGD/WGT-after contraction)
    int wdp, table;        // table != 0: contract with Wt (branch below a WGD/WGT)
    int node, pad_;        // node index of the child: its scalars {e, 1/min(e,1), ψ', c} are read from NodeK on the device
};
struct SweepStep {
    int k, first, resident, node;      // resident != 0: child[0] is the node of the previous step (its column is still in shared memory)
                                       // node: index of the node itself: {1/(1-E_k), NaN flag} are read from NodeK on the device
    double* out;
    int* out_scl;
    const uint16_t* cnt;
};
#define BLG_SWEEP_MAXSTEPS 224
#define BLG_SWEEP_MAXCHILD 400
struct SweepList {
    int nsteps;
    int pad_[3];
    SweepStep step[BLG_SWEEP_MAXSTEPS];
    SweepChild child[BLG_SWEEP_MAXCHILD];
};
static_assert(sizeof(SweepList) <= 32000, "kernel parameter space");

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
// asynchronous global→shared copy of one double (zero-filled when !take), tracked by cp.async groups
__device__ __forceinline__ void cp_async_f64(double* smem_dst, const double* gsrc, bool take) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = take ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_b32(int* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_b128(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// How the helpers below address a working column.  GCol: a generic pointer (the per-family sweep kernel, whose column may
// live in shared memory or in global scratch).  SCol: an index into the kernel's dynamic shared memory — the compiler then
// knows the address space and emits LDS/STS with 32-bit address arithmetic instead of generic LD/ST (the tile kernel, whose
// tiles are cut on the host so that they always fit their shared slice).
extern __shared__ double blg_smem[];
struct GCol {
    double* p;
    __device__ __forceinline__ double& operator[](int i) const { return p[i]; }
    __device__ __forceinline__ GCol operator+(int k) const { return GCol{p + k}; }
    __device__ __forceinline__ double* ptr(int i) const { return p + i; }
};
struct SCol {
    int off;
    __device__ __forceinline__ double& operator[](int i) const { return blg_smem[off + i]; }
    __device__ __forceinline__ SCol operator+(int k) const { return SCol{off + k}; }
    __device__ __forceinline__ double* ptr(int i) const { return &blg_smem[off + i]; }
};

// in-place table contraction (branch below a WGD/WGT): col[s] ← Σ_{j=s..min(Mw, band·s)} Wt[j*wdp+s]·col[j]
// w*(s|j) of these branches is banded: s lineages above a WGD leave at most 2s below it (3s for a WGT), so rows j > band·s
// of column s are zero (band = 2 or 3; the literal WGT table of node.jl:205-218 keeps the same support).  Table rows are
// read 16 bytes at a time (wdp is a multiple of 16 and s0 of TS: aligned).
template <int TS, class C>
__device__ __noinline__ void sweep_contract_table(C col, int Mw, const double* __restrict__ Wt, int wdp, double osc, int band) {
#pragma unroll 1
    for (int s0 = 0; s0 <= Mw; s0 += TS) {
        double acc[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) acc[k] = 0.0;
        const int jhi = min(Mw, band * (s0 + TS - 1));
#pragma unroll 1
        for (int j = s0; j <= jhi; ++j) {
            const double x = col[j * 32];
            const double2* __restrict__ w = reinterpret_cast<const double2*>(Wt + (size_t)j * wdp + s0);
#pragma unroll
            for (int k = 0; k < TS / 2; ++k) {
                const double2 ww = __ldg(w + k);
                acc[2 * k] = fma(ww.x, x, acc[2 * k]);
                acc[2 * k + 1] = fma(ww.y, x, acc[2 * k + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < TS; ++k)
            if (s0 + k <= Mw) col[(s0 + k) * 32] = acc[k] * osc;
    }
}
// in-place contraction of the 32 columns of a warp with the w* TABLE of the branch on the fp64 tensor pipe:
//     out[s][pattern] = Σ_{j >= s} w*(s|j)·col[j][pattern]      — a triangular (Mw+1)×(Mw+1) by (Mw+1)×32 product,
// mma.sync.m8n8k4.f64: A = w* (row s, column j; the table is stored transposed, Wt[j·wdp + s]), B = the warp's slice
// (row j, column = lane/pattern), C = 8 output rows × 8 patterns per instruction (256 FMA).  Output rows are produced in
// ascending blocks of 8, so the product can overwrite its input: block s0 only reads rows >= s0.  One DMMA replaces 8
// DFMA warp instructions and their loop/shared-memory traffic; for columns of a few dozen counts the contraction is the
// densest part of a pruning step.  band: 0 for an ordinary branch (dense upper triangle), 2/3 below a WGD/WGT (rows
// j > band·s of column s are zero).
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <class C>
__device__ __noinline__ void sweep_contract_mma(C col, int Mw, const double* __restrict__ Wt, int wdp, int band) {
    const unsigned lane = threadIdx.x & 31u;
    const C cb = col + (-(int)lane);               // the warp's slice: element (j, pattern) at cb[j*32 + pattern]
    const int g = (int)(lane >> 2), t = (int)(lane & 3u);
#pragma unroll 1
    for (int s0 = 0; s0 <= Mw; s0 += 8) {
        double c[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { c[nb][0] = 0.0; c[nb][1] = 0.0; }
        const int jhi = band ? min(Mw, band * (s0 + 7)) : Mw;
#pragma unroll 2
        for (int j0 = s0; j0 <= jhi; j0 += 4) {
            const int j = j0 + t;
            const bool in = j <= Mw;
            const double a = in ? __ldg(Wt + (size_t)j * wdp + s0 + g) : 0.0;
            const C row = cb + ((in ? j : Mw) * 32 + g);
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const double b = in ? row[nb * 8] : 0.0;
                dmma_m8n8k4(c[nb][0], c[nb][1], a, b);
            }
        }
        __syncwarp();                              // every lane has read rows s0..s0+7 before they are overwritten
        const int s = s0 + g;
        if (s <= Mw) {
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<double2*>(cb.ptr(s * 32 + nb * 8 + 2 * t)) = make_double2(c[nb][0], c[nb][1]);
        }
        __syncwarp();
    }
}
// in-place contraction over an ordinary branch, col[j] = ℓ[j] → col[n] = b[n] (j, n = 0..Mw; b[0] = ℓ[0]).
// The columns of w* obey v_j = T·v_{j-1}, (T v)[n] = ψ'·v[n] + c·v[n-1], v_1 = c·δ_{n1} (wstar!, node.jl:181-192), so
// b = Σ_j ℓ[j]·T^(j-1)·v_1 is evaluated by Horner's rule from j = Mw down to 1.  With r[n] = (partial b[n])/c^n a stage is
//     r[n] ← ψ'·r[n] + r[n-1]   (n descending, r[0] := ℓ[j])
// — ONE DFMA per element, in place, the shift is the FMA's addend: no register moves, no coefficient, no factorial
// scaling (r is bounded by 2^Mw·max ℓ whatever the parameters are).  TS consecutive n live in registers; a tile hands the
// old value of its top element to the tile above through the column itself: at stage j tile `base` reads col[base+j]
// (ℓ[j] for the first tile) and leaves its top element in col[base+TS+j], a slot it has already consumed.
template <int TS, class C>
__device__ __noinline__ void sweep_contract_horner(C col, int Mw, double psi, double cc, double osc) {
    double cpow = cc * osc;                            // c^n (times the pending power-of-two rescale of a resident column) at the tile's first output
#pragma unroll 1
    for (int base = 0; base < Mw; base += TS) {        // outputs n = base+1 .. base+TS
        double r[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) r[k] = 0.0;
        const int lim = Mw - base - TS;                // stages j <= lim feed the next tile
#pragma unroll 1
        for (int j = Mw - base; j >= 1; --j) {
            const double h = col[(base + j) * 32];
            const double top = r[TS - 1];
#pragma unroll
            for (int k = TS - 1; k >= 1; --k) r[k] = fma(psi, r[k], r[k - 1]);
            r[0] = fma(psi, r[0], h);
            if (j <= lim) col[(base + TS + j) * 32] = top;
        }
#pragma unroll
        for (int k = 0; k < TS; ++k) {
            const int n = base + 1 + k;
            if (n <= Mw) col[n * 32] = r[k] * cpow;
            cpow *= cc;
        }
    }
}
// single-tile in-place merge (M2w < TS).  A[0..Sw] is the "t" side, B0[0..M2w] the "s" side (both inside the
// warp's slice); the result Σ_{t+s=n} C(n,s)·E^s·A[t]·B[t,s], n = 0..Sw+M2w, is written to O[n] at step n.
// O may alias the slice as long as O+n never passes an unread A[t] (O == A, or O below A).
template <int TS, class C>
__device__ __noinline__ double sweep_merge_fast(C O, C A, int Sw, C B0, int M2w, double E, double invE,
                                                double e, double inve, double em0, double emg, int X) {
    // em0/emg: the emitted value is acc·n!·E^n·(extra factor em0·emg^n): the last merge of a node passes
    // the (1-E_k)^(-n) normalisation (and the NaN poison) through them.  Returns max_{n<=X} of the result.
    // The row is kept as B‡[t,s] = B†[t,s]·e^-(t+s): its recurrence B‡[t,s] = (s+1)·B‡[t-1,s+1] + B‡[t-1,s] is ONE
    // DFMA with an immediate multiplier (the plain form needs a DMUL and a DFMA), and the factor e^(t+s) = e^n that
    // every term of output n is missing is applied when the output is emitted (as a separate factor: multiplied into
    // n!·E^n it could underflow on short branches, where e is tiny).  |B‡| <= 2^t·e^-M2w·max|b|; inve = 1/e; the range
    // of e^-n over the count bound is part of the per-node range check (tables_kernel).
    // The accumulators slide by one output per step; that shift is fused into the accumulation of the NEXT step's terms,
    // acc[k] ← acc[k+1] + a[t+1]·B‡[t+1,k] in ascending k (in place: one DFMA, no register moves); once the "t" side is
    // exhausted the same statement runs with a = 0 and only drains the accumulators.
    double B[TS], acc[TS];
    double tf = 1.0;       // 1/(t!·E^t)
    {
        const double a0 = A[0];
        double ek = 1.0;
#pragma unroll
        for (int k = 0; k < TS; ++k) {
            B[k] = (k <= M2w) ? B0[k * 32] * (c_invfact[k] * ek) : 0.0;
            acc[k] = a0 * B[k];
            ek *= inve;
        }
    }
    const int nmax = Sw + M2w;
    const double Eg = E * emg;
    double em = em0;       // n!·E^n · em0·emg^n
    double ee = 1.0;       // e^n
    double mx = 0.0;
#pragma unroll 1
    for (int t = 0; t <= nmax; ++t) {
        const double v = (t <= X) ? (acc[0] * ee) * em : 0.0;
        O[t * 32] = v;
        mx = fmax(mx, v);                         // (a NaN column leaves mx = 0: stored as it is, exponent 0)
        em *= Eg * (double)(t + 1);
        ee *= e;
        double a = 0.0;
        if (t < Sw) {
#pragma unroll
            for (int k = 0; k < TS - 1; ++k) B[k] = fma((double)(k + 1), B[k + 1], B[k]);
            tf *= invE * c_inv[t + 1];
            a = A[(t + 1) * 32] * tf;
        }
#pragma unroll
        for (int k = 0; k < TS - 1; ++k) acc[k] = fma(a, B[k], acc[k + 1]);
        acc[TS - 1] = a * B[TS - 1];
    }
    return mx;
}
// general merge for a smaller child with >= 20 counts: 16-wide tiles from high s to low s chained
// through the history column H.  A[0..Sw]; O[0..M2w] = B[0,·], O[M2w+1..Sw+M2w] = 0 on entry; result in O.
template <class C>
__device__ __noinline__ void sweep_merge_tiled(C A, int Sw, C O, int M2w, C H, double E, double e) {
    constexpr int TS = 16;
    const int ntile = M2w / TS + 1;
    const int nmax = Sw + M2w;
    const double invE = 1.0 / E;
#pragma unroll 1
    for (int tile = ntile - 1; tile >= 0; --tile) {
        const int s0 = tile * TS;
        const bool top = (tile == ntile - 1);
        double B[TS], acc[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) {
            B[k] = (s0 + k <= M2w) ? O[(s0 + k) * 32] * c_invfact[s0 + k] : 0.0;
            acc[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < TS; ++k)
            if (s0 + k <= M2w) O[(s0 + k) * 32] = 0.0;
        double tf = 1.0;
#pragma unroll 1
        for (int t = 0; t <= Sw + TS - 1; ++t) {
            if (t <= Sw) {
                const double a = A[t * 32] * tf;
                const double halo = top ? 0.0 : H[t * 32];
                H[t * 32] = B[0];
#pragma unroll
                for (int k = 0; k < TS; ++k) acc[k] = fma(a, B[k], acc[k]);
#pragma unroll
                for (int k = 0; k < TS - 1; ++k) B[k] = fma(e, B[k], (double)(s0 + k + 1) * B[k + 1]);
                B[TS - 1] = fma(e, B[TS - 1], (double)(s0 + TS) * halo);
                tf *= invE * c_inv[t + 1];
            }
            const int n = t + s0;
            if (n <= nmax) O[n * 32] += acc[0];
#pragma unroll
            for (int k = 0; k < TS - 1; ++k) acc[k] = acc[k + 1];
            acc[TS - 1] = 0.0;
        }
    }
    double em = 1.0;
#pragma unroll 1
    for (int n = 0; n <= nmax; ++n) { O[n * 32] *= em; em *= E * (double)(n + 1); }
}

// stage (unless the column is already resident in col) + contract one child into col[0..Mw]
// have: 0 = load the child's input column; 1 = it is in col already, left there by the previous step and still to be
// multiplied by the power of two osc (applied to the outputs); 2 = it is in col already (copied by cp.async)
// (the descriptor fields travel by value: read through a reference, inside a noinline function, they were generic loads
// from the parameter bank that missed the carved-down L1 — a long-scoreboard stall at the top of every call)
struct ChildArgs { const double* ell; const double* Wt; int wdp, table; double psi, cc; };
__device__ __forceinline__ ChildArgs child_args(const SweepChild& c, double psi, double cc) { return ChildArgs{c.ell, c.Wt, c.wdp, c.table, psi, cc}; }
template <class C>
__device__ __noinline__ void sweep_child(const ChildArgs c, int f, size_t ld, int M, int Mw, C col, int have, double osc) {
    const bool resident = have != 0;
    if (c.ell == nullptr) {                        // leaf: column x_c of w* (one-hot likelihood)
        if (have) return;
        const double* __restrict__ row = c.Wt + (size_t)M * c.wdp;
#pragma unroll 1
        for (int s = 0; s <= Mw; ++s) col[s * 32] = (s <= M) ? row[s] : 0.0;
        return;
    }
    if (!resident) {
        const double* __restrict__ src = c.ell + f;
#pragma unroll 1
        for (int j = 0; j <= Mw; ++j) col[j * 32] = (j <= M) ? src[(size_t)j * ld] : 0.0;
    }
    if (Mw >= BLG_MMA_MIN && c.Wt != nullptr) {   // a few dozen counts: the table product on the tensor pipe
        if (have == 1) {                           // (the pending power of two of a resident column: exact)
#pragma unroll 1
            for (int j = 0; j <= Mw; ++j) col[j * 32] *= osc;
        }
        __syncwarp();
        sweep_contract_mma(col, Mw, c.Wt, c.wdp, c.table);
    } else if (c.table) {
        sweep_contract_table<16, C>(col, Mw, c.Wt, c.wdp, osc, c.table);
    } else {
        const double psi = c.psi, cc = c.cc;
        if (have == 1) col[0] *= osc;              // b[0] = ℓ[0]
        if (Mw <= 4) sweep_contract_horner<4, C>(col, Mw, psi, cc, osc);
        else if (Mw <= 8) sweep_contract_horner<8, C>(col, Mw, psi, cc, osc);
        else sweep_contract_horner<16, C>(col, Mw, psi, cc, osc);
    }
}

// PROF: per-(step, phase) warp-cycle totals into prof[step*8 + phase] (diagnostic build of the same kernel)
#define BLG_TICK(ph) do { if (PROF) { long long t1_ = clock64(); if (lane == 0) atomicAdd(&prof[si * 8 + (ph)], (unsigned long long)(t1_ - t0_)); t0_ = t1_; } } while (0)
template <bool PROF, int MINB, bool PERSIST>
__global__ void __launch_bounds__(128, MINB) sweep_kernel(const __grid_constant__ SweepList L, const NodeK* __restrict__ nk,
                                                       const int* __restrict__ flags, unsigned int* __restrict__ ran,
                                                       int N, size_t ld, int slice_cap,
                                                       double* gscratch, size_t gs_stride, unsigned long long* prof,
                                                       unsigned int* tile_counter) {
    extern __shared__ double smem[];
    if (flags[0] == 0) return;        // some node is outside the table-free forms' range: the general kernel (launched next) does the work
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ran[0], 1u);
    const unsigned lane = threadIdx.x & 31u;
    const int warp = threadIdx.x >> 5;
    // [warps][2 buffers][5 slots][32 lanes] ints: counts and exponents of the coming steps land here by cp.async, in the
    // buffer of the step's parity.  Counts (static data) are requested TWO steps ahead, so that a run of empty steps
    // — each far shorter than a trip to HBM — does not wait for every single one; exponents (written by earlier steps of
    // this very launch) one step ahead.  Step j commits one group {counts(j+2), scalars(j+2), exponents(j+1)}.
    // [warps][2 buffers][10] doubles behind them: the per-node scalars of the step (NodeK, device-resident: built by
    // tables_kernel, never seen by the host): child 0 {e, 1/min(e,1), ψ', c}, child 1 the same, the node {1/(1-E_k), NaN flag}.
    // (three scalar buffers, indexed by step mod 3: the request for step s+2 must not land on the scalars of step s,
    // which are read throughout the step)
    int* cb = reinterpret_cast<int*>(smem) + (size_t)warp * 320 + lane;
    double* sb = smem + (size_t)(blockDim.x >> 5) * 160 + (size_t)warp * 30;
    double* sl = smem + (size_t)(blockDim.x >> 5) * 190 + (size_t)warp * slice_cap * 32 + lane;
    const int ntiles = (N + 31) / 32;
    // persistent grid, warp-granular work stealing: tiles (32 families, heaviest first) are handed out by an atomic
    // counter, so a batch that is a little more than one wave does not pay for a nearly empty second wave
    // (PERSIST == false: one tile per warp, grid sized to the batch — better for many waves)
    for (int round = 0; PERSIST || round == 0; ++round) {
    int wg = blockIdx.x * (blockDim.x >> 5) + warp;
    if (PERSIST) {
        if (lane == 0) wg = (int)atomicAdd(tile_counter, 1u);
        wg = __shfl_sync(0xffffffffu, wg, 0);
    }
    if (wg >= ntiles) break;
    int f = wg * 32 + (int)lane;
    const bool live = f < N;
    if (!live) f = N - 1;            // idle lanes shadow the last family (loads stay in bounds; they never store)
    double* gs = gscratch ? gscratch + (size_t)wg * gs_stride + lane : nullptr;
    const int fe = f & ~1;                 // the aligned 4-byte word holding this lane's uint16 count
    const int fsh = (f & 1) * 16;

    int prevX = 0, prevScl = 0;      // count and exponent of the column left in the slice by the previous step
    double prevSc = 1.0;             // ... which is still to be multiplied by this power of two (its stored copy already is)
    double* prevR = sl;
    long long t0_ = PROF ? clock64() : 0;
    auto request_counts = [&](int s) {
        if (s < L.nsteps) {
            const SweepStep& sn = L.step[s];
            const SweepChild* cn = &L.child[sn.first];
            int* b = cb + (s & 1) * 160;
            cp_async_b32(b + 0 * 32, sn.cnt + fe);
            if (!sn.resident) cp_async_b32(b + 1 * 32, cn[0].cnt + fe);
            if (sn.k >= 2) cp_async_b32(b + 2 * 32, cn[1].cnt + fe);
            // the step's scalars, 16 bytes per lane: lanes 0,1 child 0; lanes 2,3 child 1; lane 4 the node itself
            if (lane < 5) {
                const int ci = (int)(lane >> 1);
                double* d = sb + (s % 3) * 10;
                if (lane == 4) cp_async_b128(d + 8, &nk[sn.node].inv);
                else if (ci < sn.k) cp_async_b128(d + lane * 2, &nk[cn[ci].node].e + (lane & 1) * 2);
            }
        }
    };
    auto request_exponents = [&](int s) {
        if (s < L.nsteps) {
            const SweepStep& sn = L.step[s];
            const SweepChild* cn = &L.child[sn.first];
            int* b = cb + (s & 1) * 160;
            if (!sn.resident && cn[0].scl) cp_async_b32(b + 3 * 32, cn[0].scl + f);
            if (sn.k >= 2 && cn[1].scl) cp_async_b32(b + 4 * 32, cn[1].scl + f);
        }
    };
    cp_async_wait_all();
    request_counts(0);
    cp_async_commit();
    request_counts(1);
    request_exponents(0);
    cp_async_commit();
    for (int si = 0; si < L.nsteps; ++si) {
        const SweepStep& st = L.step[si];
        const int k = st.k;
        const SweepChild* ch = &L.child[st.first];
        const bool res0 = st.resident != 0;
        const int* b = cb + (si & 1) * 160;
        const double* sc_ = sb + (si % 3) * 10;        // {e, 1/min(e,1), ψ', c} of child 0, of child 1, {1/(1-E_k), NaN flag} of the node
        cp_async_wait_but_one();                       // this step's counts and scalars are here; its exponents may still be on their way
        const int X = (b[0 * 32] >> fsh) & 0xffff;
        const int Xw = warp_max(X);
        int M0 = 0, M1 = 0, sclsum = 0;
        if (Xw != 0) {
            cp_async_wait_all();
            M0 = res0 ? prevX : ((b[1 * 32] >> fsh) & 0xffff);
            M1 = (k >= 2) ? ((b[2 * 32] >> fsh) & 0xffff) : 0;
            sclsum = (res0 ? prevScl : (ch[0].scl ? b[3 * 32] : 0)) + ((k >= 2 && ch[1].scl) ? b[4 * 32] : 0);
        }
        request_counts(si + 2);                        // into the buffer just read
        request_exponents(si + 1);
        cp_async_commit();
        if (Xw == 0) {
            // no family of this warp has a gene below the node: ℓ = [1] whatever the parameters are (n = 0 lineages
            // leave no descendants with probability 1) — or the reference's NaN, folded over the subtree by tables_kernel
            const double v = sc_[9];
            if (live) { st.out[f] = v; st.out_scl[f] = 0; }
            sl[0] = v;
            prevX = 0;
            prevScl = 0;
            prevSc = 1.0;
            prevR = sl;
            __syncwarp();
            BLG_TICK(0);
            continue;
        }
        const int M0w = warp_max(M0);
        const int M1w = warp_max(M1);
        int sumw = M0w + M1w;
        // which side of a binary merge is register-tiled: the child with the smaller bound in THIS warp
        const bool swap = (k == 2) && (M0w < M1w);
        bool tiled = (k == 2) ? (min(M0w, M1w) >= 20) : (M1w >= 20);
#pragma unroll 1
        for (int i = 2; i < k; ++i) {
            const int Miw = warp_max((int)ch[i].cnt[f]);
            sumw += Miw;
            if (Miw >= 20) tiled = true;
            if (ch[i].scl) sclsum += ch[i].scl[f];
        }
        const int need = tiled ? 3 * (sumw + 1) : sumw + k + 1;
        double* R = (need <= slice_cap) ? sl : gs;
        if (res0 && R != prevR) {
#pragma unroll 1
            for (int n = 0; n <= M0w; ++n) R[n * 32] = prevR[n * 32];
        }
        BLG_TICK(0);

        // child 1's input column is copied asynchronously while child 0 is being contracted
        const bool async1 = (k >= 2) && !tiled && R == sl;
        if (async1) {
            double* dst = R + (size_t)(M0w + 1) * 32;
            if (ch[1].ell == nullptr) {
                const double* row = ch[1].Wt + (size_t)M1 * ch[1].wdp;
#pragma unroll 1
                for (int s = 0; s <= M1w; ++s) cp_async_f64(dst + s * 32, (s <= M1) ? row + s : row, s <= M1);
            } else {
                const double* src = ch[1].ell + f;
#pragma unroll 1
                for (int j = 0; j <= M1w; ++j) cp_async_f64(dst + j * 32, (j <= M1) ? src + (size_t)j * ld : src, j <= M1);
            }
            cp_async_commit();
        }
        sweep_child(child_args(ch[0], sc_[2], sc_[3]), f, ld, M0, M0w, GCol{R}, res0 ? 1 : 0, res0 ? prevSc : 1.0);
        if (async1) cp_async_wait_all();
        BLG_TICK(1);
        int Sw = M0w;
        double mxv = 0.0;
        bool scaled = false;        // has the (1-E_k)^(-n) normalisation already been applied (and mxv taken)?
        double Eraw = sc_[0];       // product of the e's of the children merged so far (polytomies: clamped when used)
        if (!tiled) {
#pragma unroll 1
            for (int i = 1; i < k; ++i) {
                const int Mi = (i == 1) ? M1 : (int)ch[i].cnt[f];
                const int Miw = (i == 1) ? M1w : warp_max(Mi);
                double* colB = R + (size_t)(Sw + 1) * 32;
                // scalars of child i: staged for i == 1, straight from NodeK for the further children of a polytomy
                const double ei = (i == 1) ? sc_[4] : nk[ch[i].node].e;
                const double invei = (i == 1) ? sc_[5] : nk[ch[i].node].inve;
                sweep_child(child_args(ch[i], (i == 1) ? sc_[6] : nk[ch[i].node].psi, (i == 1) ? sc_[7] : nk[ch[i].node].cc), f, ld, Mi, Miw, GCol{colB},
                            (i == 1 && async1) ? 2 : 0, 1.0);
                BLG_TICK(2);
                const bool last = (i == k - 1);
                const double em0 = last ? sc_[9] : 1.0, emg = last ? sc_[8] : 1.0;
                const int Xm = last ? X : 0x7fffffff;
                // the register-tiled "s" side is child i — or, in a binary merge whose child 0 has the smaller bound in
                // this warp, child 0 (already at the front), with child i as the "t" side
                const double* tA = swap ? colB : R;
                const double* sB = swap ? R : colB;
                const int tW = swap ? Miw : Sw, sW = swap ? Sw : Miw;
                const double Epre = fmin(Eraw, 1.0);              // clamped prefix of the children merged before this one
                const double E = swap ? fmin(ei, 1.0) : Epre;
                const double e = swap ? sc_[0] : ei;
                const double invE = swap ? invei : ((k == 2) ? sc_[1] : 1.0 / E);
                const double inve = swap ? sc_[1] : invei;
                Eraw *= ei;
                double mx;
                if (sW < 4) mx = sweep_merge_fast<4>(GCol{R}, GCol{const_cast<double*>(tA)}, tW, GCol{const_cast<double*>(sB)}, sW, E, invE, e, inve, em0, emg, Xm);
                else if (sW < 8) mx = sweep_merge_fast<8>(GCol{R}, GCol{const_cast<double*>(tA)}, tW, GCol{const_cast<double*>(sB)}, sW, E, invE, e, inve, em0, emg, Xm);
                else if (sW < 12) mx = sweep_merge_fast<12>(GCol{R}, GCol{const_cast<double*>(tA)}, tW, GCol{const_cast<double*>(sB)}, sW, E, invE, e, inve, em0, emg, Xm);
                else mx = sweep_merge_fast<20>(GCol{R}, GCol{const_cast<double*>(tA)}, tW, GCol{const_cast<double*>(sB)}, sW, E, invE, e, inve, em0, emg, Xm);
                Sw += Miw;
                if (last) { mxv = mx; scaled = true; }
                BLG_TICK(3);
            }
        } else {
            const int cap = sumw + 1;
            double* rA = R;
            double* rH = R + (size_t)cap * 32;
            double* rO = R + (size_t)2 * cap * 32;
#pragma unroll 1
            for (int i = 1; i < k; ++i) {
                const int Mi = (i == 1) ? M1 : (int)ch[i].cnt[f];
                const int Miw = (i == 1) ? M1w : warp_max(Mi);
                const double ei = (i == 1) ? sc_[4] : nk[ch[i].node].e;
                sweep_child(child_args(ch[i], (i == 1) ? sc_[6] : nk[ch[i].node].psi, (i == 1) ? sc_[7] : nk[ch[i].node].cc), f, ld, Mi, Miw, GCol{rO}, 0, 1.0);
#pragma unroll 1
                for (int n = Miw + 1; n <= Sw + Miw; ++n) rO[n * 32] = 0.0;
                sweep_merge_tiled(GCol{rA}, Sw, GCol{rO}, Miw, GCol{rH}, fmin(Eraw, 1.0), ei);
                Eraw *= ei;
                Sw += Miw;
#pragma unroll 1
                for (int n = 0; n <= Sw; ++n) rA[n * 32] = rO[n * 32];     // result back to the front of R
                BLG_TICK(4);
            }
        }
        // ℓ_u[n] = Ã[n]·(1-E_k)^(-n), rescale to a power-of-two exponent, store; the column also stays
        // in R[0..Xw] (zero above the family's own count) for a parent that follows immediately
        if (scaled) {
#pragma unroll 1
            for (int n = Sw + 1; n <= Xw; ++n) R[n * 32] = 0.0;
        } else {
            double ip = sc_[9];
            const double inv = sc_[8];
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
        BLG_TICK(5);
        // exponent of the column maximum straight from the bit pattern (normal numbers; a subnormal
        // maximum keeps exponent 0, which only costs range, not correctness)
        int ex = 0;
        if (mxv > 0.0) {                               // (fmax skips NaNs; an infinite maximum keeps exponent 0)
            const int be = (__double2hiint(mxv) >> 20) & 0x7ff;
            if (be != 0 && be != 0x7ff) ex = be - 1023;
        }
        // the stored column is rescaled; the copy in R stays as it is — a parent that follows immediately applies the
        // same power of two to the outputs of its contraction (exact, so both routes give the same bits)
        const double sc = __hiloint2double((1023 - ex) << 20, 0);         // 2^-ex, exact (|ex| <= 1022)
        if (live) {
            double* o = st.out + f;
            int n = 0;
#pragma unroll 1
            for (; n < X; n += 2) {                    // two loads in flight
                const double v0 = R[n * 32], v1 = R[(n + 1) * 32];
                o[(size_t)n * ld] = v0 * sc;
                o[(size_t)(n + 1) * ld] = v1 * sc;
            }
            if (n == X) o[(size_t)n * ld] = R[n * 32] * sc;
            st.out_scl[f] = sclsum + ex;
        }
        prevSc = sc;
        prevX = X;
        prevScl = sclsum + ex;
        prevR = R;
        __syncwarp();
        BLG_TICK(6);
    }
    }   // next tile
}

// eight independent DFMA chains per thread: what the fp64 pipe sustains (blg_fp64_peak)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999 + 1e-9 * threadIdx.x, c = 1e-7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) *sink = r;           // never true: keeps the chains alive
}

// ------------------------------------------------------------------------------------------------
// root: integrate_root (model.jl:465-475), condition_oib (model.jl:492-501), weighted reduction
// ------------------------------------------------------------------------------------------------
__global__ void root_tables_kernel(ModelDev m, double* rootw, double* rootlw, int nrow) { d_root_tables(m, rootw, rootlw, nrow, (int)threadIdx.x, (int)blockDim.x); }

__global__ void __launch_bounds__(256) root_kernel(const double* __restrict__ ell, const int* __restrict__ scl,
                                                   const uint16_t* __restrict__ cnt, const double* __restrict__ rootw,
                                                   const double* __restrict__ weight, double* __restrict__ fam_ll,
                                                   double* __restrict__ partial, int N, size_t ld,
                                                   unsigned int* __restrict__ ticket, unsigned int nblocks_total,
                                                   const int* __restrict__ flags, double* __restrict__ marker, double* __restrict__ out) {
    __shared__ double red[256];
    __shared__ bool last_block;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double contrib = 0.0;
    if (f < N) {
        int X = cnt[f];
        double p = 0.0;
        for (int n = 1; n <= X; ++n) p = fma(ell[(size_t)n * ld + f], rootw[n], p);
        double l = log(p) + (double)scl[f] * 0.6931471805599453 - rootw[0];
        if (!isfinite(l)) l = -INFINITY;                                        // model.jl:366
        fam_ll[f] = l;
        contrib = weight[f] * l;
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
    if (last_block) {                       // the last block of the call's root launches (this one for the fast shard,
        __threadfence();                    // root_ragged_kernel for the heavy shard) adds the partial sums, in a fixed order
        const volatile double* vp = partial;
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)nblocks_total; i += 256) s += vp[i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int k = 128; k > 0; k >>= 1) {
            if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
            __syncthreads();
        }
        if (threadIdx.x == 0) { out[0] = red[0]; marker[0] = flags[0] ? 0.0 : 1.0; *ticket = 0u; }
    }
}

__global__ void zero_slab_kernel(double* ell, int* scl, size_t nell, int N) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < nell; k += stride) ell[k] = 0.0;
    for (size_t k = i; k < (size_t)N; k += stride) scl[k] = 0;
}

#include "general.cuh"
#include "pattern.cuh"
#include "gradient.cuh"
#include "adjoint.cuh"
#include "padjoint.cuh"
#include "ingest.cuh"
#include "sampler.cuh"

// ------------------------------------------------------------------------------------------------
// host-side state
// ------------------------------------------------------------------------------------------------
struct Slab {
    double* ell = nullptr;
    int* scl = nullptr;
    size_t nd = 0;                   // doubles in ell
    bool pat = false;                // columns are node patterns (pattern.cuh), not families
};
struct Col {
    const uint16_t* cnt = nullptr;   // device counts of this column (shared between aliases)
    int xmax = 0;
    double colsum = 0.0;             // Σ_f (x_f + 1): the algorithmic byte count, and the size of a ragged slab
    std::shared_ptr<Slab> slab;      // null: never computed ("-Inf")
    int lmax = 0x7fffffff;           // fast set: entries above this count are left to the heavy family set (see set_profiles)
    const uint32_t* off = nullptr;   // ragged (heavy) shard: per-family offsets of the column, prefix sums of (count+1); shared between aliases
};
struct HostTab {                     // host mirror of NodeTab allocations
    int wdim = 0, wdp = 0, nrow = 0, k = 0, xfast = 0;
    int cdp[BLG_MAXK] = {0};
    int mo[BLG_MAXK] = {0};          // merge order (node indices)
    double* Wt = nullptr;
    double* ipow = nullptr;
    double* coefD[BLG_MAXK] = {nullptr};
    bool built = false;
};
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) blgFree(p); }           // (also on the exception paths of the entries that use local buffers)
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) blgFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(n, 16);
        cudaError_t e = blgMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = want;
    }
    void free_() { if (p) blgFree(p); p = nullptr; cap = 0; }
};

// One of the two family sets of a handle.  The FAST shard holds the families whose root count is at most
// BLG_FAST_MAXCOUNT: node-major SoA slabs ell[n][family] padded to the shard's per-node bound, k-d family order, walked by
// sweep_kernel (32 families per warp).  The HEAVY shard holds the rest: ragged slabs (family f's column starts at
// off[f], exactly x_f+1 doubles, no padding), heaviest family first, walked by gsweep_kernel (one warp per family).
// The distinct sub-patterns of one profile column (pattern.cuh): families with the same leaf counts below the node share
// one pattern.  Keyed by the column's device count array, so the alias columns extend! creates (a WGD above node i sees
// exactly node i's sub-patterns) share it.
struct Pat {
    int U = 0;                       // distinct patterns
    size_t ld = 0;                   // column stride, U rounded up to 32
    int xmax = 0;
    double colsum = 0.0;             // Σ_p (x_p + 1)
    // the order of the patterns: first those whose second-largest child count m2 is >= 36 (general kernel, one warp each),
    // then 20 <= m2 < 36 (wide tile kernel), then the rest, grouped by which child is the larger one and sorted by counts
    int nvery = 0, nwide = 0;
    uint16_t* d_cnt = nullptr;       // [ld] count of the column per pattern (sorted descending: heaviest tiles first)
    int* d_fam2pat = nullptr;        // [N] family → pattern, on the device only for the root
    double* d_pw = nullptr;          // [ld] weight of the families with the pattern (root, gradient only)
    std::vector<int> fam2pat;        // host
    std::vector<int> rep;            // a family of each pattern
    std::vector<uint16_t> h_cnt;
    std::vector<const Pat*> kids;    // the children's patterns these were built from (a link over exactly these needs no check)
    ~Pat() {
        if (d_cnt) blgFree(d_cnt);
        if (d_fam2pat) blgFree(d_fam2pat);
        if (d_pw) blgFree(d_pw);
    }
};
// pattern of a node → pattern of each child, for one (node column, ordered child columns) combination
struct Link {
    int* d_pidx[BLG_MAXK] = {nullptr};   // nullptr: identity (single child sharing the node's patterns)
    uint16_t* d_xcnt[BLG_MAXK] = {nullptr};   // the child's count per NODE pattern (saves the dependent load cnt_child[pidx[p]])
    // tiles cut on the host for the tile kernels (w = 0: normal range, w = 1: wide range of the node's pattern order): at most
    // 32 patterns, children's maxima summing to at most the slice, the smaller maximum below the kernel's widest register
    // merge — so that every tile runs out of shared memory on the register-tiled merge.  nullptr: fixed tiles of 32.
    int* d_tiles[2] = {nullptr, nullptr};
    int ntiles[2] = {0, 0};
    int maxneed[2] = {0, 0};             // largest slice need (doubles per lane) over the tiles
    bool cut = false;                    // tiles were cut (binary merges); otherwise the generic instantiation runs
    // the inverse maps (padjoint.cuh, built at the first gradient): per child pattern the node patterns that use it, as a
    // run of consecutive positions (ascending node-pattern order inside a run), and the child patterns for phase B's warp and block paths
    int* d_inv_off[BLG_MAXK] = {nullptr};
    int* d_inv_idx[BLG_MAXK] = {nullptr};    // node pattern → its place in that order
    int* d_hit[BLG_MAXK] = {nullptr};
    int* d_mid[BLG_MAXK] = {nullptr};
    int nhit[BLG_MAXK] = {0};
    int nmid[BLG_MAXK] = {0};
    int nmid_many[BLG_MAXK] = {0};
    std::map<int, std::pair<int*, int>> big;   // (msmin, T) → the node patterns of the adjoint's warp path (padj_is_big)
    ~Link() {
        for (int* q : d_pidx) if (q) blgFree(q);
        for (int* q : d_tiles) if (q) blgFree(q);
        for (uint16_t* q : d_xcnt) if (q) blgFree(q);
        for (int* q : d_inv_off) if (q) blgFree(q);
        for (int* q : d_inv_idx) if (q) blgFree(q);
        for (int* q : d_hit) if (q) blgFree(q);
        for (int* q : d_mid) if (q) blgFree(q);
        for (auto& b : big) if (b.second.first) blgFree(b.second.first);
    }
};

struct Shard {
    bool ragged = false;
    std::vector<uint16_t> h_counts;  // host copy of the count rows [column][ld] (fast family set: pattern construction)
    std::map<const uint16_t*, std::shared_ptr<Pat>> pats;
    std::map<std::vector<const void*>, std::shared_ptr<Link>> links;
    int N = 0;               // families
    size_t ld = 0;           // stride of the count rows (and of the SoA slabs), N rounded up to 128
    std::vector<Col> Tc, Tp; // committed / proposal column tables (by id-1)
    std::vector<uint16_t*> count_allocs;
    DevBuf<double> weight, fam_ll;
    std::map<size_t, std::vector<Slab*>> pool;   // free slabs by size
    std::vector<Slab*> all_slabs;
    std::vector<uint32_t*> off_allocs;
    std::vector<int> perm;   // perm[device index] = caller's pattern index
    double weight_total = 0.0;
    struct Class { int f0, f1, maxroot; };
    std::vector<Class> classes;                  // heavy shard: launch groups by root count (shared memory per warp follows it)

    void release() {
        Tc.clear();          // the slab deleters push into `pool`: run them while it is still alive
        Tp.clear();
        pool.clear();
        for (auto* sl : all_slabs) { blgFree(sl->ell); blgFree(sl->scl); delete sl; }
        all_slabs.clear();
        for (auto* o : off_allocs) blgFree(o);
        off_allocs.clear();
        for (auto* c : count_allocs) blgFree(c);
        count_allocs.clear();
        N = 0; ld = 0; perm.clear(); classes.clear(); weight_total = 0.0;
        pats.clear(); links.clear(); h_counts.clear(); h_counts.shrink_to_fit();
    }
};

}  // namespace

struct blg_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // profiles: two family sets (see Shard); the members below alias the FAST shard (the gradient kernels work on it)
    Shard F, H;
    int& N = F.N;
    size_t& ld = F.ld;
    std::vector<Col>& Tc = F.Tc;
    std::vector<Col>& Tp = F.Tp;
    DevBuf<double>& weight = F.weight;
    DevBuf<double>& fam_ll = F.fam_ll;
    double& weight_total = F.weight_total;
    int Ntot = 0;                            // patterns of this handle (both shards)
    std::vector<int> loc;                    // caller's pattern index -> index in the fast set (which holds every family)
    std::vector<int> hloc;                   // caller's pattern index -> index in the heavy set, -1 for families below the heavy bound
    std::vector<int> H2F;                    // heavy index -> fast index
    std::map<const uint16_t*, int*> hpat;    // per column (fast set's count array): device [H.N] node pattern of every heavy family, -1 where the entry is heavy
    int heavy_min = BLG_FAST_MAXCOUNT + 1;   // families whose root count reaches this go to the heavy shard (BLG_HEAVY_MIN)
    DevBuf<double> partial, result, rootw;
    int maxcount = 0;
    DevBuf<double> pascal;
    int pdim = 0;

    // topology (host)
    int ncols_model = 0;
    std::vector<int> parent, child_pos;
    std::vector<uint8_t> kind;
    std::vector<double> t;
    std::vector<std::vector<int>> children;
    std::vector<int> postorder;
    bool have_topology = false;
    // params (host staging)
    int model_kind = BLG_BRANCH_MODEL;
    std::vector<double> lam, mu, q;
    double eta = 0.9;
    bool have_params = false;
    // device model arrays
    DevBuf<uint8_t> d_kind;
    DevBuf<int> d_parent, d_child_off, d_child_idx, d_lists, d_post;
    DevBuf<double> d_par, d_eps0, d_eps1;   // d_par: [t | λ | μ | q], one upload per parameter change
    DevBuf<NodeTab> d_tab;
    DevBuf<NodeK> d_nk;                  // per-node scalars of the pruning kernels: device-resident, never read back by a likelihood call
    DevBuf<int> d_fast_ok, d_flags;
    DevBuf<unsigned int> d_tticket;      // last-block-done counter of tables_kernel (self-resetting)
    DevBuf<unsigned int> d_ran;          // [0] sweep_kernel launches that did the work, [1] gsweep_kernel ones (cumulative)
    NodeTab* d_tab_seen = nullptr;       // device table-descriptor array the last upload went to
    std::vector<double> par_stage;
    std::vector<int> lists_stage;
    std::vector<NodeK> h_nk;             // host copy for the gradient (fetched by gradient() only)
    std::vector<double> h_eps0;
    std::vector<HostTab> tabs;
    std::vector<NodeTab> tab_stage;
    bool tabs_valid = false;
    bool rootw_valid = false;            // the root integration weights match the current parameters
    int rootw_rows = 0;
    DevBuf<unsigned int> d_ticket;       // last-block-done counter of the root kernels (self-resetting)

    // multi-GPU: this handle's rank in a communicator the library owns (nullptr: single GPU)
    NcclApi::Comm comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    DevBuf<double> d_red;                     // [1 + P] staging of the gradient all-reduce
    void allreduce_sum(double* d_buf, size_t n) {
        if (comm) NCK(g_nccl.AllReduce(d_buf, d_buf, n, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, stream));
    }

    // node patterns (pattern.cuh): the fast family set is evaluated once per distinct sub-pattern of every node.
    // pat_mode == false: one column per family (the layout the gradient kernels work on; BLG_DEDUP=0 forces it).
    bool pat_mode = true, pat_default = true;
    int gslot = 0;
    int side_min = 0;                         // BLG_SIDE: fewest tiles of a level for which the wide launch goes to the second stream (-1: never)
    bool fan_classes = true;                  // heavy set: root-count classes on their own streams (BLG_FAN=0: one after the other)
    DevBuf<double> pat_ll;                    // per root pattern: the integrated log-likelihood (written by the root's tiles)
    bool root_fused = false;
    cudaStream_t stream2 = nullptr;           // the wide tile launches of a level run beside its other launches
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    cudaStream_t gstreams[8] = {nullptr}, lstream = nullptr;   // general-kernel launches of independent family ranges
    cudaEvent_t ev_fork = nullptr, ev_join[8] = {nullptr};
    int pdebug = 0;                           // BLG_PDEBUG: time and describe every pattern launch (stderr)
    int flag_cache = -1;                      // host copy of the device-side range flag: -1 unknown (since the last ε/W build)
    std::unique_ptr<PList> plist{new PList()}, plist_w{new PList()}, plist_g{new PList()};
    double last_pattern_bytes = 0.0;          // column bytes the last call actually wrote + read (pattern columns)
    double last_pattern_cols = 0.0, last_family_cols = 0.0;   // node steps × patterns / × families of the last call

    // sweep-kernel plumbing
    int kernel_version = 2;
    bool order_resident = true;
    double link_ms = 0.0;                // host time spent building links during the current sweep (BLG_PDEBUG)
    bool family_order_always = false;    // BLG_FAMILY_ORDER=1
    bool pattern_tiebreak = true;        // BLG_PAT_TIEBREAK=0: patterns of the same shape in order of first appearance
    bool padj_constants = false;
    int padj_warp_min = 0;               // BLG_PADJ_WARP: fixed count bound of the warp kernel of the pattern adjoint (0: per depth)
    int padj_warp_big = 257;             // BLG_PADJ_BIG: the warp path's count bound at depths with many patterns
    int padj_mid_big = 0x7fff;           // BLG_PB_MIDM: phase B's warp path takes child patterns from this many counts at such depths
    long long padj_warp_level = 32768;   // BLG_PADJ_LEVEL: depths with at most this many patterns use the warp kernel
    bool grad_pat = true;                // reverse mode over node patterns (padjoint.cuh); BLG_GRAD_PAT=0: over families (adjoint.cuh)
    int grad_mode = 1;                   // 1: reverse-mode sweep where the tree allows it; 0: forward-mode tangent passes (BLG_GRAD=forward)
    int profile_sweep = 0;
    DevBuf<unsigned long long> d_prof;
    int slice_cap = 50;                       // doubles per lane of the per-warp shared-memory slice
    int sweep_wpb = 4;                        // warps per block of the sweep kernel (BLG_WPB: 1, 2 or 4)
    std::unique_ptr<SweepList> sweep_list{new SweepList()};
    std::unique_ptr<GList> glist{new GList()};
    DevBuf<unsigned int> d_tilectr, d_gctr;
    int sm_count = 148;
    DevBuf<double> gscratch;
    size_t sweep_smem_set = 0, g4_smem_set = 0, g16_smem_set = 0, ptile_smem_set[3] = {0, 0, 0};

    // accounting of the last likelihood call
    double last_bytes = 0.0, last_prune_bytes = 0.0;
    int last_launches = 0;
    bool timing = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;

    // CUDA graphs of whole likelihood calls.  A call is described launch by launch on the host (tile lists, links, slabs:
    // ~10^2 map lookups and 20-40 launches for a 100-taxon tree); the second call in an unchanged state is captured and every
    // later one is a single cudaGraphLaunch.  "Unchanged" = no device allocation or release anywhere in the library
    // (g_mem_epoch), no entry that edits the column tables, topology, profiles or table descriptors (state_epoch), and the
    // same kernel choice.  BLG_GRAPH=0 switches it off.
    struct CallGraph {
        cudaGraphExec_t exec = nullptr;
        int stage = 0;                        // 0: unseen, 1: ran launch by launch in this state, 2: captured, -1: capture failed in this state
        unsigned long long mem_epoch = 0, state_epoch = 0;
        bool general = false, timing = false, rootw_ok = false;
        // what the call leaves behind on the host
        double bytes = 0.0, prune_bytes = 0.0, pattern_bytes = 0.0, pattern_cols = 0.0, family_cols = 0.0;
        int launches = 0, rootw_rows = 0;
        bool root_fused = false, rootw_valid = false;
    };
    std::map<std::tuple<int, int, const void*>, CallGraph> call_graphs;
    unsigned long long state_epoch = 0;
    bool graph_on = true;
    unsigned long long graph_replays = 0, graph_captures = 0;
    void drop_graphs() {
        for (auto& kv : call_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        call_graphs.clear();
    }
    // an event the caller reads back (blg_last_timing): inside a capture it has to become an event-record node
    void record_timing(cudaEvent_t e) {
        if (g_capturing) CK(cudaEventRecordWithFlags(e, stream, cudaEventRecordExternal));
        else CK(cudaEventRecord(e, stream));
    }

    ~blg_handle() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        drop_graphs();
        for (auto& kv : hpat) blgFree(kv.second);
        F.release();
        H.release();
        F.weight.free_(); F.fam_ll.free_(); H.weight.free_(); H.fam_ll.free_();
        for (auto& tb : tabs) free_tab(tb);
        partial.free_(); result.free_(); rootw.free_(); pascal.free_();
        d_kind.free_(); d_parent.free_(); d_child_off.free_(); d_child_idx.free_(); d_lists.free_(); d_post.free_();
        d_par.free_(); d_eps0.free_(); d_eps1.free_(); d_tab.free_(); d_nk.free_(); d_fast_ok.free_(); d_flags.free_();
        d_tticket.free_(); d_ran.free_();
        d_prof.free_();
        gpool_release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        gscratch.free_(); pat_ll.free_(); d_tilectr.free_(); d_gctr.free_(); d_ticket.free_();
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        d_red.free_();
        for (auto& st_ : gstreams) if (st_) cudaStreamDestroy(st_);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& e : ev_join) if (e) cudaEventDestroy(e);
        if (stream2) cudaStreamDestroy(stream2);
        if (ev_main) cudaEventDestroy(ev_main);
        if (ev_side) cudaEventDestroy(ev_side);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
    static void free_tab(HostTab& tb) {
        if (tb.Wt) blgFree(tb.Wt);
        if (tb.ipow) blgFree(tb.ipow);
        for (int i = 0; i < BLG_MAXK; ++i) if (tb.coefD[i]) blgFree(tb.coefD[i]);
        tb = HostTab();
    }

    // ---- slabs ---------------------------------------------------------------------------------
    bool pattern_layout(const Shard& S) const { return pat_mode && !S.ragged; }
    Pat* pat_of(Shard& S, const Col& col) {
        auto it = S.pats.find(col.cnt);
        require(it != S.pats.end(), "internal: column without node patterns");
        return it->second.get();
    }
    size_t slab_doubles(Shard& S, const Col& col) {
        if (S.ragged) return (size_t)col.colsum + 8;
        if (pat_mode) { Pat* P = pat_of(S, col); return std::max<size_t>((size_t)(P->xmax + 1) * P->ld, 32); }
        return (size_t)(col.xmax + 1) * S.ld;
    }
    std::shared_ptr<Slab> new_slab(Shard& S, size_t nd) {
        require(!g_capturing, "internal: slab assignment inside a graph capture");
        ++state_epoch;                         // (a column changes its slab: pooled slabs come without an allocation)
        Slab* s = nullptr;
        auto it = S.pool.find(nd);
        if (it != S.pool.end() && !it->second.empty()) { s = it->second.back(); it->second.pop_back(); }
        else {
            s = new Slab();
            s->nd = nd;
            CK(blgMalloc(&s->ell, nd * sizeof(double)));
            cudaError_t e = blgMalloc(&s->scl, S.ld * sizeof(int));
            if (e != cudaSuccess) { blgFree(s->ell); delete s; throw CudaError(std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
            S.all_slabs.push_back(s);
        }
        s->pat = pattern_layout(S);
        Shard* sp = &S;
        return std::shared_ptr<Slab>(s, [sp](Slab* p) { sp->pool[p->nd].push_back(p); });
    }
    // slab the proposal may overwrite for column c (copy-on-write without the copy: a pruning step
    // rewrites the whole column for every family)
    Slab* writable(Shard& S, int c) {
        Col& col = S.Tp[c];
        if (!col.slab || col.slab.use_count() > 1 || col.slab->pat != pattern_layout(S)) col.slab = new_slab(S, slab_doubles(S, col));
        return col.slab.get();
    }
    Slab* readable(Shard& S, int c) {
        Col& col = S.Tp[c];
        if (!col.slab) {   // never computed: the reference holds -Inf there, i.e. ℓ = 0
            col.slab = new_slab(S, slab_doubles(S, col));
            zero_slab_kernel<<<296, 256, 0, stream>>>(col.slab->ell, col.slab->scl, col.slab->nd, (int)S.ld);
            CK(cudaGetLastError());
        }
        require(col.slab->pat == pattern_layout(S), "internal: column layout (node patterns / families) does not match the evaluation mode");
        return col.slab.get();
    }
    Slab* writable(int c) { return writable(F, c); }
    Slab* readable(int c) { return readable(F, c); }

    void require(bool c, const char* msg) { if (!c) throw std::runtime_error(msg); }

    // ---- topology helpers ------------------------------------------------------------------------
    bool present(int i) const { return i >= 0 && i < ncols_model && kind[i] != BLG_ABSENT; }
    // 0: ordinary branch above node c; 2 / 3: the branch below a WGD / WGT (table contraction, band of the table)
    int table_band(int c) const {
        if (kind[c] != BLG_WGDAFTER) return 0;
        return (parent[c] >= 0 && kind[parent[c]] == BLG_WGT) ? 3 : 2;
    }
    void build_children() {
        children.assign(ncols_model, {});
        std::vector<std::vector<std::pair<int, int>>> tmp(ncols_model);
        for (int i = 0; i < ncols_model; ++i) {
            if (!present(i) || parent[i] < 0) continue;
            require(present(parent[i]), "set_topology: parent id is absent");
            tmp[parent[i]].push_back({child_pos[i], i});
        }
        for (int i = 0; i < ncols_model; ++i) {
            std::sort(tmp[i].begin(), tmp[i].end());
            for (auto& pr : tmp[i]) children[i].push_back(pr.second);
            require((int)children[i].size() <= BLG_MAXK, "set_topology: more than 8 children at one node");
        }
        postorder.clear();
        require(present(0) && kind[0] == BLG_ROOT, "set_topology: id 1 must be the root");
        // iterative post-order from the root
        std::vector<std::pair<int, size_t>> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second < children[top.first].size()) { int c = children[top.first][top.second++]; st.push_back({c, 0}); }
            else { postorder.push_back(top.first); st.pop_back(); }
        }
        int np = 0;
        for (int i = 0; i < ncols_model; ++i) np += present(i) ? 1 : 0;
        require((int)postorder.size() == np, "set_topology: graph is not a tree rooted at id 1");
    }

    ModelDev model_dev() {
        ModelDev m;
        m.n = ncols_model; m.model_kind = model_kind;
        m.kind = d_kind.p; m.parent = d_parent.p; m.child_off = d_child_off.p; m.child_idx = d_child_idx.p;
        m.t = d_par.p; m.lam = d_par.p + ncols_model; m.mu = d_par.p + 2 * (size_t)ncols_model;
        m.q = d_par.p + 3 * (size_t)ncols_model; m.eta = eta;
        m.eps0 = d_eps0.p; m.eps1 = d_eps1.p; m.tab = d_tab.p; m.pascal = pascal.p; m.pdim = pdim;
        m.nk = d_nk.p; m.fast_ok = d_fast_ok.p;
        m.rootw = nullptr; m.rootlw = nullptr; m.rootw_n = 0;
        m.flags = d_flags.p; m.ticket = d_tticket.p; m.post = d_post.p; m.npost = (int)postorder.size();
        return m;
    }

    // count bound of column i in a shard (0 when the shard is empty or the column does not exist yet)
    static int xmax_of(const Shard& S, int i) { return (S.N > 0 && i < (int)S.Tp.size()) ? S.Tp[i].xmax : 0; }
    static double colsum_of(const Shard& S, int i) { return (S.N > 0 && i < (int)S.Tp.size()) ? S.Tp[i].colsum : 0.0; }

    // merge order of node i's children: the child with the larger counts first (the "t" side of the
    // merge, whose column lives in shared memory); the smaller one is the register-tiled "s" side
    std::vector<int> merge_order(int i) {
        std::vector<int> mo = children[i];
        if (mo.size() == 2) {
            double a = colsum_of(F, mo[0]);
            double b = colsum_of(F, mo[1]);
            if (b > a) std::swap(mo[0], mo[1]);
        }
        return mo;
    }
    // (re)allocate the tables of node i for the current proposal column dims.  Tables exist for the FAST shard's count
    // bounds only (<= BLG_FAST_MAXCOUNT); the branch below a WGT keeps its literal table for both shards.
    void ensure_tab(int i) {
        HostTab& tb = tabs[i];
        int k = (int)children[i].size();
        const bool havef = F.N > 0 && i < (int)F.Tp.size();
        int xm = xmax_of(F, i);
        int wdim = (kind[i] == BLG_ROOT || !havef) ? 0 : xm + 1;
        if (kind[i] == BLG_WGDAFTER && parent[i] >= 0 && kind[parent[i]] == BLG_WGT && H.N > 0 && i < (int)H.Tp.size())
            wdim = std::max(wdim, xmax_of(H, i) + 1);
        int nrow = (k > 0 && havef) ? xm + 1 : 0;
        std::vector<int> mo = merge_order(i);
        bool same = tb.wdim == wdim && tb.nrow == nrow && tb.k == k && tb.xfast == xm;
        for (int c = 0; c < k && same; ++c) {
            int ci = mo[c];
            int cx = xmax_of(F, ci);
            if (tb.mo[c] != ci) same = false;
            if (c >= 1 && nrow > 0 && tb.cdp[c] != round_up(cx + 1, BLG_TS)) same = false;
        }
        if (same && (tb.Wt || wdim == 0) && (tb.ipow || nrow == 0)) return;
        free_tab(tb);
        tb.wdim = wdim; tb.wdp = round_up(std::max(wdim, 1), 16); tb.nrow = nrow; tb.k = k; tb.xfast = xm;
        for (int c = 0; c < k; ++c) tb.mo[c] = mo[c];
        if (wdim > 0) CK(blgMalloc(&tb.Wt, (size_t)wdim * tb.wdp * sizeof(double)));
        if (nrow > 0) {
            CK(blgMalloc(&tb.ipow, (size_t)nrow * sizeof(double)));
            for (int c = 1; c < k; ++c) {
                int ci = mo[c];
                int cx = xmax_of(F, ci);
                tb.cdp[c] = round_up(cx + 1, BLG_TS);
                CK(blgMalloc(&tb.coefD[c], (size_t)nrow * tb.cdp[c] * sizeof(double)));
            }
        }
    }

    void upload_params() {
        require(have_topology, "set_params/rebuild: call blg_set_topology first");
        require(have_params, "rebuild: call blg_set_params first");
        size_t n = (size_t)ncols_model;
        d_par.ensure(4 * n); d_eps1.ensure(n); d_eps0.ensure(n);
        par_stage.resize(4 * n);
        std::copy(t.begin(), t.begin() + n, par_stage.begin());
        std::copy(lam.begin(), lam.begin() + n, par_stage.begin() + n);
        std::copy(mu.begin(), mu.begin() + n, par_stage.begin() + 2 * n);
        std::copy(q.begin(), q.begin() + n, par_stage.begin() + 3 * n);
        // pageable source: the copy is staged before the call returns
        CK(cudaMemcpyAsync(d_par.p, par_stage.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
    }

    int root_rows() const { return std::max(xmax_of(F, 0), xmax_of(H, 0)) + 1; }

    // set!(n) for the listed ids in the listed order (nodes == nullptr: set!(d), every node).  Everything stays on the
    // device: ε (eps_kernel), W / per-node scalars / range flags / NaN rule (tables_kernel); nothing is read back.
    void rebuild(const int32_t* nodes, int n) {
        require(have_topology && have_params, "rebuild: topology and params must be set first");
        std::vector<int> list;
        if (nodes == nullptr || n <= 0) list = postorder;
        else {
            for (int k = 0; k < n; ++k) {
                int i = nodes[k] - 1;
                require(present(i), "rebuild: node id not in the model");
                list.push_back(i);
            }
        }
        const bool full = list.size() == postorder.size() && (nodes == nullptr || n <= 0);
        if (!tabs_valid) require(full, "rebuild: the topology changed; a full rebuild (nodes = NULL) is required");
        if ((int)tabs.size() < ncols_model) tabs.resize(ncols_model);
        if ((int)tab_stage.size() < ncols_model) tab_stage.resize(ncols_model);
        // the tables are built once per node: a literal update! replay may list a node several times
        // (tables_kernel's blocks must not build the same table concurrently)
        std::vector<int> tlist;
        {
            std::vector<char> seen(ncols_model, 0);
            for (int i : list) if (!seen[i]) { seen[i] = 1; tlist.push_back(i); }
        }
        int maxdim = 1;
        bool tab_dirty = false;
        for (int i : tlist) {
            ensure_tab(i);
            HostTab& tb = tabs[i];
            NodeTab nt;
            memset(&nt, 0, sizeof(nt));
            nt.Wt = tb.Wt; nt.wdim = tb.wdim; nt.wdp = tb.wdp; nt.ipow = tb.ipow; nt.nrow = tb.nrow; nt.xfast = tb.xfast;
            for (int c = 0; c < BLG_MAXK; ++c) { nt.coefD[c] = tb.coefD[c]; nt.cdp[c] = tb.cdp[c]; nt.mo[c] = tb.mo[c]; }
            if (memcmp(&tab_stage[i], &nt, sizeof(nt)) != 0) { tab_stage[i] = nt; tab_dirty = true; ++state_epoch; }
            tb.built = true;
            maxdim = std::max(maxdim, tb.wdim);
        }
        const size_t nn = (size_t)ncols_model;
        d_tab.ensure(nn);
        d_nk.ensure(nn); d_fast_ok.ensure(nn); d_eps0.ensure(nn); d_eps1.ensure(nn);
        if (!d_flags.p) { d_flags.ensure(4); CK(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), stream)); }
        if (!d_tticket.p) { d_tticket.ensure(1); CK(cudaMemsetAsync(d_tticket.p, 0, sizeof(unsigned int), stream)); }
        if (!d_ran.p) { d_ran.ensure(4); CK(cudaMemsetAsync(d_ran.p, 0, 4 * sizeof(unsigned int), stream)); }
        if (tab_dirty || d_tab.p != d_tab_seen || full) {
            CK(cudaMemcpyAsync(d_tab.p, tab_stage.data(), nn * sizeof(NodeTab), cudaMemcpyHostToDevice, stream));
            d_tab_seen = d_tab.p;
        }
        // ε segments: a full rebuild groups the nodes by height above the leaves (children always sit in an
        // earlier segment); a partial list is replayed literally, one node after the other
        std::vector<int> elist = list, seg;
        if (full) {
            std::vector<int> height(ncols_model, 0);
            for (int u : postorder) {
                int hh = 0;
                for (int c : children[u]) hh = std::max(hh, height[c] + 1);
                height[u] = hh;
            }
            std::stable_sort(elist.begin(), elist.end(), [&](int a, int b) { return height[a] < height[b]; });
            seg.push_back(0);
            for (size_t i = 1; i < elist.size(); ++i) if (height[elist[i]] != height[elist[i - 1]]) seg.push_back((int)i);
            seg.push_back((int)elist.size());
        } else {
            for (size_t i = 0; i <= elist.size(); ++i) seg.push_back((int)i);
        }
        // [tlist | elist | seg] in one upload
        lists_stage.clear();
        lists_stage.insert(lists_stage.end(), tlist.begin(), tlist.end());
        lists_stage.insert(lists_stage.end(), elist.begin(), elist.end());
        lists_stage.insert(lists_stage.end(), seg.begin(), seg.end());
        d_lists.ensure(lists_stage.size());
        CK(cudaMemcpyAsync(d_lists.p, lists_stage.data(), lists_stage.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        const int* d_list_p = d_lists.p;
        const int* d_elist_p = d_lists.p + tlist.size();
        const int* d_seg_p = d_elist_p + elist.size();
        ModelDev m = model_dev();
        bool root_listed = false;
        for (int i : tlist) root_listed = root_listed || i == 0;
        if (root_listed && Ntot > 0 && !children[0].empty()) {      // the block of the root also builds the root weights
            const int nrow = root_rows();
            rootw.ensure(2 * ((size_t)nrow + 1));
            m.rootw = rootw.p;
            m.rootlw = H.N > 0 ? rootw.p + nrow + 1 : nullptr;
            m.rootw_n = nrow;
            rootw_valid = true;
            rootw_rows = nrow;
        } else {
            rootw_valid = false;
        }
        {
            const int nseg_ = (int)seg.size() - 1, nlist_ = seg.back();
            const size_t esm = (size_t)ncols_model * 61 + ((size_t)nlist_ + nseg_ + 4) * 4 + 64;
            const int stage = esm <= 46 * 1024 ? 1 : 0;
            eps_kernel<<<1, 128, stage ? esm : 0, stream>>>(m, d_elist_p, d_seg_p, nseg_, nlist_, stage);
        }
        CK(cudaGetLastError());
        const size_t tsm = std::max(2 * (size_t)maxdim * sizeof(double), (5 * nn + 8) * sizeof(int));
        tables_kernel<<<(unsigned)tlist.size(), 128, tsm, stream>>>(m, d_list_p, (int)tlist.size());
        CK(cudaGetLastError());
        if (full) tabs_valid = true;
        flag_cache = -1;
    }

    // ---- pruning -----------------------------------------------------------------------------------
    // describe the pruning step at node u against the current proposal tables of the FAST shard (allocates the output
    // slab); the table pointers are for the gradient kernels, the sweep kernel only takes the slabs from here
    NodeStep make_step(int u, bool readonly = false, bool count_bytes = true) {
        int k = (int)children[u].size();
        require(k >= 1, "prune_node: leaf");
        require(u < (int)Tp.size(), "logpdf: model has more node ids than the profile has columns (call blg_extend)");
        require(tabs_valid && tabs[u].built, "logpdf: ε/W tables are stale (call blg_rebuild)");
        require(tabs[u].k == k, "logpdf: node table does not match the topology (call blg_rebuild)");
        NodeStep st;
        memset(&st, 0, sizeof(st));
        st.k = k;
        for (int i = 0; i < k; ++i) {
            int c = tabs[u].mo[i];
            require(c >= 0 && c < (int)Tp.size(), "logpdf: child id beyond the profile columns");
            require(tabs[c].built && tabs[c].wdim >= Tp[c].xmax + 1 && tabs[c].xfast == Tp[c].xmax, "logpdf: W table of a child does not match the profile (call blg_rebuild)");
            ChildRef& cr = st.child[i];
            if (children[c].empty()) { cr.ell = nullptr; cr.scl = nullptr; if (count_bytes) last_bytes += 2.0 * N; }
            else { Slab* s = readable(c); cr.ell = s->ell; cr.scl = s->scl; if (count_bytes) last_bytes += 8.0 * Tp[c].colsum + 4.0 * N; }
            cr.cnt = Tp[c].cnt;
            cr.Wt = tabs[c].Wt;
            cr.wdp = tabs[c].wdp;
            cr.id = c;
            cr.coefD = tabs[u].coefD[i];
            cr.cdp = tabs[u].cdp[i];
            if (i >= 1) require(tabs[u].cdp[i] == round_up(Tp[c].xmax + 1, BLG_TS), "logpdf: merge table does not match the profile (call blg_rebuild)");
        }
        require(tabs[u].nrow == Tp[u].xmax + 1, "logpdf: node table does not match the profile (call blg_rebuild)");
        Slab* o = readonly ? readable(u) : writable(u);
        st.nrow = tabs[u].nrow;
        st.out = o->ell; st.out_scl = o->scl; st.cnt = Tp[u].cnt;
        st.ipow = tabs[u].ipow;
        if (count_bytes) last_bytes += 8.0 * Tp[u].colsum + 4.0 * N;
        return st;
    }
    // one launch for a whole list of nodes (chunked only if the list exceeds the parameter space): the FAST shard through
    // the table-free sweep kernel.  The kernel leaves at once when some node is outside its range (flags[0], decided on
    // the device); the general kernel launched right behind it then does the work (general_sweep(F, nodes, 1)).
    void sweep(const std::vector<int>& nodes) {
        size_t pos = 0;
        while (pos < nodes.size()) {
            SweepList& L = *sweep_list;
            memset(&L, 0, sizeof(int) * 4);
            int nch = 0, worst = 1, prev_node = -1;
            while (pos < nodes.size() && L.nsteps < BLG_SWEEP_MAXSTEPS) {
                int u = nodes[pos];
                int k = (int)children[u].size();
                if (nch + k > BLG_SWEEP_MAXCHILD) break;
                NodeStep st = make_step(u);
                // the node computed by the previous step of this launch goes to merge position 0: its column is
                // still in shared memory (and its count/exponent must not be prefetched before they exist)
                for (int c = 1; c < k; ++c)
                    if (st.child[c].id == prev_node) { std::rotate(st.child, st.child + c, st.child + c + 1); break; }
                SweepStep& ss = L.step[L.nsteps++];
                ss.k = k; ss.first = nch; ss.node = u;
                ss.resident = (L.nsteps >= 2 && prev_node == st.child[0].id && !children[prev_node].empty()) ? 1 : 0;
                prev_node = u;
                ss.out = st.out; ss.out_scl = st.out_scl; ss.cnt = st.cnt;
                int sum = 0;
                bool tiled = false;
                for (int c = 0; c < k; ++c) {
                    SweepChild& sc = L.child[nch++];
                    const ChildRef& cr = st.child[c];
                    sc.ell = cr.ell; sc.scl = cr.scl; sc.cnt = cr.cnt; sc.Wt = cr.Wt;
                    sc.wdp = cr.wdp; sc.table = table_band(cr.id);
                    sc.node = cr.id; sc.pad_ = 0;
                    int xm = Tp[cr.id].xmax;
                    sum += xm;
                    if (c >= 1 && xm >= 20) tiled = true;
                }
                if (k == 2 && std::min(Tp[st.child[0].id].xmax, Tp[st.child[1].id].xmax) < 20) tiled = false;
                worst = std::max(worst, tiled ? 3 * (sum + 1) : sum + k + 1);
                ++pos;
            }
            require(L.nsteps > 0, "sweep: a node has more children than one launch can describe");
            const int warps_per_block = sweep_wpb;            // 4 blocks of 4 warps, or 8 blocks of 2 warps, per SM
            const int per_wave = sm_count * (16 / warps_per_block);
            int nwarps = (N + 31) / 32;
            int blocks = (nwarps + warps_per_block - 1) / warps_per_block;
            // small batches (up to ~3 waves): persistent grid with tile stealing, so a fractional last wave costs a
            // fraction; large batches: plain grid (measured 10 % faster on the 5-wave benchmark)
            const bool persist = blocks <= per_wave * 3 && blocks > per_wave;
            if (persist) blocks = per_wave;
            d_tilectr.ensure(1);
            if (persist) CK(cudaMemsetAsync(d_tilectr.p, 0, sizeof(unsigned int), stream));
            int cap = std::min(slice_cap, worst);            // doubles per lane in the shared slice
            double* gs = nullptr;
            size_t gstride = 0;
            if (worst > cap) {                                // some warp may need more than its slice: global scratch
                gstride = (size_t)worst * 32;
                gscratch.ensure(gstride * (size_t)nwarps);
                gs = gscratch.p;
            }
            size_t smem = ((size_t)cap * 32 + 190) * warps_per_block * sizeof(double);
            if (smem > sweep_smem_set) {
                CK(cudaFuncSetAttribute(sweep_kernel<false, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(cudaFuncSetAttribute(sweep_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(cudaFuncSetAttribute(sweep_kernel<true, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                sweep_smem_set = smem;
            }
            if (!profile_sweep) {
                if (persist) sweep_kernel<false, 4, true><<<blocks, 32 * warps_per_block, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, N, ld, cap, gs, gstride, nullptr, d_tilectr.p);
                else sweep_kernel<false, 4, false><<<blocks, 32 * warps_per_block, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, N, ld, cap, gs, gstride, nullptr, d_tilectr.p);
                CK(cudaGetLastError());
            } else {
                d_prof.ensure((size_t)BLG_SWEEP_MAXSTEPS * 8);
                CK(cudaMemsetAsync(d_prof.p, 0, (size_t)BLG_SWEEP_MAXSTEPS * 8 * sizeof(unsigned long long), stream));
                if (persist) { blocks = (nwarps + warps_per_block - 1) / warps_per_block; }
                sweep_kernel<true, 4, false><<<blocks, 32 * warps_per_block, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, N, ld, cap, gs, gstride, d_prof.p, d_tilectr.p);
                CK(cudaGetLastError());
                std::vector<unsigned long long> hp((size_t)BLG_SWEEP_MAXSTEPS * 8);
                CK(cudaMemcpyAsync(hp.data(), d_prof.p, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                double tot[8] = {0};
                fprintf(stderr, "[blg sweep profile] warp-cycles per step: node k res | prologue child0 child1 merge tiled scale store\n");
                for (int i = 0; i < L.nsteps; ++i) {
                    double rowsum = 0;
                    for (int p = 0; p < 7; ++p) { tot[p] += (double)hp[i * 8 + p]; rowsum += (double)hp[i * 8 + p]; }
                    if (profile_sweep > 1)
                        fprintf(stderr, "  step %3d k%d r%d | %8.0f %8.0f %8.0f %8.0f %8.0f %8.0f %8.0f | %9.0f\n", i, L.step[i].k,
                                L.step[i].resident, hp[i*8]/1e3, hp[i*8+1]/1e3, hp[i*8+2]/1e3, hp[i*8+3]/1e3, hp[i*8+4]/1e3, hp[i*8+5]/1e3, hp[i*8+6]/1e3, rowsum/1e3);
                }
                double all = 0; for (int p = 0; p < 7; ++p) all += tot[p];
                fprintf(stderr, "  total (Mcycles): prologue %.1f child0 %.1f child1 %.1f merge %.1f tiled %.1f scale %.1f store %.1f | sum %.1f\n",
                        tot[0]/1e6, tot[1]/1e6, tot[2]/1e6, tot[3]/1e6, tot[4]/1e6, tot[5]/1e6, tot[6]/1e6, all/1e6);
            }
            last_launches++;
        }
    }

    // node pattern of every heavy family at column c (-1 where its entry is heavy), on the device
    const int* hpat_of(int c) {
        require(pat_mode, "heavy families need the node-pattern evaluation (BLG_DEDUP=0 and the gradient do not support them)");
        const uint16_t* key = F.Tp[c].cnt;
        auto it = hpat.find(key);
        if (it != hpat.end()) return it->second;
        Pat* P = pat_of(F, F.Tp[c]);
        std::vector<int> v(std::max(H.N, 1), -1);
        for (int i = 0; i < H.N; ++i) v[i] = P->fam2pat[H2F[i]];
        int* d = nullptr;
        CK(blgMalloc(&d, v.size() * sizeof(int)));
        CK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
        hpat[key] = d;
        return d;
    }

    // the general kernel over a shard: one warp per family, any counts, any parameters.
    // when = 0: always; when = 1: only if the device-side range flag says the sweep kernel could not run (fast shard).
    void general_sweep(Shard& S, const std::vector<int>& nodes, int when, bool count_bytes) {
        require(tabs_valid, "logpdf: ε/W tables are stale (call blg_rebuild)");
        size_t pos = 0;
        while (pos < nodes.size()) {
            GList& L = *glist;
            memset(&L, 0, sizeof(int) * 4);
            int nch = 0, prev_node = -1;
            while (pos < nodes.size() && L.nsteps < BLG_G_MAXSTEPS) {
                const int u = nodes[pos];
                const int k = (int)children[u].size();
                require(k >= 1, "prune_node: leaf");
                if (nch + k > BLG_G_MAXCHILD) break;
                require(u < (int)S.Tp.size(), "logpdf: model has more node ids than the profile has columns (call blg_extend)");
                std::vector<int> ord = children[u];
                for (int c = 1; c < k; ++c)
                    if (ord[c] == prev_node) { std::rotate(ord.begin(), ord.begin() + c, ord.begin() + c + 1); break; }
                GStep& gs = L.step[L.nsteps++];
                gs.k = k; gs.first = nch; gs.node = u;
                gs.resident = (L.nsteps >= 2 && prev_node == ord[0] && !children[prev_node].empty()) ? 1 : 0;
                for (int c : ord) {
                    require(c >= 0 && c < (int)S.Tp.size(), "logpdf: child id beyond the profile columns");
                    GChild& gc = L.child[nch++];
                    memset(&gc, 0, sizeof(gc));
                    if (children[c].empty()) { if (count_bytes) last_bytes += 2.0 * S.N; }
                    else {
                        Slab* sl = readable(S, c);
                        gc.ell = sl->ell; gc.scl = sl->scl;
                        gc.off = S.ragged ? S.Tp[c].off : nullptr;
                        gc.ld = (int)S.ld;
                        if (count_bytes) last_bytes += 8.0 * S.Tp[c].colsum + 4.0 * S.N;
                        if (S.ragged) {              // light entries of this child: the fast set's pattern columns (computed just before)
                            Pat* Pc = pat_of(F, F.Tp[c]);
                            Slab* ps = readable(F, c);
                            gc.fpat = hpat_of(c);
                            gc.pell = ps->ell; gc.pscl = ps->scl; gc.pld = (int)Pc->ld;
                        }
                    }
                    gc.cnt = S.Tp[c].cnt;
                    gc.node = c;
                    gc.mode = 0;
                    if (kind[c] == BLG_WGDAFTER) {
                        if (kind[u] == BLG_WGD) gc.mode = 1;
                        else {
                            gc.mode = 2;
                            require(tabs[c].built && tabs[c].Wt && tabs[c].wdim >= S.Tp[c].xmax + 1, "logpdf: W table below a WGT does not match the profile (call blg_rebuild)");
                            gc.Wt = tabs[c].Wt; gc.wdp = tabs[c].wdp;
                        }
                    }
                }
                Slab* o = writable(S, u);
                gs.out = o->ell; gs.out_scl = o->scl; gs.cnt = S.Tp[u].cnt;
                gs.out_off = S.ragged ? S.Tp[u].off : nullptr;
                gs.out_ld = S.ld;
                gs.fpat = S.ragged ? hpat_of(u) : nullptr;
                if (count_bytes) last_bytes += 8.0 * S.Tp[u].colsum + 4.0 * S.N;
                prev_node = u;
                ++pos;
            }
            require(L.nsteps > 0, "sweep: a node has more children than one launch can describe");
            // launch groups: the fast shard is one group; the heavy shard one per root-count class
            std::vector<Shard::Class> groups = S.classes;
            if (groups.empty()) groups.push_back({0, S.N, xmax_of(S, 0)});
            require(groups.size() <= 16, "general_sweep: too many launch groups");
            // the classes are independent family ranges: each on its own stream (the top classes hold a handful of
            // families whose columns take milliseconds — they run beside the wide classes instead of in front of them)
            const bool fan = groups.size() > 1 && fan_classes;
            if (fan && !ev_fork) {
                CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
                for (auto& e : ev_join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                for (auto& st_ : gstreams) CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
            }
            if (fan) CK(cudaEventRecord(ev_fork, stream));
            int gi = 0;
            std::vector<cudaEvent_t> dbg;                    // BLG_PDEBUG: start and end of every class on its stream
            cudaEvent_t dbg0 = nullptr;
            if (pdebug && fan) { CK(cudaEventCreate(&dbg0)); CK(cudaEventRecord(dbg0, stream)); }
            for (const auto& g : groups) {
                if (fan) { CK(cudaStreamWaitEvent(gstreams[gi % 8], ev_fork, 0)); lstream = gstreams[gi % 8]; }
                if (dbg0) { cudaEvent_t e; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, lstream)); dbg.push_back(e); }
                launch_general(L, g.f0, g.f1, g.maxroot, 1, when, gi);
                if (dbg0) { cudaEvent_t e; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, lstream)); dbg.push_back(e); }
                lstream = nullptr;
                ++gi;
            }
            if (dbg0) {
                for (int q = 0; q < std::min<int>(8, (int)groups.size()); ++q) CK(cudaStreamSynchronize(gstreams[q]));
                for (size_t q = 0; q < groups.size(); ++q) {
                    float a = 0.f, b = 0.f;
                    CK(cudaEventElapsedTime(&a, dbg0, dbg[2 * q]));
                    CK(cudaEventElapsedTime(&b, dbg0, dbg[2 * q + 1]));
                    fprintf(stderr, "[blg heavy] class %zu: families %d..%d, root count <= %d: from %.2f ms to %.2f ms after the fork\n", q, groups[q].f0,
                            groups[q].f1, groups[q].maxroot, a, b);
                }
                for (auto e : dbg) cudaEventDestroy(e);
                cudaEventDestroy(dbg0);
            }
            if (fan) {
                for (int q = 0; q < std::min<int>(8, (int)groups.size()); ++q) {
                    CK(cudaEventRecord(ev_join[q], gstreams[q]));
                    CK(cudaStreamWaitEvent(stream, ev_join[q], 0));
                }
            }
        }
    }

    // ---- node patterns (pattern.cuh) -----------------------------------------------------------------------------------
    const uint16_t* host_counts(const Shard& S, const Col& col) const {
        return S.h_counts.data() + (col.cnt - S.count_allocs[0]);
    }
    void upload_pat(Pat& P) {
        P.ld = std::max<size_t>(((size_t)P.U + 31) / 32 * 32, 32);
        std::vector<uint16_t> st(P.ld, P.h_cnt.empty() ? 0 : P.h_cnt.back());
        std::copy(P.h_cnt.begin(), P.h_cnt.end(), st.begin());
        CK(blgMalloc(&P.d_cnt, P.ld * sizeof(uint16_t)));
        CK(cudaMemcpy(P.d_cnt, st.data(), P.ld * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    // distinct sub-patterns of every column of the current topology, bottom-up: a leaf's patterns are its distinct counts,
    // a single child hands its patterns up unchanged, otherwise the distinct tuples of the children's patterns
    // the host side of one node's patterns (no CUDA calls: nodes of one tree level are built on several host threads)
    std::shared_ptr<Pat> build_pat(const Shard& S, int u) const {
        const int N = S.N;
        const Col& cu = S.Tp[u];
        const int k = (int)children[u].size();
        auto P = std::make_shared<Pat>();
        const uint16_t* hx = host_counts(S, cu);
        // entries above the column's bound belong to the heavy family set: pattern -1; so does every entry with such a child
        const int hm = cu.lmax + 1;
        std::vector<int> id(N, -1);
        int U = 0;
        std::vector<Pat*> cp;
        if (k == 0) {
            std::vector<int> tab(cu.xmax + 1, -1);
            for (int f = 0; f < N; ++f) if (hx[f] < hm) tab[hx[f]] = 0;
            for (int v = cu.xmax; v >= 0; --v) if (tab[v] == 0) tab[v] = U++;      // descending counts
            for (int f = 0; f < N; ++f) if (hx[f] < hm) id[f] = tab[hx[f]];
        } else {
            for (int c : children[u]) {
                if (c >= (int)S.Tp.size()) throw std::runtime_error("logpdf: child id beyond the profile columns");
                cp.push_back(S.pats.at(S.Tp[c].cnt).get());
            }
            for (int f = 0; f < N; ++f) {
                if (hx[f] >= hm) continue;
                bool ok = true;
                for (int j = 0; j < k; ++j) ok = ok && cp[j]->fam2pat[f] >= 0;
                if (ok) id[f] = cp[0]->fam2pat[f];
            }
            U = cp[0]->U;
            for (int j = 1; j < k; ++j) {
                const std::vector<int>& b = cp[j]->fam2pat;
                const uint64_t Ub = (uint64_t)cp[j]->U, space = (uint64_t)U * Ub;
                int next = 0;
                if (space <= std::max<uint64_t>(2ull * (uint64_t)N, 1ull << 16)) {
                    std::vector<int> tab((size_t)space, -1);
                    for (int f = 0; f < N; ++f) {
                        if (id[f] < 0) continue;
                        int& t_ = tab[(size_t)((uint64_t)id[f] * Ub + (uint64_t)b[f])];
                        if (t_ < 0) t_ = next++;
                        id[f] = t_;
                    }
                } else {
                    // open addressing, linear probing: (key + 1, value) slots, at most half full
                    size_t cap = 1024;
                    while (cap < (size_t)N * 2) cap <<= 1;
                    std::vector<uint64_t> hk(cap, 0);
                    std::vector<int> hv(cap, 0);
                    for (int f = 0; f < N; ++f) {
                        if (id[f] < 0) continue;
                        const uint64_t key = (uint64_t)id[f] * Ub + (uint64_t)b[f] + 1;
                        size_t slot = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
                        while (hk[slot] != 0 && hk[slot] != key) slot = (slot + 1) & (cap - 1);
                        if (hk[slot] == 0) { hk[slot] = key; hv[slot] = next++; }
                        id[f] = hv[slot];
                    }
                }
                U = next;
            }
        }
        // a family of each pattern, then the patterns in descending order of (own count, children's counts): the 32
        // patterns of a tile get (nearly) the same loop bounds, and the heaviest tiles start first
        {   // ids may have gaps (patterns of a child that only heavy entries of this node use): make them dense
            std::vector<int> dense(std::max(U, 1), -1);
            int nu = 0;
            for (int f = 0; f < N; ++f) if (id[f] >= 0 && dense[id[f]] < 0) dense[id[f]] = nu++;
            for (int f = 0; f < N; ++f) if (id[f] >= 0) id[f] = dense[id[f]];
            U = nu;
        }
        std::vector<int> rep(U, -1);
        for (int f = 0; f < N; ++f) if (id[f] >= 0 && rep[id[f]] < 0) rep[id[f]] = f;
        std::vector<int> ord(U);
        for (int i = 0; i < U; ++i) ord[i] = i;
        int nvery = 0, nwide = 0;
        if (k >= 2) {
            std::vector<uint64_t> key(U);
            for (int i = 0; i < U; ++i) {
                const int r = rep[i];
                int m1 = 0, m2 = 0;                     // largest and second-largest child count
                int mc[2] = {0, 0};
                for (int j = 0; j < k; ++j) {
                    const int mj = cp[j]->h_cnt[cp[j]->fam2pat[r]];
                    if (j < 2) mc[j] = mj;
                    if (mj > m1) { m2 = m1; m1 = mj; } else if (mj > m2) m2 = mj;
                }
                // class, then (own count, the larger child's count) descending: the 32 patterns of a tile get (nearly)
                // the same loop bounds AND the same register-tiled side
                // (polytomies: no wide class — their tiles run on the generic instantiation, anything beyond its register
                // merges on the general kernel)
                uint64_t cls = (m2 >= 36 || (k > 2 && m2 >= 20)) ? 3 : m2 >= 20 ? 2 : (mc[0] >= mc[1] ? 1 : 0);
                if (cls == 3) ++nvery; else if (cls == 2) ++nwide;
                key[i] = (cls << 60) | ((uint64_t)hx[r] << 32) | ((uint64_t)std::max(mc[0], mc[1]) << 16) | (uint64_t)std::min(mc[0], mc[1]);
            }
            // sorted by descending key; patterns of the same shape by the pattern index of their larger child, so that the
            // lanes of a tile gather neighbouring columns of that child (whatever order the families came in)
            std::vector<std::pair<uint64_t, uint64_t>> kv(U);
            for (int i = 0; i < U; ++i) {
                const int r = rep[i];
                const int p0 = cp[0]->fam2pat[r], p1 = cp[1]->fam2pat[r];
                const int big = pattern_tiebreak ? (cp[0]->h_cnt[p0] >= cp[1]->h_cnt[p1] ? p0 : p1) : 0;
                kv[i] = {~key[i], ((uint64_t)(uint32_t)big << 32) | (uint32_t)i};
            }
            std::sort(kv.begin(), kv.end());
            for (int i = 0; i < U; ++i) ord[i] = (int)(kv[i].second & 0xffffffffu);
        }
        P->nvery = nvery;
        P->nwide = nwide;
        P->kids.assign(cp.begin(), cp.end());
        std::vector<int> newidx(U);
        for (int i = 0; i < U; ++i) newidx[ord[i]] = i;
        P->U = U;
        P->rep.resize(U);
        P->h_cnt.resize(U);
        for (int i = 0; i < U; ++i) { P->rep[i] = rep[ord[i]]; P->h_cnt[i] = hx[rep[ord[i]]]; }
        for (int f = 0; f < N; ++f) if (id[f] >= 0) id[f] = newidx[id[f]];
        P->fam2pat.swap(id);
        P->xmax = 0;
        P->colsum = 0.0;
        for (int i = 0; i < U; ++i) { P->colsum += (double)P->h_cnt[i] + 1.0; P->xmax = std::max(P->xmax, (int)P->h_cnt[i]); }
        return P;
    }
    void ensure_pats(Shard& S) {
        // levels: a node is built after its children; single-child nodes share their child's patterns
        std::vector<int> lev(ncols_model, -1);
        std::vector<std::vector<int>> bylev, alias;
        for (int u : postorder) {
            require(u < (int)S.Tp.size(), "logpdf: model has more node ids than the profile has columns (call blg_extend)");
            if (S.pats.count(S.Tp[u].cnt)) continue;
            const int k = (int)children[u].size();
            int lv = 0;
            for (int c : children[u]) {
                require(c < (int)S.Tp.size(), "logpdf: child id beyond the profile columns");
                if (lev[c] >= 0) lv = std::max(lv, lev[c] + (k == 1 ? 0 : 1));
            }
            lev[u] = lv;
            if ((int)bylev.size() <= lv) { bylev.resize(lv + 1); alias.resize(lv + 1); }
            (k == 1 ? alias : bylev)[lv].push_back(u);
        }
        for (size_t lv = 0; lv < bylev.size(); ++lv) {
            const std::vector<int>& nodes = bylev[lv];
            std::vector<std::shared_ptr<Pat>> built(nodes.size());
            std::vector<std::string> errs(nodes.size());
            // heaviest nodes first: the threads take them one at a time
            std::atomic<size_t> next{0};
            std::vector<size_t> order(nodes.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return children[nodes[a]].size() > children[nodes[b]].size(); });
            auto work = [&]() {
                for (;;) {
                    const size_t q = next.fetch_add(1);
                    if (q >= order.size()) break;
                    const size_t i = order[q];
                    try { built[i] = build_pat(S, nodes[i]); } catch (const std::exception& e) { errs[i] = e.what(); }
                }
            };
            const unsigned nt = (unsigned)std::min<size_t>(std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())), nodes.size());
            if (nt <= 1 || S.N < 4096) work();
            else {
                std::vector<std::thread> th;
                for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
                for (auto& x : th) x.join();
            }
            for (size_t i = 0; i < nodes.size(); ++i) {
                if (!errs[i].empty()) throw std::runtime_error(errs[i]);
                const Col& cu = S.Tp[nodes[i]];
                if (S.pats.count(cu.cnt)) continue;
                upload_pat(*built[i]);
                S.pats[cu.cnt] = built[i];
            }
            for (int u : alias[lv]) {                       // (postorder inside the level: chains resolve bottom-up)
                const Col& cu = S.Tp[u];
                if (!S.pats.count(cu.cnt)) S.pats[cu.cnt] = S.pats.at(S.Tp[children[u][0]].cnt);
            }
        }
    }
    // node pattern → child pattern maps for node u with its children in the order `ord`
    Link* get_link(Shard& S, int u, const std::vector<int>& ord) {
        std::vector<const void*> key;
        key.push_back(S.Tp[u].cnt);
        for (int c : ord) key.push_back(S.Tp[c].cnt);
        auto it = S.links.find(key);
        if (it != S.links.end()) return it->second.get();
        struct Clock {
            double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
            ~Clock() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
        } clock{link_ms};
        Pat* Pu = pat_of(S, S.Tp[u]);
        auto L = std::make_shared<Link>();
        for (size_t j = 0; j < ord.size(); ++j) {
            Pat* Pc = pat_of(S, S.Tp[ord[j]]);
            if (Pc == Pu) {                       // (a single child: the node's patterns ARE the child's)
                require(ord.size() == 1, "internal: a node with several children shares a child's patterns");
                continue;
            }
            std::vector<int> px(Pu->ld, 0);
            for (int q = 0; q < Pu->U; ++q) px[q] = Pc->fam2pat[Pu->rep[q]];
            for (size_t q = Pu->U; q < Pu->ld; ++q) px[q] = Pu->U > 0 ? px[Pu->U - 1] : 0;
            // the node's pattern must determine the child's for EVERY family: it does by construction when the node's
            // patterns were built from this child's; a different tree over the same columns needs new patterns
            if (std::find(Pu->kids.begin(), Pu->kids.end(), (const Pat*)Pc) == Pu->kids.end() || Pu->kids.size() != ord.size())
                for (int f = 0; f < S.N; ++f)
                    if (Pu->fam2pat[f] >= 0 && px[Pu->fam2pat[f]] != Pc->fam2pat[f]) throw std::runtime_error("logpdf: node patterns do not match the topology (call blg_set_profiles again after changing the species tree)");
            CK(blgMalloc(&L->d_pidx[j], Pu->ld * sizeof(int)));
            CK(cudaMemcpy(L->d_pidx[j], px.data(), Pu->ld * sizeof(int), cudaMemcpyHostToDevice));
            std::vector<uint16_t> xc(Pu->ld, 0);
            for (size_t q = 0; q < Pu->ld; ++q) xc[q] = px[q] < Pc->U ? Pc->h_cnt[px[q]] : 0;     // (a node without light patterns: U = 0)
            CK(blgMalloc(&L->d_xcnt[j], Pu->ld * sizeof(uint16_t)));
            CK(cudaMemcpy(L->d_xcnt[j], xc.data(), Pu->ld * sizeof(uint16_t), cudaMemcpyHostToDevice));
        }
        // ---- cut the tiles -------------------------------------------------------------------------------------------
        const int k = (int)ord.size();
        if (k <= 2) {
            Pat* P0 = pat_of(S, S.Tp[ord[0]]);
            Pat* P1 = k == 2 ? pat_of(S, S.Tp[ord[1]]) : nullptr;
            L->cut = true;
            for (int w = 0; w < 2; ++w) {
                const int lo = w ? Pu->nvery : Pu->nvery + Pu->nwide, hi = w ? Pu->nvery + Pu->nwide : Pu->U;
                if (hi <= lo) continue;
                const int tmax = w ? 36 : 20;                                   // widest register merge of the instantiation
                const int lim = w ? 104 : std::max(56, std::min(104, Pu->xmax + 6));
                std::vector<int> tb;
                int m0 = 0, m1 = 0, cnt = 0, need = 0;
                for (int q = lo; q < hi; ++q) {
                    const int r = Pu->rep[q];
                    const int a = P0->h_cnt[P0->fam2pat[r]], b = P1 ? (int)P1->h_cnt[P1->fam2pat[r]] : 0;
                    const int n0 = std::max(m0, a), n1 = std::max(m1, b);
                    const bool fits = cnt < 32 && n0 + n1 + k + 1 <= lim && (k == 1 || std::min(n0, n1) < tmax);
                    if (cnt == 0 || !fits) {                                    // (a single pattern always fits: its class says so)
                        tb.push_back(q);
                        m0 = a; m1 = b; cnt = 1;
                    } else { m0 = n0; m1 = n1; ++cnt; }
                    need = std::max(need, m0 + m1 + k + 1);
                }
                tb.push_back(hi);
                L->ntiles[w] = (int)tb.size() - 1;
                L->maxneed[w] = need;
                CK(blgMalloc(&L->d_tiles[w], tb.size() * sizeof(int)));
                CK(cudaMemcpy(L->d_tiles[w], tb.data(), tb.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
            if (pdebug) fprintf(stderr, "[blg link] node %d U %d x %d: tiles %d (need %d) | wide tiles %d (need %d)\n", u, Pu->U, Pu->xmax,
                                L->ntiles[0], L->maxneed[0], L->ntiles[1], L->maxneed[1]);
        }
        S.links[key] = L;
        return L.get();
    }
    // algorithmic bytes of one node step in SURVEY §8(d)'s per-FAMILY accounting (whatever layout does the work)
    void count_step_bytes(Shard& S, int u) {
        for (int c : children[u]) {
            if (children[c].empty()) last_bytes += 2.0 * S.N;
            else last_bytes += 8.0 * S.Tp[c].colsum + 4.0 * S.N;
        }
        last_bytes += 8.0 * S.Tp[u].colsum + 4.0 * S.N;
        last_family_cols += (double)S.N;
    }
    // the fast family set over node patterns: the nodes of the list level by level (children in earlier launches).
    // Per level two launches: the patterns at the front of each node's order — those whose two children both have >= 20
    // counts (beyond the register tiles of the table-free merge), and, when the level is too small to fill the GPU, every
    // pattern with more than a few counts — run on the general kernel, one WARP per pattern (short latency); the rest on
    // the table-free tile kernel, one THREAD per pattern, 32 patterns of (nearly) the same shape per warp (high
    // throughput).  Single-child nodes on the same patterns (WGD-type nodes) are chained behind the node below them.
    // general == true: everything on the general kernel (parameters outside the table-free forms' range).
    void psweep(const std::vector<int>& nodes, bool general) {
        Shard& S = F;
        {
            const auto t0 = std::chrono::steady_clock::now();
            ensure_pats(S);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (pdebug && ms > 1.0) fprintf(stderr, "[blg ingest] node patterns built in %.1f ms\n", ms);
            link_ms = 0.0;
        }
        std::vector<int> level(ncols_model, -1), head(ncols_model, -1);
        std::vector<std::vector<int>> chain(ncols_model);
        int nlev = 0;
        for (int u : nodes) {
            const int k = (int)children[u].size();
            require(k >= 1, "prune_node: leaf");
            if (k == 1 && !general) {
                const int c = children[u][0];
                if (level[c] >= 0 && !children[c].empty() && pat_of(S, S.Tp[u]) == pat_of(S, S.Tp[c])) {
                    head[u] = head[c] >= 0 ? head[c] : c;
                    level[u] = level[c];
                    chain[head[u]].push_back(u);
                    continue;
                }
            }
            int lv = 0;
            for (int c : children[u]) if (level[c] >= 0) lv = std::max(lv, level[c] + 1);
            level[u] = lv;
            nlev = std::max(nlev, lv + 1);
        }
        bool first = true;
        gslot = 0;
        for (int lv = 0; lv < nlev; ++lv) {
            std::vector<int> todo;
            for (int u : nodes) if (level[u] == lv && head[u] < 0) todo.push_back(u);
            // heaviest nodes first inside a level
            std::stable_sort(todo.begin(), todo.end(), [&](int a, int b) { return pat_of(S, S.Tp[a])->colsum > pat_of(S, S.Tp[b])->colsum; });
            if (general) {
                for (int u : todo) {
                    const int k = (int)children[u].size();
                    require(tabs_valid && tabs[u].built && tabs[u].k == k, "logpdf: ε/W tables are stale (call blg_rebuild)");
                    std::vector<int> ord(tabs[u].mo, tabs[u].mo + k);
                    general_pstep(S, u, ord, pat_of(S, S.Tp[u]), get_link(S, u, ord), first);
                    first = false;
                    count_step_bytes(S, u);
                }
                continue;
            }
            size_t pos = 0;
            while (pos < todo.size()) {
                // three launches per level: [0, nvery) of every node on the general kernel (one warp per pattern),
                // [nvery, nvery + nwide) on the wide tile kernel, the rest on the tile kernel
                // tile lists: 0 = tile kernel, 1 = wide tile kernel (both over host-cut tiles, shared-memory addressing),
                // 2 = the generic instantiation (polytomies: fixed tiles, the column may spill to global scratch)
                struct TL { PList* L; std::vector<PStep> chain; int nch = 0, ntiles = 0, worst = 1; };
                TL tl[3];
                tl[0].L = plist.get(); tl[1].L = plist_w.get(); tl[2].L = plist_g.get();
                for (auto& t_ : tl) memset(t_.L, 0, sizeof(int) * 4);
                GList& G = *glist;
                memset(&G, 0, sizeof(int) * 4);
                int gch = 0, nitems = 0, gmax = 0;
                const size_t pos_before = pos;
                while (pos < todo.size()) {
                    const int u = todo[pos];
                    const int k = (int)children[u].size();
                    const int nc = (int)chain[u].size();
                    bool room = G.nsteps + 1 + nc <= BLG_G_MAXSTEPS && gch + k + nc <= BLG_G_MAXCHILD;
                    for (auto& t_ : tl) room = room && t_.L->nsteps + (int)t_.chain.size() + 1 + nc <= BLG_P_MAXSTEPS && t_.nch + k + nc <= BLG_P_MAXCHILD;
                    if (!room) break;
                    require(tabs_valid && tabs[u].built && tabs[u].k == k, "logpdf: ε/W tables are stale (call blg_rebuild)");
                    std::vector<int> ord(tabs[u].mo, tabs[u].mo + k);
                    Pat* Pu = pat_of(S, S.Tp[u]);
                    Link* lk = get_link(S, u, ord);
                    const int ng = Pu->nvery, nw = Pu->nvery + Pu->nwide;
                    Slab* o = writable(S, u);
                    std::vector<Slab*> co;
                    for (int v : chain[u]) co.push_back(writable(S, v));
                    std::vector<Slab*> csl(k, nullptr);
                    int sum = 0;
                    for (int j = 0; j < k; ++j) {
                        const int c = ord[j];
                        require(tabs[c].built && tabs[c].wdim >= S.Tp[c].xmax + 1 && tabs[c].xfast == S.Tp[c].xmax,
                                "logpdf: W table of a child does not match the profile (call blg_rebuild)");
                        Pat* Pc = pat_of(S, S.Tp[c]);
                        csl[j] = children[c].empty() ? nullptr : readable(S, c);
                        sum += Pc->xmax;
                        last_pattern_bytes += csl[j] ? 8.0 * Pc->colsum : 0.0;   // (each child column is read at least once)
                    }
                    sum = std::min(sum, k * Pu->xmax);             // (per tile the children's maxima are taken separately)
                    for (int w = 0; w < 2; ++w) {                  // w = 0: tile kernel, w = 1: wide tile kernel
                        const int lo = w ? ng : nw, hi = w ? nw : Pu->U;
                        if (hi <= lo) continue;
                        TL& T = tl[lk->cut ? w : 2];
                        require(lk->cut || w == 0, "internal: a polytomy with a wide pattern class");
                        PList& L = *T.L;
                        if (lk->cut) T.worst = std::max(T.worst, lk->maxneed[w]);                      // exact: the host cut the tiles
                        else T.worst = std::max(T.worst, std::max(3 * (sum + 1), sum + k + 1));       // a bound: chained-tile merge, 3 columns
                        PStep& ps = L.step[L.nsteps++];
                        memset(&ps, 0, sizeof(ps));
                        ps.k = k; ps.first = T.nch; ps.node = u; ps.U = hi; ps.tile0 = T.ntiles; ps.ld = (int)Pu->ld; ps.p0 = lo;
                        ps.tstart = lk->cut ? lk->d_tiles[w] : nullptr;
                        ps.cnt = Pu->d_cnt; ps.out = o->ell; ps.out_scl = o->scl;
                        ps.chain_first = (int)T.chain.size(); ps.nchain = nc;
                        if (u == 0 && rootw_valid && rootw_rows == root_rows()) {      // the root: integrate on the spot
                            pat_ll.ensure(Pu->ld);
                            ps.rootw = rootw.p; ps.pat_ll = pat_ll.p;
                            root_fused = true;
                        }
                        for (int j = 0; j < k; ++j) {
                            const int c = ord[j];
                            Pat* Pc = pat_of(S, S.Tp[c]);
                            PChild& pc = L.child[T.nch++];
                            memset(&pc, 0, sizeof(pc));
                            if (csl[j]) { pc.ell = csl[j]->ell; pc.scl = csl[j]->scl; }
                            pc.cnt = Pc->d_cnt;
                            pc.pidx = lk->d_pidx[j];
                            pc.xcnt = lk->d_xcnt[j];
                            pc.Wt = tabs[c].Wt; pc.wdp = tabs[c].wdp; pc.table = table_band(c);
                            pc.node = c; pc.ld = (int)Pc->ld;
                        }
                        int below = u;
                        for (int ci = 0; ci < nc; ++ci) {
                            const int v = chain[u][ci];
                            PStep cs;
                            memset(&cs, 0, sizeof(cs));
                            cs.k = 1; cs.first = T.nch; cs.node = v; cs.U = hi; cs.ld = (int)Pu->ld;
                            cs.cnt = Pu->d_cnt; cs.out = co[ci]->ell; cs.out_scl = co[ci]->scl;
                            PChild& pc = L.child[T.nch++];
                            memset(&pc, 0, sizeof(pc));
                            pc.cnt = Pu->d_cnt;
                            pc.ell = (below == u ? o : co[ci - 1])->ell;      // (resident: never read from memory, but not a leaf)
                            pc.scl = (below == u ? o : co[ci - 1])->scl;
                            pc.Wt = tabs[below].Wt; pc.wdp = tabs[below].wdp; pc.table = table_band(below);
                            pc.node = below; pc.ld = (int)Pu->ld;
                            T.chain.push_back(cs);
                            below = v;
                        }
                        T.ntiles += lk->cut ? lk->ntiles[w] : (hi - lo + 31) / 32;
                    }
                    if (ng > 0) {
                        GStep& gs = G.step[G.nsteps++];
                        memset(&gs, 0, sizeof(gs));
                        gs.k = k; gs.first = gch; gs.node = u; gs.resident = 0; gs.item0 = nitems; gs.nchain = nc; gs.is_chain = 0;
                        gs.out = o->ell; gs.out_scl = o->scl; gs.cnt = Pu->d_cnt; gs.out_off = nullptr; gs.out_ld = Pu->ld;
                        for (int j = 0; j < k; ++j) fill_gchild(S, G.child[gch++], u, ord[j], pat_of(S, S.Tp[ord[j]]), csl[j], lk->d_pidx[j]);
                        int below = u;
                        for (int ci = 0; ci < nc; ++ci) {
                            const int v = chain[u][ci];
                            GStep& cs = G.step[G.nsteps++];
                            memset(&cs, 0, sizeof(cs));
                            cs.k = 1; cs.first = gch; cs.node = v; cs.resident = 1; cs.item0 = nitems; cs.nchain = 0; cs.is_chain = 1;
                            cs.out = co[ci]->ell; cs.out_scl = co[ci]->scl; cs.cnt = Pu->d_cnt; cs.out_off = nullptr; cs.out_ld = Pu->ld;
                            fill_gchild(S, G.child[gch++], v, below, Pu, nullptr, nullptr);
                            G.child[gch - 1].ell = (below == u ? o : co[ci - 1])->ell;      // (resident: never read from memory)
                            G.child[gch - 1].scl = (below == u ? o : co[ci - 1])->scl;
                            below = v;
                        }
                        nitems += ng;
                        gmax = std::max(gmax, Pu->xmax);
                    }
                    last_pattern_bytes += (8.0 * Pu->colsum + 4.0 * Pu->U) * (1 + nc);
                    last_pattern_cols += (double)Pu->U * (1 + nc);
                    count_step_bytes(S, u);
                    for (int v : chain[u]) count_step_bytes(S, v);
                    ++pos;
                }
                require(pos > pos_before, "sweep: a node has more children than one launch can describe");
                if (tl[0].L->nsteps + tl[1].L->nsteps + tl[2].L->nsteps + G.nsteps == 0) continue;   // (nodes whose entries are all heavy)
                cudaEvent_t dbg[4] = {nullptr, nullptr, nullptr, nullptr};
                if (pdebug) { for (auto& e : dbg) CK(cudaEventCreate(&e)); CK(cudaEventRecord(dbg[0], stream)); }
                if (G.nsteps > 0) {
                    // (work counters of the item launches: 240 slots zeroed in one go, one slot per launch)
                    if (gslot == 0) { d_gctr.ensure(256); CK(cudaMemsetAsync(d_gctr.p + 16, 0, 240 * sizeof(unsigned int), stream)); }
                    launch_general(G, 0, nitems, gmax, 0, 0, 16 + gslot, 1);
                    gslot = (gslot + 1) % 240;
                }
                if (pdebug) CK(cudaEventRecord(dbg[1], stream));
                int caps[2] = {0, 0};
                // the wide launch goes to a second (high-priority) stream: it usually holds a few long tiles, which then run
                // beside the level's other launches instead of in front of them
                const bool side = tl[1].L->nsteps > 0 && !pdebug && side_min >= 0 && tl[0].ntiles + tl[1].ntiles >= side_min;
                if (side) {
                    if (!stream2) {
                        int lo_ = 0, hi_ = 0;
                        CK(cudaDeviceGetStreamPriorityRange(&lo_, &hi_));
                        CK(cudaStreamCreateWithPriority(&stream2, cudaStreamNonBlocking, hi_));
                        CK(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming));
                        CK(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
                    }
                    CK(cudaEventRecord(ev_main, stream));          // everything below this level is in `stream` up to here
                    CK(cudaStreamWaitEvent(stream2, ev_main, 0));
                }
                for (int w : {1, 2, 0}) {                          // the wide tiles first: they are the long ones
                    TL& T = tl[w];
                    PList& L = *T.L;
                    cudaStream_t stream = (w == 1 && side) ? stream2 : this->stream;
                    if (L.nsteps > 0) {
                        for (size_t q = 0; q < T.chain.size(); ++q) L.step[L.nsteps + q] = T.chain[q];
                        L.ntiles = T.ntiles;
                        // doubles per lane of a warp's shared slice: exactly what the launch's tiles need (w < 2); the generic
                        // instantiation sends larger tiles to global scratch
                        const int cap = w == 2 ? std::min(T.worst, std::max(slice_cap, 56)) : T.worst;
                        caps[w == 1 ? 1 : 0] = std::max(caps[w == 1 ? 1 : 0], cap);
                        const int wpb = 4;
                        double* gs = nullptr;
                        size_t gstride = 0;
                        if (T.worst > cap) {
                            gstride = (size_t)T.worst * 32;
                            gscratch.ensure(gstride * (size_t)T.ntiles);
                            gs = gscratch.p;
                        }
                        const size_t smem = (size_t)cap * 32 * wpb * sizeof(double);
                        if (smem > ptile_smem_set[w]) {
                            if (w == 1) CK(cudaFuncSetAttribute(ptile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            else if (w == 0) CK(cudaFuncSetAttribute(ptile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            else CK(cudaFuncSetAttribute(ptile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            ptile_smem_set[w] = smem;
                        }
                        const int nb = (T.ntiles + wpb - 1) / wpb;
                        if (w == 1) ptile_kernel<true, true><<<nb, 32 * wpb, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, cap, gs, gstride);
                        else if (w == 0) ptile_kernel<false, true><<<nb, 32 * wpb, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, cap, gs, gstride);
                        else ptile_kernel<false, false><<<nb, 32 * wpb, smem, stream>>>(L, d_nk.p, d_flags.p, d_ran.p, cap, gs, gstride);
                        CK(cudaGetLastError());
                        last_launches++;
                    }
                    if (pdebug && w != 2) CK(cudaEventRecord(dbg[3 - w], stream));
                }
                if (side) {                                        // join: the next level reads what the wide launch wrote
                    CK(cudaEventRecord(ev_side, stream2));
                    CK(cudaStreamWaitEvent(stream, ev_side, 0));
                }
                if (pdebug) {
                    CK(cudaEventSynchronize(dbg[3]));
                    float ms[3] = {0.f, 0.f, 0.f};
                    for (int q = 0; q < 3; ++q) CK(cudaEventElapsedTime(&ms[q], dbg[q], dbg[q + 1]));
                    for (auto& e : dbg) cudaEventDestroy(e);
                    fprintf(stderr, "[blg pattern] level %2d: warp-items %6d (%d steps) %6.1f us | wide tiles %5d (%d steps, cap %d) %6.1f us | tiles %6d (%d+%d steps, cap %d worst %d) %6.1f us |",
                            lv, nitems, G.nsteps, ms[0] * 1e3, tl[1].ntiles, tl[1].L->nsteps, caps[1], ms[1] * 1e3, tl[0].ntiles, tl[0].L->nsteps, (int)tl[0].chain.size(),
                            caps[0], tl[0].worst, ms[2] * 1e3);
                    for (int i = 0; i < std::min(tl[0].L->nsteps, 3); ++i) {
                        const PStep& q = tl[0].L->step[i];
                        fprintf(stderr, " [node %d U %d p0 %d x %d chain %d:", q.node, q.U, q.p0, pat_of(S, S.Tp[q.node])->xmax, q.nchain);
                        for (int j = 0; j < q.k; ++j) fprintf(stderr, " %d", pat_of(S, S.Tp[tl[0].L->child[q.first + j].node])->xmax);
                        fprintf(stderr, "]");
                    }
                    fprintf(stderr, "\n");
                }
            }
        }
        if (pdebug && link_ms > 1.0) fprintf(stderr, "[blg ingest] pattern links and tiles built in %.1f ms\n", link_ms);
    }
    void fill_gchild(Shard& S, GChild& gc, int u, int c, Pat* Pc, Slab* sl, const int* pidx) {
        memset(&gc, 0, sizeof(gc));
        if (sl) { gc.ell = sl->ell; gc.scl = sl->scl; }
        gc.cnt = Pc->d_cnt;
        gc.pidx = pidx;
        gc.node = c;
        gc.ld = (int)Pc->ld;
        gc.mode = 0;
        if (kind[c] == BLG_WGDAFTER) {
            if (kind[u] == BLG_WGD) gc.mode = 1;
            else {
                gc.mode = 2;
                require(tabs[c].built && tabs[c].Wt && tabs[c].wdim >= S.Tp[c].xmax + 1, "logpdf: W table below a WGT does not match the profile (call blg_rebuild)");
                gc.Wt = tabs[c].Wt; gc.wdp = tabs[c].wdp;
            }
        }
    }
    // one node over its patterns with the general kernel (one warp per pattern)
    void general_pstep(Shard& S, int u, const std::vector<int>& ord, Pat* Pu, Link* lk, bool first) {
        GList& L = *glist;
        memset(&L, 0, sizeof(int) * 4);
        L.nsteps = 1;
        GStep& gs = L.step[0];
        memset(&gs, 0, sizeof(gs));
        gs.k = (int)ord.size(); gs.first = 0; gs.node = u; gs.resident = 0;
        for (size_t j = 0; j < ord.size(); ++j) {
            const int c = ord[j];
            fill_gchild(S, L.child[j], u, c, pat_of(S, S.Tp[c]), children[c].empty() ? nullptr : readable(S, c), lk->d_pidx[j]);
        }
        Slab* o = writable(S, u);
        gs.out = o->ell; gs.out_scl = o->scl; gs.cnt = Pu->d_cnt; gs.out_off = nullptr; gs.out_ld = Pu->ld;
        launch_general(L, 0, Pu->U, Pu->xmax, first ? 1 : 0, 0, 0, 0);
        last_pattern_bytes += 8.0 * Pu->colsum + 4.0 * Pu->U;
        last_pattern_cols += (double)Pu->U;
    }
    // gsweep_kernel over items f0..f1-1 whose largest count is maxroot
    void launch_general(const GList& L, int f0, int f1, int maxroot, int count_run, int when, int counter_slot, int item_mode = 0) {
        const int nfam = f1 - f0;
        if (nfam <= 0) return;
        cudaStream_t stream = lstream ? lstream : this->stream;
        d_gctr.ensure(256);
        if (!item_mode) CK(cudaMemsetAsync(d_gctr.p + counter_slot, 0, sizeof(unsigned int), stream));
        const int cap = round_up(maxroot + 4, 2);
        const bool small = maxroot < 128;                // every merge fits 4 registers per lane
        const int cmax = small ? 4 : 16;
        const int nbuf = (maxroot >= 32 * cmax) ? 3 : 2;
        int wpb = 4;
        while (wpb > 1 && (size_t)wpb * nbuf * cap * sizeof(double) > 200 * 1024) wpb >>= 1;
        const size_t smem = (size_t)wpb * nbuf * cap * sizeof(double);
        require(smem <= 227 * 1024, "general_sweep: count bound too large for the shared-memory columns");
        int blocks = std::min((nfam + wpb - 1) / wpb, sm_count * 16);
        if (small) {
            if (smem > g4_smem_set) { CK(cudaFuncSetAttribute(gsweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); g4_smem_set = smem; }
            gsweep_kernel<4><<<blocks, 32 * wpb, smem, stream>>>(L, d_nk.p, d_flags.p, when, d_ran.p, f0, f1, count_run, cap, nbuf, d_gctr.p + counter_slot, item_mode);
        } else {
            if (smem > g16_smem_set) { CK(cudaFuncSetAttribute(gsweep_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); g16_smem_set = smem; }
            gsweep_kernel<16><<<blocks, 32 * wpb, smem, stream>>>(L, d_nk.p, d_flags.p, when, d_ran.p, f0, f1, count_run, cap, nbuf, d_gctr.p + counter_slot, item_mode);
        }
        CK(cudaGetLastError());
        last_launches++;
    }

    // ---- gradient (forward-mode, path-sparse) ---------------------------------------------------------------
    std::map<size_t, std::vector<void*>> gpool;          // scratch buffers of the tangent passes, by size
    void* galloc(size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        auto& v = gpool[bytes];
        if (!v.empty()) { void* p = v.back(); v.pop_back(); return p; }
        void* p = nullptr;
        CK(blgMalloc(&p, bytes));
        return p;
    }
    void gfree(void* p, size_t bytes) { bytes = (bytes + 255) / 256 * 256; gpool[bytes].push_back(p); }
    void gpool_release() {
        for (auto& kv : gpool) for (void* p : kv.second) blgFree(p);
        gpool.clear();
    }
    int nonwgdchild_h(int i) const { while (kind[i] == BLG_WGD || kind[i] == BLG_WGT || kind[i] == BLG_WGDAFTER) i = children[i][0]; return i; }
    int nonwgdparent_h(int i) const { while (kind[i] == BLG_WGD || kind[i] == BLG_WGT || kind[i] == BLG_WGDAFTER) i = parent[i]; return i; }

    template <int CAP> void launch_prune_dual(const NodeStep& st, const DualStep<2>& ds) {
        int blocks = (N + 63) / 64;
        prune_dual_kernel<2, CAP><<<blocks, 64, 0, stream>>>(st, ds, d_eps0.p, N, ld);
    }

    // one tangent pass with two lanes; adds Σ weight·∂logL/∂lane into out2[0..1]
    void tangent_pass(const Seeds<2>& sd, double out2[2]) {
        const int n = ncols_model;
        // ---- which quantities does this parameter group perturb? (post-order dependency walk) ----
        std::vector<char> e0d(n, 0), e1d(n, 0), wd(n, 0), sdirty(n, 0);
        auto seeded_rate = [&](int p, int c) {
            for (int t = 0; t < 2; ++t) {
                int k = sd.kind[t];
                if (k == SEED_LAM_ALL || k == SEED_MU_ALL) return true;
                if (k == SEED_LAM || k == SEED_MU) {
                    int b = nonwgdchild_h(c);
                    if (sd.node[t] == b) return true;
                    if (model_kind == BLG_NODE_MODEL && sd.node[t] == nonwgdparent_h(p)) return true;
                }
            }
            return false;
        };
        auto seeded_q = [&](int w) {
            for (int t = 0; t < 2; ++t) if (sd.kind[t] == SEED_Q && sd.node[t] == w) return true;
            return false;
        };
        std::vector<int> elist, tlist, steps;
        for (int u : postorder) {
            if (children[u].empty()) continue;
            bool any_e0 = false, any_step = false;
            const bool wg = kind[u] == BLG_WGD || kind[u] == BLG_WGT;
            for (int c : children[u]) {
                bool rate = (!wg && kind[c] != BLG_WGDAFTER) ? seeded_rate(u, c) : false;
                // ϵ[1] of c: ordinary parent → ep(λ,μ,t,ϵ_c[2]); wgd/wgt parent → polynomial in q and ϵ_c[2]
                e0d[c] = wg ? (seeded_q(u) || e1d[c]) : (rate || e1d[c]);
                // W of c: wgdafter → (q of its parent, ϵ_c[2]); ordinary → (rates, ϵ_c[2]).  (a wgd's own branch is ordinary)
                bool rate_w = (kind[c] != BLG_WGDAFTER) ? seeded_rate(u, c) : false;
                wd[c] = (kind[c] == BLG_WGDAFTER) ? (seeded_q(u) || e1d[c]) : (rate_w || e1d[c]);
                any_e0 |= e0d[c] != 0;
                any_step |= wd[c] || e0d[c] || sdirty[c];
                if (wd[c]) tlist.push_back(c);
            }
            e1d[u] = any_e0;
            if (any_e0) { elist.push_back(u); if (std::find(tlist.begin(), tlist.end(), u) == tlist.end()) tlist.push_back(u); }
            sdirty[u] = any_step;
            if (any_step) steps.push_back(u);
        }
        // a node can be in tlist both as a child (dWt) and as a parent (merge tables): dedupe, keep order
        {
            std::vector<int> uniq;
            std::vector<char> seen(n, 0);
            for (int i : tlist) if (!seen[i]) { seen[i] = 1; uniq.push_back(i); }
            tlist.swap(uniq);
        }
        ModelDev m = model_dev();
        // ---- ε tangents ----
        size_t debytes = (size_t)2 * n * sizeof(double);
        double* de0 = (double*)galloc(debytes);
        double* de1 = (double*)galloc(debytes);
        CK(cudaMemsetAsync(de0, 0, debytes, stream));
        CK(cudaMemsetAsync(de1, 0, debytes, stream));
        std::vector<std::pair<void*, size_t>> temps;
        temps.push_back({de0, debytes});
        temps.push_back({de1, debytes});
        if (!elist.empty()) {
            int* dl = (int*)galloc(elist.size() * sizeof(int));
            temps.push_back({dl, elist.size() * sizeof(int)});
            CK(cudaMemcpyAsync(dl, elist.data(), elist.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            eps_dual_kernel<2><<<1, 32, 0, stream>>>(m, sd, dl, (int)elist.size(), de0, de1);
            CK(cudaGetLastError());
        }
        // ---- table tangents ----
        std::vector<DualTab> dts(tlist.size());
        std::map<int, DualTab> dt_of;
        int maxdim = 1;
        for (size_t i = 0; i < tlist.size(); ++i) {
            int u = tlist[i];
            HostTab& tb = tabs[u];
            DualTab dt;
            memset(&dt, 0, sizeof(dt));
            if (tb.Wt) {
                size_t b = (size_t)2 * tb.wdim * tb.wdp * sizeof(double);
                dt.dWt = (double*)galloc(b);
                temps.push_back({dt.dWt, b});
                maxdim = std::max(maxdim, tb.wdim);
            }
            if (tb.ipow) {
                size_t b = (size_t)2 * tb.nrow * sizeof(double);
                dt.dipow = (double*)galloc(b);
                temps.push_back({dt.dipow, b});
                for (int c = 1; c < tb.k; ++c) {
                    size_t bc = (size_t)2 * tb.nrow * tb.cdp[c] * sizeof(double);
                    dt.dcoefD[c] = (double*)galloc(bc);
                    temps.push_back({dt.dcoefD[c], bc});
                }
            }
            dts[i] = dt;
            dt_of[u] = dt;
        }
        if (!tlist.empty()) {
            int* dl = (int*)galloc(tlist.size() * sizeof(int));
            temps.push_back({dl, tlist.size() * sizeof(int)});
            DualTab* ddt = (DualTab*)galloc(tlist.size() * sizeof(DualTab));
            temps.push_back({ddt, tlist.size() * sizeof(DualTab)});
            CK(cudaMemcpyAsync(dl, tlist.data(), tlist.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(ddt, dts.data(), tlist.size() * sizeof(DualTab), cudaMemcpyHostToDevice, stream));
            size_t smem = (size_t)2 * maxdim * 3 * sizeof(double);
            tables_dual_kernel<2><<<(unsigned)tlist.size(), 128, smem, stream>>>(m, sd, dl, ddt, de0, de1);
            CK(cudaGetLastError());
        }
        // ---- tangent columns up the perturbed steps ----
        std::map<int, std::pair<double*, size_t>> dcol;      // node -> (tangent column, stride rows*ld)
        for (int u : steps) {
            NodeStep st = make_step(u, true);
            DualStep<2> ds;
            memset(&ds, 0, sizeof(ds));
            for (int i = 0; i < st.k; ++i) {
                int c = st.child[i].id;
                auto it = dcol.find(c);
                if (it != dcol.end()) { ds.child[i].dell = it->second.first; ds.child[i].dell_stride = it->second.second; }
                auto jt = dt_of.find(c);
                if (jt != dt_of.end() && jt->second.dWt && wd[c]) { ds.child[i].dWt = jt->second.dWt; ds.child[i].dWt_stride = (size_t)tabs[c].wdim * tabs[c].wdp; }
                auto ut = dt_of.find(u);
                if (i >= 1 && ut != dt_of.end() && ut->second.dcoefD[i]) { ds.child[i].dcoefD = ut->second.dcoefD[i]; ds.child[i].dcoef_stride = (size_t)tabs[u].nrow * tabs[u].cdp[i]; }
            }
            auto ut = dt_of.find(u);
            if (ut != dt_of.end() && ut->second.dipow) ds.dipow = ut->second.dipow;
            ds.nrow = tabs[u].nrow;
            ds.de0 = de0;
            ds.de_stride = (size_t)n;
            size_t rows = (size_t)Tp[u].xmax + 1;
            size_t b = (size_t)2 * rows * ld * sizeof(double);
            ds.dout = (double*)galloc(b);
            temps.push_back({ds.dout, b});
            ds.dout_stride = rows * ld;
            ds.out_scl = st.out_scl;
            dcol[u] = {ds.dout, rows * ld};
            int need = Tp[u].xmax + 1;
            if (need <= 16) launch_prune_dual<16>(st, ds);
            else if (need <= 64) launch_prune_dual<64>(st, ds);
            else if (need <= 256) launch_prune_dual<256>(st, ds);
            else launch_prune_dual<1024>(st, ds);
            CK(cudaGetLastError());
        }
        // ---- root ----
        Slab* rs = readable(0);
        int nrow = Tp[0].xmax + 1;
        size_t rwb = (size_t)3 * nrow * sizeof(double);
        double* rw = (double*)galloc(rwb);
        temps.push_back({rw, rwb});
        root_tables_dual_kernel<2><<<1, 128, 0, stream>>>(m, sd, de0, de1, rw, nrow);
        CK(cudaGetLastError());
        int blocks = (N + 255) / 256;
        size_t pb = (size_t)2 * blocks * sizeof(double);
        double* part = (double*)galloc(pb);
        temps.push_back({part, pb});
        auto rt = dcol.find(0);
        root_dual_kernel<2><<<blocks, 256, 0, stream>>>(rs->ell, rt != dcol.end() ? rt->second.first : nullptr,
                                                        rt != dcol.end() ? rt->second.second : 0, Tp[0].cnt, rw, nrow, weight.p, part, blocks, N, ld);
        CK(cudaGetLastError());
        double* dres = (double*)galloc(2 * sizeof(double));
        temps.push_back({dres, 2 * sizeof(double)});
        grad_reduce_kernel<<<1, 256, 0, stream>>>(part, blocks, 2, dres);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out2, dres, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (auto& t : temps) gfree(t.first, t.second);
    }

    // ---- gradient, reverse mode (adjoint.cuh) -------------------------------------------------------------------
    bool adjoint_eligible() const {
        if (children[0].size() != 2) return false;
        for (int u : postorder) {
            if (children[u].size() > 2) return false;
            if (u < (int)Tp.size() && Tp[u].xmax > 255) return false;
        }
        return true;
    }
    template <int CAP> void launch_adjoint(const AdjStep& as) {
        int blocks = (N + 127) / 128;
        adjoint_node_kernel<CAP><<<blocks, 128, 0, stream>>>(as, pascal.p, pdim, N, ld);
    }
    // res[2·i + lane] = ∂logL/∂(parameter of lane `lane` of group i); the caller has just run a full logpdf
    void gradient_adjoint(const std::vector<Seeds<2>>& groups, std::vector<double>& res) {
        const int n = ncols_model;
        ModelDev m = model_dev();
        std::vector<std::pair<void*, size_t>> temps;
        auto ga = [&](size_t b) { void* p = galloc(b); temps.push_back({p, b}); return p; };
        std::vector<double*> adj(n, nullptr);
        std::vector<int*> adjs(n, nullptr);
        for (int u : postorder) {
            if (children[u].empty()) continue;
            adj[u] = (double*)ga((size_t)(Tp[u].xmax + 1) * ld * sizeof(double));
            adjs[u] = (int*)ga(ld * sizeof(int));
        }
        double* sadj = (double*)ga((size_t)n * 4 * sizeof(double));
        CK(cudaMemsetAsync(sadj, 0, (size_t)n * 4 * sizeof(double), stream));
        const int nrow = Tp[0].xmax + 1;
        double* rwbar = (double*)ga((size_t)nrow * sizeof(double));
        CK(cudaMemsetAsync(rwbar, 0, (size_t)nrow * sizeof(double), stream));
        // tangent inputs of the dual W tables below a WGD/WGT: lane 0 = ∂/∂q (no ε tangent), lane 1 = ∂/∂ε of the child
        double* dez = (double*)ga((size_t)2 * n * sizeof(double));
        CK(cudaMemsetAsync(dez, 0, (size_t)2 * n * sizeof(double), stream));
        // stage 1: root, then every internal node, parents before children
        {
            Slab* rs = readable(0);
            int blocks = (N + 255) / 256;
            double* rpart = (double*)ga((size_t)nrow * blocks * sizeof(double));
            root_adjoint_kernel<<<blocks, 256, 0, stream>>>(rs->ell, Tp[0].cnt, rootw.p, weight.p, adj[0], adjs[0], rpart, blocks, nrow, N, ld);
            CK(cudaGetLastError());
            grad_reduce_kernel<<<1, 256, 0, stream>>>(rpart, blocks, nrow, rwbar);
            CK(cudaGetLastError());
        }
        const int nblk = (N + 127) / 128;
        double* npart = (double*)ga((size_t)6 * nblk * sizeof(double));
        std::vector<std::vector<double>> keep_host;          // staging of the one-hot tangent vectors (must outlive the copies)
        std::deque<DualTab> keep_dt;                         // (deques: stable addresses for the asynchronous copies)
        std::deque<int> keep_id;
        for (auto it = postorder.rbegin(); it != postorder.rend(); ++it) {
            const int u = *it;
            if (children[u].empty()) continue;
            NodeStep ns = make_step(u, true);
            AdjStep as;
            memset(&as, 0, sizeof(as));
            as.k = ns.k;
            for (int i = 0; i < ns.k; ++i) {
                const ChildRef& cr = ns.child[i];
                AdjChild& c = as.child[i];
                c.ell = cr.ell; c.scl = cr.scl; c.cnt = cr.cnt; c.Wt = cr.Wt; c.wdp = cr.wdp; c.id = cr.id;
                c.adj = adj[cr.id]; c.adj_scl = adjs[cr.id];
                c.e = h_nk[cr.id].e; c.cc = h_nk[cr.id].cc;
                if (kind[cr.id] == BLG_WGDAFTER) {
                    const HostTab& tb = tabs[cr.id];
                    const size_t wsz = (size_t)tb.wdim * tb.wdp;
                    double* dW = (double*)ga(2 * wsz * sizeof(double));
                    DualTab dt;
                    memset(&dt, 0, sizeof(dt));
                    dt.dWt = dW;
                    // (every WGD gets its own small staging buffers: nothing is reused, so nothing has to be waited for)
                    DualTab* ddt = (DualTab*)ga(sizeof(DualTab));
                    int* dl = (int*)ga(sizeof(int));
                    double* de1w = (double*)ga((size_t)2 * n * sizeof(double));
                    keep_host.emplace_back((size_t)2 * n, 0.0);
                    keep_host.back()[(size_t)n + cr.id] = 1.0;
                    keep_dt.push_back(dt);
                    keep_id.push_back(cr.id);
                    CK(cudaMemcpyAsync(de1w, keep_host.back().data(), (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
                    CK(cudaMemcpyAsync(ddt, &keep_dt.back(), sizeof(DualTab), cudaMemcpyHostToDevice, stream));
                    CK(cudaMemcpyAsync(dl, &keep_id.back(), sizeof(int), cudaMemcpyHostToDevice, stream));
                    Seeds<2> sd;
                    sd.kind[0] = SEED_Q; sd.node[0] = u; sd.kind[1] = SEED_NONE; sd.node[1] = -1;
                    size_t smem = (size_t)2 * std::max(tb.wdim, 1) * 3 * sizeof(double);
                    tables_dual_kernel<2><<<1, 128, smem, stream>>>(m, sd, dl, ddt, dez, de1w);
                    CK(cudaGetLastError());
                    c.dW = dW;
                    c.dW_stride = wsz;
                }
            }
            as.ell = ns.out; as.scl = ns.out_scl; as.cnt = ns.cnt;
            as.adj = adj[u]; as.adj_scl = adjs[u];
            as.ipow = ns.ipow; as.inv = h_nk[u].inv; as.partial = npart;
            const int need = Tp[u].xmax + 1;
            if (need <= 8) launch_adjoint<8>(as);
            else if (need <= 16) launch_adjoint<16>(as);
            else if (need <= 32) launch_adjoint<32>(as);
            else if (need <= 64) launch_adjoint<64>(as);
            else launch_adjoint<256>(as);
            CK(cudaGetLastError());
            adjoint_reduce_kernel<<<1, 256, 0, stream>>>(npart, nblk, as.k, as.child[0].id, as.k == 2 ? as.child[1].id : 0, sadj);
            CK(cudaGetLastError());
        }
        // stage 2: chain rule per parameter group
        Seeds<2>* dg = (Seeds<2>*)ga(groups.size() * sizeof(Seeds<2>));
        int* dpost = (int*)ga(postorder.size() * sizeof(int));
        double* dout = (double*)ga(groups.size() * 2 * sizeof(double));
        CK(cudaMemcpyAsync(dg, groups.data(), groups.size() * sizeof(Seeds<2>), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(dpost, postorder.data(), postorder.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        adjoint_params_kernel<<<(unsigned)groups.size(), 32, (size_t)4 * n * sizeof(double), stream>>>(m, dg, dpost, (int)postorder.size(), sadj, rwbar,
                                                                                                   nrow, weight_total, dout);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(res.data(), dout, groups.size() * 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (auto& t : temps) gfree(t.first, t.second);
    }

    // ---- gradient, reverse mode over node patterns (padjoint.cuh) -----------------------------------------------------
    void ensure_inverse(Shard& S, Link* lk, int j, Pat* Pu, Pat* Pc) {
        if (lk->d_inv_off[j]) return;
        std::vector<int> off(Pc->U + 2, 0), idx(std::max(Pu->U, 1)), hit;
        std::vector<int> px(Pu->U);
        for (int q = 0; q < Pu->U; ++q) { px[q] = Pc->fam2pat[Pu->rep[q]]; ++off[px[q] + 1]; }
        for (int c = 0; c < Pc->U; ++c) off[c + 1] += off[c];
        std::vector<int> fill(off.begin(), off.begin() + Pc->U + 1);
        for (int q = 0; q < Pu->U; ++q) idx[fill[px[q]]++] = q;
        // the warp kernel's list: first the child patterns with many node patterns, then those with many counts (which
        // only go there at tree depths with few patterns)
        std::vector<int> mid, mid2;
        for (int c = 0; c < Pc->U; ++c) {
            const int np = off[c + 1] - off[c];
            if (Pc->h_cnt[c] == 0) continue;
            if (np > BLG_PB_BLOCK) hit.push_back(c);
            else if (np > BLG_PB_HIT) mid.push_back(c);
            else if (Pc->h_cnt[c] >= BLG_PB_MID) mid2.push_back(c);
        }
        lk->nmid_many[j] = (int)mid.size();
        mid.insert(mid.end(), mid2.begin(), mid2.end());
        lk->nmid[j] = (int)mid.size();
        if (!mid.empty()) {
            CK(blgMalloc(&lk->d_mid[j], mid.size() * sizeof(int)));
            CK(cudaMemcpy(lk->d_mid[j], mid.data(), mid.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        CK(blgMalloc(&lk->d_inv_off[j], (size_t)(Pc->U + 1) * sizeof(int)));
        CK(cudaMemcpy(lk->d_inv_off[j], off.data(), (size_t)(Pc->U + 1) * sizeof(int), cudaMemcpyHostToDevice));
        // (uploaded as the inverse permutation: where phase A puts a node pattern's b̄)
        std::vector<int> pos(std::max(Pu->U, 1), 0);
        for (int k = 0; k < Pu->U; ++k) pos[idx[k]] = k;
        CK(blgMalloc(&lk->d_inv_idx[j], pos.size() * sizeof(int)));
        CK(cudaMemcpy(lk->d_inv_idx[j], pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
        lk->nhit[j] = (int)hit.size();
        if (!hit.empty()) {
            CK(blgMalloc(&lk->d_hit[j], hit.size() * sizeof(int)));
            CK(cudaMemcpy(lk->d_hit[j], hit.data(), hit.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        (void)S;
    }
    // weight of the families behind every root pattern (summed in family order)
    double* root_pattern_weights(Shard& S, Pat* P0) {
        if (P0->d_pw) return P0->d_pw;
        std::vector<double> w(S.N), pw(P0->ld, 0.0);
        CK(cudaMemcpyAsync(w.data(), S.weight.p, (size_t)S.N * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int f = 0; f < S.N; ++f) if (P0->fam2pat[f] >= 0) pw[P0->fam2pat[f]] += w[f];
        CK(blgMalloc(&P0->d_pw, P0->ld * sizeof(double)));
        CK(cudaMemcpy(P0->d_pw, pw.data(), P0->ld * sizeof(double), cudaMemcpyHostToDevice));
        return P0->d_pw;
    }
    template <int CAP> void launch_padj(const PAStep* d, int n0, int n, int blocks, double* part, int wblocks, int wcap) {
        const size_t smem = wblocks > 0 ? (size_t)4 * (8 * (size_t)wcap + 32 * 33) * sizeof(double) : 0;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(padj_parent_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        padj_parent_kernel<CAP><<<wblocks + blocks, 128, smem, stream>>>(d + n0, n, d_nk.p, pascal.p, pdim, part, wblocks, wcap);
        CK(cudaGetLastError());
    }
    template <int CAP> void launch_pchild(const PBStep* d, int n0, int n, int blocks, int hblocks, int mblocks, int cap) {
        padj_child_kernel<CAP><<<hblocks + mblocks + blocks, 128, (size_t)4 * cap * sizeof(double), stream>>>(d + n0, n, hblocks, mblocks, cap);
        CK(cudaGetLastError());
    }
    // res[2·i + lane] = ∂logL/∂(parameter of lane `lane` of group i); the caller has just run a full logpdf in the
    // node-pattern layout
    void gradient_padjoint(const std::vector<Seeds<2>>& groups, std::vector<double>& res) {
        const auto t_enter = std::chrono::steady_clock::now();
        if (!padj_constants) {
            double r[257];
            r[0] = 0.0;
            for (int q = 1; q <= 256; ++q) r[q] = 1.0 / (double)q;
            CK(cudaMemcpyToSymbol(c_rinv, r, sizeof(r)));
            padj_constants = true;
        }
        Shard& S = F;
        const int n = ncols_model;
        ModelDev m = model_dev();
        ensure_pats(S);
        std::vector<std::pair<void*, size_t>> temps;
        auto ga = [&](size_t b) { void* p = galloc(b); temps.push_back({p, b}); return p; };
        std::vector<double*> adj(n, nullptr);
        std::vector<int> depth(n, 0);
        int maxdepth = 0;
        for (auto it = postorder.rbegin(); it != postorder.rend(); ++it) {
            const int u = *it;
            if (children[u].empty()) continue;
            depth[u] = parent[u] >= 0 ? depth[parent[u]] + 1 : 0;
            maxdepth = std::max(maxdepth, depth[u]);
            Pat* Pu = pat_of(S, S.Tp[u]);
            adj[u] = (double*)ga(std::max<size_t>((size_t)(Pu->xmax + 1) * Pu->ld, 32) * sizeof(double));
        }
        double* sadj = (double*)ga((size_t)n * 4 * sizeof(double));
        CK(cudaMemsetAsync(sadj, 0, (size_t)n * 4 * sizeof(double), stream));
        const int nrow = S.Tp[0].xmax + 1;
        double* rwbar = (double*)ga((size_t)nrow * sizeof(double));
        CK(cudaMemsetAsync(rwbar, 0, (size_t)nrow * sizeof(double), stream));
        double* dez = (double*)ga((size_t)2 * n * sizeof(double));
        CK(cudaMemsetAsync(dez, 0, (size_t)2 * n * sizeof(double), stream));
        // root
        {
            Pat* P0 = pat_of(S, S.Tp[0]);
            Slab* rs = readable(S, 0);
            const int blocks = (P0->U + 255) / 256;
            double* rpart = (double*)ga((size_t)nrow * blocks * sizeof(double));
            root_padjoint_kernel<<<blocks, 256, 0, stream>>>(rs->ell, P0->d_cnt, rootw.p, root_pattern_weights(S, P0), adj[0], rpart, blocks, nrow,
                                                             P0->U, P0->ld);
            CK(cudaGetLastError());
            grad_reduce_kernel<<<1, 256, 0, stream>>>(rpart, blocks, nrow, rwbar);
            CK(cudaGetLastError());
        }
        // the steps of every depth
        std::vector<std::vector<double>> keep_host;
        std::deque<DualTab> keep_dt;
        std::deque<int> keep_id;
        std::vector<PAStep> sa;
        std::vector<PBStep> sb;
        struct Lev { int a0, a1, b0, b1, ablk, bblk, hblk, acap, bcap, wblk, wcap, mblk, midM; };
        std::vector<Lev> levs;
        int pblk = 0;
        for (int d = 0; d <= maxdepth; ++d) {
            Lev lv;
            memset(&lv, 0, sizeof(lv));
            lv.a0 = (int)sa.size(); lv.b0 = (int)sb.size();
            // a depth is as slow as its longest thread: the patterns with many counts go to the warp path of the same launch
            // (from 16 counts at a depth with few patterns, where nothing else hides the thread path's latency; from 40
            // where the thread path's throughput is what counts)
            int T = 257, msmin = 8;
            {
                long long total = 0;
                for (int u : postorder)
                    if (!children[u].empty() && depth[u] == d) total += pat_of(S, S.Tp[u])->U;
                T = total <= padj_warp_level ? 12 : padj_warp_big;
                msmin = total <= padj_warp_level ? 0 : 8;
                if (padj_warp_min > 0) T = padj_warp_min;
                long long ctotal = 0;
                for (int u : postorder)
                    if (!children[u].empty() && depth[u] == d && children[u].size() == 2)
                        for (int c : children[u]) if (!children[c].empty()) ctotal += pat_of(S, S.Tp[c])->U;
                lv.midM = ctotal <= padj_warp_level ? BLG_PB_MID : padj_mid_big;
            }
            for (auto it = postorder.rbegin(); it != postorder.rend(); ++it) {
                const int u = *it;
                if (children[u].empty() || depth[u] != d) continue;
                const int k = (int)children[u].size();
                require(tabs_valid && tabs[u].built && tabs[u].k == k, "gradient: ε/W tables are stale (call blg_rebuild)");
                std::vector<int> ord(tabs[u].mo, tabs[u].mo + k);
                Pat* Pu = pat_of(S, S.Tp[u]);
                if (Pu->U == 0) continue;
                Link* lk = get_link(S, u, ord);
                PAStep as;
                memset(&as, 0, sizeof(as));
                as.k = k; as.node = u; as.U = Pu->U; as.ld = (int)Pu->ld;
                as.blk0 = lv.ablk; as.nblk = (Pu->U + 127) / 128; as.pblk0 = pblk;
                lv.ablk += as.nblk; pblk += as.nblk;
                as.T = T; as.msmin = msmin;
                as.wblk0 = lv.wblk; as.wpblk0 = pblk;
                lv.acap = std::max(lv.acap, Pu->xmax + 1);
                if (Pu->xmax >= T) {
                    const int key = (msmin << 16) | T;
                    auto bi = lk->big.find(key);
                    if (bi == lk->big.end()) {
                        std::vector<int> bl;
                        Pat* Pc0 = pat_of(S, S.Tp[ord[0]]);
                        Pat* Pc1 = k == 2 ? pat_of(S, S.Tp[ord[1]]) : nullptr;
                        for (int q = 0; q < Pu->U; ++q) {
                            const int r = Pu->rep[q];
                            const int m0 = Pc0->h_cnt[Pc0->fam2pat[r]], m1 = Pc1 ? (int)Pc1->h_cnt[Pc1->fam2pat[r]] : 0;
                            if (padj_is_big(k, T, msmin, Pu->h_cnt[q], m0, m1)) bl.push_back(q);
                        }
                        int* dbl = nullptr;
                        CK(blgMalloc(&dbl, std::max<size_t>(bl.size(), 1) * sizeof(int)));
                        CK(cudaMemcpy(dbl, bl.data(), bl.size() * sizeof(int), cudaMemcpyHostToDevice));
                        bi = lk->big.emplace(key, std::make_pair(dbl, (int)bl.size())).first;
                    }
                    as.big = bi->second.first; as.nbig = bi->second.second;
                    as.nwblk = (as.nbig + 3) / 4;
                    lv.wblk += as.nwblk; pblk += as.nwblk;
                    if (as.nbig > 0) lv.wcap = std::max(lv.wcap, Pu->xmax + 1);
                }
                Slab* su = readable(S, u);
                as.ell = su->ell; as.scl = su->scl; as.cnt = Pu->d_cnt; as.adj = adj[u]; as.ipow = tabs[u].ipow;
                for (int j = 0; j < k; ++j) {
                    const int c = ord[j];
                    require(tabs[c].built && tabs[c].wdim >= S.Tp[c].xmax + 1 && tabs[c].Wt, "gradient: W table of a child does not match the profile (call blg_rebuild)");
                    Pat* Pc = pat_of(S, S.Tp[c]);
                    PAChild& pc = as.child[j];
                    if (!children[c].empty()) { Slab* sc_ = readable(S, c); pc.ell = sc_->ell; pc.scl = sc_->scl; }
                    pc.cnt = Pc->d_cnt; pc.pidx = lk->d_pidx[j]; pc.xcnt = lk->d_xcnt[j];
                    pc.Wt = tabs[c].Wt; pc.wdp = tabs[c].wdp; pc.id = c; pc.ld = (int)Pc->ld;
                    if (!children[c].empty()) {
                        if (Pc == Pu) pc.adj = adj[c];
                        else {
                            ensure_inverse(S, lk, j, Pu, Pc);
                            pc.bbar = (double*)ga(std::max<size_t>((size_t)(Pc->xmax + 1) * Pu->ld, 32) * sizeof(double));
                            PBStep bs;
                            memset(&bs, 0, sizeof(bs));
                            bs.Uc = Pc->U; bs.ldc = (int)Pc->ld; bs.ldp = (int)Pu->ld; bs.blk0 = lv.bblk; bs.hblk0 = lv.hblk;
                            bs.nhit = lk->nhit[j]; bs.wdp = tabs[c].wdp;
                            pc.pos = lk->d_inv_idx[j];
                            bs.cnt = Pc->d_cnt; bs.off = lk->d_inv_off[j]; bs.hit = lk->d_hit[j];
                            bs.bbar = pc.bbar; bs.Wt = tabs[c].Wt; bs.adj = adj[c];
                            bs.midM = lv.midM;
                            bs.mblk0 = lv.mblk; bs.nmid = lv.midM < 0x7fff ? lk->nmid[j] : lk->nmid_many[j]; bs.mid = lk->d_mid[j];
                            lv.mblk += (bs.nmid + 3) / 4;
                            lv.bblk += (Pc->U + 127) / 128;
                            lv.hblk += bs.nhit;
                            lv.bcap = std::max(lv.bcap, Pc->xmax + 1);
                            sb.push_back(bs);
                        }
                    }
                    if (kind[c] == BLG_WGDAFTER) {
                        const HostTab& tb = tabs[c];
                        const size_t wsz = (size_t)tb.wdim * tb.wdp;
                        double* dW = (double*)ga(2 * wsz * sizeof(double));
                        DualTab dt;
                        memset(&dt, 0, sizeof(dt));
                        dt.dWt = dW;
                        DualTab* ddt = (DualTab*)ga(sizeof(DualTab));
                        int* dl = (int*)ga(sizeof(int));
                        double* de1w = (double*)ga((size_t)2 * n * sizeof(double));
                        keep_host.emplace_back((size_t)2 * n, 0.0);
                        keep_host.back()[(size_t)n + c] = 1.0;
                        keep_dt.push_back(dt);
                        keep_id.push_back(c);
                        CK(cudaMemcpyAsync(de1w, keep_host.back().data(), (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
                        CK(cudaMemcpyAsync(ddt, &keep_dt.back(), sizeof(DualTab), cudaMemcpyHostToDevice, stream));
                        CK(cudaMemcpyAsync(dl, &keep_id.back(), sizeof(int), cudaMemcpyHostToDevice, stream));
                        Seeds<2> sd;
                        sd.kind[0] = SEED_Q; sd.node[0] = u; sd.kind[1] = SEED_NONE; sd.node[1] = -1;
                        size_t smem = (size_t)2 * std::max(tb.wdim, 1) * 3 * sizeof(double);
                        tables_dual_kernel<2><<<1, 128, smem, stream>>>(m, sd, dl, ddt, dez, de1w);
                        CK(cudaGetLastError());
                        pc.dW = dW;
                        pc.dW_stride = wsz;
                    }
                }
                sa.push_back(as);
            }
            lv.a1 = (int)sa.size(); lv.b1 = (int)sb.size();
            levs.push_back(lv);
        }
        PAStep* dsa = (PAStep*)ga(std::max<size_t>(sa.size(), 1) * sizeof(PAStep));
        PBStep* dsb = (PBStep*)ga(std::max<size_t>(sb.size(), 1) * sizeof(PBStep));
        double* part = (double*)ga((size_t)std::max(pblk, 1) * 6 * sizeof(double));
        if (!sa.empty()) CK(cudaMemcpyAsync(dsa, sa.data(), sa.size() * sizeof(PAStep), cudaMemcpyHostToDevice, stream));
        if (!sb.empty()) CK(cudaMemcpyAsync(dsb, sb.data(), sb.size() * sizeof(PBStep), cudaMemcpyHostToDevice, stream));
        const auto t_launch = std::chrono::steady_clock::now();
        std::vector<std::pair<std::string, cudaEvent_t>> marks;
        auto mark = [&](const std::string& what) {
            if (!pdebug) return;
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            CK(cudaEventRecord(e, stream));
            marks.push_back({what, e});
        };
        mark("start");
        int lvno = 0;
        for (const Lev& lv : levs) {
            const std::string tag = "depth " + std::to_string(lvno++) + " ";
            if (lv.a1 > lv.a0) {
                const int na = lv.a1 - lv.a0;
                const int wcap = round_up(lv.wcap + 1, 2);
                if (lv.acap <= 8) launch_padj<8>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                else if (lv.acap <= 16) launch_padj<16>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                else if (lv.acap <= 32) launch_padj<32>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                else if (lv.acap <= 64) launch_padj<64>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                else if (lv.acap <= 128) launch_padj<128>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                else launch_padj<256>(dsa, lv.a0, na, lv.ablk, part, lv.wblk, wcap);
                mark(tag + "A: warp blocks " + std::to_string(lv.wblk) + " + thread blocks " + std::to_string(lv.ablk) + ", cap " + std::to_string(lv.acap));
            }
            if (lv.b1 > lv.b0) {
                const int nb = lv.b1 - lv.b0;
                const int bcap = std::min(lv.bcap, lv.midM);          // (the thread path's columns: below the warp path's bound)
                const int cap = round_up(lv.bcap + 1, 2);
                if (bcap <= 16) launch_pchild<16>(dsb, lv.b0, nb, lv.bblk, lv.hblk, lv.mblk, cap);
                else if (bcap <= 32) launch_pchild<32>(dsb, lv.b0, nb, lv.bblk, lv.hblk, lv.mblk, cap);
                else if (bcap <= 64) launch_pchild<64>(dsb, lv.b0, nb, lv.bblk, lv.hblk, lv.mblk, cap);
                else if (bcap <= 128) launch_pchild<128>(dsb, lv.b0, nb, lv.bblk, lv.hblk, lv.mblk, cap);
                else launch_pchild<256>(dsb, lv.b0, nb, lv.bblk, lv.hblk, lv.mblk, cap);
                mark(tag + "B: block " + std::to_string(lv.hblk) + " + warp blocks " + std::to_string(lv.mblk) + " + thread blocks " + std::to_string(lv.bblk));
            }
        }
        if (!sa.empty()) {
            padj_reduce_kernel<<<(unsigned)sa.size(), 256, 0, stream>>>(part, dsa, sadj);
            CK(cudaGetLastError());
            mark("reduce");
        }
        // stage 2: chain rule per parameter group
        Seeds<2>* dg = (Seeds<2>*)ga(groups.size() * sizeof(Seeds<2>));
        int* dpost = (int*)ga(postorder.size() * sizeof(int));
        double* dout = (double*)ga(groups.size() * 2 * sizeof(double));
        CK(cudaMemcpyAsync(dg, groups.data(), groups.size() * sizeof(Seeds<2>), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(dpost, postorder.data(), postorder.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        adjoint_params_kernel<<<(unsigned)groups.size(), 32, (size_t)4 * n * sizeof(double), stream>>>(m, dg, dpost, (int)postorder.size(), sadj, rwbar,
                                                                                                   nrow, weight_total, dout);
        CK(cudaGetLastError());
        mark("parameters");
        CK(cudaMemcpyAsync(res.data(), dout, groups.size() * 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (pdebug) {
            const double host_ms = std::chrono::duration<double, std::milli>(t_launch - t_enter).count();
            fprintf(stderr, "[blg padjoint] host: steps built in %.3f ms\n", host_ms);
            for (size_t i = 1; i < marks.size(); ++i) {
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second));
                fprintf(stderr, "[blg padjoint] %-40s %8.1f us\n", marks[i].first.c_str(), ms * 1e3);
            }
            for (auto& mk : marks) cudaEventDestroy(mk.second);
        }
        for (auto& t : temps) gfree(t.first, t.second);
    }

    // g: [λ_1..λ_n, μ_1..μ_n, q (ascending wgd id), η]  (asvector layout, model.jl:541-546) or [λ, μ] for cr
    void gradient(double* g, int len, double* logpdf_out, bool cr) {
        require(have_topology && have_params && tabs_valid, "gradient: model not set (topology, params, rebuild)");
        int nn = 0;
        while (nn < ncols_model && present(nn) && (kind[nn] == BLG_ROOT || kind[nn] == BLG_SP)) ++nn;
        std::vector<int> wg;
        for (int i = nn; i < ncols_model; ++i) {
            if (!present(i)) continue;
            require(kind[i] != BLG_ROOT && kind[i] != BLG_SP, "gradient: non-WGD ids must be 1..n (call reindex!)");
            if (kind[i] == BLG_WGD || kind[i] == BLG_WGT) wg.push_back(i);
        }
        const int P = cr ? 2 : 2 * nn + (int)wg.size() + 1;
        require(len >= P, "gradient: output vector too short");
        for (int i = 0; i < P; ++i) g[i] = 0.0;
        if (Ntot == 0) {
            double ll = 0.0;
            if (comm) {                          // an empty shard still takes part in the sums
                logpdf(0, 1, result.p);
                CK(cudaMemcpyAsync(&ll, result.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
            }
            if (logpdf_out) *logpdf_out = ll;
            return;
        }
        require(H.N == 0, "gradient: families with a root count above the fast bound are not supported by the gradient kernels "
                          "(raise BLG_HEAVY_MIN up to 256 to keep them in the table-based shard)");
        // fresh evaluation that leaves the committed/proposal state alone (the reference uses a fresh L, model.jl:395)
        // (the gradient kernels work on one column per FAMILY: the evaluation below is a private per-family sweep)
        // (the reverse-mode sweep runs on the node patterns; the forward-mode passes and BLG_GRAD_PAT=0 work on one column
        // per FAMILY: their evaluation below is then a private per-family sweep)
        std::vector<Col> saved = Tp;
        const bool pm = pat_mode;
        const bool padj = pat_mode && grad_pat && grad_mode == 1 && adjoint_eligible();
        if (!padj) pat_mode = false;
        try {
            if (padj && pattern_fast_path() && flag_cache < 0) fetch_flag();
            logpdf(0, 1, result.p);
            double ll;
            CK(cudaMemcpyAsync(&ll, result.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (logpdf_out) *logpdf_out = ll;
            // the per-node scalars live on the device; the gradient's host-side step descriptions need a copy
            if (!padj) {
                h_nk.resize(ncols_model);
                CK(cudaMemcpyAsync(h_nk.data(), d_nk.p, (size_t)ncols_model * sizeof(NodeK), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
            }
            // parameter groups (two tangent lanes each) and where their results go
            std::vector<Seeds<2>> groups;
            std::vector<std::pair<int, int>> dest;
            auto add_group = [&](int k0, int n0, int d0, int k1, int n1, int d1) {
                Seeds<2> sd;
                sd.kind[0] = k0; sd.node[0] = n0; sd.kind[1] = k1; sd.node[1] = n1;
                groups.push_back(sd);
                dest.push_back({d0, d1});
            };
            if (cr) {
                add_group(SEED_LAM_ALL, -1, 0, SEED_MU_ALL, -1, 1);
            } else {
                for (int b = 0; b < nn; ++b) {
                    if (model_kind == BLG_BRANCH_MODEL && b == 0) continue;     // root rates are unused by Branch models
                    add_group(SEED_LAM, b, b, SEED_MU, b, nn + b);
                }
                for (size_t j = 0; j < wg.size(); ++j) add_group(SEED_Q, wg[j], 2 * nn + (int)j, SEED_NONE, -1, -1);
                add_group(SEED_ETA, 0, P - 1, SEED_NONE, -1, -1);
            }
            std::vector<double> res(2 * groups.size(), 0.0);
            if (padj) {
                gradient_padjoint(groups, res);
            } else if (grad_mode == 1 && adjoint_eligible()) {
                gradient_adjoint(groups, res);
            } else {
                for (size_t i = 0; i < groups.size(); ++i) tangent_pass(groups[i], &res[2 * i]);
            }
            for (size_t i = 0; i < groups.size(); ++i) {
                if (dest[i].first >= 0) g[dest[i].first] = res[2 * i];
                if (dest[i].second >= 0) g[dest[i].second] = res[2 * i + 1];
            }
        } catch (...) { Tp = saved; pat_mode = pm; throw; }
        Tp = saved;
        pat_mode = pm;
    }

    void root_reduce(double* d_out) {
        require(tabs_valid, "logpdf: ε/W tables are stale (call blg_rebuild)");
        require(!children[0].empty(), "logpdf: the root is a leaf");
        const int nrow = root_rows();
        if (!rootw_valid || rootw_rows != nrow) {   // normally built by tables_kernel together with the root's tables
            rootw.ensure(2 * ((size_t)nrow + 1));
            ModelDev m = model_dev();
            root_tables_kernel<<<1, 128, 0, stream>>>(m, rootw.p, H.N > 0 ? rootw.p + nrow + 1 : nullptr, nrow);
            CK(cudaGetLastError());
            rootw_valid = true;
            rootw_rows = nrow;
            last_launches += 1;
        }
        const int bF = (F.N + 255) / 256, bH = (H.N + 7) / 8;
        partial.ensure((size_t)(bF + bH));
        if (!d_ticket.p) { d_ticket.ensure(1); CK(cudaMemsetAsync(d_ticket.p, 0, sizeof(unsigned int), stream)); }
        double* marker = result.p + 8;
        if (F.N > 0) {
            Slab* s = readable(F, 0);
            if (pat_mode) {
                ensure_pats(F);
                Pat* P0 = pat_of(F, F.Tp[0]);
                if (!P0->d_fam2pat) {
                    CK(blgMalloc(&P0->d_fam2pat, F.ld * sizeof(int)));
                    CK(cudaMemcpy(P0->d_fam2pat, P0->fam2pat.data(), (size_t)F.N * sizeof(int), cudaMemcpyHostToDevice));
                }
                root_pattern_kernel<<<bF, 256, 0, stream>>>(s->ell, s->scl, P0->d_cnt, P0->d_fam2pat, rootw.p, F.weight.p, F.fam_ll.p, partial.p, F.N,
                                                            P0->ld, d_ticket.p, (unsigned)(bF + bH), d_flags.p, d_ran.p,
                                                            root_fused ? pat_ll.p : nullptr, P0->nvery, marker, d_out);
                last_pattern_bytes += 8.0 * P0->colsum + 4.0 * P0->U + 12.0 * F.N;
            } else {
                root_kernel<<<bF, 256, 0, stream>>>(s->ell, s->scl, F.Tp[0].cnt, rootw.p, F.weight.p, F.fam_ll.p, partial.p, F.N, F.ld, d_ticket.p,
                                                    (unsigned)(bF + bH), d_flags.p, marker, d_out);
            }
            CK(cudaGetLastError());
            last_bytes += 8.0 * F.Tp[0].colsum + 4.0 * F.N + 8.0 * F.N;
            last_launches += 1;
        }
        if (H.N > 0) {
            Slab* s = readable(H, 0);
            root_ragged_kernel<<<bH, 256, 0, stream>>>(s->ell, s->scl, H.Tp[0].cnt, H.Tp[0].off, rootw.p + nrow + 1, rootw.p, H.weight.p, H.fam_ll.p,
                                                       partial.p, bF, H.N, d_ticket.p, (unsigned)(bF + bH), d_flags.p, marker, d_out);
            CK(cudaGetLastError());
            last_launches += 1;
        }
    }
    // any post-order is valid; visit the child with the larger counts LAST so that its column is still in shared
    // memory when the parent (the very next step) needs it
    std::vector<int> resident_order(const Shard& S) {
        std::vector<int> nodes;
        std::vector<std::pair<int, size_t>> stk;
        std::vector<std::vector<int>> ord(ncols_model);
        stk.push_back({0, 0});
        while (!stk.empty()) {
            int u = stk.back().first;
            if (stk.back().second == 0 && ord[u].empty() && !children[u].empty()) {
                ord[u] = children[u];
                std::stable_sort(ord[u].begin(), ord[u].end(), [&](int a, int b) { return colsum_of(S, a) < colsum_of(S, b); });
            }
            if (stk.back().second < ord[u].size()) { int c = ord[u][stk.back().second++]; stk.push_back({c, 0}); }
            else { if (!children[u].empty()) nodes.push_back(u); stk.pop_back(); }
        }
        return nodes;
    }
    // force_general: run the general kernel on the fast family set whatever the range flag says (the caller has seen the flag)
    void logpdf(int mode, int node, double* d_out, bool force_general = false) {
        // mode 0 full, 1 from node, 2 root only
        require(have_topology, "logpdf: no topology");
        if (Ntot == 0) {                                                                  // profile.jl:56
            last_bytes = 0.0;
            last_launches = 0;
            last_pattern_bytes = last_pattern_cols = last_family_cols = 0.0;
            root_fused = false;
            CK(cudaMemsetAsync(d_out, 0, sizeof(double), stream));
            CK(cudaMemsetAsync(result.p + 8, 0, sizeof(double), stream));
            allreduce_sum(d_out, 1);             // (an empty shard still takes part in the sum)
            return;
        }
        require(have_params && tabs_valid, "logpdf: ε/W tables are stale (call blg_rebuild)");
        if (timing) for (auto& e : ev) if (!e) CK(cudaEventCreate(&e));
        const bool general = force_general || kernel_version == 1 || flag_cache == 0;
        const bool graphable = graph_on && !pdebug && pat_mode && H.N == 0 && F.N > 0 && !profile_sweep;
        CallGraph* E = nullptr;
        if (graphable) {
            if (call_graphs.size() > 1024) drop_graphs();
            E = &call_graphs[std::make_tuple(mode, node, (const void*)d_out)];
            const bool rootw_ok = rootw_valid && rootw_rows == root_rows();
            const bool match = E->stage != 0 && E->mem_epoch == g_mem_epoch.load() && E->state_epoch == state_epoch &&
                               E->general == general && E->timing == timing && E->rootw_ok == rootw_ok;
            if (!match) {
                if (E->exec) { cudaGraphExecDestroy(E->exec); E->exec = nullptr; }
                E->stage = 0;
            }
            if (E->stage == 1) {                 // second call in this state: capture it
                cudaGraph_t g = nullptr;
                bool ok = cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
                if (ok) {
                    g_capturing = true;
                    try { logpdf_launches(mode, node, d_out, general); } catch (const std::exception&) { ok = false; }
                    g_capturing = false;
                    if (cudaStreamEndCapture(stream, &g) != cudaSuccess) ok = false;
                    ok = ok && g != nullptr && E->mem_epoch == g_mem_epoch.load() && E->state_epoch == state_epoch;
                    if (ok && cudaGraphInstantiate(&E->exec, g, 0) != cudaSuccess) { ok = false; E->exec = nullptr; }
                    if (g) cudaGraphDestroy(g);
                }
                (void)cudaGetLastError();
                if (ok) {
                    E->stage = 2;
                    E->bytes = last_bytes; E->prune_bytes = last_prune_bytes; E->pattern_bytes = last_pattern_bytes;
                    E->pattern_cols = last_pattern_cols; E->family_cols = last_family_cols; E->launches = last_launches;
                    E->root_fused = root_fused; E->rootw_valid = rootw_valid; E->rootw_rows = rootw_rows;
                    ++graph_captures;
                } else {
                    E->stage = -1;               // this state does not capture: launch by launch until it changes
                }
            }
            if (E->stage == 2) {
                last_bytes = E->bytes; last_prune_bytes = E->prune_bytes; last_pattern_bytes = E->pattern_bytes;
                last_pattern_cols = E->pattern_cols; last_family_cols = E->family_cols; last_launches = E->launches;
                root_fused = E->root_fused; rootw_valid = E->rootw_valid; rootw_rows = E->rootw_rows;
                CK(cudaGraphLaunch(E->exec, stream));
                ++graph_replays;
                allreduce_sum(d_out, 1);
                if (timing) { CK(cudaEventRecord(ev[2], stream)); ev_valid = true; }
                return;
            }
        }
        logpdf_launches(mode, node, d_out, general);
        if (E && E->stage == 0) {                // (epochs AFTER the call: its first run allocates slabs, patterns and links)
            E->stage = 1;
            E->mem_epoch = g_mem_epoch.load(); E->state_epoch = state_epoch;
            E->general = general; E->timing = timing;
            E->rootw_ok = rootw_valid && rootw_rows == root_rows();     // (what the next call in this state starts from)
        }
        allreduce_sum(d_out, 1);                 // Σ over the ranks' shards, on the same stream (8 bytes over NVLink)
        if (timing) { CK(cudaEventRecord(ev[2], stream)); ev_valid = true; }
    }
    // the launches of one likelihood call up to the rank's own scalar in d_out (everything a graph of the call holds)
    void logpdf_launches(int mode, int node, double* d_out, bool general) {
        last_bytes = 0.0;
        last_launches = 0;
        last_pattern_bytes = last_pattern_cols = last_family_cols = 0.0;
        root_fused = false;
        if (timing) record_timing(ev[0]);
        require(pat_mode || H.N == 0, "families at or above the heavy bound need the node-pattern evaluation (BLG_DEDUP=0 and the gradient do not support them)");
        if (H.N > 0) ensure_pats(F);
        for (Shard* S : {&F, &H}) {
            if (S->N == 0) continue;
            require((int)S->Tp.size() >= 1, "logpdf: no profile columns");
            std::vector<int> nodes;
            if (mode == 0 && (!order_resident || pattern_layout(*S))) {
                for (int u : postorder) if (!children[u].empty()) nodes.push_back(u);
            } else if (mode == 0) {
                nodes = resident_order(*S);
            } else if (mode == 1) {
                int u = node - 1;
                require(present(u), "logpdf_from: node id not in the model");
                while (u >= 0) { if (!children[u].empty()) nodes.push_back(u); u = parent[u]; }
            }
            if (nodes.empty()) continue;
            if (S == &H) general_sweep(H, nodes, 0, false);     // (the fast set's accounting already covers every family)
            else if (pat_mode) psweep(nodes, general);
            else if (kernel_version == 1) general_sweep(F, nodes, 0, true);
            else {
                // the table-free kernel, and right behind it the general one, which only works when the device-side
                // range flag says the first could not (no host round trip decides this)
                sweep(nodes);
                general_sweep(F, nodes, 1, false);
            }
        }
        if (timing) record_timing(ev[1]);
        last_prune_bytes = last_bytes;
        root_reduce(d_out);
    }
    // the pattern path runs OPTIMISTICALLY on the table-free kernel: the kernels leave at once when the device-side range
    // flag of the last ε/W build says no, and the root kernel hands the flag back next to the result.
    bool pattern_fast_path() const { return pat_mode && F.N > 0 && kernel_version != 1; }
    void fetch_flag() {
        int fl = 1;
        CK(cudaMemcpyAsync(&fl, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        flag_cache = fl ? 1 : 0;
    }
};

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define BLG_ENTER(h)                                                    \
    if (!(h)) return 1;                                                 \
    try {                                                               \
        CK(cudaSetDevice((h)->device));
#define BLG_LEAVE(h)                                                    \
    }                                                                   \
    catch (const std::exception& e) { (h)->err = e.what(); return 2; }  \
    return 0;

extern "C" {

int blg_create(int device, void* stream, blg_handle** out) {
    if (!out) return 1;
    *out = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU path");
        if (device < 0 || device >= ndev) throw std::runtime_error("blg_create: bad device index");
        CK(cudaSetDevice(device));
        auto* h = new blg_handle();
        h->device = device;
        if (stream) h->stream = (cudaStream_t)stream;
        else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
        h->result.ensure(16);
        { int smc = 0; if (cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && smc > 0) h->sm_count = smc; }
        {   // 1/n! (table-free kernel) and 1/n (both pruning kernels); per device, idempotent
            const int n = BLG_FAST_MAXCOUNT + 72;
            std::vector<double> fi(n);
            long double acc = 1.0L;
            for (int i = 0; i < n; ++i) {
                if (i > 0) acc *= (long double)i;
                fi[i] = (double)(1.0L / acc);
            }
            CK(cudaMemcpyToSymbol(c_invfact, fi.data(), n * sizeof(double)));
            const int ni = BLG_MAXCOUNT + 9;
            std::vector<double> iv(ni);
            for (int i = 0; i < ni; ++i) iv[i] = i > 0 ? 1.0 / (double)i : 0.0;
            CK(cudaMemcpyToSymbol(c_inv, iv.data(), ni * sizeof(double)));
        }
        if (const char* kv = getenv("BLG_KERNEL")) h->kernel_version = atoi(kv) == 1 ? 1 : 2;
        if (const char* o = getenv("BLG_ORDER")) h->order_resident = atoi(o) != 0;
        if (const char* o = getenv("BLG_PROF")) h->profile_sweep = atoi(o);
        if (const char* o = getenv("BLG_GRAD")) h->grad_mode = (strcmp(o, "forward") == 0) ? 0 : 1;
        if (const char* o = getenv("BLG_GRAD_PAT")) h->grad_pat = atoi(o) != 0;
        if (const char* o = getenv("BLG_FAMILY_ORDER")) h->family_order_always = atoi(o) != 0;
        if (const char* o = getenv("BLG_PAT_TIEBREAK")) h->pattern_tiebreak = atoi(o) != 0;
        if (const char* o = getenv("BLG_PADJ_WARP")) h->padj_warp_min = std::max(0, atoi(o));
        if (const char* o = getenv("BLG_PADJ_LEVEL")) h->padj_warp_level = atoll(o);
        if (const char* o = getenv("BLG_PADJ_BIG")) h->padj_warp_big = std::max(1, atoi(o));
        if (const char* o = getenv("BLG_PB_MIDM")) h->padj_mid_big = std::max(BLG_PB_MID, atoi(o));
        if (const char* w = getenv("BLG_WPB")) { int v = atoi(w); if (v == 1 || v == 2 || v == 4) h->sweep_wpb = v; }
        if (const char* sc = getenv("BLG_SLICE")) h->slice_cap = std::max(8, std::min(224, atoi(sc)));
        // families whose root count reaches this go to the heavy shard; above BLG_FAST_MAXCOUNT+1 the fast shard leaves the
        // table-free kernel's range (general kernel on SoA slabs) but stays inside the table-based gradient kernels' (<= 256)
        if (const char* pd = getenv("BLG_PDEBUG")) h->pdebug = atoi(pd);
        if (const char* gr = getenv("BLG_GRAPH")) h->graph_on = atoi(gr) != 0;
        if (const char* fn = getenv("BLG_FAN")) h->fan_classes = atoi(fn) != 0;
        if (const char* sd = getenv("BLG_SIDE")) h->side_min = atoi(sd);
        if (const char* dd = getenv("BLG_DEDUP")) h->pat_mode = h->pat_default = atoi(dd) != 0;
        if (const char* hm = getenv("BLG_HEAVY_MIN")) h->heavy_min = std::max(1, std::min(256, atoi(hm)));
        *out = h;
    } catch (const std::exception& e) { g_create_error = e.what(); return 2; }
    return 0;
}
int blg_destroy(blg_handle* h) { delete h; return 0; }
const char* blg_last_error(blg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// ----------------------------------------------------------------------------------------------------
// perm: on entry the rows of X (caller's pattern indices) that form the set; on return the same rows in device order
static void family_order(const int32_t* X, int ncols, std::vector<int>& perm) {
    // device order: families that share a warp should have similar count vectors at the heavy nodes, because
    // every loop of a warp runs to the warp's maximum count.  k-d ordering: recursively split the family set
    // at a multiple of 32 along the heavy column with the largest spread, until 32 families remain per leaf.
    const int npatterns = (int)perm.size();
    if (npatterns > 32) {
        // candidate columns: largest Σx² first, identical columns (wgd / wgdafter copies) dropped
        std::vector<double> s1(ncols, 0.0), s2(ncols, 0.0);
        for (int f : perm) {
            const int32_t* x = X + (size_t)f * ncols;
            for (int c = 0; c < ncols; ++c) { s1[c] += x[c]; s2[c] += (double)x[c] * x[c]; }
        }
        std::vector<int> byw(ncols);
        for (int c = 0; c < ncols; ++c) byw[c] = c;
        std::stable_sort(byw.begin(), byw.end(), [&](int a, int b) { return s2[a] > s2[b]; });
        std::vector<int> keys;
        for (int c : byw) {
            bool dup = false;
            for (int k : keys) if (s1[k] == s1[c] && s2[k] == s2[c]) { dup = true; break; }
            if (!dup && s2[c] > 0.0) keys.push_back(c);
            if (keys.size() == 32) break;
        }
        const int nk = (int)keys.size();
        if (nk > 0) {
            // K is indexed by position in the ORIGINAL perm; `idx` is what gets permuted
            std::vector<float> K((size_t)npatterns * nk);
            for (int i = 0; i < npatterns; ++i)
                for (int k = 0; k < nk; ++k) K[(size_t)i * nk + k] = (float)X[(size_t)perm[i] * ncols + keys[k]];
            std::vector<int> idx(npatterns);
            for (int i = 0; i < npatterns; ++i) idx[i] = i;
            std::vector<std::pair<int, int>> stk;
            stk.push_back({0, npatterns});
            std::vector<double> m1(nk), m2(nk);
            while (!stk.empty()) {
                auto seg = stk.back();
                stk.pop_back();
                const int lo = seg.first, hi = seg.second, n = hi - lo;
                if (n <= 32) continue;
                std::fill(m1.begin(), m1.end(), 0.0);
                std::fill(m2.begin(), m2.end(), 0.0);
                for (int i = lo; i < hi; ++i) {
                    const float* kf = &K[(size_t)idx[i] * nk];
                    for (int k = 0; k < nk; ++k) { m1[k] += kf[k]; m2[k] += (double)kf[k] * kf[k]; }
                }
                int best = 0;
                double bestv = -1.0;
                for (int k = 0; k < nk; ++k) {
                    double var = m2[k] / n - (m1[k] / n) * (m1[k] / n);
                    if (var > bestv) { bestv = var; best = k; }
                }
                if (bestv <= 0.0) continue;             // identical on every key: nothing to gain
                int half = ((n / 2 + 31) / 32) * 32;
                if (half >= n) half = n / 2;
                std::nth_element(idx.begin() + lo, idx.begin() + lo + half, idx.begin() + hi, [&](int a, int b) {
                    float va = K[(size_t)a * nk + best], vb = K[(size_t)b * nk + best];
                    return va != vb ? va > vb : a < b;
                });
                stk.push_back({lo + half, hi});
                stk.push_back({lo, lo + half});
            }
            std::vector<int> out(npatterns);
            for (int i = 0; i < npatterns; ++i) out[i] = perm[idx[i]];
            perm.swap(out);
        }
    }
}

// upload one family set: counts [column][family] (uint16), weights, column tables
// runs body(t, lo, hi) over [0, n) in contiguous chunks on up to 16 host threads
static void host_parallel(size_t n, size_t grain, const std::function<void(unsigned, size_t, size_t)>& body) {
    unsigned nt = std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, n / std::max<size_t>(grain, 1)));
    if (nt <= 1) { body(0u, (size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t lo = std::min(n, (size_t)t * chunk), hi = std::min(n, lo + chunk);
        th.emplace_back([&body, t, lo, hi]() { body(t, lo, hi); });
    }
    for (auto& x : th) x.join();
}

// X: row-major [pattern][column] (colmajor = false) or column-major [column][pattern]
static void upload_shard(blg_handle* h, Shard& S, const int32_t* X, bool colmajor, int npatterns, const int64_t* weight, int ncols,
                         const std::vector<char>& isheavy) {
    S.N = (int)S.perm.size();
    S.ld = ((size_t)S.N + 127) / 128 * 128;
    if (S.N == 0) return;
    const size_t ld = S.ld;
    std::vector<uint16_t> stage((size_t)ncols * ld, 0);
    std::vector<Col> cols(ncols);
    const int N = S.N;
    const int* perm = S.perm.data();
    const bool ragged = S.ragged;
    const int heavy_min = h->heavy_min;
    // one column per task: conversion to 16 bits in device order, the column's bounds and byte count
    host_parallel((size_t)ncols, 1, [&](unsigned, size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; ++c) {
            int mx = 0;
            double cs = 0.0;
            uint16_t* dst = stage.data() + c * ld;
            int fastmax = 0;                 // largest count of this column among the families below the heavy bound
            for (int f = 0; f < N; ++f) {
                const int32_t v = colmajor ? X[c * (size_t)npatterns + perm[f]] : X[(size_t)perm[f] * ncols + c];
                dst[f] = (uint16_t)v;
                if (!isheavy[perm[f]]) fastmax = std::max(fastmax, (int)v);
                mx = std::max(mx, (int)v);
                cs += (double)v + 1.0;
            }
            if (!ragged) {
                // the fast set holds every family, but a heavy family's entries are evaluated there (as node patterns) only up
                // to the count the other families reach at this column (at least 32): the column's bound — and with it the
                // range of the table-free merge, whose scalings grow with the bound — stays what the fast families need
                const int lmax = std::min(heavy_min - 1, std::max(fastmax, 32));
                cols[c].lmax = lmax;
                mx = 0;
                for (int f = 0; f < N; ++f) if ((int)dst[f] <= lmax) mx = std::max(mx, (int)dst[f]);
            }
            cols[c].xmax = mx;
            cols[c].colsum = cs;
        }
    });
    uint16_t* dcnt = nullptr;
    CK(blgMalloc(&dcnt, stage.size() * sizeof(uint16_t)));
    S.count_allocs.push_back(dcnt);
    CK(cudaMemcpyAsync(dcnt, stage.data(), stage.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    for (int c = 0; c < ncols; ++c) cols[c].cnt = dcnt + (size_t)c * ld;
    if (S.ragged) {
        // ragged slabs: family f's column starts at off[f] and holds exactly count+1 doubles
        std::vector<uint32_t> ostage((size_t)ncols * ld, 0);
        for (int c = 0; c < ncols; ++c) {
            if (cols[c].colsum >= 4.0e9) throw std::runtime_error("set_profiles: heavy column too large for 32-bit offsets");
            const uint16_t* src = stage.data() + (size_t)c * ld;
            uint32_t* dst = ostage.data() + (size_t)c * ld;
            uint32_t run = 0;
            for (int f = 0; f < S.N; ++f) { dst[f] = run; run += (uint32_t)src[f] + 1u; }
        }
        uint32_t* doff = nullptr;
        CK(blgMalloc(&doff, ostage.size() * sizeof(uint32_t)));
        S.off_allocs.push_back(doff);
        CK(cudaMemcpy(doff, ostage.data(), ostage.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        for (int c = 0; c < ncols; ++c) cols[c].off = doff + (size_t)c * ld;
    }
    std::vector<double> w(ld, 0.0);
    S.weight_total = 0.0;
    for (int f = 0; f < S.N; ++f) { w[f] = (double)weight[S.perm[f]]; S.weight_total += w[f]; }
    S.weight.ensure(ld);
    S.fam_ll.ensure(ld);
    CK(cudaMemcpyAsync(S.weight.p, w.data(), ld * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(S.fam_ll.p, 0, ld * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));       // the staging vectors die here
    if (!S.ragged) S.h_counts = std::move(stage);   // (kept on the host: the node patterns are built from it)
    S.Tc = cols;
    S.Tp = cols;
}

static int set_profiles_impl(blg_handle* h, const int32_t* X, bool colmajor, const int64_t* weight, int ncols, int npatterns) {
    BLG_ENTER(h)
    const auto t_0 = std::chrono::steady_clock::now();
    h->require(ncols >= 0 && npatterns >= 0, "set_profiles: negative size");
    CK(cudaStreamSynchronize(h->stream));
    h->drop_graphs();
    ++h->state_epoch;
    h->F.release();
    h->H.release();
    h->F.ragged = false;
    h->H.ragged = true;
    h->Ntot = npatterns;
    h->loc.clear();
    h->tabs_valid = false;
    h->rootw_valid = false;
    for (auto& tb : h->tabs) blg_handle::free_tab(tb);
    if (npatterns == 0) return 0;
    h->require(X && weight && ncols >= 1, "set_profiles: null pointer");
    // split by the family's largest count (its root count, for an extended profile)
    std::vector<int> rootc(npatterns, 0);
    std::vector<int> bad(16, 0);
    if (colmajor) {
        host_parallel((size_t)npatterns, 1 << 14, [&](unsigned t, size_t f0, size_t f1) {
            for (int c = 0; c < ncols; ++c) {
                const int32_t* x = X + (size_t)c * npatterns;
                for (size_t f = f0; f < f1; ++f) {
                    if (x[f] < 0 || x[f] > BLG_MAXCOUNT) bad[t] = 1;
                    rootc[f] = std::max(rootc[f], (int)x[f]);
                }
            }
        });
    } else {
        host_parallel((size_t)npatterns, 1 << 14, [&](unsigned t, size_t f0, size_t f1) {
            for (size_t f = f0; f < f1; ++f) {
                int mx = 0;
                const int32_t* x = X + f * ncols;
                for (int c = 0; c < ncols; ++c) {
                    if (x[c] < 0 || x[c] > BLG_MAXCOUNT) bad[t] = 1;
                    mx = std::max(mx, (int)x[c]);
                }
                rootc[f] = mx;
            }
        });
    }
    for (int b : bad) if (b) throw std::runtime_error("set_profiles: count outside [0, " + std::to_string(BLG_MAXCOUNT) + "]");
    h->F.perm.resize(npatterns);
    for (int f = 0; f < npatterns; ++f) {
        h->F.perm[f] = f;                                          // the fast set holds EVERY family (its entries below the heavy bound)
        if (rootc[f] >= h->heavy_min) h->H.perm.push_back(f);      // the heavy set: the families with an entry at or above it
    }
    const auto t_1 = std::chrono::steady_clock::now();
    // the k-d family order serves the kernels that give every FAMILY a lane (BLG_DEDUP=0, the forward-mode gradient of trees
    // with polytomies, BLG_GRAD_PAT=0: 2.4x faster with it); node patterns have their own order per node (ensure_pats), the
    // families stay as the caller gave them unless BLG_FAMILY_ORDER=1 asks for the order (0.5-1 s of host time for 10^6 rows)
    if (!h->pat_mode || h->family_order_always) {
        if (colmajor) {
            std::vector<int32_t> rows((size_t)npatterns * ncols);
            for (int c = 0; c < ncols; ++c)
                for (int f = 0; f < npatterns; ++f) rows[(size_t)f * ncols + c] = X[(size_t)c * npatterns + f];
            family_order(rows.data(), ncols, h->F.perm);
        } else {
            family_order(X, ncols, h->F.perm);
        }
    }
    const auto t_2 = std::chrono::steady_clock::now();
    // heavy shard: heaviest first (the general kernel hands families out one per warp), launch groups by root count
    std::stable_sort(h->H.perm.begin(), h->H.perm.end(), [&](int a, int b) { return rootc[a] > rootc[b]; });
    {
        static const int bounds[] = {4095, 2047, 1023, 511, 255, 127};
        size_t pos = 0;
        for (int bi = 0; bi < 6 && pos < h->H.perm.size(); ++bi) {
            const int lo = bi + 1 < 6 ? bounds[bi + 1] : -1;      // this class holds root counts in (lo, bounds[bi]]
            size_t end = pos;
            while (end < h->H.perm.size() && rootc[h->H.perm[end]] > lo) ++end;
            if (end > pos) h->H.classes.push_back({(int)pos, (int)end, rootc[h->H.perm[pos]]});
            pos = end;
        }
    }
    std::vector<char> isheavy(npatterns, 0);
    for (int f : h->H.perm) isheavy[f] = 1;
    upload_shard(h, h->F, X, colmajor, npatterns, weight, ncols, isheavy);
    upload_shard(h, h->H, X, colmajor, npatterns, weight, ncols, isheavy);
    const auto t_3 = std::chrono::steady_clock::now();
    h->loc.assign(npatterns, 0);
    h->hloc.assign(npatterns, -1);
    for (int i = 0; i < h->F.N; ++i) h->loc[h->F.perm[i]] = i;
    h->H2F.assign(h->H.N, 0);
    for (int i = 0; i < h->H.N; ++i) { h->hloc[h->H.perm[i]] = i; h->H2F[i] = h->loc[h->H.perm[i]]; }
    for (auto& kv : h->hpat) blgFree(kv.second);
    h->hpat.clear();
    // Pascal triangle for the merge tables of the gradient kernels (fast shard bounds only)
    int gmax = 0;
    for (const Col& c : h->F.Tp) gmax = std::max(gmax, c.xmax);
    h->maxcount = gmax;
    int pd = gmax + 1;
    if (pd > h->pdim) {
        h->pascal.ensure((size_t)pd * pd);
        h->pdim = pd;
        pascal_kernel<<<1, 256, 2 * (size_t)pd * sizeof(double), h->stream>>>(h->pascal.p, pd);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    if (h->pdebug) {
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[blg set_profiles] %d patterns x %d columns (%s): bounds %.1f ms | family order %.1f ms | staging + upload %.1f ms | total %.1f ms\n",
                npatterns, ncols, colmajor ? "column-major" : "row-major", ms(t_0, t_1), ms(t_1, t_2), ms(t_2, t_3), ms(t_0, std::chrono::steady_clock::now()));
    }
    BLG_LEAVE(h)
}
int blg_set_profiles(blg_handle* h, const int32_t* X, const int64_t* weight, int ncols, int npatterns) {
    return set_profiles_impl(h, X, false, weight, ncols, npatterns);
}
int blg_set_profiles_columns(blg_handle* h, const int32_t* Xc, const int64_t* weight, int ncols, int npatterns) {
    return set_profiles_impl(h, Xc, true, weight, ncols, npatterns);
}

int blg_set_topology(blg_handle* h, int ncols, const int32_t* parent, const uint8_t* kind, const int32_t* child_pos) {
    BLG_ENTER(h)
    h->require(ncols >= 1 && parent && kind && child_pos, "set_topology: bad arguments");
    h->drop_graphs();
    ++h->state_epoch;
    h->ncols_model = ncols;
    h->parent.assign(ncols, -1);
    h->kind.assign(kind, kind + ncols);
    h->child_pos.assign(child_pos, child_pos + ncols);
    for (int i = 0; i < ncols; ++i) {
        if (kind[i] == BLG_ABSENT) continue;
        h->require(kind[i] <= BLG_WGDAFTER, "set_topology: unknown node kind");
        h->require(parent[i] >= 0 && parent[i] <= ncols, "set_topology: parent id out of range");
        h->parent[i] = parent[i] - 1;
    }
    h->build_children();
    for (int i = 0; i < ncols; ++i) {
        if (!h->present(i)) continue;
        uint8_t kd = h->kind[i];
        if (kd == BLG_WGD || kd == BLG_WGT || kd == BLG_WGDAFTER) h->require(h->children[i].size() == 1, "set_topology: wgd/wgt/wgdafter nodes have exactly one child");
        if (kd == BLG_WGDAFTER) h->require(h->kind[h->parent[i]] == BLG_WGD || h->kind[h->parent[i]] == BLG_WGT, "set_topology: wgdafter must sit below a wgd/wgt node");
    }
    size_t n = (size_t)ncols;
    std::vector<int> off(n + 1, 0), idx;
    for (int i = 0; i < ncols; ++i) { off[i + 1] = off[i] + (int)h->children[i].size(); for (int c : h->children[i]) idx.push_back(c); }
    h->d_kind.ensure(n); h->d_parent.ensure(n); h->d_child_off.ensure(n + 1); h->d_child_idx.ensure(std::max<size_t>(idx.size(), n));
    CK(cudaMemcpyAsync(h->d_kind.p, h->kind.data(), n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_parent.p, h->parent.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_child_off.p, off.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!idx.empty()) CK(cudaMemcpyAsync(h->d_child_idx.p, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->d_post.ensure(h->postorder.size());
    CK(cudaMemcpyAsync(h->d_post.p, h->postorder.data(), h->postorder.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_topology = true;
    h->have_params = false;
    h->tabs_valid = false;
    h->rootw_valid = false;
    for (auto& tb : h->tabs) tb.built = false;
    BLG_LEAVE(h)
}

int blg_set_params(blg_handle* h, int model_kind, const double* lambda, const double* mu, const double* q, const double* t, double eta) {
    BLG_ENTER(h)
    h->require(h->have_topology, "set_params: call blg_set_topology first");
    h->require(model_kind == BLG_NODE_MODEL || model_kind == BLG_BRANCH_MODEL, "set_params: bad model kind");
    h->require(lambda && mu && q && t, "set_params: null pointer");
    int n = h->ncols_model;
    h->model_kind = model_kind;
    h->lam.assign(lambda, lambda + n);
    h->mu.assign(mu, mu + n);
    h->q.assign(q, q + n);
    h->t.assign(t, t + n);
    h->eta = eta;
    h->have_params = true;
    h->rootw_valid = false;
    h->upload_params();
    BLG_LEAVE(h)
}

int blg_rebuild(blg_handle* h, const int32_t* nodes, int n) {
    BLG_ENTER(h)
    h->rebuild(nodes, n);
    BLG_LEAVE(h)
}

static int logpdf_sync(blg_handle* h, int mode, int node, double* out) {
    BLG_ENTER(h)
    h->require(out != nullptr, "logpdf: null output");
    // with a communicator every rank must issue the same number of collectives: no optimistic run + rerun, the range
    // flag is looked at first (one 4-byte read per ε/W build)
    if (h->comm && h->Ntot > 0 && h->pattern_fast_path() && h->flag_cache < 0 && h->tabs_valid) h->fetch_flag();
    h->logpdf(mode, node, h->result.p);
    double r[9] = {0};
    CK(cudaMemcpyAsync(r, h->result.p, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!h->comm && h->Ntot > 0 && h->pattern_fast_path() && h->flag_cache != 0) {
        h->flag_cache = (r[8] != 0.0) ? 0 : 1;
        if (h->flag_cache == 0) {            // outside the table-free kernel's range: once more, with the general kernel
            h->logpdf(mode, node, h->result.p, true);
            CK(cudaMemcpyAsync(r, h->result.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
    }
    *out = r[0];
    BLG_LEAVE(h)
}
static int logpdf_async(blg_handle* h, int mode, int node, double* d_out) {
    BLG_ENTER(h)
    h->require(d_out != nullptr, "logpdf: null output");
    // nobody looks at the result on the host here: the range flag must be known before the kernels are chosen
    if (h->Ntot > 0 && h->pattern_fast_path() && h->flag_cache < 0 && h->tabs_valid) h->fetch_flag();
    h->logpdf(mode, node, d_out);
    BLG_LEAVE(h)
}
int blg_logpdf_full(blg_handle* h, double* out) { return logpdf_sync(h, 0, 1, out); }
int blg_logpdf_from(blg_handle* h, int node, double* out) { return logpdf_sync(h, 1, node, out); }
int blg_logpdf_root(blg_handle* h, double* out) { return logpdf_sync(h, 2, 1, out); }
int blg_logpdf_full_async(blg_handle* h, double* d_out) { return logpdf_async(h, 0, 1, d_out); }
int blg_logpdf_from_async(blg_handle* h, int node, double* d_out) { return logpdf_async(h, 1, node, d_out); }
int blg_logpdf_root_async(blg_handle* h, double* d_out) { return logpdf_async(h, 2, 1, d_out); }

// with a communicator: Σ over the ranks of logL (inside the evaluation) and of the gradient vector (here, P doubles)
static void gradient_allreduce(blg_handle* h, double* g, int P) {
    if (!h->comm || P <= 0) return;
    h->d_red.ensure((size_t)P);
    CK(cudaMemcpyAsync(h->d_red.p, g, (size_t)P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->allreduce_sum(h->d_red.p, (size_t)P);
    CK(cudaMemcpyAsync(g, h->d_red.p, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}
int blg_gradient(blg_handle* h, double* g, int len, double* logpdf_out) {
    BLG_ENTER(h)
    h->require(g != nullptr, "gradient: null output");
    ++h->state_epoch;                    // (the gradient may switch the fast set between node patterns and family columns)
    h->gradient(g, len, logpdf_out, false);
    ++h->state_epoch;
    gradient_allreduce(h, g, len);
    BLG_LEAVE(h)
}
int blg_gradient_cr(blg_handle* h, double* g2, double* logpdf_out) {
    BLG_ENTER(h)
    h->require(g2 != nullptr, "gradient_cr: null output");
    ++h->state_epoch;
    h->gradient(g2, 2, logpdf_out, true);
    ++h->state_epoch;
    gradient_allreduce(h, g2, 2);
    BLG_LEAVE(h)
}

// ---- multi-GPU: one process per GPU, the library owns the NCCL communicator -------------------------------------------
int blg_comm_unique_id(uint8_t* id128) {
    if (!id128) return 1;
    try {
        if (!g_nccl.load()) throw std::runtime_error("NCCL not available (" + g_nccl.why + "); set BLG_NCCL_LIB");
        NcclApi::UniqueId id;
        NCK(g_nccl.GetUniqueId(&id));
        memcpy(id128, id.internal, 128);
    } catch (const std::exception& e) { g_create_error = e.what(); return 2; }
    return 0;
}
int blg_comm_init(blg_handle* h, int nranks, int rank, const uint8_t* id128) {
    BLG_ENTER(h)
    h->require(id128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad arguments");
    h->require(h->comm == nullptr, "comm_init: this handle already has a communicator");
    ++h->state_epoch;
    if (!g_nccl.load()) throw std::runtime_error("NCCL not available (" + g_nccl.why + "); set BLG_NCCL_LIB");
    NcclApi::UniqueId id;
    memcpy(id.internal, id128, 128);
    NCK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->comm_rank = rank;
    h->comm_size = nranks;
    BLG_LEAVE(h)
}
int blg_comm_size(blg_handle* h) { return h ? h->comm_size : -1; }

int blg_commit(blg_handle* h) {
    BLG_ENTER(h)
    ++h->state_epoch;
    h->F.Tc = h->F.Tp;
    h->H.Tc = h->H.Tp;
    BLG_LEAVE(h)
}
int blg_revert(blg_handle* h) {
    BLG_ENTER(h)
    ++h->state_epoch;
    h->F.Tp = h->F.Tc;
    h->H.Tp = h->H.Tc;
    BLG_LEAVE(h)
}
int blg_extend(blg_handle* h, int i) {
    BLG_ENTER(h)
    ++h->state_epoch;
    if (h->Ntot == 0) return 0;
    for (Shard* S : {&h->F, &h->H}) {
        if (S->N == 0) continue;
        h->require(i >= 1 && i <= (int)S->Tp.size(), "extend: column out of range");
        Col c = S->Tp[i - 1];          // counts (and the ragged offsets, which follow the counts) are shared with column i
        c.slab.reset();
        S->Tp.push_back(c);
        S->Tp.push_back(c);
    }
    BLG_LEAVE(h)
}
int blg_shrink(blg_handle* h, int i) {
    BLG_ENTER(h)
    ++h->state_epoch;
    if (h->Ntot == 0) return 0;
    for (Shard* S : {&h->F, &h->H}) {
        if (S->N == 0) continue;
        h->require(i >= 1 && i + 1 <= (int)S->Tp.size(), "shrink: column out of range");
        S->Tp.erase(S->Tp.begin() + (i - 1), S->Tp.begin() + (i + 1));
    }
    BLG_LEAVE(h)
}

int blg_get_column(blg_handle* h, int pattern, int node, int which, double* out, int cap, int* rows) {
    BLG_ENTER(h)
    h->require(pattern >= 0 && pattern < h->Ntot, "get_column: pattern out of range");
    // every family lives in the fast set; its entries at or above the heavy bound are held by the heavy set
    h->require(node >= 1 && node <= (int)(which ? h->F.Tp : h->F.Tc).size(), "get_column: node out of range");
    CK(cudaStreamSynchronize(h->stream));
    uint16_t x;
    CK(cudaMemcpy(&x, (which ? h->F.Tp : h->F.Tc)[node - 1].cnt + h->loc[pattern], sizeof(uint16_t), cudaMemcpyDeviceToHost));
    bool heavy = false;
    if (h->hloc[pattern] >= 0) {          // a family of the heavy set: is THIS entry one the fast set's patterns cover?
        h->ensure_pats(h->F);
        heavy = h->pat_of(h->F, (which ? h->F.Tp : h->F.Tc)[node - 1])->fam2pat[h->loc[pattern]] < 0;
    }
    Shard& S = heavy ? h->H : h->F;
    const int slot = heavy ? h->hloc[pattern] : h->loc[pattern];
    std::vector<Col>& T = which ? S.Tp : S.Tc;
    Col& c = T[node - 1];
    // rows of the column as the reference holds it: count bound of this node over ALL patterns + 1
    int xm = 0;
    if (h->F.N > 0) xm = std::max(xm, (which ? h->F.Tp : h->F.Tc)[node - 1].xmax);
    if (h->H.N > 0) xm = std::max(xm, (which ? h->H.Tp : h->H.Tc)[node - 1].xmax);
    if (rows) *rows = xm + 1;
    for (int n = 0; n < cap; ++n) out[n] = -INFINITY;
    bool leaf = h->have_topology && node - 1 < h->ncols_model && h->present(node - 1) && h->children[node - 1].empty();
    if (leaf) { if (x < cap) out[x] = 0.0; return 0; }     // model.jl:409-410
    if (!c.slab) return 0;
    int sc;
    CK(cudaMemcpy(&sc, c.slab->scl + slot, sizeof(int), cudaMemcpyDeviceToHost));
    size_t base = (size_t)slot, stride = S.ld;
    if (c.slab->pat) {                   // node patterns: the column of this family's sub-pattern at the node
        h->ensure_pats(S);
        Pat* P = h->pat_of(S, c);
        base = (size_t)P->fam2pat[slot];
        stride = P->ld;
        CK(cudaMemcpy(&sc, c.slab->scl + base, sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (S.ragged) {
        h->require(c.off != nullptr, "get_column: column has no offsets");
        uint32_t o;
        CK(cudaMemcpy(&o, c.off + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        base = o;
        stride = 1;
    }
    for (int n = 0; n <= (int)x && n < cap; ++n) {
        double v;
        CK(cudaMemcpy(&v, c.slab->ell + base + (size_t)n * stride, sizeof(double), cudaMemcpyDeviceToHost));
        out[n] = std::log(v) + (double)sc * 0.6931471805599453;
    }
    BLG_LEAVE(h)
}
int blg_get_family_logpdf(blg_handle* h, double* out, int n) {
    BLG_ENTER(h)
    h->require(n <= h->Ntot, "get_family_logpdf: n exceeds the pattern count");
    CK(cudaStreamSynchronize(h->stream));
    if (n > 0) {
        std::vector<double> tf(h->F.N), th(h->H.N);
        if (h->F.N) CK(cudaMemcpy(tf.data(), h->F.fam_ll.p, (size_t)h->F.N * sizeof(double), cudaMemcpyDeviceToHost));
        if (h->H.N) CK(cudaMemcpy(th.data(), h->H.fam_ll.p, (size_t)h->H.N * sizeof(double), cudaMemcpyDeviceToHost));
        for (int f = 0; f < n; ++f) out[f] = h->hloc[f] >= 0 ? th[h->hloc[f]] : tf[h->loc[f]];
    }
    BLG_LEAVE(h)
}
int blg_get_eps(blg_handle* h, double* eps_top, double* eps_node, int ncols) {
    BLG_ENTER(h)
    h->require(ncols <= h->ncols_model && h->d_eps0.p, "get_eps: nothing built");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(eps_top, h->d_eps0.p, (size_t)ncols * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(eps_node, h->d_eps1.p, (size_t)ncols * sizeof(double), cudaMemcpyDeviceToHost));
    BLG_LEAVE(h)
}
int blg_get_W(blg_handle* h, int node, double* out, int cap, int* dim) {
    BLG_ENTER(h)
    h->require(node >= 1 && node <= (int)h->tabs.size(), "get_W: node out of range");
    HostTab& tb = h->tabs[node - 1];
    if (dim) *dim = tb.wdim;
    h->require(tb.Wt != nullptr && tb.built, "get_W: no table");
    h->require(cap >= tb.wdim * tb.wdim, "get_W: buffer too small");
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> wt((size_t)tb.wdim * tb.wdp);
    CK(cudaMemcpy(wt.data(), tb.Wt, wt.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int n = 0; n < tb.wdim; ++n)
        for (int j = 0; j < tb.wdim; ++j) out[(size_t)n * tb.wdim + j] = wt[(size_t)j * tb.wdp + n];
    BLG_LEAVE(h)
}
int blg_ncols(blg_handle* h, int which) {
    if (!h) return -1;
    const Shard& S = h->F.N > 0 ? h->F : h->H;
    return (int)(which ? S.Tp.size() : S.Tc.size());
}
int blg_npatterns(blg_handle* h) { return h ? h->Ntot : -1; }
int blg_family_order(const int32_t* X, int ncols, int npatterns, int32_t* perm_out) {
    if (!X || !perm_out || ncols < 0 || npatterns < 0) return 1;
    try {
        std::vector<int> perm(npatterns);
        for (int f = 0; f < npatterns; ++f) perm[f] = f;
        family_order(X, ncols, perm);
        for (int f = 0; f < npatterns; ++f) perm_out[f] = perm[f];
    } catch (...) { return 1; }
    return 0;
}
int blg_last_algorithmic(blg_handle* h, double* bytes_total, double* bytes_prune, int* launches) {
    if (!h) return 1;
    if (bytes_total) *bytes_total = h->last_bytes;
    if (bytes_prune) *bytes_prune = h->last_prune_bytes;
    if (launches) *launches = h->last_launches;
    return 0;
}
// how many launches of the table-free sweep kernel / of the general kernel have done the work so far (cumulative; both
// kernels decide on the device, from the range flag of the last ε/W build, which of them runs), and how the patterns
// are split between the fast (SoA, 32 families per warp) and the heavy (ragged, one warp per family) family sets
int blg_kernel_runs(blg_handle* h, unsigned int* sweep_runs, unsigned int* general_runs, int* nfast, int* nheavy) {
    BLG_ENTER(h)
    unsigned int r[4] = {0, 0, 0, 0};
    if (h->d_ran.p) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(r, h->d_ran.p, sizeof(r), cudaMemcpyDeviceToHost));
    }
    if (sweep_runs) *sweep_runs = r[0];
    if (general_runs) *general_runs = r[1];
    if (nfast) *nfast = h->F.N - h->H.N;      // (the fast set holds every family; these are the ones entirely inside it)
    if (nheavy) *nheavy = h->H.N;
    BLG_LEAVE(h)
}
// what the last likelihood call did on the fast family set: bytes of pattern columns actually written + read, node steps ×
// distinct patterns evaluated, node steps × families they stand for (diagnostic; the algorithmic figure of
// blg_last_algorithmic stays SURVEY §8(d)'s per-family definition)
int blg_graph_stats(blg_handle* h, unsigned long long* captures, unsigned long long* replays) {
    if (!h) return 1;
    if (captures) *captures = h->graph_captures;
    if (replays) *replays = h->graph_replays;
    return 0;
}
int blg_pattern_stats(blg_handle* h, double* bytes_moved, double* pattern_columns, double* family_columns) {
    if (!h) return 1;
    if (bytes_moved) *bytes_moved = h->last_pattern_bytes;
    if (pattern_columns) *pattern_columns = h->last_pattern_cols;
    if (family_columns) *family_columns = h->last_family_cols;
    return 0;
}
// rand(model, N; condition) (src/sim.jl:17-27) at the count level, on the device (csrc/sampler.cuh), for the handle's current
// topology and parameters (blg_set_topology + blg_set_params; no profiles needed).  condition: 0 none, 1 at least one gene
// (the reference's default), 2 at least one gene in every clade below the root (rootclades); max_leaf / max_rootsum: also
// reject families with a larger leaf count / total count (-1: no bound).  counts_out: N × nleaves, row-major, leaf columns
// by ascending leaf id (leaf_ids_out).  draws_out: how many families were drawn to accept N.
int blg_rand_profiles(blg_handle* h, int N, unsigned long long seed, int condition, int max_leaf, int max_rootsum,
                      int32_t* leaf_ids_out, int* nleaves_out, uint16_t* counts_out, long long* draws_out) {
    BLG_ENTER(h)
    h->require(h->have_topology && h->have_params, "rand_profiles: set the topology and the parameters first");
    h->require(N >= 0 && counts_out && leaf_ids_out && nleaves_out, "rand_profiles: bad arguments");
    // pre-order over the model, children in the reference's iteration order
    std::vector<int> pre, pos(h->ncols_model, -1);
    {
        std::vector<int> stk{0};
        while (!stk.empty()) {
            const int u = stk.back(); stk.pop_back();
            pos[u] = (int)pre.size(); pre.push_back(u);
            for (auto it = h->children[u].rbegin(); it != h->children[u].rend(); ++it) stk.push_back(*it);
        }
    }
    std::vector<int> leaf_ids;
    for (int i = 0; i < h->ncols_model; ++i) if (h->present(i) && h->children[i].empty()) leaf_ids.push_back(i);
    const int nleaves = (int)leaf_ids.size();
    std::vector<int> leafcol(h->ncols_model, -1);
    for (int k = 0; k < nleaves; ++k) { leafcol[leaf_ids[k]] = k; leaf_ids_out[k] = leaf_ids[k] + 1; }
    *nleaves_out = nleaves;
    const int nclades = (int)h->children[0].size();
    h->require(nclades <= 30, "rand_profiles: too many clades below the root");
    std::vector<SimNode> sn(pre.size());
    for (size_t k = 0; k < pre.size(); ++k) {
        const int u = pre[k];
        SimNode& s_ = sn[k];
        memset(&s_, 0, sizeof(s_));
        s_.leaf = leafcol[u];
        if (u == 0) { s_.parent = -1; s_.clade = -1; continue; }
        const int p = h->parent[u];
        s_.parent = pos[p];
        s_.clade = (p == 0) ? (int)(std::find(h->children[0].begin(), h->children[0].end(), u) - h->children[0].begin()) : sn[pos[p]].clade;
        s_.ret = h->kind[p] == BLG_WGD ? 2 : h->kind[p] == BLG_WGT ? 3 : 0;
        s_.q = s_.ret ? h->q[p] : 0.0;
        const double t = h->t[u];
        if (t > 0.0) {
            // getλμ(nonwgdparent(node), nonwgdchild(c)) as sim.jl:108 calls it (node.jl:220-230); ϕ, ψ of utils.jl:8-9
            const int a = h->nonwgdparent_h(p), b = h->nonwgdchild_h(u);
            double l = h->lam[b], m = h->mu[b];
            if (h->model_kind == BLG_NODE_MODEL) { l = 0.5 * (h->lam[a] + h->lam[b]); m = 0.5 * (h->mu[a] + h->mu[b]); }
            if (std::fabs(l - m) <= 1.4901161193847656e-8 * std::max(std::fabs(l), std::fabs(m))) {
                s_.phi = s_.psi = l * t / (1.0 + l * t);
            } else {
                const double e = std::exp(t * (l - m));
                s_.phi = m * (e - 1.0) / (l * e - m);
                s_.psi = (l / m) * s_.phi;
            }
        }
    }
    long long draws = 0;
    if (N > 0) {
        const int chunk = 1 << 19;
        DevBuf<SimNode> d_nodes;
        DevBuf<int> d_cnt;
        DevBuf<uint16_t> d_leaves, d_out;
        DevBuf<unsigned int> d_keep, d_bsum;
        d_nodes.ensure(sn.size());
        d_cnt.ensure(sn.size() * (size_t)chunk);
        d_leaves.ensure((size_t)chunk * nleaves);
        d_out.ensure((size_t)N * nleaves);
        const int nb = (chunk + 1023) / 1024;
        d_keep.ensure((size_t)chunk);
        d_bsum.ensure((size_t)nb + 4);
        CK(cudaMemcpyAsync(d_nodes.p, sn.data(), sn.size() * sizeof(SimNode), cudaMemcpyHostToDevice, h->stream));
        long long have = 0;
        int rounds = 0;
        while (have < N) {
            h->require(++rounds <= 100000, "rand_profiles: the condition is (almost) never met");
            rand_counts_kernel<<<(chunk + 255) / 256, 256, 0, h->stream>>>(d_nodes.p, (int)sn.size(), nleaves, nclades, h->eta, seed, (unsigned long long)draws, chunk,
                                                                         condition, max_leaf, max_rootsum, d_cnt.p, d_leaves.p, d_keep.p);
            sim_block_sums<<<nb, 256, 0, h->stream>>>(d_keep.p, chunk, d_bsum.p);
            sim_scan_sums<<<1, 32, 0, h->stream>>>(d_bsum.p, nb, d_bsum.p + nb);
            sim_compact<<<nb, 256, 0, h->stream>>>(d_keep.p, chunk, d_bsum.p, d_leaves.p, nleaves, d_out.p, have, (long long)N);
            CK(cudaGetLastError());
            unsigned int got = 0;
            CK(cudaMemcpyAsync(&got, d_bsum.p + nb, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            have += got;
            draws += chunk;
        }
        CK(cudaMemcpy(counts_out, d_out.p, (size_t)N * nleaves * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        d_nodes.free_(); d_cnt.free_(); d_leaves.free_(); d_out.free_(); d_keep.free_(); d_bsum.free_();
    }
    if (draws_out) *draws_out = draws;
    BLG_LEAVE(h)
}

// profile(t, l, df), first half (src/model.jl:40-41): the site patterns of a count table and their multiplicities, in
// order of first appearance.  counts: N × T, row-major (one row per family, one column per taxon), host memory.
// rep_out[k] = index of the first family with pattern k, weight_out[k] = how many families share it, k < *npatterns.
// Runs on the given device (hash grouping, csrc/ingest.cuh); the host only sorts the representatives.
extern "C++" {
template <class V>
static int site_patterns_impl(int device, const V* counts, int N, int T, int32_t* rep_out, int64_t* weight_out, int* npatterns) {
    if (!counts || !rep_out || !weight_out || !npatterns || N < 0 || T < 1) return 1;
    try {
        *npatterns = 0;
        if (N == 0) return 0;
        CK(cudaSetDevice(device));
        unsigned int nslots = 1024;
        while (nslots < 2u * (unsigned int)N) nslots <<= 1;
        DevBuf<V> d_counts;
        DevBuf<unsigned int> d_u;           // [slot_rep | slot_min | slot_cnt] (nslots each), fam_slot (N), out_rep, out_cnt (N each), counter
        d_counts.ensure((size_t)N * T);
        d_u.ensure((size_t)3 * nslots + (size_t)3 * N + 16);
        unsigned int* slot_rep = d_u.p;
        unsigned int* slot_min = slot_rep + nslots;
        unsigned int* slot_cnt = slot_min + nslots;
        unsigned int* fam_slot = slot_cnt + nslots;
        unsigned int* out_rep = fam_slot + N;
        unsigned int* out_cnt = out_rep + N;
        unsigned int* counter = out_cnt + N;
        CK(cudaMemcpy(d_counts.p, counts, (size_t)N * T * sizeof(V), cudaMemcpyHostToDevice));
        CK(cudaMemset(slot_rep, 0xff, (size_t)2 * nslots * sizeof(unsigned int)));      // rep and min: 0xffffffff
        CK(cudaMemset(slot_cnt, 0, (size_t)nslots * sizeof(unsigned int)));
        CK(cudaMemset(counter, 0, sizeof(unsigned int)));
        const int nb = (N + 255) / 256;
        ingest_insert_kernel<V><<<nb, 256>>>(d_counts.p, N, T, slot_rep, fam_slot, nslots - 1);
        ingest_group_kernel<<<nb, 256>>>(N, fam_slot, slot_min, slot_cnt);
        ingest_compact_kernel<<<(nslots + 255) / 256, 256>>>(nslots, slot_min, slot_cnt, out_rep, out_cnt, counter);
        CK(cudaGetLastError());
        unsigned int U = 0;
        CK(cudaMemcpy(&U, counter, sizeof(unsigned int), cudaMemcpyDeviceToHost));
        std::vector<unsigned int> rep(U), cnt(U);
        CK(cudaMemcpy(rep.data(), out_rep, (size_t)U * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cnt.data(), out_cnt, (size_t)U * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        d_counts.free_();
        d_u.free_();
        std::vector<unsigned int> ord(U);
        for (unsigned int i = 0; i < U; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](unsigned int a, unsigned int b) { return rep[a] < rep[b]; });
        for (unsigned int i = 0; i < U; ++i) { rep_out[i] = (int32_t)rep[ord[i]]; weight_out[i] = (int64_t)cnt[ord[i]]; }
        *npatterns = (int)U;
    } catch (const std::exception& e) { g_create_error = e.what(); return 2; }
    return 0;
}
}  // extern "C++"
int blg_site_patterns(int device, const int32_t* counts, int N, int T, int32_t* rep_out, int64_t* weight_out, int* npatterns) {
    return site_patterns_impl<int32_t>(device, counts, N, T, rep_out, weight_out, npatterns);
}
int blg_site_patterns_u16(int device, const uint16_t* counts, int N, int T, int32_t* rep_out, int64_t* weight_out, int* npatterns) {
    return site_patterns_impl<uint16_t>(device, counts, N, T, rep_out, weight_out, npatterns);
}
int blg_set_timing(blg_handle* h, int on) {
    if (!h) return 1;
    h->timing = on != 0;
    h->ev_valid = false;
    return 0;
}
int blg_last_timing(blg_handle* h, float* prune_ms, float* root_ms) {
    BLG_ENTER(h)
    h->require(h->ev_valid, "last_timing: no timed likelihood call yet (blg_set_timing)");
    CK(cudaEventSynchronize(h->ev[2]));
    if (prune_ms) CK(cudaEventElapsedTime(prune_ms, h->ev[0], h->ev[1]));
    if (root_ms) CK(cudaEventElapsedTime(root_ms, h->ev[1], h->ev[2]));
    BLG_LEAVE(h)
}
int blg_sync(blg_handle* h) {
    BLG_ENTER(h)
    CK(cudaStreamSynchronize(h->stream));
    BLG_LEAVE(h)
}
// fp64 FMA throughput of this GPU, measured: the denominator of the "fraction of the fp64 roofline" figure
int blg_fp64_peak(blg_handle* h, double* tflops) {
    BLG_ENTER(h)
    h->require(tflops != nullptr, "fp64_peak: null output");
    DevBuf<double> sink;
    sink.ensure(1);
    cudaEvent_t a = nullptr, b = nullptr;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const int iters = 8192, blocks = h->sm_count * 8, threads = 256;
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(sink.p, 256, 1.0);                 // warm-up
    CK(cudaEventRecord(a, h->stream));
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(sink.p, iters, 1.0);
    CK(cudaEventRecord(b, h->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    sink.free_();
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    BLG_LEAVE(h)
}

}  // extern "C"
