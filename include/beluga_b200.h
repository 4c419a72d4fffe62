/*
 * beluga_b200.h — C ABI of the B200-native gene-family likelihood library.
 *
 * Drop-in boundary for ONE hot path of arzwa/Beluga.jl: the Csűrös–Miklós conditional-survival
 * pruning likelihood behind logpdf!(model, profile) for the DLWGD model.  The reference has no
 * FFI boundary of its own (pure Julia); each entry below names the Julia method it backs
 * (file:line under the reference tree).  The Julia-side `ccall` stubs are in INTEGRATION.md and
 * julia/BelugaB200.jl; the Python mirror used by the tests is beluga.jl_b200/.
 *
 * Conventions
 *   - plain C: opaque handle, pointers and sizes only; no exceptions or aborts cross the ABI.
 *   - every entry returns int: 0 = ok, non-zero = error (message via blg_last_error).
 *   - numerical failures are VALUES, not errors: a family whose likelihood is not finite
 *     contributes -Inf (model.jl:351,366,384,391).
 *   - node ids are the reference's: 1-based, pre-order over the Newick tree (root = 1), WGD
 *     pairs appended (wgd = max+1, wgdafter = max+2, model.jl:106-138).  Arrays "by id" are
 *     indexed id-1 and have `ncols` entries (ncols = largest id in use).
 *   - one handle drives ONE GPU on ONE stream.  Families shard across GPUs by giving each
 *     process/handle a contiguous block of patterns (profile.jl:27,34-37 is the same DP); the
 *     only cross-GPU step is the sum of the scalars / gradient vectors, which the library does
 *     itself once the handles share a communicator (blg_comm_init).
 *   - host pointers are borrowed for the duration of the call only.  `d_*` pointers are device
 *     pointers on the handle's device.
 *   - not re-entrant per handle; independent handles are independent.
 */
#ifndef BELUGA_B200_H
#define BELUGA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blg_handle blg_handle;

/* node kinds: node.jl:38-66 (:root, :sp, :wgd, :wgt, :wgdafter); BLG_ABSENT marks an id that is
 * not in model.nodes (possible between removewgd!(…, reindex=false) and reindex!, rjmcmc.jl:602) */
enum { BLG_ROOT = 0, BLG_SP = 1, BLG_WGD = 2, BLG_WGT = 3, BLG_WGDAFTER = 4, BLG_ABSENT = 255 };
/* node.jl:12-26: rates on nodes (branch rate = mean of flanking nodes) or on branches */
enum { BLG_NODE_MODEL = 0, BLG_BRANCH_MODEL = 1 };

/* --- lifetime ------------------------------------------------------------------------------
 * backs DLWGD(nw, df, …) (model.jl:22-28) / the finalizer of the PArray replacement.
 * `stream`: a cudaStream_t to run on (e.g. torch's current stream), or NULL to create one. */
int blg_create(int device, void* stream, blg_handle** out);
int blg_destroy(blg_handle* h);
/* h may be NULL: message of the last failed blg_create on this thread */
const char* blg_last_error(blg_handle* h);

/* --- profiles: Profile(X, n) (profile.jl:36-37) -----------------------------------------------
 * X is permutedims(M) of model.jl:48: ncols × npatterns, node (id-1) fastest; `weight` the
 * site-pattern multiplicities (x1 of model.jl:41).  The library re-lays the counts out
 * node-major / family-contiguous on the device and owns all partial-likelihood storage
 * (committed L and proposal Lp, profile.jl:14-20).  npatterns == 0 is the mock profile
 * (profile.jl:28,31): every likelihood entry then returns 0. */
int blg_set_profiles(blg_handle* h, const int32_t* X, const int64_t* weight, int ncols, int npatterns);
/* the same with X in column-major order (Xc[c * npatterns + f]: one contiguous run per node id — what a column-wise
 * construction of the extended profile, leaves first, produces; no transposition on either side) */
int blg_set_profiles_columns(blg_handle* h, const int32_t* Xc, const int64_t* weight, int ncols, int npatterns);

/* --- rand(model, N; condition) (sim.jl:17-27), count level, on the device: families are drawn from the handle's current
 * topology and parameters (no profiles needed) until N are accepted.  condition: 0 none, 1 non-extinct (the reference's
 * default), 2 rootclades (sim.jl:51); max_leaf / max_rootsum: extra rejection bounds (-1: none).  counts_out: N × nleaves
 * uint16, row-major, columns by ascending leaf id (leaf_ids_out, room for ncols entries).  Reproducible for a given seed. */
int blg_rand_profiles(blg_handle* h, int N, unsigned long long seed, int condition, int max_leaf, int max_rootsum,
                      int32_t* leaf_ids_out, int* nleaves_out, uint16_t* counts_out, long long* draws_out);

/* --- profile(t, l, df), first half (model.jl:40-41): the site patterns of a count table (N families × T taxa, row-major,
 * host memory) and their multiplicities, in order of first appearance, grouped on the device by a content-keyed hash
 * table.  rep_out[k] = first family with pattern k, weight_out[k] = its multiplicity (both with room for N entries). */
int blg_site_patterns(int device, const int32_t* counts, int N, int T, int32_t* rep_out, int64_t* weight_out,
                      int* npatterns);
/* the same for a table of 16-bit counts (what blg_rand_profiles writes): half the bytes to the device, no widening copy */
int blg_site_patterns_u16(int device, const uint16_t* counts, int N, int T, int32_t* rep_out, int64_t* weight_out,
                          int* npatterns);

/* --- model graph: initmodel (model.jl:227-244), addwgd!/addwgt!/removewgd!/reindex!
 * (model.jl:106-183) -----------------------------------------------------------------------------
 * parent[i]: id of the parent (0 for the root); kind[i]: BLG_*; child_pos[i]: position of node
 * i+1 in its parent's child vector (the iteration order of n.c).  A topology change invalidates
 * every ε/W table: follow with blg_set_params and blg_rebuild(h, NULL, 0). */
int blg_set_topology(blg_handle* h, int ncols, const int32_t* parent, const uint8_t* kind,
                     const int32_t* child_pos);

/* --- parameters: set!(d) (model.jl:56-60), update!(n, …) (node.jl:99-116), setrates!
 * (model.jl:69-76) -------------------------------------------------------------------------------
 * blg_set_params uploads θ (arrays by id: λ/μ are read at non-WGD ids, q at wgd/wgt ids, t — the
 * branch length above the node, θ[:t] — at every id).
 * blg_rebuild runs the on-device ε/W build — set!(n) = setϵ!(n); setW!(n) (node.jl:118-179) —
 * for the listed ids in the listed order; nodes == NULL or n == 0 means all nodes in post-order
 * (set!(d)).  update!(n) on a Branch model is the path n→root (setabove!, node.jl:123-128); on a
 * Node model the children of n first (setbelow!, node.jl:130-134). */
int blg_set_params(blg_handle* h, int model_kind, const double* lambda, const double* mu,
                   const double* q, const double* t, double eta);
int blg_rebuild(blg_handle* h, const int32_t* nodes, int n);

/* --- likelihood ----------------------------------------------------------------------------------
 * All three write the proposal state (Lp) and return Σ weight·logL over this handle's patterns;
 * -Inf if any family is not finite.  The *_async forms leave the scalar in device memory
 * (d_out, 8 bytes) on the handle's stream without synchronising, so an NCCL all-reduce can be
 * issued behind it on the same stream. */
int blg_logpdf_full(blg_handle* h, double* out);            /* logpdf!(d, p)      profile.jl:56-57 */
int blg_logpdf_from(blg_handle* h, int node, double* out);  /* logpdf!(d[i], p)   profile.jl:59-60 */
int blg_logpdf_root(blg_handle* h, double* out);            /* logpdfroot(n, p)   profile.jl:62-63 */
int blg_logpdf_full_async(blg_handle* h, double* d_out);
int blg_logpdf_from_async(blg_handle* h, int node, double* d_out);
int blg_logpdf_root_async(blg_handle* h, double* d_out);

/* --- gradient: gradient(d, p) / gradient_cr(d, p) (profile.jl:71-75, model.jl:517-533) -----------
 * g has 2n+k+1 entries laid out as asvector (model.jl:541-546): [λ_1..λ_n, μ_1..μ_n, q…, η] with
 * n = number of non-WGD nodes; the q block is ordered by ASCENDING wgd id (the order model(θ),
 * model.jl:320-322, reads it; asvector itself lists them descending — see INTEGRATION.md).
 * Does not disturb the committed or proposal state (the reference uses a fresh L, model.jl:395). */
int blg_gradient(blg_handle* h, double* g, int len, double* logpdf_out);
int blg_gradient_cr(blg_handle* h, double* g2, double* logpdf_out);

/* --- multi-GPU: families shard across GPUs, one process and one handle per GPU (profile.jl:27-37,56-75 is a mapreduce
 * over worker processes holding contiguous blocks of families).  The library owns the NCCL communicator: rank 0 asks for
 * a unique id (128 bytes), the host program hands it to every rank by its own means (MPI / Distributed / torch.distributed)
 * and each rank joins with blg_comm_init.  From then on every likelihood and gradient entry of the handle returns the SUM
 * over all ranks (one ncclAllReduce of 8 bytes — 8·P for a gradient — on the handle's stream, over NVLink); each rank
 * passes its own contiguous block of patterns to blg_set_profiles (an empty block is fine).  All ranks must make the same
 * sequence of likelihood / gradient calls.  NCCL is bound at run time (libnccl.so.2, or $BLG_NCCL_LIB). */
int blg_comm_unique_id(uint8_t* id128);
int blg_comm_init(blg_handle* h, int nranks, int rank, const uint8_t* id128);
int blg_comm_size(blg_handle* h);

/* --- accept / reject / reversible jump: set!, rev!, extend!, shrink! (profile.jl:79-121) --------
 * commit: proposal becomes committed (Lp→L, xp→x); revert: the reverse.  Both are O(#nodes)
 * pointer edits, no bulk copies.  extend appends two columns (ids ncols+1, ncols+2) whose counts
 * equal column i's and whose partial likelihoods are "-Inf"; shrink deletes columns i and i+1 and
 * renumbers — in the proposal state only. */
int blg_commit(blg_handle* h);
int blg_revert(blg_handle* h);
int blg_extend(blg_handle* h, int i);
int blg_shrink(blg_handle* h, int i);

/* --- inspection (tests; test/tests.jl:9-12 compares L[:,1]) --------------------------------------
 * blg_get_column: log-space column of one pattern/node as the reference would hold it in Lp
 * (which = 1) or L (which = 0); entries above the pattern's count are -Inf.  Returns the count
 * bound+1 of that column in *rows. */
int blg_get_column(blg_handle* h, int pattern, int node, int which, double* out, int cap, int* rows);
int blg_get_family_logpdf(blg_handle* h, double* out, int n);      /* per-pattern logL of the last call */
int blg_get_eps(blg_handle* h, double* eps_top, double* eps_node, int ncols);  /* ϵ[1], ϵ[2] by id */
int blg_get_W(blg_handle* h, int node, double* out, int cap, int* dim);        /* row-major [n][j] */
int blg_ncols(blg_handle* h, int which);
int blg_npatterns(blg_handle* h);
/* host-only diagnostic: the device order blg_set_profiles gives the rows of X (perm_out[slot] = row); the
 * reference keeps families in input order (src/model.jl:40-49) — the order is invisible through the ABI */
int blg_family_order(const int32_t* X, int ncols, int npatterns, int32_t* perm_out);
/* algorithmic bytes (definition of SURVEY.md §8(d): each internal column written once and read once by
 * its parent, 2 B per leaf count, 4 B per stored scale exponent, 8 B per family result) — in total and
 * for the pruning launches alone — and the number of kernels the last likelihood call launched */
int blg_last_algorithmic(blg_handle* h, double* bytes_total, double* bytes_prune, int* launches);
/* diagnostic: launches of the table-free sweep kernel / of the general kernel that did the work so far (the choice is made
 * on the device from the range flags of the last ε/W build), and the split of the patterns between the fast family set
 * (root count <= 100: SoA slabs, 32 families per warp) and the heavy one (ragged slabs, one warp per family) */
int blg_kernel_runs(blg_handle* h, unsigned int* sweep_runs, unsigned int* general_runs, int* nfast, int* nheavy);
/* diagnostic: node patterns (families are deduplicated per node by the counts below it, csrc/pattern.cuh) — bytes of
 * pattern columns the last likelihood call wrote + read, node steps × distinct patterns evaluated, node steps × families */
int blg_pattern_stats(blg_handle* h, double* bytes_moved, double* pattern_columns, double* family_columns);
/* diagnostic: likelihood calls captured as CUDA graphs so far / replayed from a captured graph.  The second call in an
 * unchanged state (same entry, same output pointer; no allocation, commit/revert/extend/shrink, topology, profile or table
 * descriptor change in between) is captured, later ones are one cudaGraphLaunch (+ the all-reduce).  BLG_GRAPH=0 disables it. */
int blg_graph_stats(blg_handle* h, unsigned long long* captures, unsigned long long* replays);
int blg_sync(blg_handle* h);
/* measurement: when on, every likelihood call records CUDA events on the handle's stream around its
 * pruning launches and around its root/reduction launches; blg_last_timing waits for them and returns
 * the two device durations of the last call in milliseconds. */
int blg_set_timing(blg_handle* h, int on);
int blg_last_timing(blg_handle* h, float* prune_ms, float* root_ms);
/* measured fp64 FMA throughput of the device (TFLOP/s, 8 independent DFMA chains per thread): the fp64-roofline
 * denominator SURVEY.md §8(d) asks for; not part of the reference's API */
int blg_fp64_peak(blg_handle* h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
