// ingest.cuh — site-pattern compression on the device.  Included by beluga_b200.cu.
//
// profile(t, l, df) (src/model.jl:40-49) first groups identical rows of the count table (`by(df, names(df), …)`: the site
// patterns and their multiplicities) and then extends each pattern to the internal nodes.  For 10^6 families × 100 taxa
// the grouping is the expensive part (a lexicographic sort of 10^6 rows on the host).  Here every family inserts its row
// into an open-addressing hash table keyed by the ROW CONTENT (64-bit hash for the slot, full-row comparison against the
// slot's representative on a hit), keeps the smallest family index of its group as the representative (first appearance,
// the order the reference's grouping returns) and counts the group's size.  One thread per family, one pass; the host
// sorts the representatives (a few 10^5 integers).
#pragma once

__device__ __forceinline__ unsigned long long ing_mix(unsigned long long h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}
// slot_rep[s]: representative family of the group living in slot s (0xffffffff: empty); fam_slot[f]: the slot of family f
template <class V>
__global__ void __launch_bounds__(256) ingest_insert_kernel(const V* __restrict__ counts, int N, int T, unsigned int* __restrict__ slot_rep,
                                                            unsigned int* __restrict__ fam_slot, unsigned int nslots_mask) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= N) return;
    const V* row = counts + (size_t)f * T;
    unsigned long long h = 0x9e3779b97f4a7c15ULL;
    for (int j = 0; j < T; ++j) h = ing_mix(h ^ (unsigned long long)(unsigned int)row[j]) + 0x9e3779b97f4a7c15ULL * (unsigned long long)(j + 1);
    unsigned int s = (unsigned int)(h >> 17) & nslots_mask;
    for (;;) {
        unsigned int rep = atomicCAS(&slot_rep[s], 0xffffffffu, (unsigned int)f);
        if (rep == 0xffffffffu) { fam_slot[f] = s; return; }        // claimed an empty slot: f is (for now) the representative
        // occupied: the same pattern?  (the representative's row is immutable input: safe to read at any time)
        const V* other = counts + (size_t)rep * T;
        bool same = true;
        for (int j = 0; j < T; ++j) if (row[j] != other[j]) { same = false; break; }
        if (same) { fam_slot[f] = s; return; }
        s = (s + 1) & nslots_mask;                                   // another pattern lives here: linear probing
    }
}
// smallest family index and size of every group
__global__ void __launch_bounds__(256) ingest_group_kernel(int N, const unsigned int* __restrict__ fam_slot, unsigned int* __restrict__ slot_min,
                                                           unsigned int* __restrict__ slot_cnt) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= N) return;
    const unsigned int s = fam_slot[f];
    atomicMin(&slot_min[s], (unsigned int)f);
    atomicAdd(&slot_cnt[s], 1u);
}
// compact the occupied slots: (first family, multiplicity) pairs, in slot order (the host puts them in first-appearance order)
__global__ void __launch_bounds__(256) ingest_compact_kernel(unsigned int nslots, const unsigned int* __restrict__ slot_min, const unsigned int* __restrict__ slot_cnt,
                                                             unsigned int* __restrict__ out_rep, unsigned int* __restrict__ out_cnt, unsigned int* __restrict__ counter) {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots || slot_cnt[s] == 0u) return;
    const unsigned int k = atomicAdd(counter, 1u);
    out_rep[k] = slot_min[s];
    out_cnt[k] = slot_cnt[s];
}
