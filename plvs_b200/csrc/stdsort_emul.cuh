// Exact emulation of libstdc++'s std::sort(first, last, comp) (bits/stl_algo.h, GCC 4.x .. 14) for a
// comparator that looks at a 32-bit key only.  Needed because the reference sorts the expandable quadtree nodes
// with a comparator that is NOT a total order -- (size, UL.x), src/ORBextractor.cc:594-609 -- so the order of equal
// keys is whatever the algorithm does: introsort (median-of-3 pivot moved to *first, unguarded Hoare partition,
// depth limit 2*floor(log2 n), heapsort fallback) down to partitions of <= 16 elements, then one insertion sort
// (guarded for the first 16 elements, unguarded for the rest).  Elements are 64-bit: key in the high word (compared),
// payload in the low word (carried).  Host-checked against std::sort on tie-heavy inputs (tests/test_host_logic.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PLVS_SORT_HD __host__ __device__ inline
#else
#define PLVS_SORT_HD inline
#endif

namespace plvs {
namespace stdsort {

typedef unsigned long long elem_t;
PLVS_SORT_HD bool less_key(elem_t a, elem_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
PLVS_SORT_HD void swap_e(elem_t* a, elem_t* b) { elem_t t = *a; *a = *b; *b = t; }

// std::__adjust_heap + std::__push_heap
PLVS_SORT_HD void adjust_heap(elem_t* first, int hole, int len, elem_t value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less_key(first[child], first[child - 1])) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && less_key(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

// std::__partial_sort(first, last, last) == __heap_select (make_heap over the whole range) + __sort_heap
PLVS_SORT_HD void heap_sort(elem_t* first, int len)
{
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            elem_t v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            --parent;
        }
    }
    for (int last = len; last > 1;) {
        --last;
        elem_t v = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, v);
    }
}

PLVS_SORT_HD void move_median_to_first(elem_t* result, elem_t* a, elem_t* b, elem_t* c)
{
    if (less_key(*a, *b)) {
        if (less_key(*b, *c)) swap_e(result, b);
        else if (less_key(*a, *c)) swap_e(result, c);
        else swap_e(result, a);
    } else if (less_key(*a, *c)) swap_e(result, a);
    else if (less_key(*b, *c)) swap_e(result, c);
    else swap_e(result, b);
}

PLVS_SORT_HD int unguarded_partition(elem_t* base, int first, int last, int pivot)
{
    for (;;) {
        while (less_key(base[first], base[pivot])) ++first;
        --last;
        while (less_key(base[pivot], base[last])) --last;
        if (!(first < last)) return first;
        swap_e(&base[first], &base[last]);
        ++first;
    }
}

PLVS_SORT_HD void unguarded_linear_insert(elem_t* base, int last)
{
    elem_t val = base[last];
    int next = last - 1;
    while (less_key(val, base[next])) { base[last] = base[next]; last = next; --next; }
    base[last] = val;
}

PLVS_SORT_HD void insertion_sort(elem_t* base, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less_key(base[i], base[first])) {
            elem_t val = base[i];
            for (int k = i; k > first; --k) base[k] = base[k - 1];       // move_backward(first, i, i+1)
            base[first] = val;
        } else unguarded_linear_insert(base, i);
    }
}

// std::sort on base[0..n)
PLVS_SORT_HD void sort(elem_t* base, int n)
{
    if (n <= 1) return;
    // __introsort_loop with an explicit stack of (first, last, depth) -- recursion goes to the RIGHT part, loop on the left
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int lg = 0; for (int t = n; t > 1; t >>= 1) ++lg;
    int first = 0, last = n, depth = 2 * lg;
    for (;;) {
        while (last - first > 16) {
            if (depth == 0) { heap_sort(base + first, last - first); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            move_median_to_first(&base[first], &base[first + 1], &base[mid], &base[last - 1]);
            const int cut = unguarded_partition(base, first + 1, last, first);
            // recursive call on [cut, last) happens BEFORE continuing with [first, cut): emulate with a stack that
            // finishes the right part first (order of the two sub-sorts does not change the result: disjoint ranges)
            stack_first[sp] = first; stack_last[sp] = cut; stack_depth[sp] = depth; ++sp;
            first = cut;
        }
        if (sp == 0) break;
        --sp; first = stack_first[sp]; last = stack_last[sp]; depth = stack_depth[sp];
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(base, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(base, i);
    } else insertion_sort(base, 0, n);
}

// ---- the same result, computed range by range ----------------------------------------------------------------------------------------------
// __introsort_loop recurses into [cut, last) and loops on [first, cut): two DISJOINT ranges, so the two sub-sorts commute and the recursion
// tree can be processed level by level, every range of a level by a different thread.  __final_insertion_sort then only ever moves an element
// inside the leaf range it ended up in (<= 16 elements, or a heap-sorted range): after a Hoare partition everything left of the cut is <= pivot
// <= everything right of it, so `less(val, *prev)` fails at the first element of the previous leaf -- and the guarded variant used for the
// first 16 positions moves to the front exactly when a linear insert would get there too (the prefix is sorted).  Hence: leaf-wise guarded
// linear insertion, one thread per leaf.  tests/test_host_logic.py checks this formulation against std::sort as well.
PLVS_SORT_HD int introsort_step(elem_t* base, int first, int last, int depth)        // one iteration of the loop on a range of > 16 elements
{
    if (depth == 0) { heap_sort(base + first, last - first); return -1; }          // sorted: no children
    const int mid = first + (last - first) / 2;
    move_median_to_first(&base[first], &base[first + 1], &base[mid], &base[last - 1]);
    return unguarded_partition(base, first + 1, last, first);                       // children [first, cut) and [cut, last), depth - 1
}

PLVS_SORT_HD void leaf_insertion(elem_t* base, int a, int b)
{
    for (int i = a + 1; i < b; ++i) {
        const elem_t val = base[i];
        int k = i;
        while (k > a && less_key(val, base[k - 1])) { base[k] = base[k - 1]; --k; }
        base[k] = val;
    }
}

// __unguarded_partition as a statement about ranks: the scans never revisit an element that was swapped (it is behind the pointers), so the i-th
// swap exchanges the i-th position from the left whose ORIGINAL key is not less than the pivot's with the i-th position from the right whose
// original key is not greater, for as long as the left one lies below the right one; the cut is where the last left scan stops: the next such left
// position or the right partner of the last swap, whichever comes first.  The two
// position lists come from two independent scans and the swaps touch disjoint positions: a warp does them with ballots (warp_partition below).
// This serial form is what the host test checks against std::sort.  lpos / rpos: scratch, indexed first .. last-1.
PLVS_SORT_HD int partition_by_ranks(elem_t* base, int first /*pivot*/, int last, int* lpos, int* rpos)
{
    const elem_t pv = base[first];
    int nl = 0, nr = 0;
    for (int p = first + 1; p < last; ++p) if (!less_key(base[p], pv)) lpos[first + nl++] = p;
    for (int p = last - 1; p >= first; --p) if (!less_key(pv, base[p])) rpos[first + nr++] = p;
    int ns = 0;
    while (ns < nl && ns < nr && lpos[first + ns] < rpos[first + ns]) { swap_e(&base[lpos[first + ns]], &base[rpos[first + ns]]); ++ns; }
    // the last left scan stops at the next original stopper OR at the right partner of the last swap (which now holds a key >= the pivot's), whichever comes first
    const int next_l = ns < nl ? lpos[first + ns] : last;
    return ns > 0 && rpos[first + ns - 1] < next_l ? rpos[first + ns - 1] : next_l;
}

PLVS_SORT_HD int depth_limit(int n) { int lg = 0; for (int t = n; t > 1; t >>= 1) ++lg; return 2 * lg; }

#if defined(__CUDACC__) || defined(PLVS_CUDA_EMU)
// std::sort on base[0..n) by a whole CTA (every thread calls it).  scratch: 2 * n ints for the leaves, 6 * (n / 16 + 2) ints for the two range
// lists, 4 counters -- see sort_cta_scratch_ints().
__host__ __device__ inline int sort_cta_scratch_ints(int n) { return 4 * (n + 1) + 6 * (n / 16 + 2) + 4; }

// partition_by_ranks by one warp (all 32 lanes call it; pivot already at base[first])
__device__ inline int warp_partition(elem_t* base, int first, int last, int lane, int* lpos, int* rpos)
{
    const uint32_t pk = (uint32_t)(base[first] >> 32);
    const uint32_t below = (1u << lane) - 1u;
    int nl = 0, nr = 0;
    for (int p0 = first + 1; p0 < last; p0 += 32) {
        const int p = p0 + lane;
        const bool f = p < last && !((uint32_t)(base[p] >> 32) < pk);
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (f) lpos[first + nl + __popc(m & below)] = p;
        nl += __popc(m);
    }
    for (int p0 = last - 1; p0 >= first; p0 -= 32) {
        const int p = p0 - lane;
        const bool f = p >= first && !(pk < (uint32_t)(base[p] >> 32));
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (f) rpos[first + nr + __popc(m & below)] = p;
        nr += __popc(m);
    }
    __syncwarp();
    const int nm = nl < nr ? nl : nr;
    int ns = 0;
    for (int i0 = 0; i0 < nm; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < nm && lpos[first + i] < rpos[first + i];
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (f) { const int a = lpos[first + i], b = rpos[first + i]; const elem_t t = base[a]; base[a] = base[b]; base[b] = t; }
        ns += __popc(m);
        if (m != 0xffffffffu) break;                  // the predicate holds for a prefix (left positions ascend, right positions descend)
    }
    __syncwarp();
    const int next_l = ns < nl ? lpos[first + ns] : last;
    return ns > 0 && rpos[first + ns - 1] < next_l ? rpos[first + ns - 1] : next_l;
}
__device__ inline void sort_cta(elem_t* base, int n, int* scratch)
{
    if (n <= 1) return;
    const int tid = threadIdx.x, R = n / 16 + 2;
    int* leaf_a = scratch; int* leaf_b = leaf_a + (n + 1);
    int* lpos = leaf_b + (n + 1); int* rpos = lpos + (n + 1);
    int* rng[2] = {rpos + (n + 1), rpos + (n + 1) + 3 * R};
    int* cnt = rng[1] + 3 * R;                       // [0] ranges of the current level, [1] of the next, [2] leaves
    if (tid == 0) {
        cnt[0] = cnt[1] = cnt[2] = 0;
        if (n > 16) { rng[0][0] = 0; rng[0][1] = n; rng[0][2] = depth_limit(n); cnt[0] = 1; }
        else { leaf_a[0] = 0; leaf_b[0] = n; cnt[2] = 1; }
    }
    __syncthreads();
    int cur = 0;
    for (;;) {
        const int ncur = cnt[0];
        if (ncur == 0) break;
        __syncthreads();                              // everyone has read cnt[0]
        for (int r = tid >> 5; r < ncur; r += (int)(blockDim.x >> 5)) {        // one warp per range of this level
            const int lane = tid & 31;
            const int first = rng[cur][3 * r], last = rng[cur][3 * r + 1], depth = rng[cur][3 * r + 2];
            int cut = -1;
            if (depth == 0) { if (lane == 0) heap_sort(base + first, last - first); __syncwarp(); }          // sorted: no children
            else {
                if (lane == 0) move_median_to_first(&base[first], &base[first + 1], &base[first + (last - first) / 2], &base[last - 1]);
                __syncwarp();
                cut = warp_partition(base, first, last, lane, lpos, rpos);
            }
            if (lane != 0) continue;
            if (cut < 0) { const int l = atomicAdd(&cnt[2], 1); leaf_a[l] = first; leaf_b[l] = last; continue; }
            const int lo[2] = {first, cut}, hi[2] = {cut, last};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (hi[c] - lo[c] > 16) { const int k = atomicAdd(&cnt[1], 1); rng[cur ^ 1][3 * k] = lo[c]; rng[cur ^ 1][3 * k + 1] = hi[c]; rng[cur ^ 1][3 * k + 2] = depth - 1; }
                else if (hi[c] - lo[c] > 1) { const int l = atomicAdd(&cnt[2], 1); leaf_a[l] = lo[c]; leaf_b[l] = hi[c]; }
            }
        }
        __syncthreads();
        if (tid == 0) { cnt[0] = cnt[1]; cnt[1] = 0; }
        cur ^= 1;
        __syncthreads();
    }
    const int nleaf = cnt[2];
    for (int l = tid; l < nleaf; l += blockDim.x) leaf_insertion(base, leaf_a[l], leaf_b[l]);
    __syncthreads();
}
#endif

}  // namespace stdsort
}  // namespace plvs
