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

}  // namespace stdsort
}  // namespace plvs
