// Deterministic re-implementation of the element movements of libstdc++'s std::sort
// (introsort: median-of-three to front, unguarded Hoare partition, 16-element threshold,
// heapsort fallback at depth 2*floor(log2 n), final insertion sort) for 64-bit elements
// compared on their high 32 bits only.
//
// Why: scipy's coo -> csc conversion (scipy/sparse/sparsetools/csr.h csr_sort_indices) sorts
// the (row, value) pairs of every column with std::sort and then sums equal rows left to
// right (csr_sum_duplicates).  std::sort is not stable, so the summation order of the
// duplicates -- and with it which floating-point cancellation residues survive
// eliminate_zeros in qubit_operator_sparse (linalg/sparse_tools.py:190-193) -- is defined by
// these exact element movements.  Elements here are (row << 32 | term_id): the row is the
// key, the term id rides along.  Usable from host code (unit-tested against std::sort in
// tests/test_stdsort_emulation.py) and from CUDA device code (csc_exact.cu).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OFB_HD __host__ __device__ __forceinline__
#else
#define OFB_HD inline
#endif

namespace ofb_sort {

// Strided element accessor: element i lives at base[i * stride].
struct Seq {
    uint64_t *base;
    int64_t stride;
    OFB_HD uint64_t get(int64_t i) const { return base[i * stride]; }
    OFB_HD void set(int64_t i, uint64_t v) const { base[i * stride] = v; }
    OFB_HD void swap(int64_t i, int64_t j) const {
        uint64_t a = base[i * stride], b = base[j * stride];
        base[i * stride] = b;
        base[j * stride] = a;
    }
};

OFB_HD bool key_less(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }

OFB_HD void unguarded_linear_insert(const Seq &s, int64_t last) {
    uint64_t val = s.get(last);
    int64_t next = last - 1;
    while (key_less(val, s.get(next))) {
        s.set(last, s.get(next));
        last = next;
        --next;
    }
    s.set(last, val);
}

OFB_HD void insertion_sort(const Seq &s, int64_t first, int64_t last) {
    if (first == last) return;
    for (int64_t i = first + 1; i != last; ++i) {
        if (key_less(s.get(i), s.get(first))) {
            uint64_t val = s.get(i);
            for (int64_t k = i; k > first; --k) s.set(k, s.get(k - 1));  // move_backward
            s.set(first, val);
        } else {
            unguarded_linear_insert(s, i);
        }
    }
}

OFB_HD void push_heap(const Seq &s, int64_t first, int64_t hole, int64_t top, uint64_t value) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && key_less(s.get(first + parent), value)) {
        s.set(first + hole, s.get(first + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    s.set(first + hole, value);
}

OFB_HD void adjust_heap(const Seq &s, int64_t first, int64_t hole, int64_t len, uint64_t value) {
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (key_less(s.get(first + child), s.get(first + (child - 1)))) child--;
        s.set(first + hole, s.get(first + child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        s.set(first + hole, s.get(first + (child - 1)));
        hole = child - 1;
    }
    push_heap(s, first, hole, top, value);
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
OFB_HD void heap_sort(const Seq &s, int64_t first, int64_t last) {
    int64_t len = last - first;
    if (len >= 2) {
        int64_t parent = (len - 2) / 2;
        while (true) {
            uint64_t value = s.get(first + parent);
            adjust_heap(s, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        uint64_t value = s.get(last);
        s.set(last, s.get(first));
        adjust_heap(s, first, 0, last - first, value);
    }
}

OFB_HD void move_median_to_first(const Seq &s, int64_t result, int64_t a, int64_t b, int64_t c) {
    uint64_t va = s.get(a), vb = s.get(b), vc = s.get(c);
    if (key_less(va, vb)) {
        if (key_less(vb, vc)) s.swap(result, b);
        else if (key_less(va, vc)) s.swap(result, c);
        else s.swap(result, a);
    } else if (key_less(va, vc)) s.swap(result, a);
    else if (key_less(vb, vc)) s.swap(result, c);
    else s.swap(result, b);
}

OFB_HD int64_t unguarded_partition(const Seq &s, int64_t first, int64_t last, int64_t pivot) {
    while (true) {
        while (key_less(s.get(first), s.get(pivot))) ++first;
        --last;
        while (key_less(s.get(pivot), s.get(last))) --last;
        if (!(first < last)) return first;
        s.swap(first, last);
        ++first;
    }
}

// std::sort(first, last) on elements [0, n)
OFB_HD void std_sort(const Seq &s, int64_t n) {
    if (n <= 0) return;
    // __introsort_loop with an explicit stack for the recursion on the right part
    int64_t stack_first[72], stack_last[72];
    int stack_depth[72];
    int sp = 0;
    int lg = 0;
    for (int64_t m = n; m > 1; m >>= 1) ++lg;
    int64_t first = 0, last = n;
    int depth = 2 * lg;
    while (true) {
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort(s, first, last);
                break;
            }
            --depth;
            int64_t mid = first + (last - first) / 2;
            move_median_to_first(s, first, first + 1, mid, last - 1);
            int64_t cut = unguarded_partition(s, first + 1, last, first);
            // recurse on [cut, last) later; continue with [first, cut)
            stack_first[sp] = cut;
            stack_last[sp] = last;
            stack_depth[sp] = depth;
            ++sp;
            last = cut;
        }
        if (sp == 0) break;
        --sp;
        first = stack_first[sp];
        last = stack_last[sp];
        depth = stack_depth[sp];
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(s, 0, 16);
        for (int64_t i = 16; i != n; ++i) unguarded_linear_insert(s, i);
    } else {
        insertion_sort(s, 0, n);
    }
}

}  // namespace ofb_sort
