// TEST INFRASTRUCTURE ONLY: checks openfermion_b200/csrc/stdsort_emul.h against the real
// libstdc++ std::sort on (key, id) pairs compared by key only.  Exit code 0 = identical
// element order on every trial.  Built and run by tests/test_stdsort_emulation.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>

#include "../openfermion_b200/csrc/stdsort_emul.h"

struct Less {
    bool operator()(const std::pair<int, double> &a, const std::pair<int, double> &b) const {
        return a.first < b.first;
    }
};

int main(int argc, char **argv) {
    int trials = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 rng(12345);
    long checked = 0;
    for (int trial = 0; trial < trials; ++trial) {
        int n = (trial % 7 == 0) ? (int)(rng() % 40) : (int)(rng() % 3000);
        if (trial % 97 == 0) n = 20000 + (int)(rng() % 20000);
        int distinct = 1 + (int)(rng() % (trial % 3 == 0 ? 8 : (trial % 3 == 1 ? 64 : 4096)));
        int stride = 1 + (int)(rng() % 3);
        std::vector<std::pair<int, double>> ref(n);
        std::vector<uint64_t> mine((size_t)n * stride + 1, 0);
        int pattern = trial % 5;
        for (int i = 0; i < n; ++i) {
            int key;
            switch (pattern) {
                case 0: key = (int)(rng() % distinct); break;
                case 1: key = i % distinct; break;                 // periodic
                case 2: key = (i * 7919) % distinct; break;
                case 3: key = (n - i) % distinct; break;           // descending runs
                default: key = (int)((i ^ (i >> 3)) % distinct); break;  // xor pattern
            }
            ref[i] = {key, (double)i};
            mine[(size_t)i * stride] = ((uint64_t)(uint32_t)key << 32) | (uint32_t)i;
        }
        std::sort(ref.begin(), ref.end(), Less());
        ofb_sort::Seq s{mine.data(), stride};
        ofb_sort::std_sort(s, n);
        for (int i = 0; i < n; ++i) {
            uint64_t e = mine[(size_t)i * stride];
            if ((int)(e >> 32) != ref[i].first || (int)(uint32_t)e != (int)ref[i].second) {
                printf("MISMATCH trial %d n %d pos %d\n", trial, n, i);
                return 1;
            }
        }
        checked += n;
    }
    // heapsort fallback (std::__partial_sort(first, last, last) = make_heap + sort_heap)
    for (int trial = 0; trial < 300; ++trial) {
        int n = 1 + (int)(rng() % 4000);
        int distinct = 1 + (int)(rng() % 50);
        std::vector<uint64_t> a(n), r;
        for (int i = 0; i < n; ++i) a[i] = ((uint64_t)(rng() % distinct) << 32) | (uint32_t)i;
        r = a;
        auto cmp = [](uint64_t x, uint64_t y) { return (uint32_t)(x >> 32) < (uint32_t)(y >> 32); };
        std::make_heap(r.begin(), r.end(), cmp);
        std::sort_heap(r.begin(), r.end(), cmp);
        ofb_sort::Seq s{a.data(), 1};
        ofb_sort::heap_sort(s, 0, n);
        if (r != a) {
            printf("HEAP MISMATCH trial %d\n", trial);
            return 1;
        }
    }
    printf("OK %d trials, %ld elements\n", trials, checked);
    return 0;
}
