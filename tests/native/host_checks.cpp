// Host-side checks of product code that needs no GPU: the keypoint distributor (orb_distribute.hpp)
// and the glibc sincosf port (libm_sincosf.cuh), exposed with C linkage for the CPU test-suite.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../plvs_b200/csrc/orb_distribute.hpp"
#include "../../plvs_b200/csrc/libm_sincosf.cuh"
#include "../../plvs_b200/csrc/stdsort_emul.cuh"
#include <algorithm>
#include <utility>

extern "C" {

int chk_distribute(int n, const int* xs, const int* ys, const int* resp, int minX, int maxX, int minY, int maxY, int N, int* sel)
{
    plvs::orb::Distributor d;
    std::vector<int> out;
    d.run(n, xs, ys, resp, minX, maxX, minY, maxY, N, out);
    for (size_t i = 0; i < out.size(); ++i) sel[i] = out[i];
    return (int)out.size();
}

// std::sort emulation vs the real std::sort with a key-only (non-total) comparator; returns #position mismatches
int chk_stdsort(int n, const uint32_t* keys, int* order_out)
{
    std::vector<std::pair<uint32_t, int>> ref(n);
    std::vector<plvs::stdsort::elem_t> em(n);
    for (int i = 0; i < n; ++i) { ref[i] = std::make_pair(keys[i], i); em[i] = ((plvs::stdsort::elem_t)keys[i] << 32) | (uint32_t)i; }
    std::sort(ref.begin(), ref.end(), [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) { return a.first < b.first; });
    plvs::stdsort::sort(em.data(), n);
    int bad = 0;
    for (int i = 0; i < n; ++i) { if ((int)(uint32_t)em[i] != ref[i].second) ++bad; if (order_out) order_out[i] = (int)(uint32_t)em[i]; }
    return bad;
}

// the range-by-range formulation the device runs (stdsort_emul.cuh: introsort_step per range, level by level, then leaf-wise insertion) vs std::sort
int chk_stdsort_ranges(int n, const uint32_t* keys)
{
    using namespace plvs::stdsort;
    std::vector<std::pair<uint32_t, int>> ref(n);
    std::vector<elem_t> em(n);
    for (int i = 0; i < n; ++i) { ref[i] = std::make_pair(keys[i], i); em[i] = ((elem_t)keys[i] << 32) | (uint32_t)i; }
    std::sort(ref.begin(), ref.end(), [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) { return a.first < b.first; });
    struct R { int first, last, depth; };
    std::vector<R> cur, next; std::vector<std::pair<int, int>> leaves;
    std::vector<int> lpos(n + 1), rpos(n + 1);
    if (n > 16) cur.push_back(R{0, n, depth_limit(n)}); else if (n > 1) leaves.push_back(std::make_pair(0, n));
    while (!cur.empty()) {
        next.clear();
        for (size_t r = cur.size(); r-- > 0;) {                 // any order inside a level: here back to front
            const R g = cur[r];
            int cut;
            if (g.depth == 0) { heap_sort(em.data() + g.first, g.last - g.first); cut = -1; }
            else {      // the step as the device's warps do it: median to the front, then the partition as a statement about ranks
                move_median_to_first(&em[g.first], &em[g.first + 1], &em[g.first + (g.last - g.first) / 2], &em[g.last - 1]);
                cut = partition_by_ranks(em.data(), g.first, g.last, lpos.data(), rpos.data());
            }
            if (cut < 0) { leaves.push_back(std::make_pair(g.first, g.last)); continue; }
            const int lo[2] = {g.first, cut}, hi[2] = {cut, g.last};
            for (int c = 0; c < 2; ++c) {
                if (hi[c] - lo[c] > 16) next.push_back(R{lo[c], hi[c], g.depth - 1});
                else if (hi[c] - lo[c] > 1) leaves.push_back(std::make_pair(lo[c], hi[c]));
            }
        }
        cur.swap(next);
    }
    for (size_t l = leaves.size(); l-- > 0;) leaf_insertion(em.data(), leaves[l].first, leaves[l].second);
    int bad = 0;
    for (int i = 0; i < n; ++i) if ((int)(uint32_t)em[i] != ref[i].second) ++bad;
    return bad;
}

// sweep floats in [lo_bits, hi_bits] with the given stride; returns the number of (cos,sin) mismatches vs libm
long chk_sincosf_sweep(uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, float* first_bad)
{
    long bad = 0;
    for (uint64_t b = lo_bits; b <= hi_bits; b += stride) {
        uint32_t u = (uint32_t)b; float y; std::memcpy(&y, &u, 4);
        float c, s; plvs::libm_sincosf(y, &c, &s);
        const float lc = cosf(y), ls = sinf(y);
        if (std::memcmp(&c, &lc, 4) || std::memcmp(&s, &ls, 4)) { if (!bad && first_bad) *first_bad = y; ++bad; }
    }
    return bad;
}

}
