// Host-side keypoint distributor (quad-tree) -- product implementation of
// PLVS2::ORBextractor::DistributeOctTree (reference: src/ORBextractor.cc:611-865, with
// ExtractorNode::DivideNode :536-592 and compareNodes :594-609).
//
// The reference keeps a std::list of nodes, inserts children at the FRONT, erases the parent and
// finally emits, per surviving node in list order, the first maximum-response keypoint.  That
// order is observable (it is the order of the returned keypoints), so it is reproduced here --
// but with flat arrays instead of a linked list of vector-owning nodes:
//   * candidates live once in a permutation array; a node is (bounds, [begin,count)) and a split
//     is a stable 4-way partition of its range (input order inside a node is preserved, which is
//     what makes "first maximum wins" agree);
//   * the list is stored back-to-front in a vector (front of the list == back of the vector), so
//     push_front == push_back and a full sweep rebuilds it as  kept-leaves ++ new-children;
//   * only the three corner values the reference's arithmetic ever reads are kept
//     (UL.x, UL.y, UR.x, BR.y; its BL/BR.x fields are written but never read).
// The partial-expansion phase sorts (size, UL.x) with the reference's comparator through
// libstdc++'s std::sort: the comparator is not a total order, so the tie permutation is defined
// by the algorithm, and using the same library routine on the same sequence reproduces it.
//
// Round-1 placement: this runs on the host between the FAST/compaction kernels and the
// orientation/descriptor kernel (the reference's own CUDA build does the same,
// src/ORBextractor.cc:1010-1013).  DESIGN.md lists the device version as the next step.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace plvs {
namespace orb {

struct QuadNode {
    int ulx, uly, urx, bry;
    int begin, count;
    bool leaf;      // bNoMore
    bool dead;      // erased from the list
};

class Distributor {
public:
    // xs/ys: candidate coordinates relative to (minX,minY) (integer valued), resp: FAST score.
    // Appends the indices of the selected candidates, in the reference's output order, to `out`.
    void run(int n, const int* xs, const int* ys, const int* resp,
             int minX, int maxX, int minY, int maxY, int N, std::vector<int>& out)
    {
        out.clear();
        const int nIni = (int)std::round((float)(maxX - minX) / (float)(maxY - minY));
        if (nIni == 0 || n == 0) return;
        const float hX = (float)(maxX - minX) / (float)nIni;
        x_ = xs; y_ = ys;
        perm_.resize(n); tmp_.resize(n);
        nodes_.clear(); order_.clear();

        // root nodes + bucket the candidates (stable counting sort by root index)
        std::vector<int> rootOf(n), cnt(nIni + 1, 0);
        for (int i = 0; i < n; ++i) { rootOf[i] = (int)((float)xs[i] / hX); ++cnt[rootOf[i] + 1]; }
        for (int r = 0; r < nIni; ++r) cnt[r + 1] += cnt[r];
        {
            std::vector<int> cur(cnt.begin(), cnt.end() - 1);
            for (int i = 0; i < n; ++i) perm_[cur[rootOf[i]]++] = i;
        }
        // the reference emplaces roots at the BACK (list order = root 0 .. nIni-1); empty roots are erased
        std::vector<int> fwd;
        for (int r = 0; r < nIni; ++r) {
            const int c = cnt[r + 1] - cnt[r];
            if (c == 0) continue;
            QuadNode q{(int)(hX * (float)r), 0, (int)(hX * (float)(r + 1)), maxY - minY, cnt[r], c, c == 1, false};
            nodes_.push_back(q);
            fwd.push_back((int)nodes_.size() - 1);
        }
        order_.assign(fwd.rbegin(), fwd.rend());      // stored back-to-front
        int live = (int)order_.size();

        std::vector<std::pair<int, int>> expandable, sorted;   // (size, node id) in creation order
        std::vector<int> next;
        bool done = false;
        while (!done) {
            const int prev = live;
            int nToExpand = 0;
            expandable.clear();
            next.clear();
            // sweep front->back == order_ back->front; leaves keep their relative order
            for (int i = 0; i < (int)order_.size(); ++i) if (nodes_[order_[i]].leaf) next.push_back(order_[i]);
            for (int i = (int)order_.size() - 1; i >= 0; --i) {
                const int id = order_[i];
                if (nodes_[id].leaf) continue;
                int kid[4];
                split(id, kid);
                for (int qd = 0; qd < 4; ++qd) {
                    if (kid[qd] < 0) continue;
                    next.push_back(kid[qd]);
                    if (nodes_[kid[qd]].count > 1) { ++nToExpand; expandable.emplace_back(nodes_[kid[qd]].count, kid[qd]); }
                }
            }
            order_.swap(next);
            live = (int)order_.size();
            if (live >= N || live == prev) {
                done = true;
            } else if (live + nToExpand * 3 > N) {
                while (!done) {
                    const int prev2 = live;
                    sorted = expandable;
                    expandable.clear();
                    std::sort(sorted.begin(), sorted.end(), [this](const std::pair<int, int>& a, const std::pair<int, int>& b) {
                        if (a.first < b.first) return true;
                        if (a.first > b.first) return false;
                        return nodes_[a.second].ulx < nodes_[b.second].ulx;
                    });
                    for (int j = (int)sorted.size() - 1; j >= 0; --j) {
                        const int id = sorted[j].second;
                        int kid[4];
                        split(id, kid);
                        for (int qd = 0; qd < 4; ++qd) {
                            if (kid[qd] < 0) continue;
                            order_.push_back(kid[qd]);
                            ++live;
                            if (nodes_[kid[qd]].count > 1) expandable.emplace_back(nodes_[kid[qd]].count, kid[qd]);
                        }
                        nodes_[id].dead = true;
                        --live;
                        if (live >= N) break;
                    }
                    if (live >= N || live == prev2) done = true;
                }
            }
        }
        for (int i = (int)order_.size() - 1; i >= 0; --i) {
            const QuadNode& q = nodes_[order_[i]];
            if (q.dead) continue;
            int best = perm_[q.begin];
            int br = resp[best];
            for (int k = 1; k < q.count; ++k) {
                const int c = perm_[q.begin + k];
                if (resp[c] > br) { best = c; br = resp[c]; }
            }
            out.push_back(best);
        }
    }

private:
    // stable 4-way partition of node `id`; kid[q] = new node id or -1 if the quadrant is empty.
    // Quadrant order n1 (UL), n2 (UR), n3 (BL), n4 (BR) as in DivideNode.
    void split(int id, int kid[4])
    {
        const QuadNode p = nodes_[id];
        const int halfX = (int)std::ceil((float)(p.urx - p.ulx) / 2);
        const int halfY = (int)std::ceil((float)(p.bry - p.uly) / 2);
        const int mx = p.ulx + halfX, my = p.uly + halfY;
        int c[4] = {0, 0, 0, 0};
        for (int k = 0; k < p.count; ++k) {
            const int i = perm_[p.begin + k];
            const int qd = (x_[i] < mx ? 0 : 1) + (y_[i] < my ? 0 : 2);
            tmp_[p.begin + k] = qd;
            ++c[qd];
        }
        int o[4] = {p.begin, p.begin + c[0], p.begin + c[0] + c[1], p.begin + c[0] + c[1] + c[2]};
        int w[4] = {o[0], o[1], o[2], o[3]};
        scratch_.resize(p.count);
        for (int k = 0; k < p.count; ++k) scratch_[k] = perm_[p.begin + k];
        for (int k = 0; k < p.count; ++k) perm_[w[tmp_[p.begin + k]]++] = scratch_[k];
        const int bx[4][2] = {{p.ulx, mx}, {mx, p.urx}, {p.ulx, mx}, {mx, p.urx}};
        const int by[4][2] = {{p.uly, my}, {p.uly, my}, {my, p.bry}, {my, p.bry}};
        for (int qd = 0; qd < 4; ++qd) {
            if (c[qd] == 0) { kid[qd] = -1; continue; }
            nodes_.push_back(QuadNode{bx[qd][0], by[qd][0], bx[qd][1], by[qd][1], o[qd], c[qd], c[qd] == 1, false});
            kid[qd] = (int)nodes_.size() - 1;
        }
    }

    const int* x_ = nullptr;
    const int* y_ = nullptr;
    std::vector<int> perm_, tmp_, scratch_;
    std::vector<QuadNode> nodes_;
    std::vector<int> order_;
};

}  // namespace orb
}  // namespace plvs
