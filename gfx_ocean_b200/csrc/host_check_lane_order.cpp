// Host-side check of the cross-lane ordering of ocean_update_overlapped (host/lane_order.hpp) against a
// happens-before model. Nodes: the column kernel of every frame (the only kernel that writes maps) and every
// main-stream entry point. Edges, exactly what the library enqueues:
//   * stream order inside a lane (frame -> next frame of the lane),
//   * LaneOrder::enqueue() == true: the other lane's latest frame -> this frame's column kernel
//     (the other lane's `done` event; by stream order it covers that lane's earlier frames too),
//   * a main-stream entry point: the latest frame of every busy lane -> the entry point (join), and the entry point
//     -> the next frame of either lane (the lanes wait for ev_main before their next frame).
// Property: two frames on different lanes that write a common tile are ordered (the earlier call happens before the
// later one), so the map always ends up holding the LATEST update -- for random call sequences and for the
// returning-frame case the first version of the bookkeeping missed. The latest-frame-only rule is checked to FAIL.
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "host/lane_order.hpp"

namespace {

struct Frame {
    int lane;
    uint32_t first, count;
    std::vector<uint8_t> before;   // before[j] = 1: node j happens before this node
};

// Returns the number of unordered cross-lane pairs writing a common tile. `latest_only`: the rule the entry point first
// shipped with (look at the other lane's latest frame only).
int run(uint32_t n_tiles, const std::vector<int>& kind, const std::vector<uint32_t>& firsts, const std::vector<uint32_t>& counts,
        bool latest_only)
{
    ocean::LaneOrder lo;
    lo.resize(n_tiles);
    std::vector<Frame> nodes;                    // frames and entry points (lane = -1)
    int last_of_lane[2] = {-1, -1}, last_main = -1;
    bool main_dirty = true;
    int next_lane = 0;
    uint32_t lf[2] = {0, 0}, lc[2] = {0, 0};
    auto add_edge = [&](Frame& to, int from) {
        if (from < 0) return;
        to.before[from] = 1;
        for (size_t j = 0; j < nodes[from].before.size(); ++j)
            if (nodes[from].before[j]) to.before[j] = 1;
    };
    for (size_t i = 0; i < kind.size(); ++i) {
        Frame f{-1, 0, 0, std::vector<uint8_t>(kind.size(), 0)};
        if (kind[i] == 0) {                      // a main-stream entry point (download, upload, plain update, ocean_join ...)
            for (int l = 0; l < 2; ++l)
                if (lo.busy[l]) add_edge(f, last_of_lane[l]);
            add_edge(f, last_main);
            lo.main_joined();
            main_dirty = true;
            nodes.push_back(f);
            last_main = int(nodes.size()) - 1;
            continue;
        }
        const int li = next_lane;
        f.lane = li;
        f.first = firsts[i];
        f.count = counts[i];
        if (main_dirty) {                        // both lanes wait for ev_main
            lo.lanes_resumed();
            main_dirty = false;
        }
        add_edge(f, last_main);                  // (every lane waited for the latest ev_main at some point before this frame)
        const bool other_busy = lo.busy[li ^ 1];
        bool wait = lo.enqueue(li, f.first, f.count);
        if (latest_only) wait = other_busy && f.first < lf[li ^ 1] + lc[li ^ 1] && lf[li ^ 1] < f.first + f.count;
        add_edge(f, last_of_lane[li]);
        if (wait) add_edge(f, last_of_lane[li ^ 1]);
        lf[li] = f.first;
        lc[li] = f.count;
        nodes.push_back(f);
        last_of_lane[li] = int(nodes.size()) - 1;
        next_lane = li ^ 1;
    }
    int unordered = 0;
    for (size_t b = 0; b < nodes.size(); ++b)
        for (size_t a = 0; a < b; ++a) {
            const Frame &x = nodes[a], &y = nodes[b];
            if (x.lane < 0 || y.lane < 0 || x.lane == y.lane) continue;
            if (x.first < y.first + y.count && y.first < x.first + x.count && !y.before[a]) ++unordered;
        }
    return unordered;
}

}  // namespace

int main()
{
    // the returning-frame case: lane 0: A (tiles 0..6), lane 1: B (7), lane 0: C (7), lane 1: D (0)
    {
        const std::vector<int> kind = {1, 1, 1, 1};
        const std::vector<uint32_t> firsts = {0, 7, 7, 0}, counts = {7, 1, 1, 1};
        const int ok = run(8, kind, firsts, counts, false), old = run(8, kind, firsts, counts, true);
        std::printf("returning frame: per-tile bookkeeping %d unordered pairs, latest-frame-only rule %d\n", ok, old);
        if (ok != 0 || old == 0) return 1;
    }
    std::mt19937 rng(12345);
    long total_pairs_old = 0;
    for (int trial = 0; trial < 4000; ++trial) {
        const uint32_t n_tiles = 1 + rng() % 9;
        const size_t len = 2 + rng() % 40;
        std::vector<int> kind(len);
        std::vector<uint32_t> firsts(len), counts(len);
        for (size_t i = 0; i < len; ++i) {
            kind[i] = (rng() % 8) ? 1 : 0;
            firsts[i] = rng() % n_tiles;
            counts[i] = (rng() % 3) ? 1 : 1 + rng() % (n_tiles - firsts[i]);
        }
        const int bad = run(n_tiles, kind, firsts, counts, false);
        if (bad) {
            std::printf("trial %d: %d unordered cross-lane writers of a common tile\n", trial, bad);
            return 2;
        }
        total_pairs_old += run(n_tiles, kind, firsts, counts, true);
    }
    std::printf("4000 random call sequences: 0 unordered pairs (the latest-frame-only rule leaves %ld)\n", total_pairs_old);
    return total_pairs_old > 0 ? 0 : 3;
}
