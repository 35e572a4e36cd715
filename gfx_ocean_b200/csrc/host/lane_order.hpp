// Which cross-lane waits ocean_update_overlapped has to enqueue -- pure bookkeeping, no CUDA, so that it can be
// checked on the host against a happens-before model (csrc/host_check_lane_order.cpp, tests/test_lane_order.py).
//
// Two lanes = two streams with their own row-pass intermediates; frames alternate between them. Within a lane stream
// order serialises everything. Across lanes only the MAPS are shared: the column kernel of a frame must be ordered
// behind every frame of the other lane that wrote one of its tiles and that this lane is not ordered behind yet --
// its latest frame or an older one (with disjoint tile ranges the lanes run free of each other, and a later frame may
// return to a tile an older frame of the other lane wrote). The other lane's `done` event is recorded behind all of
// its frames, so one wait covers them all.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ocean {

struct LaneOrder {
    std::vector<uint8_t> wrote[2];   // [n_tiles] maps written by frames of the lane that the OTHER lane is not yet ordered behind
    bool busy[2] = {false, false};   // frames in flight that the main stream is not yet ordered behind

    void resize(size_t n_tiles)
    {
        for (auto& w : wrote)
            if (w.size() != n_tiles) w.assign(n_tiles, 0);
    }
    // A frame of `lane` over tiles [first, first + count) is about to be enqueued. Returns true when its column kernel
    // has to wait for the other lane's `done` event (recorded behind every frame enqueued on that lane so far).
    bool enqueue(int lane, uint32_t first, uint32_t count)
    {
        auto& mine = wrote[lane];
        auto& theirs = wrote[lane ^ 1];
        bool wait = false;
        for (uint32_t t = first; t < first + count && !wait; ++t) wait = theirs[t] != 0;
        if (wait) std::fill(theirs.begin(), theirs.end(), uint8_t(0));   // this lane's later work is behind that event now
        std::fill(mine.begin() + first, mine.begin() + first + count, uint8_t(1));
        busy[lane] = true;
        return wait;
    }
    // An entry point on the main stream made it wait for the `done` event of every busy lane.
    void main_joined() { busy[0] = busy[1] = false; }
    // Both lanes are about to wait for an event recorded on the main stream. If the main stream was joined with both
    // lanes before (no frame since), that event is behind every frame enqueued so far: the lanes start afresh.
    void lanes_resumed()
    {
        if (!busy[0] && !busy[1])
            for (auto& w : wrote) std::fill(w.begin(), w.end(), uint8_t(0));
    }
};

}  // namespace ocean
