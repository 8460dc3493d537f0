// ddz_search.cuh -- the search bot's view of a move list: server/mcts/get_moves.py:36-69 ranks the
// moves of a list longer than 10 by  cards_value[move] - 0.1 * (cards the hand keeps)  (server/mcts/evaluator.py:17-57),
// drops the rocket-kicker moves (get_moves.py:22-34) and keeps the lowest- and highest-valued thirds, interleaved.
// Here: the value of a packed move as an integer sort key (value_key), and the p-th entry of the pruned list without
// building it (pruned_pick).  Everything is integer work; the one floating-point expression of the reference is replaced
// by the rank table make_value_rank.py computes with the reference's own two IEEE operations.
#pragma once
#include "ddz_device.cuh"

namespace ddz {
namespace search {

constexpr int kC2Min = -14, kC2Max = 57, kLeftMax = 20;    // 2 x cards_value in [-14, 57]; the hand keeps 0..20 cards
constexpr uint32_t kDropped = 0xFFFFu;                     // key of a move get_moves.py:56-57 skips
constexpr int kPruneAbove = 10;                            // get_moves.py:51: lists of up to 10 moves stay as they are

#ifndef DDZ_HOST_HARNESS
__device__
#endif
const uint16_t g_value_rank[(kC2Max - kC2Min + 1) * (kLeftMax + 1)] = {
#include "ddz_value_rank.inc"
};

// sum over the set bits i of x of (i + 1)
DDZ_DEV int weighted_bits(uint32_t x) {
    int s = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) s += __popc(x >> j);
    return s;
}

// 2 x cards_value[move] (evaluator.py:17-57); char2val - 10 = rank index - 7.  The table is keyed by count vectors,
// built in card.py's action order: a[0] is the lowest main rank, a[-1] the highest card of a sequence -- and, for the
// planes with kickers, the highest KICKER (the kickers are written after the mains, card.py:117,130).
DDZ_DEV int value2(const Trick& t, const Masks& a) {
    const int v = t.val - 7;
    switch (t.cat) {
        case 0: return 0;
        case 1: case 13: case 14: return 2 * v;
        case 2: case 5: case 6: return v > 0 ? 3 * v : 2 * v;
        case 3: return v > 0 ? 4 * v : 2 * v;
        case 4: return 18;
        case 7: case 8: case 9: return max(0, t.val + t.len - 1 - 7);
        case 12: return 24;
        default: {                                          // 10, 11: planes with solo / pair kickers
            const uint32_t kick = (t.cat == 10 ? a.g1 : a.g2) & ~a.g3;
            const int top = 31 - __clz(kick) - 7;
            return max(0, top) + (t.cat == 10 ? 2 : 3) * weighted_bits(kick >> 8);
        }
    }
}

// the sort key of `mv` played from a hand of `handnum` cards: dense rank of the reference's double, or kDropped for
// 4 + {BJ, RJ} and 33 3 44 4 + {BJ, RJ} (sidaihuojian / sandaihuojian, get_moves.py:22-34)
DDZ_DEV uint32_t value_key(uint64_t mv, int handnum) {
    const Trick t = classify(mv);
    const Masks a = masks_of(mv);
    if ((a.g1 & kRocket) == kRocket && (t.cat == 13 || (t.cat == 10 && t.len == 2))) return kDropped;
    const int left = min(max(handnum - card_count(mv), 0), kLeftMax);     // a hand of one deck keeps 0..20 cards
    return g_value_rank[(value2(t, a) - kC2Min) * (kLeftMax + 1) + left];
}

// number of entries of the pruned list of a list of n moves (get_moves.py:61: int(length / 3 + 1) rounds of two)
DDZ_DEV int pruned_size(int n) { return n <= kPruneAbove ? n : 2 * (n / 3 + 1); }

// Which move of the full list is entry j of the pruned list?  keys[0..n) = value_key of every move (n > kPruneAbove).
// Entry 2k is the k-th lowest kept move, entry 2k+1 the k-th highest, in the order of a STABLE ascending sort
// (get_moves.py:60-63); past the kept moves the position is clamped (the reference would raise; cannot happen for hands
// of one deck).  Binary search over the key for the position's value class, then a scan for the tie: no sort.
template <class K>
DDZ_DEV int pruned_pick(const K* keys, int n, int j) {
    int m = 0;
    for (int i = 0; i < n; i++) m += keys[i] != (K)kDropped;
    const int k = j >> 1;
    int p = (j & 1) ? m - 1 - k : k;
    p = max(0, min(p, m - 1));
    uint32_t lo = 0, hi = kDropped - 1;                     // smallest key x with  #{key <= x} > p
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        int c = 0;
        for (int i = 0; i < n; i++) c += (uint32_t)keys[i] <= mid;
        if (c > p) hi = mid; else lo = mid + 1;
    }
    int below = 0;
    for (int i = 0; i < n; i++) below += (uint32_t)keys[i] < lo;
    int tie = p - below;                                    // the tie-th move of that class, in list order
    for (int i = 0; i < n; i++)
        if ((uint32_t)keys[i] == lo && tie-- == 0) return i;
    return 0;
}

}  // namespace search
}  // namespace ddz
