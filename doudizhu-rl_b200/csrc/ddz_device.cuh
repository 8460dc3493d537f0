// ddz_device.cuh -- device-side Doudizhu rules for sm_100a: rank masks, move classification, legal-move
// counting (closed form) and enumeration (canonical order), deal, state transition, Philox.
//
// Representation.  A hand / move is 15 per-rank counts packed as nibbles in a uint64 ("packed").  All rule
// logic runs on four dense 15-bit rank masks g1..g4 (bit r of gk <=> count[r] >= k), i.e. the thermometer
// planes of the reference's one-hot (envi.py:140-146).  Sets of ranks are then popc / shift-AND / ffs work:
//   runs (straights, pair-straights, airplanes)  = shift-AND chains over a mask,
//   kicker choices                               = lexicographic k-subsets of a mask,
//   counts                                       = popc x binomial table.
// Canonical order of the move list = index order of the reference's action space
// (rule_based/utils/card.py:34-159), pass first when following, the 24 rocket-kicker moves of
// server/mcts/get_moves.py:22-34 where unfiltered itertools.combinations would put them.
#pragma once
#include <cstdint>

#ifdef DDZ_HOST_HARNESS
// tests/host_harness compiles this rule code for the CPU with g++ (the per-hand logic has no warp intrinsics), so that
// the enumeration can be checked against the oracle where there is no GPU.  Never defined in the product build.
#include <algorithm>
#define DDZ_DEV inline
#define __constant__ static const
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
using std::max;
using std::min;
#else
#define DDZ_DEV __device__ __forceinline__
#endif

namespace ddz {

constexpr uint64_t kDeckPacked = 0x0114444444444444ull;  // 4 of ranks 0..12, one of each joker
constexpr uint32_t kLineMask = 0x0FFFu;                  // ranks 3..A may appear in sequences (card.py:86-130)
constexpr uint32_t kRocket = 0x6000u;

struct Masks { uint32_t g1, g2, g3, g4; };

// bits at 0,4,..,28 of a 32-bit word -> dense 8 bits
DDZ_DEV uint32_t compress8(uint32_t s) {
    s = (s | (s >> 3)) & 0x03030303u;
    s = (s | (s >> 6)) & 0x000F000Fu;
    s = (s | (s >> 12)) & 0xFFu;
    return s;
}
// dense 8 bits -> bits at 0,4,..,28
DDZ_DEV uint32_t spread8(uint32_t x) {
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}
DDZ_DEV uint32_t plane15(uint32_t lo, uint32_t hi, int bit) {
    return compress8((lo >> bit) & 0x11111111u) | (compress8((hi >> bit) & 0x11111111u) << 8);
}
DDZ_DEV Masks masks_of(uint64_t packed) {
    uint32_t lo = (uint32_t)packed, hi = (uint32_t)(packed >> 32);
    uint32_t b0 = plane15(lo, hi, 0), b1 = plane15(lo, hi, 1), b2 = plane15(lo, hi, 2);
    Masks m;
    m.g4 = b2;                 // counts are 0..4: 4 = 0b100
    m.g3 = b2 | (b1 & b0);
    m.g2 = b2 | b1;
    m.g1 = b2 | b1 | b0;
    return m;
}
// packed move = mult x (one per rank of `main`) + kmult x (one per rank of `kick`)
DDZ_DEV uint64_t pack_move(uint32_t main, uint32_t mult, uint32_t kick, uint32_t kmult) {
    uint32_t lo = spread8(main & 0xFFu) * mult + spread8(kick & 0xFFu) * kmult;
    uint32_t hi = spread8(main >> 8) * mult + spread8(kick >> 8) * kmult;
    return ((uint64_t)hi << 32) | lo;
}
DDZ_DEV int card_count(uint64_t packed) {  // sum of the 15 nibbles (<= 20)
    uint64_t x = (packed & 0x0F0F0F0F0F0F0F0Full) + ((packed >> 4) & 0x0F0F0F0F0F0F0F0Full);
    return (int)((x * 0x0101010101010101ull) >> 56);
}
DDZ_DEV uint32_t above(int v) { return ~((1u << v) - 1u); }  // ranks >= v, v in 0..31

// what has to be beaten: category / length / main rank of the last non-pass hand-out
// (CardGroup type/len/value, card.py:372-527); cat 0 = free lead
struct Trick { int cat, len, val; };

DDZ_DEV Trick classify(uint64_t move) {
    Trick t; t.cat = 0; t.len = 1; t.val = 0;
    if (move == 0) return t;
    Masks a = masks_of(move);
    int n = __popc(a.g1) + __popc(a.g2) + __popc(a.g3) + __popc(a.g4);
    if (a.g4) { t.val = __ffs(a.g4) - 1; t.cat = (n == 4) ? 4 : (n == 6 ? 13 : 14); }
    else if (a.g3) {
        int k = __popc(a.g3); t.val = __ffs(a.g3) - 1; t.len = k;
        if (n == 3 * k) t.cat = (k == 1) ? 3 : 9;
        else if (n == 4 * k) t.cat = (k == 1) ? 5 : 10;
        else t.cat = (k == 1) ? 6 : 11;
    } else if (a.g2) {
        int k = __popc(a.g2); t.val = __ffs(a.g2) - 1; t.len = k; t.cat = (k == 1) ? 2 : 8;
    } else {
        t.val = __ffs(a.g1) - 1;
        if (n == 1) t.cat = 1;
        else if (n == 2) t.cat = 12;
        else { t.cat = 7; t.len = n; }
    }
    return t;
}

// the trick the player to move must beat: previous player's hand-out, else the one before (envi.py:103-110)
DDZ_DEV uint64_t last_move(uint64_t rec_prev, uint64_t rec_prevprev) { return rec_prev ? rec_prev : rec_prevprev; }

// C(n, k) for n <= 15, k <= 5 (kicker choices)
__constant__ uint16_t kBinom[16][6] = {
    {1, 0, 0, 0, 0, 0},      {1, 1, 0, 0, 0, 0},       {1, 2, 1, 0, 0, 0},        {1, 3, 3, 1, 0, 0},
    {1, 4, 6, 4, 1, 0},      {1, 5, 10, 10, 5, 1},     {1, 6, 15, 20, 15, 6},     {1, 7, 21, 35, 35, 21},
    {1, 8, 28, 56, 70, 56},  {1, 9, 36, 84, 126, 126}, {1, 10, 45, 120, 210, 252}, {1, 11, 55, 165, 330, 462},
    {1, 12, 66, 220, 495, 792}, {1, 13, 78, 286, 715, 1287}, {1, 14, 91, 364, 1001, 2002},
    {1, 15, 105, 455, 1365, 3003}};
DDZ_DEV int binom(int n, int k) { return (n < 0 || k > n) ? 0 : (int)kBinom[n][k]; }

// Which categories may be played on `t`, and from which main rank (CardGroup.bigger_than, card.py:307-325):
// lead: everything but pass.  follow: pass, same category with same len and higher value, any bomb (a higher
// one on a bomb), the rocket; nothing but pass on a rocket.
struct Rule {
    bool lead; int cat, len, val;
    DDZ_DEV bool allowed(int c) const { return lead || (cat != 12 && (c == cat || c == 4 || c == 12)); }
    DDZ_DEV uint32_t from(int c) const { return (!lead && c == cat) ? above(val + 1) : 0xFFFFFFFFu; }
    DDZ_DEV bool len_ok(int c, int L) const { return lead || c != cat || L == len; }
};
DDZ_DEV Rule rule_of(uint64_t last) {
    Trick t = classify(last);
    Rule r; r.lead = (last == 0); r.cat = t.cat; r.len = t.len; r.val = t.val;
    return r;
}

// ------------------------------------------------------------------------------------------------
// number of legal moves, closed form (popc x binomial); must equal what enumerate() emits
// ------------------------------------------------------------------------------------------------
template <int LMIN, int LMAX>
DDZ_DEV int count_lines(uint32_t src, const Rule& ru, int cat) {
    uint32_t R = src & kLineMask, t = R, from = ru.from(cat);
    int n = 0;
#pragma unroll
    for (int L = 2; L <= LMAX; L++) {
        t &= R >> (L - 1);
        if (L >= LMIN && ru.len_ok(cat, L)) n += __popc(t & from);
    }
    return n;
}
template <int LMAX>
DDZ_DEV int count_planes(uint32_t g3, int nkick, const Rule& ru, int cat) {
    uint32_t R = g3 & kLineMask, t = R, from = ru.from(cat);
    int n = 0;
#pragma unroll
    for (int L = 2; L <= LMAX; L++) {
        t &= R >> (L - 1);
        if (ru.len_ok(cat, L)) n += __popc(t & from) * binom(nkick - L, L);
    }
    return n;
}
DDZ_DEV int count_legal(const Masks& m, const Rule& ru, bool has_last) {
    if (m.g1 == 0) return has_last ? 1 : 0;  // empty hand (finished env): pass only / nothing
    int n = ru.lead ? 0 : 1;             // pass
    int n1 = __popc(m.g1), n2 = __popc(m.g2);
    if (ru.allowed(1)) n += __popc(m.g1 & ru.from(1));
    if (ru.allowed(2)) n += __popc(m.g2 & ru.from(2));
    if (ru.allowed(3)) n += __popc(m.g3 & ru.from(3));
    if (ru.allowed(4)) n += __popc(m.g4 & ru.from(4));
    if (ru.allowed(5)) n += __popc(m.g3 & ru.from(5)) * (n1 - 1);
    if (ru.allowed(6)) n += __popc(m.g3 & ru.from(6)) * (n2 - 1);
    if (ru.allowed(7)) n += count_lines<5, 12>(m.g1, ru, 7);
    if (ru.allowed(8)) n += count_lines<3, 10>(m.g2, ru, 8);
    if (ru.allowed(9)) n += count_lines<2, 6>(m.g3, ru, 9);
    if (ru.allowed(10)) n += count_planes<5>(m.g3, n1, ru, 10);
    if (ru.allowed(11)) n += count_planes<4>(m.g3, n2, ru, 11);
    if (ru.allowed(12)) n += ((m.g1 & kRocket) == kRocket);
    if (ru.allowed(13)) n += __popc(m.g4 & ru.from(13)) * binom(n1 - 1, 2);
    if (ru.allowed(14)) n += __popc(m.g4 & ru.from(14)) * binom(n2 - 1, 2);
    return n;
}

// ------------------------------------------------------------------------------------------------
// enumeration in canonical order
// ------------------------------------------------------------------------------------------------
// lexicographic successor of the k-subset c of S (the order itertools.combinations yields, card.py:115);
// returns 0 when c was the last one
DDZ_DEV uint32_t next_combo(uint32_t c, uint32_t S) {
    uint32_t U = S & ~c;
    if (U == 0) return 0;
    int pu = 31 - __clz(U);
    uint32_t top = ~((2u << pu) - 1u);
    uint32_t rest = c & ~top;
    if (rest == 0) return 0;
    int j = __popc(c & top);
    int q = 31 - __clz(rest);
    uint32_t up = S & ~((2u << q) - 1u);
    uint32_t nb = up & (0u - up);
    uint32_t out = (rest & ~(1u << q)) | nb;
    uint32_t rem = up & ~nb;
    for (int i = 0; i < j; i++) { out |= rem & (0u - rem); rem &= rem - 1; }
    return out;
}
// idx-th k-subset of S in lexicographic order (idx 0 = the k lowest elements)
DDZ_DEV uint32_t unrank_combo(uint32_t S, int k, int idx) {
    uint32_t out = 0;
    int n = __popc(S);
    for (int i = 0; i < k; i++) {
        for (;;) {
            const int c = binom(n - 1, k - i - 1);     // subsets whose next element is the lowest one left
            const uint32_t low = S & (0u - S);
            S ^= low; n--;
            if (idx < c) { out |= low; break; }
            idx -= c;
        }
    }
    return out;
}

DDZ_DEV uint32_t first_combo(uint32_t S, int k) {
    uint32_t c = 0;
    for (int i = 0; i < k; i++) { c |= S & (0u - S); S &= S - 1; }
    return c;
}
template <class F>
DDZ_DEV void each_kicker_set(uint32_t main, uint32_t mult, uint32_t S, int k, uint32_t kmult, F& f) {
    if (__popc(S) < k) return;
    uint64_t base = pack_move(main, mult, 0, 0);
    for (uint32_t c = first_combo(S, k); c; c = next_combo(c, S)) f(base + pack_move(c, kmult, 0, 0));
}
// MULT cards of one rank: the nibble goes straight to its place (no mask spreading)
DDZ_DEV uint64_t pack_rank(int rank, uint32_t mult) { return (uint64_t)mult << (4 * rank); }
// MULT cards of each of the L consecutive ranks s .. s+L-1 (1 <= L <= 12)
DDZ_DEV uint64_t pack_run(int s, int L, uint32_t mult) { return ((0x1111111111111111ull >> (64 - 4 * L)) * mult) << (4 * s); }
template <int MULT, class F>
DDZ_DEV void each_rank(uint32_t mask, F& f) {
    while (mask) { const int r = __ffs(mask) - 1; mask &= mask - 1; f(pack_rank(r, MULT)); }
}
template <int MULT, int LMIN, int LMAX, class F>
DDZ_DEV void each_line(uint32_t src, const Rule& ru, int cat, F& f) {
    uint32_t R = src & kLineMask;
    // starts of runs of at least LMIN
    uint32_t t = R;
#pragma unroll
    for (int L = 2; L <= LMIN; L++) t &= R >> (L - 1);
    t &= ru.from(cat);
    while (t) {
        int s = __ffs(t) - 1; t &= t - 1;
        uint32_t run = ((1u << LMIN) - 1u) << s;
        for (int L = LMIN; L <= LMAX; L++) {
            if ((R & run) != run) break;
            if (ru.len_ok(cat, L)) f(pack_run(s, L, MULT));
            run |= run << 1;
        }
    }
}
template <int LMAX, int KMULT, class F>
DDZ_DEV void each_plane(uint32_t g3, uint32_t kicksrc, const Rule& ru, int cat, F& f) {
    uint32_t R = g3 & kLineMask;
    uint32_t t = R & (R >> 1) & ru.from(cat);
    while (t) {
        int s = __ffs(t) - 1; t &= t - 1;
        uint32_t run = 3u << s;
        for (int L = 2; L <= LMAX; L++) {
            if ((R & run) != run) break;
            if (ru.len_ok(cat, L)) each_kicker_set(run, 3, kicksrc & ~run, L, KMULT, f);
            run |= run << 1;
        }
    }
}
// one env per lane, one loop nest per category with compile-time constants (the fast path for short lists);
// f(move) is called in canonical order
template <class F>
DDZ_DEV void enumerate_legal(const Masks& m, const Rule& ru, bool has_last, F& f) {
    if (m.g1 == 0) { if (has_last) f(0ull); return; }
    if (!ru.lead) f(0ull);
    if (ru.allowed(1)) each_rank<1>(m.g1 & ru.from(1), f);
    if (ru.allowed(2)) each_rank<2>(m.g2 & ru.from(2), f);
    if (ru.allowed(3)) each_rank<3>(m.g3 & ru.from(3), f);
    if (ru.allowed(4)) each_rank<4>(m.g4 & ru.from(4), f);
    if (ru.allowed(5)) {
        uint32_t mains = m.g3 & ru.from(5);
        while (mains) {
            const int r = __ffs(mains) - 1; mains &= mains - 1;
            const uint64_t base = pack_rank(r, 3);
            uint32_t ks = m.g1 & ~(1u << r);
            while (ks) { const int q = __ffs(ks) - 1; ks &= ks - 1; f(base + pack_rank(q, 1)); }
        }
    }
    if (ru.allowed(6)) {
        uint32_t mains = m.g3 & ru.from(6);
        while (mains) {
            const int r = __ffs(mains) - 1; mains &= mains - 1;
            const uint64_t base = pack_rank(r, 3);
            uint32_t ks = m.g2 & ~(1u << r);
            while (ks) { const int q = __ffs(ks) - 1; ks &= ks - 1; f(base + pack_rank(q, 2)); }
        }
    }
    if (ru.allowed(7)) each_line<1, 5, 12>(m.g1, ru, 7, f);
    if (ru.allowed(8)) each_line<2, 3, 10>(m.g2, ru, 8, f);
    if (ru.allowed(9)) each_line<3, 2, 6>(m.g3, ru, 9, f);
    if (ru.allowed(10)) each_plane<5, 1>(m.g3, m.g1, ru, 10, f);
    if (ru.allowed(11)) each_plane<4, 2>(m.g3, m.g2, ru, 11, f);
    if (ru.allowed(12) && (m.g1 & kRocket) == kRocket) f(pack_move(kRocket, 1, 0, 0));
    if (ru.allowed(13)) {
        uint32_t mains = m.g4 & ru.from(13);
        while (mains) { uint32_t b = mains & (0u - mains); mains ^= b; each_kicker_set(b, 4, m.g1 & ~b, 2, 1, f); }
    }
    if (ru.allowed(14)) {
        uint32_t mains = m.g4 & ru.from(14);
        while (mains) { uint32_t b = mains & (0u - mains); mains ^= b; each_kicker_set(b, 4, m.g2 & ~b, 2, 2, f); }
    }
}

#ifndef DDZ_HOST_HARNESS
// ------------------------------------------------------------------------------------------------
// warp-cooperative enumeration of ONE env (same list, same order), for envs with many legal moves.
// All 32 lanes walk the canonical group sequence with identical control flow; a group of c moves starting at
// list index `off` is expanded by the lanes in parallel (lane j -> j-th set bit / j-th chunk of the kicker
// combinations, found by unranking), and f(index, packed_move) stores it.  A thread-per-env enumeration of a
// 300-move hand is a 300-step dependent chain in one lane; this makes it ~10 steps per lane.
// ------------------------------------------------------------------------------------------------

// c kicker sets of size k out of S for one main group, split into 32 contiguous chunks
template <class F>
DDZ_DEV void coop_kicker_sets(uint32_t main, uint32_t mult, uint32_t S, int k, uint32_t kmult, int off, int lane, F& f) {
    const int c = binom(__popc(S), k);
    const int per = (c + 31) >> 5;
    const int i0 = lane * per, i1 = min(c, i0 + per);
    if (i0 >= i1) return;
    const uint64_t base = pack_move(main, mult, 0, 0);
    uint32_t combo = unrank_combo(S, k, i0);
    for (int i = i0; i < i1; i++) { f(off + i, base + pack_move(combo, kmult, 0, 0)); combo = next_combo(combo, S); }
}
template <int MULT, class F>
DDZ_DEV int coop_ranks(uint32_t mask, int off, int lane, F& f) {
    // lane l owns rank l: its place in the list is the number of set bits below it (no search for the j-th bit)
    if ((mask >> lane) & 1u) f(off + __popc(mask & ((1u << lane) - 1u)), pack_rank(lane, MULT));
    return __popc(mask);
}
template <class F>
DDZ_DEV int coop_main_plus_one(uint32_t mains, uint32_t kicksrc, uint32_t kmult, int off, int lane, F& f) {
    // uniform loop over the trios; lane l owns kicker rank l (its list index = set bits below it)
    int n = 0;
    while (mains) {
        const int mr = __ffs(mains) - 1; mains &= mains - 1;
        const uint32_t ks = kicksrc & ~(1u << mr);
        if ((ks >> lane) & 1u) f(off + n + __popc(ks & ((1u << lane) - 1u)), pack_rank(mr, 3) + pack_rank(lane, kmult));
        n += __popc(ks);
    }
    return n;
}
template <int MULT, int LMIN, int LMAX, class F>
DDZ_DEV int coop_lines(uint32_t src, const Rule& ru, int cat, int off, int lane, F& f) {
    const uint32_t R = src & kLineMask;
    uint32_t t = R;
#pragma unroll
    for (int L = 2; L <= LMIN; L++) t &= R >> (L - 1);
    t &= ru.from(cat);
    const bool same = !ru.lead && cat == ru.cat;
    int n = 0;
    while (t) {                                            // uniform: every lane sees the same starts
        const int s = __ffs(t) - 1; t &= t - 1;
        const int run = __ffs(~(R >> s)) - 1;              // length of the run of ones starting at s
        const int maxL = min(run, LMAX);
        const int c = same ? ((ru.len >= LMIN && ru.len <= maxL) ? 1 : 0) : (maxL - LMIN + 1);
        if (lane < c) {
            const int L = same ? ru.len : LMIN + lane;
            f(off + n + lane, pack_run(s, L, MULT));
        }
        n += c;
    }
    return n;
}
template <int LMAX, int KMULT, class F>
DDZ_DEV int coop_planes(uint32_t g3, uint32_t kicksrc, const Rule& ru, int cat, int off, int lane, F& f) {
    const uint32_t R = g3 & kLineMask;
    uint32_t t = R & (R >> 1) & ru.from(cat);
    int n = 0;
    while (t) {
        const int s = __ffs(t) - 1; t &= t - 1;
        uint32_t run = 3u << s;
        for (int L = 2; L <= LMAX; L++) {
            if ((R & run) != run) break;
            if (ru.len_ok(cat, L)) {
                const uint32_t S = kicksrc & ~run;
                coop_kicker_sets(run, 3, S, L, KMULT, off + n, lane, f);
                n += binom(__popc(S), L);
            }
            run |= run << 1;
        }
    }
    return n;
}
template <int KMULT, class F>
DDZ_DEV int coop_four_two(uint32_t mains, uint32_t kicksrc, int off, int lane, F& f) {
    int n = 0;
    while (mains) {
        const uint32_t b = mains & (0u - mains); mains ^= b;
        const uint32_t S = kicksrc & ~b;
        coop_kicker_sets(b, 4, S, 2, KMULT, off + n, lane, f);
        n += binom(__popc(S), 2);
    }
    return n;
}
// returns the number of moves (must equal count_legal); f(index, move) is called once per move by some lane
template <class F>
DDZ_DEV int enumerate_legal_warp(const Masks& m, const Rule& ru, bool has_last, int lane, F& f) {
    if (m.g1 == 0) { if (has_last && lane == 0) f(0, 0ull); return has_last ? 1 : 0; }
    int off = 0;
    if (!ru.lead) { if (lane == 0) f(0, 0ull); off = 1; }
    if (ru.allowed(1)) off += coop_ranks<1>(m.g1 & ru.from(1), off, lane, f);
    if (ru.allowed(2)) off += coop_ranks<2>(m.g2 & ru.from(2), off, lane, f);
    if (ru.allowed(3)) off += coop_ranks<3>(m.g3 & ru.from(3), off, lane, f);
    if (ru.allowed(4)) off += coop_ranks<4>(m.g4 & ru.from(4), off, lane, f);
    if (ru.allowed(5)) off += coop_main_plus_one(m.g3 & ru.from(5), m.g1, 1, off, lane, f);
    if (ru.allowed(6)) off += coop_main_plus_one(m.g3 & ru.from(6), m.g2, 2, off, lane, f);
    if (ru.allowed(7)) off += coop_lines<1, 5, 12>(m.g1, ru, 7, off, lane, f);
    if (ru.allowed(8)) off += coop_lines<2, 3, 10>(m.g2, ru, 8, off, lane, f);
    if (ru.allowed(9)) off += coop_lines<3, 2, 6>(m.g3, ru, 9, off, lane, f);
    if (ru.allowed(10)) off += coop_planes<5, 1>(m.g3, m.g1, ru, 10, off, lane, f);
    if (ru.allowed(11)) off += coop_planes<4, 2>(m.g3, m.g2, ru, 11, off, lane, f);
    if (ru.allowed(12) && (m.g1 & kRocket) == kRocket) { if (lane == 0) f(off, pack_move(kRocket, 1, 0, 0)); off++; }
    if (ru.allowed(13)) off += coop_four_two<1>(m.g4 & ru.from(13), m.g1, off, lane, f);
    if (ru.allowed(14)) off += coop_four_two<2>(m.g4 & ru.from(14), m.g2, off, lane, f);
    return off;
}

#endif  // DDZ_HOST_HARNESS

// ------------------------------------------------------------------------------------------------
// idx-th legal move in canonical order WITHOUT enumerating the list: walk the categories with their closed-form
// counts, then unrank inside the one group that holds it (O(#groups), not O(#moves)).  idx must be < count_legal.
// Used by the playout kernel (a random rollout needs one move per decision, not the list).
// ------------------------------------------------------------------------------------------------
#ifdef DDZ_HOST_HARNESS
DDZ_DEV uint32_t nth_bit(uint32_t mask, int j) { while (j-- > 0) mask &= mask - 1; return mask & (0u - mask); }
#else
DDZ_DEV uint32_t nth_bit(uint32_t mask, int j) { return 1u << __fns(mask, 0, j + 1); }   // j-th set bit (0-based)
#endif
template <int MULT, int LMIN, int LMAX>
DDZ_DEV bool select_line(uint32_t src, const Rule& ru, int cat, int& idx, uint64_t& mv) {
    const uint32_t R = src & kLineMask;
    uint32_t t = R;
#pragma unroll
    for (int L = 2; L <= LMIN; L++) t &= R >> (L - 1);
    t &= ru.from(cat);
    const bool same = !ru.lead && cat == ru.cat;
    while (t) {
        const int s = __ffs(t) - 1; t &= t - 1;
        const int maxL = min(__ffs(~(R >> s)) - 1, LMAX);
        const int c = same ? ((ru.len >= LMIN && ru.len <= maxL) ? 1 : 0) : (maxL - LMIN + 1);
        if (idx < c) { const int L = same ? ru.len : LMIN + idx; mv = pack_move(((1u << L) - 1u) << s, MULT, 0, 0); return true; }
        idx -= c;
    }
    return false;
}
template <int LMAX, int KMULT>
DDZ_DEV bool select_plane(uint32_t g3, uint32_t kicksrc, const Rule& ru, int cat, int& idx, uint64_t& mv) {
    const uint32_t R = g3 & kLineMask;
    uint32_t t = R & (R >> 1) & ru.from(cat);
    while (t) {
        const int s = __ffs(t) - 1; t &= t - 1;
        uint32_t run = 3u << s;
        for (int L = 2; L <= LMAX; L++, run |= run << 1) {
            if ((R & run) != run) break;
            if (!ru.len_ok(cat, L)) continue;
            const uint32_t S = kicksrc & ~run;
            const int c = binom(__popc(S), L);
            if (idx < c) { mv = pack_move(run, 3, unrank_combo(S, L, idx), KMULT); return true; }
            idx -= c;
        }
    }
    return false;
}
template <int KMULT>
DDZ_DEV bool select_four_two(uint32_t mains, uint32_t kicksrc, int nk, int& idx, uint64_t& mv) {
    const int per = binom(nk - 1, 2), c = __popc(mains) * per;     // every bomb rank is itself in kicksrc
    if (idx >= c) { idx -= c; return false; }
    const int mi = idx / per;
    const uint32_t main = nth_bit(mains, mi);
    mv = pack_move(main, 4, unrank_combo(kicksrc & ~main, 2, idx - mi * per), KMULT);
    return true;
}
DDZ_DEV uint64_t select_legal(const Masks& m, const Rule& ru, bool has_last, int idx) {
    if (m.g1 == 0) return 0ull;
    if (!ru.lead) { if (idx == 0) return 0ull; idx--; }
    uint64_t mv = 0;
    const int n1 = __popc(m.g1), n2 = __popc(m.g2);
#define DDZ_SELECT_RANKS(CAT, MASK, MULT)                                                     \
    if (ru.allowed(CAT)) {                                                                    \
        const uint32_t mk = (MASK) & ru.from(CAT); const int c = __popc(mk);                  \
        if (idx < c) return pack_move(nth_bit(mk, idx), MULT, 0, 0);                          \
        idx -= c;                                                                             \
    }
    DDZ_SELECT_RANKS(1, m.g1, 1) DDZ_SELECT_RANKS(2, m.g2, 2) DDZ_SELECT_RANKS(3, m.g3, 3) DDZ_SELECT_RANKS(4, m.g4, 4)
#undef DDZ_SELECT_RANKS
#define DDZ_SELECT_MAIN_PLUS_ONE(CAT, KSRC, NK, KMULT)                                        \
    if (ru.allowed(CAT)) {                                                                    \
        const uint32_t mains = m.g3 & ru.from(CAT); const int nk = (NK) - 1, c = __popc(mains) * nk; \
        if (idx < c) {                                                                        \
            const int mi = idx / nk; const uint32_t main = nth_bit(mains, mi);                \
            return pack_move(main, 3, nth_bit((KSRC) & ~main, idx - mi * nk), KMULT);         \
        }                                                                                     \
        idx -= c;                                                                             \
    }
    DDZ_SELECT_MAIN_PLUS_ONE(5, m.g1, n1, 1) DDZ_SELECT_MAIN_PLUS_ONE(6, m.g2, n2, 2)
#undef DDZ_SELECT_MAIN_PLUS_ONE
    if (ru.allowed(7) && select_line<1, 5, 12>(m.g1, ru, 7, idx, mv)) return mv;
    if (ru.allowed(8) && select_line<2, 3, 10>(m.g2, ru, 8, idx, mv)) return mv;
    if (ru.allowed(9) && select_line<3, 2, 6>(m.g3, ru, 9, idx, mv)) return mv;
    if (ru.allowed(10) && select_plane<5, 1>(m.g3, m.g1, ru, 10, idx, mv)) return mv;
    if (ru.allowed(11) && select_plane<4, 2>(m.g3, m.g2, ru, 11, idx, mv)) return mv;
    if (ru.allowed(12) && (m.g1 & kRocket) == kRocket) { if (idx == 0) return pack_move(kRocket, 1, 0, 0); idx--; }
    if (ru.allowed(13) && select_four_two<1>(m.g4 & ru.from(13), m.g1, n1, idx, mv)) return mv;
    if (ru.allowed(14) && select_four_two<2>(m.g4 & ru.from(14), m.g2, n2, idx, mv)) return mv;
    return ~0ull;   // idx out of range
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, counter (env_lo, env_hi, step, 0), key (seed_lo, seed_hi): the action-index stream
// ------------------------------------------------------------------------------------------------
DDZ_DEV uint32_t philox(uint64_t seed, uint64_t env, uint32_t step) {
    uint32_t c0 = (uint32_t)env, c1 = (uint32_t)(env >> 32), c2 = step, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// ------------------------------------------------------------------------------------------------
// per-env state in registers
// ------------------------------------------------------------------------------------------------
struct Env {
    uint64_t hand[3], hist[3], recent[3];
    uint32_t meta;
    DDZ_DEV int cur() const { return meta & 3; }
    DDZ_DEV bool done() const { return (meta >> 2) & 1; }
};
struct StateView {  // SoA arrays inside the caller's state buffer
    uint64_t* f;    // [9][B]
    uint32_t* meta; // [B]
    int B;
};
DDZ_DEV StateView view_of(void* state, int B) {
    StateView v; v.f = (uint64_t*)state; v.meta = (uint32_t*)(v.f + 9 * (size_t)B); v.B = B;
    return v;
}
DDZ_DEV Env load_env(const StateView& v, int b) {
    Env e;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        e.hand[q] = v.f[(size_t)q * v.B + b];
        e.hist[q] = v.f[(size_t)(3 + q) * v.B + b];
        e.recent[q] = v.f[(size_t)(6 + q) * v.B + b];
    }
    e.meta = v.meta[b];
    return e;
}
DDZ_DEV void store_env(const StateView& v, int b, const Env& e) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
        v.f[(size_t)q * v.B + b] = e.hand[q];
        v.f[(size_t)(3 + q) * v.B + b] = e.hist[q];
        v.f[(size_t)(6 + q) * v.B + b] = e.recent[q];
    }
    v.meta[b] = e.meta;
}
template <class T>
DDZ_DEV T pick3(int i, T a, T b, T c) { return i == 0 ? a : (i == 1 ? b : c); }

DDZ_DEV uint64_t hand_to_move(const Env& e) { int s = e.cur(); return pick3(s, e.hand[0], e.hand[1], e.hand[2]); }
DDZ_DEV uint64_t trick_of(const Env& e) {
    int s = e.cur();
    uint64_t p1 = pick3((s + 2) % 3, e.recent[0], e.recent[1], e.recent[2]);
    uint64_t p2 = pick3((s + 1) % 3, e.recent[0], e.recent[1], e.recent[2]);
    return last_move(p1, p2);
}

#ifndef DDZ_HOST_HARNESS
// Warp-cooperative deal.  `need` = ballot of the lanes whose env must be re-dealt; for each of them the
// whole warp reads the 54-byte permutation row (lane l < 27 loads cards 2l and 2l+1 as one 16-bit word), builds the
// four piles with hardware warp reductions (nibble sums never carry: a rank has at most 4 cards) and checks that the
// row is a permutation of 0..53.  One HBM round trip and ~60 instructions per deal, no per-lane 54-step chain.
// Returns the ballot of lanes whose deal was refused (bad permutation / lord_pile): their env is left untouched.
DDZ_DEV unsigned int warp_deal(unsigned int need, Env& e, const int8_t* __restrict__ perm,
                               const int8_t* __restrict__ lord_pile, int pool_games, int B, int b, int lane) {
    unsigned int bad = 0;
    const unsigned long long myrow = (unsigned long long)((e.meta >> 8) % (uint32_t)pool_games) * (unsigned long long)B + (unsigned long long)b;
    while (need) {
        const int src = __ffs(need) - 1; need &= need - 1;
        const unsigned long long row = __shfl_sync(0xFFFFFFFFu, myrow, src);
        int lp = 0;
        if (lane == src && lord_pile) lp = lord_pile[row];
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, seen_lo = 0, seen_hi = 0;
        bool ok = true;
        if (lane < 27) {
            const uint32_t v = reinterpret_cast<const uint16_t*>(perm + 54 * row)[lane];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int i = 2 * lane + c;
                int id = (int)(int8_t)((v >> (8 * c)) & 0xFFu);
                ok = ok && (id >= 0 && id < 54);
                id = (id < 0) ? 0 : (id > 53 ? 53 : id);
                if (id < 32) seen_lo |= 1u << id; else seen_hi |= 1u << (id - 32);
                const int rank = id < 52 ? (id >> 2) : id - 39;
                const int p = i < 17 ? 0 : (i < 34 ? 1 : (i < 51 ? 2 : 3));
                const uint32_t nib = 1u << (4 * (rank & 7));
#pragma unroll
                for (int q = 0; q < 4; q++) if (p == q) { if (rank < 8) lo[q] += nib; else hi[q] += nib; }
            }
        }
        uint64_t pile[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            pile[q] = ((uint64_t)__reduce_add_sync(0xFFFFFFFFu, hi[q]) << 32) | __reduce_add_sync(0xFFFFFFFFu, lo[q]);
        seen_lo = __reduce_or_sync(0xFFFFFFFFu, seen_lo);
        seen_hi = __reduce_or_sync(0xFFFFFFFFu, seen_hi);
        ok = __all_sync(0xFFFFFFFFu, ok) && seen_lo == 0xFFFFFFFFu && seen_hi == 0x003FFFFFu;
        if (lane == src) {
            ok = ok && lp >= 0 && lp <= 2;
            if (ok) {
                e.hand[1] = pick3(lp, pile[0], pile[1], pile[2]) + pile[3];
                e.hand[2] = pick3((lp + 1) % 3, pile[0], pile[1], pile[2]);
                e.hand[0] = pick3((lp + 2) % 3, pile[0], pile[1], pile[2]);
#pragma unroll
                for (int q = 0; q < 3; q++) { e.hist[q] = 0; e.recent[q] = 0; }
                const uint32_t games = (e.meta >> 8) + 1;
                e.meta = 1u | (e.meta & 0x20u) | (games << 8);  // lord to move, not done, keep sticky error
            } else {    // refuse: leave an empty, finished env with the sticky error bit
#pragma unroll
                for (int q = 0; q < 3; q++) { e.hand[q] = e.hist[q] = e.recent[q] = 0; }
                e.meta = (e.meta & 0xFFFFFF00u) | 1u | 4u | 0x20u;
                bad |= 1u << src;
            }
        }
    }
    return __reduce_or_sync(0xFFFFFFFFu, bad);
}

#endif  // DDZ_HOST_HARNESS

struct StepOut { int r, done, cat; float reward[3]; bool applied, pass; int winner; };

// envi.py:38-43 _update + native step_manual: remove cards, record, terminal, rotate
DDZ_DEV StepOut apply_move(Env& e, uint64_t move, const int32_t* rewards) {
    StepOut o; o.r = 0; o.cat = classify(move).cat; o.applied = true; o.pass = (move == 0); o.winner = -1;
    o.reward[0] = o.reward[1] = o.reward[2] = 0.f;
    int s = e.cur();
#pragma unroll
    for (int q = 0; q < 3; q++)
        if (q == s) { e.hand[q] -= move; e.hist[q] += move; e.recent[q] = move; }
    uint64_t h = pick3(s, e.hand[0], e.hand[1], e.hand[2]);
    uint32_t meta = e.meta;
    if (move != 0 && h == 0) {
        meta |= 4u | ((uint32_t)s << 3);
        o.r = (s == 1) ? -1 : 1; o.winner = s;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            bool win = (s == 1) ? (q == 1) : (q != 1);
            o.reward[q] = win ? (float)rewards[q] : -(float)rewards[q];
        }
    }
    meta = (meta & ~3u) | (uint32_t)((s + 1) % 3);
    e.meta = meta;
    o.done = (meta >> 2) & 1;
    return o;
}

}  // namespace ddz
