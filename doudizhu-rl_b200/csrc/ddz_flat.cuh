// ddz_flat.cuh -- flat (move-parallel) enumeration of long legal-move lists: r.get_moves (reference envi.py:111,
// server/mcts/get_moves.py:22-34) for hands with hundreds of moves (BASELINE config 5).
//
// The per-category walks of ddz_device.cuh cost a fixed price per category and hand, paid by a whole warp, with few
// lanes busy.  Here the list of a hand is described first and expanded afterwards:
//
//   phase A (one lane per hand)  walk the categories once and emit one 8-byte DESCRIPTOR per group of moves that share
//                                their main cards: (first list index, kind, main run, kicker source, kicker count).
//                                A 497-move hand has ~50 groups.
//   phase B (one lane per MOVE)  move i of the tile belongs to the last descriptor whose first index is <= i; inside its
//                                group it is number j, and its cards are  main run + the j-th k-subset of the kicker
//                                source  in the order itertools.combinations yields them (card.py:115).
//
// The j-th k-subset is a table lookup: lexicographic index j over an n-set = colexicographic index C(n,k)-1-j counted
// from the TOP of the set, and colexicographic k-subsets of {0..n-1} are simply the integers with k bits in ascending
// order -- one table per k serves every n (909 entries x 2 bytes, staged in shared memory).  The kicker source is the
// hand's thermometer plane (count >= 1 or >= 2) without the main ranks, which are consecutive in it: position q of the
// source is position q (+ run length if q is at or above the run) of the plane's descending rank list, kept per hand
// in shared memory.  Every lane does the same work whatever category its move is in; there is no per-category cost.
#pragma once
#include "ddz_device.cuh"

namespace ddz {
namespace flat {

// colex k-subset tables: T_k = integers with k bits below 2^nmax(k), ascending; T_0 = {0}.  nmax bounds the kicker
// sources of ANY 15-rank hand: k=1 rank groups (15), k=2 four-with-two (14), k=3/4/5 airplanes of 3/4/5 trios (15 - k).
constexpr int kNmax[6] = {0, 15, 14, 12, 11, 10};
constexpr int kTableSize = 909;   // 1 + 15 + C(14,2) + C(12,3) + C(11,4) + C(10,5)

struct SubsetTable {
    uint16_t v[kTableSize + 3];
    constexpr SubsetTable() : v() {
        int idx = 0;
        v[idx++] = 0;
        for (int k = 1; k <= 5; k++) {
            uint32_t m = (1u << k) - 1u;
            const uint32_t lim = 1u << kNmax[k];
            while (m < lim) {                                  // Gosper: next integer with the same number of bits
                v[idx++] = (uint16_t)m;
                const uint32_t c = m & (0u - m), r = m + c;
                m = (((r ^ m) >> 2) / c) | r;
            }
        }
    }
};

constexpr int kListStride = 17;          // bytes per descending rank list (16 entries + 1: no bank conflicts between hands)
constexpr int kListsPerHand = 4;         // planes count >= 1, 2, 3, 4

// descriptor word `prm`
//   bit 0      1 = line group (straight / pair straight / trio straight from one start), 0 = main + kicker subsets
//   bits 1-4   s      lowest main rank
//   kicker groups: bits 5-7 L (main run length, >= 1), 8-10 mult (cards per main rank; 0 = no main: pass),
//                  11-17 list (4 * hand slot + plane; a kicker rank contributes plane + 1 cards), 18-21 ap (descending
//                  position of the main run in the list; 15 = the run is not in the list), 22-31 first entry of the
//                  subset table T_k (k = kickers per move)
//   line groups:   bits 5-8 lmin (length of the group's first move), 9-11 mult
constexpr int kTableFirst[6] = {0, 1, 16, 107, 327, 657};
DDZ_DEV uint32_t kick_prm(int s, int L, int mult, int list, int ap, int k) {
    const int toff = k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 16 : k == 3 ? 107 : k == 4 ? 327 : 657;
    return ((uint32_t)s << 1) | ((uint32_t)L << 5) | ((uint32_t)mult << 8) | ((uint32_t)list << 11) |
           ((uint32_t)ap << 18) | ((uint32_t)toff << 22);
}
DDZ_DEV uint32_t line_prm(int s, int lmin, int mult) {
    return 1u | ((uint32_t)s << 1) | ((uint32_t)lmin << 5) | ((uint32_t)mult << 9);
}

// Groups of the legal-move list of one hand in canonical order (the order of enumerate_legal, ddz_device.cuh).
// sk.add(prm, cnt) is called once per group with cnt > 0 moves.
template <int MULT, int LMIN, int LMAX, class Sink>
DDZ_DEV void line_groups(uint32_t src, const Rule& ru, int cat, Sink& sk) {
    const uint32_t R = src & kLineMask;
    uint32_t t = R;
#pragma unroll
    for (int L = 2; L <= LMIN; L++) t &= R >> (L - 1);
    t &= ru.from(cat);
    const bool same = !ru.lead && cat == ru.cat;
    while (t) {
        const int s = __ffs(t) - 1; t &= t - 1;
        const int maxL = min(__ffs(~(R >> s)) - 1, LMAX);      // length of the run of ones starting at s
        if (same) { if (ru.len >= LMIN && ru.len <= maxL) sk.add(line_prm(s, ru.len, MULT), 1); }
        else sk.add(line_prm(s, LMIN, MULT), maxL - LMIN + 1);
    }
}
template <int LMAX, class Sink>
DDZ_DEV void plane_groups(uint32_t g3, uint32_t ksrc, int list, const Rule& ru, int cat, Sink& sk) {
    const uint32_t R = g3 & kLineMask;
    uint32_t t = R & (R >> 1) & ru.from(cat);
    const int nk = __popc(ksrc);
    while (t) {
        const int s = __ffs(t) - 1; t &= t - 1;
        uint32_t run = 3u << s;
        for (int L = 2; L <= LMAX; L++, run |= run << 1) {
            if ((R & run) != run) break;
            if (!ru.len_ok(cat, L)) continue;
            const int cnt = binom(nk - L, L);                  // the trios' own ranks are in ksrc (g3 within g2 within g1)
            if (cnt > 0) sk.add(kick_prm(s, L, 3, list, __popc(ksrc >> (s + L)), L), cnt);
        }
    }
}
template <class Sink>
DDZ_DEV void walk_groups(const Masks& m, const Rule& ru, bool has_last, int list0, Sink& sk) {
    const uint32_t kPass = kick_prm(0, 1, 0, list0, 15, 0);
    if (m.g1 == 0) { if (has_last) sk.add(kPass, 1); return; }
    if (!ru.lead) sk.add(kPass, 1);
    const int n1 = __popc(m.g1), n2 = __popc(m.g2);
    // solo, pair, trio, bomb: the ranks of plane c at or above the trick's value = the top cnt entries of its list
#pragma unroll
    for (int c = 1; c <= 4; c++)
        if (ru.allowed(c)) {
            const int cnt = __popc((c == 1 ? m.g1 : c == 2 ? m.g2 : c == 3 ? m.g3 : m.g4) & ru.from(c));
            if (cnt > 0) sk.add(kick_prm(0, 1, 0, list0 + c - 1, 15, 1), cnt);
        }
#pragma unroll
    for (int v = 0; v < 2; v++)                               // trio + solo, trio + pair
        if (ru.allowed(5 + v)) {
            const uint32_t ksrc = v ? m.g2 : m.g1;
            const int cnt = (v ? n2 : n1) - 1;
            uint32_t mains = m.g3 & ru.from(5 + v);
            while (mains && cnt > 0) {
                const int r = __ffs(mains) - 1; mains &= mains - 1;
                sk.add(kick_prm(r, 1, 3, list0 + v, __popc(ksrc >> (r + 1)), 1), cnt);
            }
        }
    if (ru.allowed(7)) line_groups<1, 5, 12>(m.g1, ru, 7, sk);
    if (ru.allowed(8)) line_groups<2, 3, 10>(m.g2, ru, 8, sk);
    if (ru.allowed(9)) line_groups<3, 2, 6>(m.g3, ru, 9, sk);
    if (ru.allowed(10)) plane_groups<5>(m.g3, m.g1, list0, ru, 10, sk);
    if (ru.allowed(11)) plane_groups<4>(m.g3, m.g2, list0 + 1, ru, 11, sk);
    if (ru.allowed(12) && (m.g1 & kRocket) == kRocket) sk.add(kick_prm(13, 2, 1, list0, 15, 0), 1);
#pragma unroll
    for (int v = 0; v < 2; v++)                               // four with two solos / two pairs
        if (ru.allowed(13 + v)) {
            const uint32_t ksrc = v ? m.g2 : m.g1;
            const int cnt = binom((v ? n2 : n1) - 1, 2);
            uint32_t mains = m.g4 & ru.from(13 + v);
            while (mains && cnt > 0) {
                const int b = __ffs(mains) - 1; mains &= mains - 1;
                sk.add(kick_prm(b, 1, 4, list0 + v, __popc(ksrc >> (b + 1)), 2), cnt);
            }
        }
}

// number of groups walk_groups() emits, in closed form (the number of moves is count_legal(), ddz_device.cuh)
template <int LMIN, int LMAX>
DDZ_DEV int count_line_groups(uint32_t src, const Rule& ru, int cat) {
    const uint32_t R = src & kLineMask;
    const bool same = !ru.lead && cat == ru.cat;
    if (same && (ru.len < LMIN || ru.len > LMAX)) return 0;
    const int need = same ? ru.len : LMIN;                     // a start counts if its run is at least this long
    uint32_t t = R;
    for (int L = 2; L <= need; L++) t &= R >> (L - 1);
    return __popc(t & ru.from(cat));
}
template <int LMAX>
DDZ_DEV int count_plane_groups(uint32_t g3, int nk, const Rule& ru, int cat) {
    const uint32_t R = g3 & kLineMask, from = ru.from(cat);
    uint32_t t = R;
    int ng = 0;
#pragma unroll
    for (int L = 2; L <= LMAX; L++) {
        t &= R >> (L - 1);
        if (ru.len_ok(cat, L) && nk - L >= L) ng += __popc(t & from);
    }
    return ng;
}
DDZ_DEV int count_groups(const Masks& m, const Rule& ru, bool has_last) {
    if (m.g1 == 0) return has_last ? 1 : 0;
    int ng = ru.lead ? 0 : 1;
    const int n1 = __popc(m.g1), n2 = __popc(m.g2);
    if (ru.allowed(1)) ng += (m.g1 & ru.from(1)) != 0;
    if (ru.allowed(2)) ng += (m.g2 & ru.from(2)) != 0;
    if (ru.allowed(3)) ng += (m.g3 & ru.from(3)) != 0;
    if (ru.allowed(4)) ng += (m.g4 & ru.from(4)) != 0;
    if (ru.allowed(5) && n1 > 1) ng += __popc(m.g3 & ru.from(5));
    if (ru.allowed(6) && n2 > 1) ng += __popc(m.g3 & ru.from(6));
    if (ru.allowed(7)) ng += count_line_groups<5, 12>(m.g1, ru, 7);
    if (ru.allowed(8)) ng += count_line_groups<3, 10>(m.g2, ru, 8);
    if (ru.allowed(9)) ng += count_line_groups<2, 6>(m.g3, ru, 9);
    if (ru.allowed(10)) ng += count_plane_groups<5>(m.g3, n1, ru, 10);
    if (ru.allowed(11)) ng += count_plane_groups<4>(m.g3, n2, ru, 11);
    if (ru.allowed(12)) ng += ((m.g1 & kRocket) == kRocket);
    if (ru.allowed(13) && n1 - 1 >= 2) ng += __popc(m.g4 & ru.from(13));
    if (ru.allowed(14) && n2 - 1 >= 2) ng += __popc(m.g4 & ru.from(14));
    return ng;
}

struct CountSink {   // number of moves and of groups
    int n = 0, ng = 0;
    DDZ_DEV void add(uint32_t, int cnt) { n += cnt; ng++; }
};
struct Desc { uint32_t start, prm; };   // 8 bytes: one LDS.64
struct WriteSink {   // descriptors [d, dlim) of the arena; `start` = list index of the group's first move
    Desc* arena; int d, dlim; uint32_t start;
    DDZ_DEV void add(uint32_t prm, int cnt) {
        if (d >= 0 && d < dlim) { Desc x; x.start = start; x.prm = prm; arena[d] = x; }
        d++; start += (uint32_t)cnt;
    }
};

// the four descending rank lists of a hand: lists[(list0 + c) * kListStride + q] = 4 * (q-th highest rank with count > c),
// i.e. the shift that puts a card of that rank into its nibble
DDZ_DEV void write_lists(const Masks& m, int list0, uint8_t* lists) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t mask = c == 0 ? m.g1 : c == 1 ? m.g2 : c == 2 ? m.g3 : m.g4;
        uint8_t* l = lists + (list0 + c) * kListStride;
        while (mask) { const int r = 31 - __clz(mask); *l++ = (uint8_t)(4 * r); mask ^= 1u << r; }
    }
}

// the main cards of a kicker group: mult cards of each of the L <= 5 ranks s .. s+L-1 (20 bits before the shift)
DDZ_DEV uint64_t main_cards(uint32_t prm) {
    const uint32_t run = (0x11111u >> (20 - 4 * ((prm >> 5) & 7))) * ((prm >> 8) & 7);
    return (uint64_t)run << (4 * ((prm >> 1) & 15));
}
// move number j of a group of cnt moves
DDZ_DEV uint64_t decode(uint32_t prm, int j, int cnt, const uint16_t* __restrict__ table, const uint8_t* __restrict__ lists) {
    if (prm & 1u) return pack_run((prm >> 1) & 15, (int)((prm >> 5) & 15) + j, (prm >> 9) & 7);
    uint64_t mv = main_cards(prm);
    const int L = (prm >> 5) & 7, ap = (prm >> 18) & 15;
    const uint32_t list = (prm >> 11) & 127;
    const uint64_t kmult = (list & 3) + 1;
    const uint8_t* lst = lists + kListStride * (int)list;
    uint32_t P = table[(int)(prm >> 22) + cnt - 1 - j];
    while (P) {
        int q = __ffs(P) - 1; P &= P - 1;
        if (q >= ap) q += L;
        mv += kmult << lst[q];
    }
    return mv;
}
// two moves at once (the pair a lane owns): one loop over the kickers of both, so that their shared-memory loads and
// shift-adds interleave
DDZ_DEV void decode2(uint32_t prm0, int j0, int cnt0, bool v0, uint32_t prm1, int j1, int cnt1, bool v1,
                     const uint16_t* __restrict__ table, const uint8_t* __restrict__ lists, uint64_t& mv0, uint64_t& mv1) {
    const bool k0 = v0 && !(prm0 & 1u), k1 = v1 && !(prm1 & 1u);
    mv0 = (v0 && (prm0 & 1u)) ? pack_run((prm0 >> 1) & 15, (int)((prm0 >> 5) & 15) + j0, (prm0 >> 9) & 7) : (k0 ? main_cards(prm0) : 0ull);
    mv1 = (v1 && (prm1 & 1u)) ? pack_run((prm1 >> 1) & 15, (int)((prm1 >> 5) & 15) + j1, (prm1 >> 9) & 7) : (k1 ? main_cards(prm1) : 0ull);
    const int L0 = (prm0 >> 5) & 7, ap0 = (prm0 >> 18) & 15, L1 = (prm1 >> 5) & 7, ap1 = (prm1 >> 18) & 15;
    const uint32_t list0 = (prm0 >> 11) & 127, list1 = (prm1 >> 11) & 127;
    const uint64_t km0 = (list0 & 3) + 1, km1 = (list1 & 3) + 1;
    const uint8_t* l0 = lists + kListStride * (int)list0;
    const uint8_t* l1 = lists + kListStride * (int)list1;
    uint32_t P0 = k0 ? table[(int)(prm0 >> 22) + cnt0 - 1 - j0] : 0u;
    uint32_t P1 = k1 ? table[(int)(prm1 >> 22) + cnt1 - 1 - j1] : 0u;
    while (P0 | P1) {
        if (P0) { int q = __ffs(P0) - 1; P0 &= P0 - 1; q += (q >= ap0) ? L0 : 0; mv0 += km0 << l0[q]; }
        if (P1) { int q = __ffs(P1) - 1; P1 &= P1 - 1; q += (q >= ap1) ? L1 : 0; mv1 += km1 << l1[q]; }
    }
}

}  // namespace flat
}  // namespace ddz
