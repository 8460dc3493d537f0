// flat_host.cpp -- CPU harness for the flat long-list enumerator (doudizhu-rl_b200/csrc/ddz_flat.cuh).
// TEST INFRASTRUCTURE ONLY: compiles the product's per-hand rule code (descriptor walk, decode, subset table) for the host
// with g++ (-DDDZ_HOST_HARNESS) and replays what one warp of k_legal_flat does -- counts, arena rounds, the 64-move
// windows with their boundary bitmaps -- with the warp primitives written as loops over 32 lanes.  tests/ compares its
// output with the oracle where there is no GPU.  Nothing in the product loads this.
#define DDZ_HOST_HARNESS
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../doudizhu-rl_b200/csrc/ddz_flat.cuh"

using namespace ddz;

static const flat::SubsetTable kTable;

extern "C" int flat_host_table(uint16_t* out) { memcpy(out, kTable.v, sizeof(uint16_t) * flat::kTableSize); return flat::kTableSize; }

// one tile of nh <= 32 hands; out[cap] receives the concatenated lists, offsets[nh+1] the CSR offsets; returns total
extern "C" long long flat_host_tile(const uint64_t* hands, const uint64_t* lasts, int nh, uint64_t* out, long long cap,
                                    int32_t* offsets, int arena_size, long long base, int* rounds_out) {
    if (nh < 0 || nh > 32 || arena_size < 160) return -1;
    Masks hm[32]; Rule ru[32]; bool hl[32]; int n[32], ng[32], local[33], dloc[33];
    std::vector<uint8_t> lists(32 * flat::kListsPerHand * flat::kListStride, 0xEE);
    for (int l = 0; l < 32; l++) {
        n[l] = ng[l] = 0;
        if (l < nh) {
            hm[l] = masks_of(hands[l]); ru[l] = rule_of(lasts[l]); hl[l] = lasts[l] != 0;
            flat::CountSink cs; flat::walk_groups(hm[l], ru[l], hl[l], 4 * l, cs);
            n[l] = cs.n; ng[l] = cs.ng;
            if (n[l] != count_legal(hm[l], ru[l], hl[l])) return -2;     // the closed forms must agree with the walk
            if (ng[l] != flat::count_groups(hm[l], ru[l], hl[l])) return -9;
            flat::write_lists(hm[l], 4 * l, lists.data());
        }
    }
    local[0] = dloc[0] = 0;
    for (int l = 0; l < 32; l++) { local[l + 1] = local[l] + n[l]; dloc[l + 1] = dloc[l] + ng[l]; }
    const int total = local[32];
    for (int l = 0; l <= nh; l++) offsets[l] = (int32_t)(base + local[l]);
    std::vector<flat::Desc> arena(arena_size + 1);
    int h0 = 0, rounds = 0;
    while (h0 < 32) {
        const int D0 = dloc[h0];
        int h1 = h0;
        while (h1 < 32 && dloc[h1 + 1] - D0 <= arena_size) h1++;
        if (h1 == h0) return -3;
        const int m0 = local[h0], m1 = local[h1], nd = dloc[h1] - D0;
        for (int l = h0; l < h1; l++) {
            if (l >= nh) continue;
            flat::WriteSink ws{arena.data(), dloc[l] - D0, arena_size, (uint32_t)local[l]};
            flat::walk_groups(hm[l], ru[l], hl[l], 4 * l, ws);
            if (ws.d != dloc[l + 1] - D0 || ws.start != (uint32_t)local[l + 1]) return -4;
        }
        arena[nd].start = (uint32_t)m1; arena[nd].prm = 0;
        rounds++;
        long long lim = cap - base; if (lim > total) lim = total; if (lim < 0) lim = 0;
        const int mEnd = (int)std::min<long long>(m1, lim);
        const int sh = (int)((base + m0) & 1);
        int gcur = 0;
        for (int w0 = m0 - sh; w0 < mEnd; w0 += 64) {
            uint32_t blo = 0, bhi = 0;
            for (int gq = gcur + 1;; gq += 32) {              // boundary bitmap of the window, 32 descriptors per pass
                bool more = false;
                for (int lane = 0; lane < 32; lane++) {
                    const int idx = gq + lane;
                    const uint32_t rel = idx <= nd ? arena[idx].start - (uint32_t)w0 : 64u;
                    if (rel < 32u) blo |= 1u << rel; else if (rel < 64u) bhi |= 1u << (rel - 32);
                    if (lane == 31) more = rel < 64u;
                }
                if (!more) break;
            }
            for (int lane = 0; lane < 32; lane++) {
                const int x = 2 * lane;
                const uint32_t mlo = x < 31 ? (2u << x) - 1u : 0xFFFFFFFFu, mhi = x < 32 ? 0u : (2u << (x - 32)) - 1u;
                const int g0 = gcur + __popc(blo & mlo) + __popc(bhi & mhi);
                const int b1 = (x + 1 < 32) ? (blo >> (x + 1)) & 1 : (bhi >> (x + 1 - 32)) & 1;
                const int g1 = g0 + b1;
                const int i0 = w0 + x, i1 = i0 + 1;
                const bool v0 = i0 >= m0 && i0 < mEnd, v1 = i1 < mEnd;
                if (!v0 && !v1) continue;
                const flat::Desc d0 = arena[g0], d1 = arena[g1];
                const int c0 = (int)(arena[g0 + 1].start - d0.start), c1 = (int)(arena[g1 + 1].start - d1.start);
                if (v0 && !(d0.start <= (uint32_t)i0 && (uint32_t)i0 < arena[g0 + 1].start)) return -5;
                if (v1 && !(d1.start <= (uint32_t)i1 && (uint32_t)i1 < arena[g1 + 1].start)) return -6;
                uint64_t mv0, mv1;                             // the kernel's pair decode, cross-checked with the single one
                flat::decode2(d0.prm, i0 - (int)d0.start, c0, v0, d1.prm, i1 - (int)d1.start, c1, v1, kTable.v, lists.data(), mv0, mv1);
                if (v0) { if (mv0 != flat::decode(d0.prm, i0 - (int)d0.start, c0, kTable.v, lists.data())) return -7; out[base + i0] = mv0; }
                if (v1) { if (mv1 != flat::decode(d1.prm, i1 - (int)d1.start, c1, kTable.v, lists.data())) return -8; out[base + i1] = mv1; }
            }
            gcur += __popc(blo) + __popc(bhi);
        }
        h0 = h1;
    }
    if (rounds_out) *rounds_out = rounds;
    return total;
}
