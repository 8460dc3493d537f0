// search_host.cpp -- CPU harness for the search bot's move pruning (doudizhu-rl_b200/csrc/ddz_search.cuh).
// TEST INFRASTRUCTURE ONLY: compiles the product's value key, rank table and pruned_pick for the host with g++
// (-DDDZ_HOST_HARNESS) so that tests/ can compare them with the oracle and the golden vectors where there is no GPU.
// This is the per-lane path of k_playout<true>; k_mcts_moves (one warp per list) has its own -m gpu test.
#define DDZ_HOST_HARNESS
#include <cstdint>
#include <vector>
#include "../../doudizhu-rl_b200/csrc/ddz_search.cuh"

using namespace ddz;

struct Collect {
    std::vector<uint64_t>* v;
    void operator()(uint64_t mv) { v->push_back(mv); }
};

extern "C" int search_host_value2(uint64_t mv) { return search::value2(classify(mv), masks_of(mv)); }
extern "C" uint32_t search_host_value_key(uint64_t mv, int handnum) { return search::value_key(mv, handnum); }

// the pruned list of (hand, last), entry by entry through pruned_pick, as the playout draws from it; returns its length
extern "C" int search_host_pruned_list(uint64_t hand, uint64_t last, uint64_t* out, int cap) {
    std::vector<uint64_t> moves;
    Collect c{&moves};
    const Masks hm = masks_of(hand);
    const Rule ru = rule_of(last);
    enumerate_legal(hm, ru, last != 0, c);
    const int n = (int)moves.size();
    if (n != count_legal(hm, ru, last != 0)) return -1;
    const int size = search::pruned_size(n);
    if (size > cap) return -2;
    if (n <= search::kPruneAbove) { for (int i = 0; i < n; i++) out[i] = moves[i]; return n; }
    std::vector<uint16_t> keys(n);
    const int handnum = card_count(hand);
    for (int i = 0; i < n; i++) keys[i] = (uint16_t)search::value_key(moves[i], handnum);
    for (int j = 0; j < size; j++) {
        const int i = search::pruned_pick(keys.data(), n, j);
        if (select_legal(hm, ru, last != 0, i) != moves[i]) return -3;      // what the playout then plays
        out[j] = moves[i];
    }
    return size;
}
