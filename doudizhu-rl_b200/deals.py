"""Synthetic inputs of the benchmarks and tests: deck permutations and the adversarial (hand, last) pairs of BASELINE
config 5.  numpy only -- no torch, no native library -- so that `bench.py --impl reference` (the CPU arm) and the product
build their workloads from the same code without the CPU arm ever loading libddz_b200.so.  bench.py loads this file by
path (importing the package would load the library)."""
import os

import numpy as np

DEAL_SEED0 = 20260101                    # default deal stream: PCG64(DEAL_SEED0 + game).permutation(54)
_HERE = os.path.dirname(os.path.abspath(__file__))
POOL_FILE = os.path.join(_HERE, "data", "adversarial_pool.npz")

# the 497-move hand, 5-trio airplanes, bombs + rocket, 12-straights + pairs, ...
ADVERSARIAL_POOL = np.array([
    [1, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0], [3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0],
    [4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1], [3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 2, 2, 1, 1],
    [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1], [2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0],
    [3, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 0, 0], [4, 4, 0, 0, 0, 0, 2, 2, 2, 2, 0, 0, 0, 1, 1],
    [1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 0, 0, 0], [2, 2, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],
], np.int8)


def default_deals(first_game, n):
    """The documented default shuffle stream: game g -> PCG64(20260101+g).permutation(54), lord_pile 0."""
    perm = np.stack([np.random.Generator(np.random.PCG64(DEAL_SEED0 + first_game + i)).permutation(54)
                     for i in range(n)]).astype(np.int8)
    return perm, np.zeros(n, np.int8)


def random_deals(n, seed, pool_games=1):
    """Vectorised synthetic deals for large batches: int8 [pool_games*n, 54] permutations, random landlord pile."""
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.permuted(np.tile(np.arange(54, dtype=np.int8), (pool_games * n, 1)), axis=1)
    lord = rng.integers(0, 3, size=pool_games * n, dtype=np.int8)
    return perm, lord


def pack_counts_np(counts):
    """int [...,15] per-rank counts -> uint64 packed nibbles (the library's hand / move format)"""
    c = np.asarray(counts).astype(np.uint64)
    return (c << (np.arange(15, dtype=np.uint64) * np.uint64(4))).sum(-1, dtype=np.uint64)


def adversarial_pairs(n, seed=5):
    """BASELINE config 5 input: n (hand, last) pairs as packed uint64 -- hands drawn from ADVERSARIAL_POOL, half of them
    leading (last = 0), half following a random legal lead move of the pool hands (the committed list in POOL_FILE, made
    by tests/golden/make_adversarial_pool.py).  Returns (hands uint64[n], lasts uint64[n])."""
    lead_moves = np.load(POOL_FILE)["lead_moves"]
    rng = np.random.default_rng(seed)
    hands = pack_counts_np(ADVERSARIAL_POOL)[rng.integers(0, len(ADVERSARIAL_POOL), n)]
    lasts = np.zeros(n, np.uint64)
    follow = rng.random(n) < 0.5
    lasts[follow] = lead_moves[rng.integers(0, len(lead_moves), int(follow.sum()))]
    return hands, lasts
