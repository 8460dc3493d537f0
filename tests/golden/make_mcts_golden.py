"""Generate tests/golden/mcts_moves.npz from the REFERENCE's search-bot code, run unmodified:

    python tests/golden/make_mcts_golden.py [/root/reference]

  cards_value   server/mcts/evaluator.py:17-57 -- the whole table (13 527 keys), keys as count vectors + values
  get_moves     server/mcts/get_moves.py:36-69 -- the pruned move list for positions cut from random games and for the
                adversarial hands of doudizhu-rl_b200/data/adversarial_pool.npz; `r.get_moves` underneath is the oracle's
                definitional generator (oracle/pyshim/r.py: the reference's native `r` is absent), so what is pinned is the
                pruning rule -- filter, value, stable sort, lowest/highest interleave -- not the native's list order

Only input/output vectors are stored; nothing of the reference is copied.
"""
import importlib.util
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyshim"))
sys.path.insert(0, os.path.join(REF, "server"))        # the `mcts` package
sys.path.insert(0, REF)                                # envi.py and ITS config.py first

np.int = int
np.bool = bool

from mcts.evaluator import cards_value  # noqa: E402
from mcts.get_moves import get_moves  # noqa: E402
from oracle import ddz_oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("ddz_deals", os.path.join(ROOT, "doudizhu-rl_b200", "deals.py"))
deals = importlib.util.module_from_spec(spec)
spec.loader.exec_module(deals)

INDEX = [str(i) for i in range(3, 14)] + ["1", "2", "14", "15"]       # get_moves.py:39


def ref_list(hand, last):
    handcards = dict(zip(INDEX, [int(x) for x in hand]))
    lastcards = [int(INDEX[r]) for r in range(15) for _ in range(int(last[r]))]
    out = get_moves(handcards, lastcards)
    return np.array([[d[k] for k in INDEX] for d in out], np.int8).reshape(-1, 15)


def main():
    keys = np.array(list(cards_value.keys()), np.int8)
    vals = np.array([float(v) for v in cards_value.values()], np.float64)

    hands, lasts = [], []
    rng = np.random.default_rng(20261018)
    B = 48
    rb = O.RefBatch(B, 0)
    perm, lord = deals.random_deals(B, seed=5)
    rb.deal(perm, lord)
    for t in range(60):                                    # positions along random games (index stream from rng)
        offs, acts, _, _ = rb.observe(want_f32=False, want_face=False)
        for b in range(B):
            if rb.envs["done"][b] or rng.random() > 0.25:
                continue
            e = rb.envs[b]
            cur = int(e["cur"])
            prev, pp = e["recent"][(cur + 2) % 3], e["recent"][(cur + 1) % 3]
            hands.append(e["hand"][cur].copy())
            lasts.append((prev if prev.any() else pp).copy())
        rb.step(rng.integers(0, 1 << 30, B).astype(np.int32), mode=1)
    ah, al = deals.adversarial_pairs(96, seed=3)           # the long lists
    hands += list(O.unpack(ah))
    lasts += list(O.unpack(al))
    # both jokers + bombs: the dropped rocket-kicker moves
    for h, l in (([4, 4, 4, 0, 1, 1, 0, 0, 0, 0, 0, 0, 2, 1, 1], [0] * 15),
                 ([4, 4, 0, 3, 3, 0, 0, 2, 0, 0, 0, 0, 0, 1, 1], [0] * 15),
                 ([0, 4, 4, 4, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 1], [4, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0]),
                 ([3, 3, 3, 0, 0, 1, 1, 2, 0, 0, 0, 0, 0, 1, 1], [0] * 15)):
        hands.append(np.array(h, np.int8)); lasts.append(np.array(l, np.int8))
    hands, lasts = np.array(hands, np.int8), np.array(lasts, np.int8)
    lists = [ref_list(h, l) for h, l in zip(hands, lasts)]
    offsets = np.zeros(len(lists) + 1, np.int32)
    offsets[1:] = np.cumsum([len(x) for x in lists])
    full = np.array([len(O.get_moves(h, l, fast=True)) for h, l in zip(hands, lasts)], np.int32)
    path = os.path.join(HERE, "mcts_moves.npz")
    np.savez_compressed(path, value_keys=keys, value_vals=vals, hands=hands, lasts=lasts, offsets=offsets,
                        moves=np.concatenate(lists), full_counts=full)
    print("wrote", path, os.path.getsize(path), "bytes;", len(keys), "values,", len(lists), "positions,",
          int((full > 10).sum()), "pruned,", int(offsets[-1]), "moves")


if __name__ == "__main__":
    main()
