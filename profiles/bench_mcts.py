"""The search bot on the device: wall time of mcts(payload) for the reference's budget of 1000 UCT
iterations (server/mcts/interface.py:37) at several playout widths, on a mid-game position and on an opening, with the
bot's pruned move lists (the reference's configuration) and with the full lists."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


def main():
    rng = np.random.default_rng(3)
    deck = [min(i // 4 + 3, 15) if i < 52 else 16 + (i - 52) for i in range(54)]
    cards = [int(c) for c in rng.permutation(deck)[:31]]
    payload = {"role_id": 1, "hand_card": {0: sorted(cards[:10]), 1: sorted(cards[10:22]), 2: sorted(cards[22:31])},
               "last_taken": {0: [], 1: [], 2: []}}
    cards = [int(c) for c in rng.permutation(deck)]
    opening = {"role_id": 1, "hand_card": {0: sorted(cards[:17]), 1: sorted(cards[17:37]), 2: sorted(cards[37:54])},
               "last_taken": {0: [], 1: [], 2: []}}
    out = {"runs": []}
    D.mcts(payload, computation_budget=50)                      # warm-up (allocations, first launches)
    D.mcts(payload, computation_budget=50, prune=False)
    for name, pl in (("mid-game 10/12/9 cards", payload), ("opening 17/20/17 cards", opening)):
        for prune in (True, False):                             # True = the reference's own configuration
            for width in (1, 16, 256):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                move, s = D.mcts(pl, computation_budget=1000, width=width, seed=5, return_search=True, prune=prune)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                moves, visits, rate = s.root_table()
                out["runs"].append({"position": name, "pruned_lists": prune, "width": width, "seconds": dt,
                                    "iterations": s.iterations, "playouts": s.playouts,
                                    "us_per_iteration": dt / s.iterations * 1e6, "move": move,
                                    "root_moves": int(len(moves)), "best_win_rate": float(rate.max())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
