"""BASELINE config 5: legal-move generation microbench on adversarial synthetic hands (one GPU per process).

131 072 (hand, last) pairs per call drawn from an adversarial pool (the 497-move hand, 5-trio airplanes, bombs + rocket,
12-straights + pairs, ...), 50 % leads / 50 % following a random legal previous move.  Output = the packed CSR list only
(ddz_legal_moves).  Algorithmic bytes per call = 16 (hand, last) + 8 N (list) + 4 (offset) per pair.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D

POOL = [
    [1, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0], [3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0],
    [4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1], [3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 2, 2, 1, 1],
    [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1], [2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0],
    [3, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 0, 0], [4, 4, 0, 0, 0, 0, 2, 2, 2, 2, 0, 0, 0, 1, 1],
    [1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 0, 0, 0], [2, 2, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],
]


def main():
    n, reps = 131072, 50
    rng = np.random.default_rng(5)
    pool = np.array(POOL, np.int8)
    # previous moves to follow: legal lead moves of the pool hands themselves (computed on the GPU)
    lead_moves, lead_off = D.get_moves(pool, np.zeros_like(pool))
    lead_moves = lead_moves.cpu().numpy()
    hands = pool[rng.integers(0, len(pool), n)]
    lasts = np.zeros(n, np.int64)
    follow = rng.random(n) < 0.5
    lasts[follow] = lead_moves[rng.integers(0, len(lead_moves), int(follow.sum()))]
    hp = D.pack_counts(torch.as_tensor(hands).cuda()).contiguous()
    lp = torch.as_tensor(lasts).cuda()
    gen = D.MoveGenerator(n)
    for _ in range(5):
        gen.generate(hp, lp)
    torch.cuda.synchronize()
    total = int(gen.offsets[n].item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gen.generate(hp, lp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = n * 20 + 8 * total
    peak = 6534.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"workload": "config 5: %d adversarial (hand,last) pairs, 50%% lead" % n, "ms_per_call": ms,
                      "hands_per_s": n / ms * 1e3, "moves_per_s": total / ms * 1e3, "mean_moves": total / n,
                      "algorithmic_GBs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                      "errors": int(gen.stats[7].item())}))


if __name__ == "__main__":
    main()
