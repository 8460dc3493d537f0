"""Where does a 20-step window of the four-group replay spend its time?  Per group and per graph replay (two env-steps), the
completion time relative to the start of the window, for several windows in a row -- each opened, as bench.py does, right
after a device-wide synchronize.  Shows whether the first replays of a window are slower than the later ones (the groups
start in phase) and how the windows differ from one another."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddz_b200 as D


def main():
    B, NG, P, SPG, K = 131072, 4, 8, 2, 20
    windows = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    perm, lord = D.random_deals(B, seed=20260101, pool_games=P)
    ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=NG, seed=20260101, max_actions_per_env=160)
    ge.prepare(perm, lord, pool_games=P)
    for _ in range(150):
        ge.rollout_step()
    ge.join()
    ge.capture(steps_per_graph=SPG)
    for _ in range(12):
        ge.replay()
    torch.cuda.synchronize()
    out = {"windows": []}
    n = K // SPG
    for w in range(windows):
        for _ in range(3):
            ge.replay()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n)] for _ in range(NG)]
        e0.record()
        for i in range(n):
            ge.replay()
            for g in range(NG):
                ev[g][i].record(ge.streams[g])
        torch.cuda.synchronize()
        t = [[e0.elapsed_time(ev[g][i]) for i in range(n)] for g in range(NG)]
        end = max(t[g][-1] for g in range(NG))
        per_replay = [[round(t[g][i] - (t[g][i - 1] if i else 0.0), 4) for i in range(n)] for g in range(NG)]
        out["windows"].append({"ms_per_step": end / K, "replay_ms_by_group": per_replay})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
