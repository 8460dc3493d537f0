"""20-step windows of the four-group replay, each opened right after a device-wide synchronize as bench.py's are, with the
groups' FIRST launches staggered on the host by 0 / 10 / 20 / 30 / 40 us: ms per step, several windows per setting, interleaved.
Needs the experiment's `stagger_us` argument of GroupedEnv.replay (a host busy-wait between two groups' graph launches), which
is not in the shipped env.py: the result was "no effect" (profiles/r3x_window_start_probes.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddz_b200 as D


def main():
    B, NG, P, SPG, K = 131072, 4, 8, 2, 20
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
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
    settings = (0.0, 10.0, 20.0, 30.0, 40.0)
    res = {str(s): [] for s in settings}
    for _ in range(rounds):
        for s in settings:
            for _ in range(3):
                ge.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K // SPG):
                ge.replay(stagger_us=s if i == 0 else 0.0)
            ge.join()
            e1.record()
            torch.cuda.synchronize()
            res[str(s)].append(e0.elapsed_time(e1) / K)
    out = {k: {"mean": sum(v) / len(v), "min": min(v), "max": max(v), "all": [round(x, 4) for x in v]} for k, v in res.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
