"""Steady-state DRAM traffic of the hot kernels, for `roofline.traffic` (profiles/r2_traffic.json).

Run under ncu in RANGE replay mode, so that the counters cover >= 10 consecutive steps with every env group in flight
(concurrent launches on four streams) instead of one isolated, cold-cache launch whose last rows never leave the L2:

    ncu --replay-mode range --profile-from-start off --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file out.csv \
        python profiles/traffic_probe.py config4 [steps]      # or: config5;  DDZ_NO_COMPRESSION=1 for plain row buffers

The profiled range is bracketed by cudaProfilerStart/Stop; the script prints how many steps it holds and the algorithmic
bytes per step, so that  traffic per step = (read + write) / steps.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


def config4(steps, groups=4):
    B, P, SEED = 131072, 8, 20260101
    perm, lord = D.random_deals(B, seed=SEED, pool_games=P)
    ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=groups, seed=SEED, max_actions_per_env=160)
    ge.prepare(perm, lord, pool_games=P)
    for _ in range(150):
        ge.rollout_step()
    ge.join(); torch.cuda.synchronize()
    s0 = ge.stats.clone()
    torch.cuda.profiler.start()
    for _ in range(steps):
        ge.rollout_step()
    ge.join(); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    d = (ge.stats - s0).cpu().numpy()
    nbar = float(d[8]) / float(d[4])
    alg = B * (2 * 76 + 8 + 8 + 15 + 4 + 240 * 9 + 248 * nbar)
    return {"workload": "config 4: %d envs, %d groups in flight" % (B, groups), "steps_in_range": steps, "launches_in_range": steps * groups,
            "mean_legal_moves": nbar, "algorithmic_bytes_per_step": alg,
            "row_buffers": "compressible" if hasattr(ge.envs[0]._face, "_ddz_rows") else "plain", "errors": int(ge.stats[7].item())}


def config5(steps):
    n = 131072
    h, l = D.adversarial_pairs(n, seed=5)
    hands, lasts = torch.as_tensor(h.view(np.int64)).cuda(), torch.as_tensor(l.view(np.int64)).cuda()
    gen = D.MoveGenerator(n)
    for _ in range(5):
        gen.generate(hands, lasts)
    torch.cuda.synchronize()
    total = int(gen.offsets[n].item())
    torch.cuda.profiler.start()
    for _ in range(steps):
        gen.generate(hands, lasts)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return {"workload": "config 5: %d adversarial pairs" % n, "steps_in_range": steps, "launches_in_range": steps,
            "moves_per_step": total, "algorithmic_bytes_per_step": 20.0 * n + 8.0 * total, "errors": int(gen.stats[7].item())}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "config4"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    print(json.dumps(config4(steps) if which == "config4" else config5(steps)))
