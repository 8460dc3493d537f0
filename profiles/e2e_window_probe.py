"""Where does a short (20-step) e2e window spend its time?  Per step: host time to issue it, and the device time at which
the step's kernels of group 0 and the step's D2H are done (events), relative to the window start."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


def main(K=20, NG=4, refill_at=0):
    B, P, SEED = 131072, 8, 20260101
    dev = torch.device("cuda", 0)
    perm, lord = D.random_deals(B, seed=SEED, pool_games=P)
    ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=NG, seed=SEED, device=dev, max_actions_per_env=160)
    ge.prepare(perm, lord, pool_games=P)
    for _ in range(150):
        ge.rollout_step()
    ge.join(); torch.cuda.synchronize()
    host = D.HostRolloutGroups(ge)
    rng = np.random.default_rng(1)
    ent = [torch.as_tensor(rng.integers(0, 1 << 31, B, dtype=np.int64).astype(np.int32)).pin_memory() for _ in range(4)]
    pp, ll = D.random_deals(B, seed=5)
    pp, ll = torch.as_tensor(pp).pin_memory(), torch.as_tensor(ll).pin_memory()
    pending = [None] * 4
    out = {}
    for trial in ("warm", "timed"):
        for i in range(8):
            pending[i % 4] = host.step(ent[i % 4])
        for r in pending:
            D.HostRolloutGroups.wait(r)
        ge.join(); torch.cuda.synchronize()
        x0 = torch.cuda.Event(enable_timing=True); x0.record()
        ge._fork()
        evs, host_us = [], []
        t0 = time.perf_counter()
        for i in range(K):
            ta = time.perf_counter()
            if i == refill_at:
                host.refill(1, pp, ll)
            old = pending[i % 4]
            if old is not None:
                D.HostRolloutGroups.wait(old)
            pending[i % 4] = host.step(ent[i % 4])
            e = torch.cuda.Event(enable_timing=True); e.record(ge.streams[0]); evs.append(e)
            host_us.append((time.perf_counter() - ta) * 1e6)
        tb = time.perf_counter()
        for r in pending:
            D.HostRolloutGroups.wait(r)
        tc = time.perf_counter()
        ge.join()
        x1 = torch.cuda.Event(enable_timing=True); x1.record()
        torch.cuda.synchronize()
        out[trial] = {"window_ms": x0.elapsed_time(x1), "ms_per_step": x0.elapsed_time(x1) / K,
                      "host_issue_us_per_step": [round(h, 1) for h in host_us], "host_issue_total_ms": (tb - t0) * 1e3,
                      "host_final_wait_ms": (tc - tb) * 1e3,
                      "group0_kernel_done_ms": [round(x0.elapsed_time(e), 3) for e in evs]}
    print(json.dumps(out))


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
