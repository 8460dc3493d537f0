"""Long soak at full size: ALL 131 072 envs of BASELINE config 4's per-GPU slice, N steps (default 2000, ~44 games per env,
the deal pool wrapping around many times) through the shipped launch shape (four env groups, CUDA-graph replays), then the
complete state of every env and the counters against the oracle's replay of the same run (ddz_ref_rollout_export, all host
threads).  One wrong legal list, index or transition anywhere in 2.6e8 env-steps changes a state."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D
from oracle import ddz_oracle as O


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    B, P, seed = 131072, 8, 31415
    perm, lord = D.random_deals(B, seed=17, pool_games=P)
    ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=4, seed=seed, max_actions_per_env=160)
    ge.prepare(perm, lord, pool_games=P)
    ge.capture(steps_per_graph=2)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps // 2):
        ge.replay()
    ge.join()
    torch.cuda.synchronize()
    gpu_s = time.time() - t0
    t0 = time.time()
    n, ostats, f, meta = O.rollout_export(B, steps, 2, seed, perm, lord, P, os.cpu_count() or 1)
    cpu_s = time.time() - t0
    gf = torch.cat([e._fields()[0] for e in ge.envs], dim=1).cpu().numpy().view(np.uint64)
    gm = torch.cat([e._fields()[1] for e in ge.envs]).cpu().numpy().view(np.uint32)
    st = ge.stats.cpu().numpy()
    ok = bool(np.array_equal(gf, f) and np.array_equal(gm, meta) and
              np.array_equal(st[[0, 1, 2, 3, 4, 5, 6, 9]], ostats[[0, 1, 2, 3, 4, 5, 6, 9]]) and st[7] == 0)
    print(json.dumps({"envs": B, "steps": steps, "env_steps": int(n), "games_finished": int(st[0]), "errors": int(st[7]),
                      "lord_win_rate": float(st[1]) / max(1, int(st[0])), "all_states_and_counters_equal_the_oracle": ok,
                      "gpu_seconds": round(gpu_s, 3), "oracle_seconds": round(cpu_s, 2), "oracle_threads": os.cpu_count()}))
    assert ok


if __name__ == "__main__":
    main()
