"""The B = 1 drop-in views driven exactly like the reference's loop (game.py:259-275 with step_random for every seat):
per decision env.face, env.valid_actions(), env.step_random().  Decisions per second of ONE env -- the latency-bound
regime a user of the unmodified game.py sees (the reference: 1.96e3/s on one core, profiles/r1n_reference_envi_*)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


def main(games=40):
    env = D.EnvCooperation(seed=1)
    steps = 0
    env.reset(); env.prepare()
    for _ in range(30):
        env.face; env.valid_actions(); env.step_random()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g in range(games):
        env.reset()
        env.prepare()
        done = False
        while not done:
            face = env.face
            actions = env.valid_actions()
            _, done, _ = env.step_random()
            steps += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"workload": "EnvCooperation (B=1 view): face + valid_actions + step_random per decision",
           "games": games, "decisions": steps, "decisions_per_s": steps / dt, "us_per_decision": dt / steps * 1e6}
    # the agent's path (game.py:95-104): step_manual with a row of valid_actions(), as dqn.py returns it
    steps2 = 0
    t0 = time.perf_counter()
    for g in range(games):
        env.reset()
        env.prepare()
        done = False
        while not done:
            face = env.face
            actions = env.valid_actions()
            _, done, _ = env.step_manual(actions[(steps2 * 7) % actions.shape[0]])
            steps2 += 1
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
    out["step_manual_us_per_decision"] = dt2 / steps2 * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
