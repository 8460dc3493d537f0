"""Build-container only (needs /root/reference): time the reference's UNMODIFIED Python env wrapper (envi.py) on top of
the oracle's stand-ins for its absent natives -- the loop of Game.compete with random seats (game.py:259-275,
envi.py:79-85): per decision env.face + env.valid_actions() + env.step_random().  One process, one core, like the
reference.  This is the closest thing to "the reference's own CPU path" that can be run anywhere; the C port timed by
bench.py's cpu_baseline has none of this interpreter overhead."""
import json
import os
import sys
import time

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyshim"))
sys.path.insert(0, REF)
np.int = int
np.bool = bool
import envi  # noqa: E402
import env as shim_env  # noqa: E402

shim_env.reset_deal_stream(0)
e = envi.EnvCooperation()
games, steps = 60, 0
t0 = time.perf_counter()
for g in range(games):
    e.reset()
    e.prepare()
    done = False
    while not done:
        face = e.face                       # what the agent would consume (game.py:95)
        acts = e.valid_actions()            # tensor form (game.py:101)
        _, done, _ = e.step_random()
        steps += 1
sec = time.perf_counter() - t0
print(json.dumps({"what": "unmodified reference envi.EnvCooperation over oracle-backed natives: face + valid_actions + step_random per decision",
                  "games": games, "env_steps": steps, "seconds": sec, "env_steps_per_s_one_core": steps / sec,
                  "where": "build container CPU, 1 process"}))
