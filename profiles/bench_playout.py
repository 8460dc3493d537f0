"""Random playouts (MCTS default policy) with ddz_playout: 131 072 fresh deals played to the end in one launch."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D

B = 131072
perm, lord = D.random_deals(B, seed=1, pool_games=4)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
env = D.BatchedEnv(B, seed=2)
times = []
for rep in range(4):
    env.prepare(pd, ld, pool_games=4)
    s0 = env.stats.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    env.playout(max_steps=400)
    e1.record(); torch.cuda.synchronize()
    d = (env.stats - s0).cpu().numpy()
    times.append((e0.elapsed_time(e1), int(d[4]), int(d[0]), int(d[1])))
ms, steps, games, lordw = min(times)
print(json.dumps({"workload": "%d random playouts from fresh deals to the end of the game, one launch" % B, "ms": ms,
                  "env_steps": steps, "env_steps_per_s": steps / ms * 1e3, "games_per_s": games / ms * 1e3,
                  "mean_decisions_per_game": steps / games, "lord_win_rate": lordw / games}))
