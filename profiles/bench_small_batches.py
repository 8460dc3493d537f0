"""Latency-bound regime: env-steps/s of one chain of CUDA-graph-replayed steps for small batches (BASELINE config 2 is
4096 envs).  One launch of k_env per step; below ~30 000 envs the step time is the kernel's fixed latency."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D

out = []
for B in (1, 256, 4096, 16384, 65536):
    perm, lord = D.random_deals(B, seed=1, pool_games=4)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    env = D.BatchedEnvCooperation(B, seed=2)
    env.prepare(pd, ld, pool_games=4)
    for _ in range(150):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=4)
    gr = D.GraphedRollout(env, pd, ld, 4)
    for _ in range(5):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 400
    out.append({"envs": B, "us_per_step": ms * 1e3, "env_steps_per_s": B / ms * 1e3, "errors": int(env.stats[7].item())})
    print(json.dumps(out[-1]), flush=True)
