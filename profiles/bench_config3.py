"""BASELINE config 3: 65 536 envs, DQN lord self-play with Q-scoring of every legal action, farmers random (1 GPU).

The lord's decision = argmax_a Q(face, a) over its legal moves with a network of the reference's NetCooperation contract
(10 x 15 x 4 input, 256-wide (1,k)-stride-(1,4) convolutions, net.py:125-139; random-init weights, no checkpoint ships).
Only the envs where it is the lord's turn are scored (about a third of them per step); the farmers' envs take a random
legal move.
Reports env-steps/s and how the step time splits between the env kernel and the network.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddz_b200 as D
from qnet_like import QNetLike


class Autocast(torch.nn.Module):
    """the same network under torch.autocast (bf16 tensor cores) -- a consumer-side choice, reported separately"""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward_state_action(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return self.inner.forward_state_action(x).float()


def main(precision="fp32"):
    B, P, steps, warm = 65536, 8, 20, 60
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (precision == "tf32")
    net = QNetLike(9, width=256).cuda().eval()
    if precision == "bf16":
        net = Autocast(net)
    policy = D.BatchedGreedyPolicy(net, chunk_actions=1 << 16)
    perm, lord = D.random_deals(B, seed=11, pool_games=P)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    env = D.BatchedEnvCooperation(B, seed=5, max_actions_per_env=160)
    env.prepare(pd, ld, pool_games=P)
    for _ in range(warm):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)

    def step():
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
        t0.record()
        is_lord = env.get_role_ID() == 2
        q = policy.q_values(env, is_lord)
        greedy = policy.select(env, q)
        off = env.offsets
        cnt = (off[1:] - off[:-1])
        ent = torch.randint(0, 1 << 30, (B,), device="cuda", dtype=torch.int32, generator=gen)
        choice = torch.where(is_lord, greedy, ent % cnt.clamp(min=1)).to(torch.int32)
        t1.record()
        env.rollout_step(choice, mode=D.native.CHOICE_INDEX, perm=pd, lord_pile=ld, pool_games=P)
        t2.record()
        return t0, t1, t2

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    evs = [step() for _ in range(steps)]
    torch.cuda.synchronize()
    net_ms = sum(a.elapsed_time(b) for a, b, _ in evs) / steps
    env_ms = sum(b.elapsed_time(c) for _, b, c in evs) / steps
    st = env.stats.cpu().numpy()
    print(json.dumps({"workload": "config 3: %d envs, lord = argmax Q (NetCooperation-shaped random-init net, %s), farmers random" % (B, precision),
                      "env_steps_per_s": B / ((net_ms + env_ms) * 1e-3), "ms_per_step": net_ms + env_ms,
                      "network_and_selection_ms": net_ms, "env_kernel_ms": env_ms,
                      "actions_scored_per_step": int(env.num_actions), "lord_win_rate": float(st[1]) / max(1, st[0]),
                      "errors": int(st[7])}))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fp32")
