"""Example / smoke of the trainer-driver pieces: a batched DQN landlord learns against random farmers.

Mirrors train.py:18-31 + Game.train (game.py:183-238) in batched form: EnvCooperation features, the lord acts
epsilon-greedily on Q(face, action), farmers play random legal moves, transitions are assembled with the reference's
delayed-feedback rule (TransitionCollector), stored in a GPU replay buffer and learned with the reference's TD step.
The network is a small random-init net with the reference's input contract (tests/qnet_like.py)."""
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


def main(B=2048, iters=600, updates_per_iter=2, batch=256, gamma=0.95):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    policy, target = QNetLike(9, width=64).to(dev), QNetLike(9, width=64).to(dev)
    target.load_state_dict(policy.state_dict())
    opt = torch.optim.Adam(policy.parameters(), 1e-4)
    env = D.BatchedEnvCooperation(B, seed=1)
    P = 8
    perm, lord = D.random_deals(B, seed=2, pool_games=P)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    env.prepare(pd, ld, pool_games=P)
    col = D.TransitionCollector(env, role=1, reward=100.0)
    rb = D.ReplayBuffer(200_000, 9, dev)
    explore = D.BatchedGreedyPolicy(policy, epsilon=0.5, seed=3)
    greedy = D.BatchedGreedyPolicy(policy, epsilon=0.0)
    gen = torch.Generator(device="cuda"); gen.manual_seed(4)
    log = []
    for it in range(iters):
        explore.epsilon = 0.01 + 0.49 * float(np.exp(-it / (iters / 5)))           # dqn.py:73-76 schedule, compressed
        acts, offs = env.valid_actions()
        off = offs.to(torch.int64)
        cnt = off[1:] - off[:-1]
        q = explore.q_values(env)
        k_lord, k_greedy = explore.select(env, q).to(torch.int64), greedy.select(env, q).to(torch.int64)
        rnd = torch.randint(0, 1 << 30, (B,), device=dev, generator=gen) % cnt.clamp(min=1)
        is_lord = env.get_role_ID() == 2
        choice = torch.where(is_lord, k_lord, rnd).clamp(min=0)
        live = (cnt > 0)[:, None, None]
        last = max(int(off[-1]) - 1, 0)
        a0 = acts[(off[:-1] + choice).clamp(max=last)] * live
        a1 = acts[(off[:-1] + k_greedy.clamp(min=0)).clamp(max=last)] * live
        rb.append(*col.on_turn(a0, a1)[:6])
        was_done = env.is_done.clone()
        env.step(choice.to(torch.int32))
        env.observe()
        rb.append(*col.on_step_done(env.is_done & ~was_done)[:6])
        env.prepare(pd, ld, only_done=True, pool_games=P)
        if len(rb) >= batch:
            for _ in range(updates_per_iter):
                td_loss = D.td_step(policy, target, opt, rb.sample(batch, gen), gamma)
        if it % 20 == 19:
            target.load_state_dict(policy.state_dict())                              # config.py:14
        if it % 100 == 99:
            st = env.stats.cpu().numpy()
            log.append({"iter": it + 1, "games": int(st[0]), "lord_win_rate_so_far": float(st[1]) / max(1, int(st[0])),
                        "replay": len(rb), "td_loss": float(td_loss)})
            print(json.dumps(log[-1]), flush=True)
    return log


if __name__ == "__main__":
    main()
