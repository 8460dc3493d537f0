"""Developer probe: host time per e2e step (HostRollout on the native pipe, 2 groups) vs device time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D

B, NG, P = 131072, 2, 8
perm, lord = D.random_deals(B, seed=1, pool_games=P)
ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=NG, seed=3, max_actions_per_env=160)
ge.prepare(perm, lord, pool_games=P)
for _ in range(150):
    ge.rollout_step()
ge.join(); torch.cuda.synchronize()
hosts = []
for g in range(NG):
    with torch.cuda.stream(ge.streams[g]):
        hosts.append(D.HostRollout(ge.envs[g], ge._pool[g][0], ge._pool[g][1], P))
Bg = B // NG
ent = [[torch.as_tensor(np.random.default_rng(i).integers(0, 1 << 31, Bg).astype(np.int32)).pin_memory() for i in range(4)] for _ in range(NG)]
N = 300
for mode in ("submit_only", "with_wait_and_read"):
    pending = [[None, None] for _ in range(NG)]
    torch.cuda.synchronize()
    t0 = time.perf_counter(); host = 0.0
    for i in range(N):
        h0 = time.perf_counter()
        for g in range(NG):
            old = pending[g][i & 1]
            if mode != "submit_only" and old is not None:
                D.HostRollout.wait(old)
                x = int(old.done_np[0]) + int(old.r_np[-1])
            pending[g][i & 1] = hosts[g].step(ent[g][i % 4])
        host += time.perf_counter() - h0
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(mode, "host issue %.1f us/step, wall %.1f us/step" % (t_issue / N * 1e6, t_all / N * 1e6))

# the same two groups without any copies: direct launches of the fused step from Python (no graph), device entropy
dev_ent = [torch.as_tensor(np.random.default_rng(9 + g).integers(0, 1 << 31, Bg).astype(np.int32)).cuda() for g in range(NG)]
for label in ("direct_launch_mod_entropy", "direct_launch_philox"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        for g, env in enumerate(ge.envs):
            with torch.cuda.stream(ge.streams[g]):
                pg, lg, _ = ge._pool[g]
                if label.endswith("philox"):
                    env.rollout_step(perm=pg, lord_pile=lg, pool_games=P)
                else:
                    env.rollout_step(dev_ent[g], mode=D.native.CHOICE_MOD, perm=pg, lord_pile=lg, pool_games=P)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(label, "host issue %.1f us/step, wall %.1f us/step" % (t_issue / N * 1e6, t_all / N * 1e6))
