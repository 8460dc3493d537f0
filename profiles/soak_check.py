"""Long soak: 131 072 envs x N steps of the fused kernel (two env groups, CUDA-graph replays between checks), with a
2048-env sample followed step by step on the CPU oracle (legal lists, faces, results, state).  Not part of the test
suite (takes minutes); run on the GPU box: python profiles/soak_check.py [steps]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D
from oracle import ddz_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B, P, seed = 131072, 8, 424242
perm, lord = D.random_deals(B, seed=9, pool_games=P)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
env = D.BatchedEnvCooperation(B, seed=seed, max_actions_per_env=160)
env.prepare(pd, ld, pool_games=P)
sample = np.sort(np.random.default_rng(3).choice(B, 2048, replace=False))
sp = perm.reshape(P, B, 54)[:, sample].reshape(-1, 54)
sl = lord.reshape(P, B)[:, sample].reshape(-1)
ref = O.RefBatch(len(sample), 2)
ref.deal(sp, sl, pool_games=P)
st = torch.as_tensor(sample).cuda()
t0 = time.time()
checked = 0
for t in range(steps):
    o_off, o_au, _, o_face = ref.observe(want_f32=False)
    off = env.offsets.to(torch.int64)
    cnt = (off[1:] - off[:-1])[st].cpu().numpy()
    assert np.array_equal(cnt, np.diff(o_off)), ("counts", t)
    starts = off[:-1][st]
    idx = torch.repeat_interleave(starts, torch.as_tensor(cnt).cuda()) + \
        (torch.arange(int(cnt.sum()), device="cuda") - torch.repeat_interleave(torch.as_tensor(np.concatenate([[0], np.cumsum(cnt)[:-1]])).cuda(), torch.as_tensor(cnt).cuda()))
    assert np.array_equal(env.actions_packed[idx].cpu().numpy().view(np.uint64), o_au), ("lists", t)
    assert np.array_equal(env.face[st].cpu().numpy(), o_face), ("face", t)
    checked += len(o_au)
    r, done, cat = env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
    ent = np.array([O.philox(seed, int(b), t) for b in sample], dtype=np.uint32)
    rr, rd, rc, _ = ref.step(ent.view(np.int32), mode=1)
    assert np.array_equal(r[st].cpu().numpy(), rr) and np.array_equal(done[st].cpu().numpy(), rd) and np.array_equal(cat[st].cpu().numpy(), rc), ("results", t)
    ref.deal(sp, sl, only_done=True, pool_games=P)
stats = env.stats.cpu().numpy()
print(json.dumps({"steps": steps, "envs": B, "sample": len(sample), "moves_checked": checked, "games_finished": int(stats[0]),
                  "errors": int(stats[7]), "sample_games": int(ref.stats[0]), "seconds": time.time() - t0}))
