"""How long do the pieces of an env-step take as separate kernels (131 072 envs, steady state)?
(a) rule-only fused step: transition + re-deal + packed lists, no face / one-hot rows
(b) ddz_encode_face   (c) ddz_encode_actions over the step's moves   vs (d) the fused kernel."""
import ctypes, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D
N = D.native
B, P = 131072, 8
perm, lord = D.random_deals(B, seed=11, pool_games=P)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
env = D.BatchedEnvCooperation(B, seed=5, max_actions_per_env=160)
env.prepare(pd, ld, pool_games=P)
for _ in range(150):
    env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
torch.cuda.synchronize()
p = lambda t: None if t is None else t.data_ptr()
st = torch.cuda.current_stream().cuda_stream

def rule_only():
    nxt = 1 - env._cur
    N.check(N.lib.ddz_rollout_step(p(env._state), p(env._ws), env.VARIANT, p(env._offsets[env._cur]), p(env._actions_u64[env._cur]),
            None, N.CHOICE_PHILOX, env.seed, env.env0, env._stepno, env._rewards.data_ptr(), p(pd), p(ld), P,
            p(env.r), p(env.done), p(env.cat), None, p(env._offsets[nxt]), p(env._actions_u64[nxt]), None, env.cap,
            None, p(env.stats), B, st), "rule")
    env._cur = nxt; env._stepno += 1

def face_only():
    N.check(N.lib.ddz_encode_face(p(env._state), env.VARIANT, p(env._face), B, st), "face")

def timeit(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

out = {}
out["fused_ms"] = timeit(lambda: env.rollout_step(perm=pd, lord_pile=ld, pool_games=P))
out["rule_only_ms"] = timeit(rule_only)
out["face_only_ms"] = timeit(face_only)
n = int(env._offsets[env._cur][B].item())
packed = env._actions_u64[env._cur]
def acts_only():
    N.check(N.lib.ddz_encode_actions(p(packed), n, p(env._actions_f32), st), "acts")
out["actions_only_ms"] = timeit(acts_only)
out["moves"] = n
out["face_GBs"] = B * 9 * 240 / out["face_only_ms"] / 1e6
out["actions_GBs"] = n * 240 / out["actions_only_ms"] / 1e6
def split():
    rule_only(); face_only(); acts_only()
out["rule_face_actions_back_to_back_ms"] = timeit(split)
print(json.dumps(out))
