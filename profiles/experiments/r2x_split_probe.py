import sys, os, torch
sys.path.insert(0, os.getcwd())
import ddz_b200 as D
D.native.set_row_writer(sys.argv[1])
B, P = 131072, 8
perm, lord = D.random_deals(B, seed=1, pool_games=P)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
env = D.BatchedEnvCooperation(B, seed=1, max_actions_per_env=160)
env.prepare(pd, ld, pool_games=P)
for _ in range(150):
    env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(6):
    env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", int(env.stats[7]))
