"""Developer tool: per-warp phase timeline of k_env (needs the -DDDZ_TRACE build of the library).

    nvcc ... -DDDZ_TRACE -o gpurun_out/libddz_trace.so doudizhu-rl_b200/csrc/ddz_kernels.cu
    DDZ_LIB=gpurun_out/libddz_trace.so python profiles/trace_phases.py
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D

B, G = (int(sys.argv[1]) if len(sys.argv) > 1 else 131072), 8
perm, lord = D.random_deals(B, seed=1, pool_games=G)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
env = D.BatchedEnvCooperation(B, seed=3, max_actions_per_env=160)
env.prepare(pd, ld, pool_games=G)
for _ in range(150):
    env.rollout_step(perm=pd, lord_pile=ld, pool_games=G)
nt = (B + 31) // 32
buf = torch.zeros(nt * 8, dtype=torch.int64, device="cuda")
D.native.lib.ddz_debug_set_trace.argtypes = [ctypes.c_void_p]
assert D.native.lib.ddz_debug_set_trace(buf.data_ptr()) == 0
torch.cuda.synchronize()
env.rollout_step(perm=pd, lord_pile=ld, pool_games=G)
torch.cuda.synchronize()
tr = buf.cpu().numpy().reshape(nt, 8).astype(np.float64)
t0 = tr[:, 0].min()
tr = (tr - t0) / 1e3    # microseconds
tr[tr < 0] = np.nan
names = ["start", "phaseA_end", "face_begin", "face_end", "act_begin", "lookback_end", "act_end", "enum0_end"]
out = {}
for i, nme in enumerate(names):
    col = tr[:, i]
    out[nme] = {"min": float(np.nanmin(col)), "p50": float(np.nanmedian(col)), "p90": float(np.nanpercentile(col, 90)),
                "max": float(np.nanmax(col))}
dur = {"phaseA": tr[:, 1] - tr[:, 0], "face_rows": tr[:, 3] - tr[:, 2], "enum_window0": tr[:, 7] - tr[:, 4],
       "lookback_wait": tr[:, 5] - tr[:, 7],
       "enum_and_action_rows": tr[:, 6] - tr[:, 5], "warp_lifetime": np.maximum(tr[:, 3], tr[:, 6]) - tr[:, 0]}
for k, v in dur.items():
    out["dur_" + k] = {"mean": float(np.nanmean(v)), "p50": float(np.nanmedian(v)), "p99": float(np.nanpercentile(v, 99)),
                       "max": float(np.nanmax(v))}
end = np.fmax(tr[:, 3], tr[:, 6])
out["warp_end_percentiles_us"] = {str(q): float(np.nanpercentile(end, q)) for q in (1, 10, 25, 50, 75, 90, 99, 100)}
out["kernel_span_us"] = float(np.nanmax(tr))
even = np.arange(nt) % 2 == 0
odd = np.arange(nt) % 2 == 1
for nme, sel in (("even", ~odd), ("odd", odd)):
    out[nme + "_tiles"] = {k: float(np.nanmean(v[sel])) for k, v in dur.items()}
out["even_tiles_end_p50"] = float(np.nanmedian(np.maximum(tr[even, 3], tr[even, 6])))
out["odd_tiles_end_p50"] = float(np.nanmedian(np.maximum(tr[~even, 3], tr[~even, 6])))
# what makes phase A slow, and who waits for whom: deals per tile, legal moves per tile
done_after = env.done.cpu().numpy().astype(np.int64)            # envs that finished with this step (and were re-dealt)
pad = (-B) % 32
deals = np.pad(done_after, (0, pad)).reshape(nt, 32).sum(1)
off = env.offsets.cpu().numpy().astype(np.int64)
moves = np.add.reduceat(np.diff(off), np.arange(0, B, 32))
out["phaseA_by_deals_in_tile"] = {str(k): {"tiles": int((deals == k).sum()), "mean_us": float(np.nanmean(dur["phaseA"][deals == k]))}
                                  for k in range(0, 6) if (deals == k).any()}
a_end = tr[:, 1]
running_max = np.fmax.accumulate(np.nan_to_num(a_end, nan=0.0))
out["phaseA_end_running_max_percentiles_us"] = {str(q): float(np.percentile(running_max, q)) for q in (10, 50, 90, 100)}
out["lookback_end_minus_running_max_p50_us"] = float(np.nanmedian(tr[:, 5] - running_max))
out["corr_moves_vs_lifetime"] = float(np.corrcoef(moves, np.nan_to_num(dur["warp_lifetime"]))[0, 1])
print(json.dumps(out, indent=1))
