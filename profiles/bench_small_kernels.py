"""The other kernels behind the C-ABI, each timed alone at BASELINE sizes with its algorithmic bytes against the HBM copy
peak: ddz_encode_state_actions (config 3's network input, built in place), ddz_encode_actions, ddz_encode_face,
ddz_select_actions, ddz_legal_count, ddz_observe (lists + rows without a step), ddz_reset."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6534.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    B, P = 65536, 4
    perm, lord = D.random_deals(B, seed=3, pool_games=P)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    env = D.BatchedEnvCooperation(B, seed=1, max_actions_per_env=160)
    env.prepare(pd, ld, pool_games=P)
    for _ in range(60):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
    n = env.num_actions
    out = {"envs": B, "legal_moves": n, "peak_GBs": peak, "kernels": {}}

    def rec(name, ms, nbytes, note=""):
        out["kernels"][name] = {"ms": ms, "algorithmic_MB": nbytes / 1e6, "GBs": nbytes / ms / 1e6, "frac_of_copy_peak": nbytes / ms / 1e6 / peak, "note": note}

    N = D.native
    st = torch.cuda.current_stream().cuda_stream
    x = torch.empty((n, env.C + 1, 15, 4), device="cuda")
    off = env.offsets
    au = env._actions_u64[env._cur]
    rec("ddz_encode_state_actions (all moves)", timed(lambda: N.lib.ddz_encode_state_actions(env._state.data_ptr(), 2, off.data_ptr(), au.data_ptr(), None, None, x.data_ptr(), B, st)),
        n * 240 * (env.C + 1) + 8 * n + 80 * B, "[n, C+1, 15, 4] fp32 written; state + lists read")
    acts = torch.empty((n, 15, 4), device="cuda")
    rec("ddz_encode_actions", timed(lambda: N.lib.ddz_encode_actions(au.data_ptr(), n, acts.data_ptr(), st)), n * 248)
    face = torch.empty((B, env.C, 15, 4), device="cuda")
    rec("ddz_encode_face", timed(lambda: N.lib.ddz_encode_face(env._state.data_ptr(), 2, face.data_ptr(), B, st)), B * (76 + 240 * env.C))
    q = torch.randn(n, device="cuda")
    choice = torch.empty(B, dtype=torch.int32, device="cuda")
    rec("ddz_select_actions", timed(lambda: N.lib.ddz_select_actions(q.data_ptr(), off.data_ptr(), 0.0, 1, 0, 0, choice.data_ptr(), B, st)), 4 * n + 8 * B)
    counts = torch.empty(B, dtype=torch.int32, device="cuda")
    rec("ddz_legal_count", timed(lambda: N.lib.ddz_legal_count(env._state.data_ptr(), counts.data_ptr(), B, st)), 80 * B)
    rec("ddz_observe (lists + rows, no step)", timed(lambda: env.observe()), B * (76 + 4 + 240 * env.C) + 248 * n)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
