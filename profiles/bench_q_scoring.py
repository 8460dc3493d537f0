"""Q-scoring of BASELINE config 3's decisions (65 536 envs, the landlord's legal moves, NetCooperation's layer sizes):
where the time goes.  ddz_q_features alone (float32 and bfloat16 rows; achieved GB/s of the rows it writes), the fc1 / fc2
GEMMs behind it at three precisions, and the torch module on the materialised [n, C+1, 15, 4] input for comparison."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddz_b200 as D
from qnet_like import QNetLike


def timed(fn, reps=5):
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
    B, P = 65536, 8
    perm, lord = D.random_deals(B, seed=20260101, pool_games=P)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    D.native.set_tile_order("ticket")
    env = D.BatchedEnvCooperation(B, seed=1, max_actions_per_env=160)
    env.prepare(pd, ld, pool_games=P)
    for _ in range(60):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=P)
    mask = env.get_role_ID() == 2
    net = QNetLike(9, 256, 256, seed=0).eval().cuda()
    out = {"envs": B, "landlord_decisions": int(mask.sum()), "runs": {}}
    for precision in ("fp32", "tf32", "bf16"):
        sc = D.FusedQScorer(net, 9, precision=precision, chunk_rows=1 << 20)
        q = sc.q_values(env, mask)
        rows = int((q != 0).sum())
        total = timed(lambda: sc.q_values(env, mask))
        # the kernel alone: one chunk covers everything here
        off = env._offsets[env._cur]
        cnt = (off[1:] - off[:-1]) * mask
        dst = torch.zeros(B + 1, dtype=torch.int32, device="cuda"); dst[1:] = torch.cumsum(cnt, 0)
        n = int(dst[-1])
        feat = sc._feat[:n]
        m8 = mask.to(torch.uint8)

        def kernel():
            D.native.check(D.native.lib.ddz_q_features(env._state.data_ptr(), env.VARIANT, off.data_ptr(),
                                                       env._actions_u64[env._cur].data_ptr(), m8.data_ptr(), dst.data_ptr(), 0, B, 0,
                                                       sc.T.data_ptr(), sc.rank_bias.data_ptr(), sc.L.data_ptr(),
                                                       sc.line_bias.data_ptr(), sc.W, feat.data_ptr(), int(precision == "bf16"), B,
                                                       torch.cuda.current_stream().cuda_stream), "ddz_q_features")
        torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
        k_ms = timed(kernel)
        g_ms = timed(lambda: torch.mv(torch.addmm(sc.b1, feat, sc.w1t).relu_().float(), sc.w2))
        row_bytes = 19 * 256 * (2 if precision == "bf16" else 4)
        out["runs"][precision] = {"rows": n, "q_values_ms": total, "ddz_q_features_ms": k_ms,
                                  "rows_written_GBs": n * row_bytes / (k_ms * 1e-3) / 1e9, "fc1_fc2_gemm_ms": g_ms,
                                  "gemm_TFLOPs": 2.0 * n * 19 * 256 * 256 / (g_ms * 1e-3) / 1e12}
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    pol = D.BatchedGreedyPolicy(net, chunk_actions=1 << 16)
    out["runs"]["torch module, fp32"] = {"q_values_ms": timed(lambda: pol.q_values(env, mask), reps=2),
                                         "encode_state_actions_ms": timed(lambda: env.state_actions(mask))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
