"""train.py of the reference (train.py:1-31: a DQN landlord, `Game(EnvCooperation, ...).train(...)`) on the batched
driver: `BatchedGame(BatchedEnvCooperation, ...)` with 2048 envs, the lord learning with the reference's delayed-feedback
transitions / replay / TD step, farmers playing random legal moves.  Prints the win-rate curve."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddz_b200 as D
from qnet_like import QNetLike


def main(num_envs=2048, episodes=24000, width=64, hidden=64, fused=False, precision="fp32"):
    """width = hidden = 256 are the layer sizes of the reference's NetCooperation (net.py:125-139); fused=True acts through
    agent.FusedQScorer (ddz_q_features + two GEMMs) instead of the module's forward -- the TD updates use the module either way"""
    torch.manual_seed(0)
    net = lambda: QNetLike(9, width=width, hidden=hidden)
    dqn = lambda net_cls, channels, device: D.BatchedDQN(net_cls, channels, device, replay_size=200_000, decay=120,
                                                         update_target_every=20, fused=fused, precision=precision)
    game = D.BatchedGame(D.BatchedEnvCooperation, {"lord": net, "down": None, "up": None},
                         {"lord": dqn, "down": None, "up": None}, reward_dict={"lord": 100, "down": None, "up": None},
                         train_dict={"lord": True, "up": False, "down": False}, seed=1, num_envs=num_envs,
                         updates_per_step=2)
    hist = game.train(episodes, log_every=episodes // 8)
    out = [{"episodes": h["episodes"], "iterations": h["iterations"], "seconds": round(h["seconds"], 2),
            "lord_recent_win": round(h["lord"]["recent_win"], 4), "lord_mean_loss": round(h["lord"]["mean_loss"], 2)}
           for h in hist]
    for o in out:
        print(json.dumps(o))
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--episodes", type=int, default=24000)
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--precision", default="fp32")
    a = ap.parse_args()
    import time
    t0 = time.time()
    main(a.envs, a.episodes, a.width, a.hidden, a.fused, a.precision)
    torch.cuda.synchronize()
    print(json.dumps({"envs": a.envs, "episodes": a.episodes, "width": a.width, "hidden": a.hidden, "fused": a.fused,
                      "precision": a.precision, "wall_seconds": round(time.time() - t0, 2)}))
