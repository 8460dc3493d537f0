#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched Doudizhu env (legal moves + state/action encode + step + re-deal).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused env-step of every env of the rank's slice (ddz_rollout_step: apply the chosen move,
re-deal finished envs, generate the legal moves of the new state, write face [B,9,15,4] and the action one-hots).
Workload: BASELINE.json config 4's per-GPU slice -- 131 072 envs per GPU (weak scaling), EnvCooperation face
(C=9, reference train.py:4-5), all three seats play uniformly random legal moves from the Philox index stream
(config 2's lord-vs-random policy), synthetic random deals.  Inputs per step are ~0.5 GB of output traffic, i.e.
larger than the 126 MB L2, so no flush between iterations is needed.

--impl reference times the CPU restatement of the reference env (oracle/, kind "port": the reference's own native
modules are not shipped) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (legal moves + state encode)"
UNIT = "env-steps/s"
VARIANT, CHANNELS = 2, 9            # EnvCooperation
STATE_BYTES = 76                    # 9 x uint64 + uint32 per env (include/ddz_b200.h)
POOL_GAMES = 8
STATS_EVERY = 64                    # env-steps between two statistics all-reduces (multi-GPU runs)
SEED = 20260101


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=131072, help="envs per GPU")
    ap.add_argument("--prefill", type=int, default=150, help="untimed env-steps that bring the games to their steady-state mix")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--groups", type=int, default=4, help="stream-parallel env groups per GPU (1 = one chain of launches)")
    ap.add_argument("--no-single", action="store_true", help="skip the single-group information run")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(envs, groups, compressed):
    """DRAM bytes per step (all launches of the step) from the committed ncu --set full captures of the same workload,
    or None when there is no capture for this shape."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if envs != 131072:
            return None
        key = {(1, False): "single_group", (2, True): "two_groups", (2, False): "two_groups_plain_memory",
               (4, True): "four_groups"}[(groups, compressed)]
        return d[key]["traffic_bytes_per_step"]
    except Exception:
        return None


def bytes_per_env_step(nbar, C=CHANNELS):
    """ALGORITHMIC bytes of one env-step = one env's share of one k_env launch: state read + write,
    previous offsets (2 x 4) and chosen move (8) read, r/done/cat/reward (15) written, new offset (4), packed list
    (8 N) and one-hot rows (240 N) written, face (240 C) written.  Deck permutations of re-dealt envs (55 B x ~1.6 %)
    and the look-back words are not counted."""
    return 2 * STATE_BYTES + 8 + 8 + 15 + 4 + 240 * C + 248 * nbar


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_rollout(envs, warm, steps, threads):
    import ddz_b200  # noqa: F401  (only for random_deals; pure numpy)
    from oracle import ddz_oracle as O
    perm, lord = ddz_b200.random_deals(envs, seed=SEED, pool_games=POOL_GAMES)
    n, sec, stats, cs = O.rollout(envs, warm, steps, VARIANT, SEED, perm, lord, POOL_GAMES, threads)
    return n, sec, stats


def cpu_baseline(target_seconds):
    """oracle port timed on this box's host cores over a bounded sample of the same workload."""
    threads = os.cpu_count() or 1
    envs = 4096 * threads
    n, sec, _ = cpu_rollout(envs, 100, 20, threads)                 # calibration
    rate = n / max(sec, 1e-9)
    steps = max(20, min(int(target_seconds * rate / envs), 20000))
    n, sec, stats = cpu_rollout(envs, 100, steps, threads)
    return {"value": n / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d env-steps after 100 warm-up steps (%.1f s), C oracle (oracle/ddz_oracle.c), "
                      "mean legal moves %.2f" % (envs, steps, sec, stats[8] / max(n, 1))}


def run_reference(args):
    """Reference arm: the reference's CPU path = oracle port on all host threads (its natives are absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    envs = 4096 * threads
    warm_prefill = 100
    # W warm-up steps then exactly K timed steps, each one env-step of the bounded sample of `envs` envs
    n, sec, stats = cpu_rollout(envs, warm_prefill + args.warmup, args.steps, threads)
    value = n / sec
    nbar = stats[8] / max(n, 1)
    sample = "%d envs (of the %d-per-GPU workload) x %d env-steps, C oracle port, %d threads" % (envs, args.envs, args.steps, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "BASELINE config 4 per-GPU slice: %d envs/GPU, lord-vs-random rollout (all seats uniform "
                                   "random legal move, Philox stream), EnvCooperation face C=9, fused step+redeal+legal+encode"
                                   % args.envs,
                       "envs_per_gpu": args.envs, "face_channels": CHANNELS, "sample_envs": envs, "mean_legal_moves": nbar,
                       "prefill_steps": warm_prefill, "pool_games": POOL_GAMES,
                       "note": "CPU arm: the reference's natives are absent, so this is the C oracle port of the same "
                               "env-step (legal moves + face + action one-hots + step + re-deal) on all host threads, "
                               "over a bounded sample of the workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ddz_b200 as D

    rank, local, world = D.sharding.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the env has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K, W, P, NG = args.envs, args.steps, args.warmup, POOL_GAMES, args.groups
    env0 = rank * B                                   # global env ids keep results independent of the GPU count
    perm, lord = D.random_deals(B, seed=SEED + 1000 * rank, pool_games=P)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)       # samples from here to the end of the e2e region: the same kernel runs throughout
    if rank == 0:
        sampler.start()

    # ---------------- the rank's envs as NG stream-parallel groups (one group's load-only prologue and tail
    # overlap the other group's store phase); device-resident deal pool, Philox moves on the device
    ge = D.GroupedEnv(D.BatchedEnvCooperation, B, groups=NG, seed=SEED, device=dev, env0=env0, max_actions_per_env=160)
    compressed = hasattr(ge.envs[0]._face, "_ddz_rows")          # row buffers in compressible memory (the default where supported)
    ge.prepare(perm, lord, pool_games=P)
    for _ in range(args.prefill):
        ge.rollout_step()
    ge.capture(steps_per_graph=2)       # the loop is launch-bound from Python: replay the ping-pong pair as a CUDA graph
    for _ in range(max(W, 3)):
        ge.replay()
    ge.join()
    torch.cuda.synchronize(dev)
    if int(ge.stats[7].item()):
        raise SystemExit("env reported errors during warm-up")

    # ---------------- timed region: exactly K env-steps of every env, CUDA events on the launching (current) stream
    stats0 = ge.stats.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = torch.cuda.Stream(dev) if world > 1 else None
    if side is not None:                # warm the collective up on its stream
        with torch.cuda.stream(side):
            dist.all_reduce(ge.stats.clone())
        side.synchronize()
    barrier()
    e0.record()
    ge._fork()
    reduced = []
    for i in range(K // 2):
        ge.replay()                     # per group: two launches of k_env<cooperation, step+observe>
        if side is not None and (2 * i) % STATS_EVERY == 0:
            # the path's only collective: win/return statistics, all-reduced on a side stream while the envs keep
            # stepping (reference game.py:209-226 logs these counters every log_every episodes)
            side.wait_stream(ge.streams[0])
            with torch.cuda.stream(side):
                snap = torch.stack([e.stats for e in ge.envs]).sum(0)
                dist.all_reduce(snap)
                reduced.append(snap)
    if K % 2:
        ge.rollout_step()
    ge.join()
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    dstats_t = (ge.stats - stats0).clone()
    dstats = dstats_t.cpu().numpy()
    if int(ge.stats[7].item()):
        raise SystemExit("env reported errors during the timed region")
    local_steps = int(dstats[4])
    assert local_steps == B * K, (local_steps, B * K)   # every env applied one move per step (finished ones re-dealt)
    nbar = float(dstats[8]) / local_steps

    # ---------------- for information: one group (a single chain of launches), graph replay and eager Python loop
    info = {}
    if NG > 1 and not args.no_single:
        perm_d, lord_d = torch.as_tensor(perm).to(dev), torch.as_tensor(lord).to(dev)
        one = D.BatchedEnvCooperation(B, seed=SEED, device=dev, env0=env0, max_actions_per_env=160)
        one.prepare(perm_d, lord_d, pool_games=P)
        kw = dict(perm=perm_d, lord_pile=lord_d, pool_games=P)
        for _ in range(args.prefill):
            one.rollout_step(**kw)
        gr = D.GraphedRollout(one, perm_d, lord_d, P)
        for _ in range(3):
            gr.replay()
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize(dev)
        g0.record()
        for _ in range(K // 2):
            gr.replay()
        g1.record()
        for _ in range(K):
            one.rollout_step(**kw)
        g2.record()
        torch.cuda.synchronize(dev)
        info = {"single_group_graph_ms_per_step": g0.elapsed_time(g1) / (2 * (K // 2)),
                "single_group_eager_python_ms_per_step": g1.elapsed_time(g2) / K}
        del one, gr, perm_d, lord_d
        torch.cuda.empty_cache()

    # ---------------- e2e: the same env-step through the host-facing API (HostRollout, one per group) with HOST buffers
    # every step: H2D of the step's entropy (int32 [B], pinned) and of the deal-pool refill (one slot of host-made
    # permutations every REFILL steps = about one game length, the rate at which deals are consumed), D2H of the step's
    # results (r, done, cat, reward) into pinned memory.  All copies are inside the timed region.
    R, REFILL, Bg = 4, 64, B // NG
    rng = np.random.default_rng(SEED + rank)
    ent_h = [[torch.as_tensor(rng.integers(0, 1 << 31, Bg, dtype=np.int64).astype(np.int32)).pin_memory() for _ in range(R)]
             for _ in range(NG)]
    pool_h = []
    for i in range(2):
        pp, ll = D.random_deals(Bg, seed=SEED + 77 + i + 1000 * rank)
        pool_h.append((torch.as_tensor(pp).pin_memory(), torch.as_tensor(ll).pin_memory()))
    hosts = []
    for g, env in enumerate(ge.envs):
        with torch.cuda.stream(ge.streams[g]):
            pg, lg, _ = ge._pool[g]
            hosts.append(D.HostRollout(env, pg, lg, P))
    sink = [0]
    DEPTH = D.native.PIPE_DEPTH                                   # the host reads step i's results while step i+3 is issued
    pending = [[None] * DEPTH for _ in range(NG)]

    def e2e_step(i):
        for g in range(NG):                                       # each HostRollout issues on its group's stream
            if i % REFILL == 0:
                pp, ll = pool_h[(i // REFILL) % 2]
                hosts[g].refill((i // REFILL) % P, pp, ll)
            old = pending[g][i % DEPTH]
            if old is not None:                                   # the host reads the results of step i-DEPTH
                D.HostRollout.wait(old)
                sink[0] += int(old.done_np[0]) + int(old.r_np[-1])
            pending[g][i % DEPTH] = hosts[g].step(ent_h[g][i % R])

    for i in range(max(W, 4)):
        e2e_step(i)
    ge.join()
    barrier()
    stats_e0 = ge.stats.clone()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    ge._fork()
    for i in range(K):
        e2e_step(i)
    for g in range(NG):
        for res in pending[g]:
            D.HostRollout.wait(res)
    ge.join()
    x1.record()
    barrier()
    clocks = sampler.stop()
    e2e_ms = x0.elapsed_time(x1)
    e2e_K = K
    e2e_steps = int((ge.stats - stats_e0)[4].item())
    assert e2e_steps == B * e2e_K and sink[0] > -(1 << 40)
    h2d = 4 * B + (55 * B * ((e2e_K + REFILL - 1) // REFILL)) // e2e_K
    d2h = NG * hosts[0].d2h_bytes
    if int(ge.stats[7].item()):
        raise SystemExit("env reported errors during the e2e region")

    # ---------------- reduce over ranks: MAX of the device times, SUM of the work
    total_ms = D.sharding.max_over_ranks(total_ms, dev)
    e2e_ms = D.sharding.max_over_ranks(e2e_ms, dev)
    gstats = D.sharding.allreduce_stats(dstats_t, side_stream=torch.cuda.Stream(dev) if world > 1 else None)
    torch.cuda.synchronize(dev)
    all_steps = int(gstats[4].item())

    if rank == 0:
        peak, peak_src = peaks()
        eb = bytes_per_env_step(nbar)
        value = all_steps / (total_ms * 1e-3)
        step_ms = total_ms / K
        kern_gbs = B * eb / (step_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": dict({"workload": "BASELINE config 4 per-GPU slice: %d envs/GPU, lord-vs-random rollout (all seats uniform "
                                        "random legal move, Philox stream), EnvCooperation face C=9, fused step+redeal+legal+encode" % B,
                            "envs_per_gpu": B, "env_groups_per_gpu": NG, "face_channels": CHANNELS, "mean_legal_moves": nbar,
                            "prefill_steps": args.prefill, "pool_games": P, "parallelism": "env-shard x%d" % world,
                            "launch": "CUDA graph replay of the 2-launch ping-pong pair, one chain per env group",
                            "collectives": ("none (single GPU)" if world == 1 else
                                            "stats all-reduce (int64[16], NCCL) every %d steps on a side stream, inside the timed region"
                                            % STATS_EVERY),
                            "l2_policy": "per-step output (%.0f MB) exceeds the 126 MB L2; no flush" % (B * eb / 1e6),
                            "row_buffers": ("compressible device memory (CU_MEM_ALLOCATION_COMP_GENERIC: the L2 compresses "
                                            "the 0/1 thermometer rows on their way to HBM)" if compressed else "plain device memory"),
                            "games_finished": int(gstats[0].item()),
                            "lord_win_rate": float(gstats[1].item()) / max(1, int(gstats[0].item()))}, **info),
            "roofline": {"bound": "hbm", "kernel": "k_env<2,step+observe>", "achieved": kern_gbs, "peak": peak,
                         "unit": "GB/s", "frac": kern_gbs / peak, "traffic": ncu_traffic(B, NG, compressed),
                         "peak_source": peak_src, "algorithmic_bytes_per_env": eb,
                         "algorithmic_bytes_per_step": B * eb, "ms_per_step": step_ms, "launches_in_flight": NG,
                         "note": "achieved = algorithmic bytes of one step of all envs / wall time per step; the step is "
                                 "%d concurrent launches of the same kernel (one per env group); traffic = ncu DRAM bytes "
                                 "of those launches (profiles/r1_traffic.json)%s" % (
                                     NG, "; it is far below the algorithmic bytes because the row buffers are compressible memory"
                                     if compressed else ""),
                         "single_launch": ({"ms_per_launch": info["single_group_graph_ms_per_step"],
                                            "achieved": B * eb / (info["single_group_graph_ms_per_step"] * 1e-3) / 1e9,
                                            "frac": B * eb / (info["single_group_graph_ms_per_step"] * 1e-3) / 1e9 / peak,
                                            "note": "one launch over all %d envs, one chain (no env groups)" % B}
                                           if "single_group_graph_ms_per_step" in info else None)},
            "e2e": {"value": B * e2e_K * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_K, "steps": e2e_K,
                    "api": "HostRollout.step(entropy_host) per env group (native ddz_pipe_step: H2D entropy -> k_env -> D2H "
                           "r/done/cat -- the reference step's return tuple -- on copy streams), results of step t-4 read by the host every step (ring of 4 pinned result buffers), "
                           "refill(slot, perm_host, lord_host) uploads host-made deals every %d steps" % REFILL},
            "gpu_launches": K * NG,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ensure_built():
    """a fresh checkout has no built artefacts (*.so are git-ignored): build them in-tree before importing the package"""
    if not os.path.exists(os.path.join(ROOT, "doudizhu-rl_b200", "libddz_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


def main():
    args = parse_args()
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ensure_built()
    else:                                   # the other ranks wait for rank 0's build instead of racing it
        lib = os.path.join(ROOT, "doudizhu-rl_b200", "libddz_b200.so")
        for _ in range(600):
            if os.path.exists(lib) and time.time() - os.path.getmtime(lib) > 2.0:
                break
            time.sleep(0.5)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
