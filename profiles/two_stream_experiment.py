"""Experiment: 131 072 envs on one GPU as TWO independent env groups of 65 536 on two streams.  Each group's step
kernel has a load-only prologue (no stores for ~8 us) and a tail; with two unsynchronised chains one group's prologue
and tail overlap the other group's store phase."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddz_b200 as D


class Chain:
    def __init__(self, B, seed, stream, steps_per_graph):
        self.stream = stream
        with torch.cuda.stream(stream):
            perm, lord = D.random_deals(B, seed=seed, pool_games=8)
            self.pd, self.ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
            self.env = D.BatchedEnvCooperation(B, seed=seed, max_actions_per_env=160)
            self.env.prepare(self.pd, self.ld, pool_games=8)
            for _ in range(150):
                self.env.rollout_step(perm=self.pd, lord_pile=self.ld, pool_games=8)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            self.n = steps_per_graph
            self.env._ensure()
            with torch.cuda.graph(self.graph, stream=stream):
                for _ in range(steps_per_graph):
                    self.env.rollout_step(perm=self.pd, lord_pile=self.ld, pool_games=8, auto_step=True)

    def replay(self):
        with torch.cuda.stream(self.stream):
            self.graph.replay()


def run(groups, envs_per_group, steps_per_graph, reps):
    streams = [torch.cuda.Stream() for _ in range(groups)]
    chains = [Chain(envs_per_group, 100 + i, streams[i], steps_per_graph) for i in range(groups)]
    torch.cuda.synchronize()
    for c in chains:
        c.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in chains:
        c.stream.wait_event(e0)
    for _ in range(reps):
        for c in chains:
            c.replay()
    main = torch.cuda.current_stream()
    for c in chains:
        ev = torch.cuda.Event(); ev.record(c.stream); main.wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = groups * envs_per_group * steps_per_graph * reps
    errs = sum(int(c.env.stats[7].item()) for c in chains)
    return {"groups": groups, "envs_per_group": envs_per_group, "steps_per_graph": steps_per_graph,
            "env_steps_per_s": steps / ms * 1e3, "ms_per_step_of_all_envs": ms / (steps_per_graph * reps), "errors": errs}


if __name__ == "__main__":
    out = [run(1, 131072, 8, 25), run(2, 65536, 8, 25), run(4, 32768, 8, 25), run(2, 131072, 8, 25)]
    for o in out:
        print(json.dumps(o))
