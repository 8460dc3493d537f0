"""CPU, world_size 2, gloo: the N>1 path.  Envs shard with no data-path collective; the only exchange is the stats
all-reduce.  Each rank runs its env slice on the oracle (no GPU here) keyed by GLOBAL env ids, and the all-reduced
statistics must equal a single-process run over all envs -- i.e. results do not depend on the number of ranks."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS, SEED, G = 96, 150, 4242, 4


def _run_slice(lo, hi):
    sys.path.insert(0, ROOT)
    import ddz_b200 as D
    from oracle import ddz_oracle as O
    perm, lord = D.random_deals(TOTAL, seed=17, pool_games=G)
    perm = perm.reshape(G, TOTAL, 54)[:, lo:hi].reshape(-1, 54)
    lord = lord.reshape(G, TOTAL)[:, lo:hi].reshape(-1)
    rb = O.RefBatch(hi - lo, 2)
    rb.deal(perm, lord, pool_games=G)
    for t in range(STEPS):
        rb.observe(want_f32=False, want_face=False)
        rb.step(mode=2, seed=SEED, env0=lo, step=t)
        rb.deal(perm, lord, only_done=True, pool_games=G)
    return rb.stats.copy()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import ddz_b200 as D
    r, _, w = D.sharding.init_distributed(backend="gloo")
    lo, hi = D.sharding.shard_range(r, w, TOTAL)
    local = torch.as_tensor(_run_slice(lo, hi))
    total = D.sharding.allreduce_stats(local)
    assert torch.equal(local, torch.as_tensor(_run_slice(lo, hi)))      # input left per-rank
    slowest = D.sharding.max_over_ranks(1.0 + r, torch.device("cpu"))
    if r == 0:
        out.put((total.numpy(), slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_world2_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    total, slowest = q.get()
    single = _run_slice(0, TOTAL)
    assert np.array_equal(total, single)
    assert single[0] > 0 and single[4] > 0
    assert slowest == 2.0
