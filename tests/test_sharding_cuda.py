"""GPU, world_size 2: the N>1 path on the CUDA env.  Two processes each run THEIR slice of the envs (global env ids via
env0, their rows of the deal pool) as sm_100a kernels -- on two GPUs when the box has them, else both on cuda:0 (the
slices are independent launches; nothing waits across processes) -- and all-reduce the statistics (NCCL with >= 2 GPUs,
gloo on host copies otherwise).  The reduced statistics and every rank's final state must equal a single-process CUDA run
over all envs and the oracle: results do not depend on the rank layout."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS, SEED, G = 1024, 120, 4242, 4


def _deals():
    sys.path.insert(0, ROOT)
    import ddz_b200 as D
    perm, lord = D.random_deals(TOTAL, seed=17, pool_games=G)
    return D, perm.reshape(G, TOTAL, 54), lord.reshape(G, TOTAL)


def _run_cuda_slice(D, perm, lord, lo, hi, device):
    env = D.BatchedEnvCooperation(hi - lo, seed=SEED, device=device, env0=lo)
    pd = torch.as_tensor(np.ascontiguousarray(perm[:, lo:hi])).to(device)
    ld = torch.as_tensor(np.ascontiguousarray(lord[:, lo:hi])).to(device)
    env.prepare(pd, ld, pool_games=G)
    for _ in range(STEPS):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=G)
    torch.cuda.synchronize(device)
    f, m = env._fields()
    return env.stats.clone(), f.cpu().numpy().copy(), m.cpu().numpy().copy()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D, perm, lord = _deals()
    ngpu = torch.cuda.device_count()
    device = torch.device("cuda", rank % ngpu)
    torch.cuda.set_device(device)
    backend = "nccl" if ngpu >= world else "gloo"
    os.environ["LOCAL_RANK"] = str(rank % ngpu)
    r, _, w = D.sharding.init_distributed(backend=backend)
    lo, hi = D.sharding.shard_range(r, w, TOTAL)
    stats, f, m = _run_cuda_slice(D, perm, lord, lo, hi, device)
    local = stats if backend == "nccl" else stats.cpu()
    total = D.sharding.allreduce_stats(local, side_stream=torch.cuda.Stream(device) if backend == "nccl" else None)
    torch.cuda.synchronize(device)
    out.put((r, total.cpu().numpy(), f, m, backend))
    dist.barrier()
    dist.destroy_process_group()


def test_cuda_env_slices_world2_match_single_process_and_oracle(oracle):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get() for _ in range(2)]
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got.sort(key=lambda x: x[0])
    D, perm, lord = _deals()
    stats1, f1, m1 = _run_cuda_slice(D, perm, lord, 0, TOTAL, torch.device("cuda", 0))
    keep = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    for r, total, _, _, backend in got:
        assert np.array_equal(total[keep], stats1.cpu().numpy()[keep]), (r, backend)
    assert np.array_equal(np.concatenate([g[2] for g in got], 1), f1) and np.array_equal(np.concatenate([g[3] for g in got]), m1)
    # ... and the oracle over all envs in one process
    rb = oracle.RefBatch(TOTAL, 2)
    rb.deal(perm.reshape(-1, 54), lord.reshape(-1), pool_games=G)
    for t in range(STEPS):
        rb.observe(want_f32=False, want_face=False)
        rb.step(mode=2, seed=SEED, env0=0, step=t)
        rb.deal(perm.reshape(-1, 54), lord.reshape(-1), only_done=True, pool_games=G)
    fo, mo = rb.export()
    assert np.array_equal(f1.view(np.uint64), fo) and np.array_equal(m1.view(np.uint32), mo)
    assert stats1[7] == 0 and int(stats1[4]) == rb.stats[4] and int(stats1[0]) == rb.stats[0] > 0
