"""CPU: the C-ABI library loads and exports every symbol include/ddz_b200.h declares; host-side logic
(converters, packing, deals, sharding); the product fails loudly without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ddz_b200 as D
    hdr = open(os.path.join(ROOT, "include", "ddz_b200.h")).read()
    declared = set(re.findall(r"\b(ddz_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(D.native.EXPORTS)
    lib = ctypes.CDLL(D.native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert D.native.lib.ddz_abi_version() == 3
    assert [D.native.lib.ddz_face_channels(v) for v in range(4)] == [4, 7, 9, 6]
    assert D.native.lib.ddz_face_channels(4) == D.native.E_ARG
    assert D.native.lib.ddz_state_bytes(1000) == 1000 * 76 and D.native.lib.ddz_state_bytes(0) == 0
    assert D.native.lib.ddz_workspace_bytes(4096) == 256 + 8 * 128
    # argument checks come before any CUDA call (no compute without a GPU): null buffers, empty batches, a width the kernel
    # does not handle
    L, E = D.native.lib, D.native.E_ARG
    assert L.ddz_mcts_moves(None, None, None, None, 4, None) == E and L.ddz_mcts_moves(1, 1, 1, 1, 0, None) == E
    assert L.ddz_playout_pruned(None, 10, 0, 0, 0, None, None, None, 4, None) == E
    assert L.ddz_playout(1, -1, 0, 0, 0, None, None, None, 4, None) == E
    assert L.ddz_q_features(1, 2, 1, 1, None, None, 0, 4, 0, 1, 1, 1, 1, 260, 1, 0, 4, None) == E     # width > 256
    assert L.ddz_q_features(1, 2, 1, 1, None, None, 0, 4, 0, 1, 1, 1, 1, 30, 1, 0, 4, None) == E      # not a multiple of 4
    assert L.ddz_q_features(1, 2, 1, 1, None, None, 2, 4, 0, 1, 1, 1, 1, 256, 1, 0, 4, None) == E     # env range past B
    assert L.ddz_q_features(1, 7, 1, 1, None, None, 0, 4, 0, 1, 1, 1, 1, 256, 1, 0, 4, None) == E     # no such face variant


def test_sass_is_sm100a_only():
    import subprocess, shutil
    import ddz_b200 as D
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", D.native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    import torch
    import ddz_b200 as D
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(D.native.DdzError):
        D.BatchedEnv(4)
    with pytest.raises(D.native.DdzError):
        D.get_moves(np.zeros((1, 15), np.int8), np.zeros((1, 15), np.int8))
    # the product never imports the oracle
    import sys
    for name, mod in list(sys.modules.items()):
        if name.startswith("doudizhu-rl_b200"):
            src = getattr(mod, "__file__", None)
            if src and src.endswith(".py"):
                assert "oracle" not in open(src).read().replace("the CPU oracle", ""), name


def test_converters_match_envi_golden(golden):
    import ddz_b200 as D
    g = golden.converters
    for i, a in enumerate(g["arrs"]):
        cards = g["cards_flat"][g["cards_off"][i]:g["cards_off"][i + 1]]
        assert np.array_equal(D.Env.arr2cards(a), cards)                 # envi.py:119-130
        assert np.array_equal(D.Env.cards2arr(cards), a)                 # envi.py:133-137
        assert np.array_equal(D.Env.onehot2arr(g["onehot"][i]), a)       # envi.py:149-157
    assert np.array_equal(D.Env.batch_arr2onehot(g["arrs"]), g["onehot"])  # envi.py:140-146


def test_pack_unpack_roundtrip(oracle):
    import torch
    import ddz_b200 as D
    cnt = oracle.universe()[0]
    packed = D.pack_counts(torch.as_tensor(cnt))
    assert np.array_equal(packed.numpy().view(np.uint64), oracle.pack(cnt))
    assert np.array_equal(D.unpack_counts(packed).numpy(), cnt)


def test_deal_streams():
    import ddz_b200 as D
    perm, lord = D.default_deals(3, 4)
    assert perm.shape == (4, 54) and (np.sort(perm, 1) == np.arange(54)).all() and not lord.any()
    want = np.random.Generator(np.random.PCG64(20260101 + 5)).permutation(54)
    assert np.array_equal(perm[2], want)
    perm, lord = D.random_deals(100, seed=1, pool_games=3)
    assert perm.shape == (300, 54) and perm.dtype == np.int8 and (np.sort(perm, 1) == np.arange(54)).all()
    assert set(np.unique(lord)) <= {0, 1, 2}


def test_shard_range_partitions():
    import ddz_b200 as D
    for total, world in [(1048576, 8), (10, 3), (7, 8), (0, 2)]:
        spans = [D.sharding.shard_range(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.sharding.shard_range(3, 2, 10)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU oracle port on all host threads) prints one JSON line with the contract's keys."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_replay_buffer_and_td_step_on_cpu():
    """trainer.ReplayBuffer = bounded deque semantics (dqn.py:14); td_step = the reference's target and loss
    (dqn.py:39-47); BatchedDQN schedules (dqn.py:73-80).  Pure torch host logic: runs without a GPU."""
    import torch
    import ddz_b200 as D
    rb = D.ReplayBuffer(5, 2, "cpu")
    mk = lambda n, v: (torch.full((n, 2, 15, 4), float(v)), torch.full((n, 15, 4), float(v)), torch.full((n,), float(v)),
                       torch.full((n, 2, 15, 4), float(v) + 0.5), torch.full((n, 15, 4), float(v) + 0.5), torch.zeros(n))
    rb.append(*mk(3, 1))
    assert len(rb) == 3
    rb.append(*mk(4, 2))                                   # wraps: the two oldest entries are overwritten
    assert len(rb) == 5 and sorted(rb.r[:, 0].tolist()) == [1.0, 2.0, 2.0, 2.0, 2.0]
    rb.append(*mk(9, 3))                                   # more than the capacity at once: the last five survive
    assert len(rb) == 5 and rb.r[:, 0].tolist() == [3.0] * 5
    s0, a0, r, s1, a1, done = rb.sample(64, torch.Generator().manual_seed(0))
    assert s0.shape == (64, 2, 15, 4) and r.shape == (64, 1) and done.shape == (64, 1)

    class Q(torch.nn.Module):                              # Q(s, a) = w * (mean(s) + mean(a))
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(2.0))

        def forward(self, face, actions):
            return (self.w * (face.flatten(1).mean(1) + actions.flatten(1).mean(1))).unsqueeze(1)

    pol, tgt = Q(), Q()
    with torch.no_grad():
        tgt.w.fill_(3.0)
    opt = torch.optim.SGD(pol.parameters(), lr=0.0)
    batch = (torch.ones(4, 2, 15, 4), torch.ones(4, 15, 4), torch.tensor([[0.], [0.], [100.], [-100.]]),
             torch.ones(4, 2, 15, 4) * 2, torch.ones(4, 15, 4) * 2, torch.tensor([[0.], [0.], [1.], [1.]]))
    loss = D.td_step(pol, tgt, opt, batch, gamma=0.95)
    y = torch.tensor([0.95 * 12, 0.95 * 12, 100.0, -100.0])          # r + (1 - done) * gamma * Q_target(s1, a1)
    assert torch.allclose(loss, ((torch.full((4,), 4.0) - y) ** 2).mean())

    dqn = D.BatchedDQN(Q, 2, "cpu", replay_size=64, batch_size=8)
    assert dqn.epsilon == 0.5 and dqn.perceive(*mk(4, 1)) is None     # fewer than a minibatch: no update yet
    assert dqn.perceive(*mk(6, 2)) is not None and len(dqn.replay_buffer) == 10
    dqn.update_epsilon(1066)
    assert abs(dqn.epsilon - (0.01 + 0.49 * np.exp(-1.0))) < 1e-9      # config.py:9-10,13 + dqn.py:73-76
    with torch.no_grad():
        dqn.policy_net.w.fill_(7.0)
    dqn.update_target(19)
    assert float(dqn.target_net.w.detach()) != 7.0
    dqn.update_target(20)
    assert float(dqn.target_net.w.detach()) == 7.0


def test_reference_arm_never_loads_the_product(oracle):
    """bench.py --impl reference is the CPU arm: numpy + oracle/libddz_oracle.so only.  It must not import the product
    package or map libddz_b200.so (the synthetic inputs come from doudizhu-rl_b200/deals.py, loaded by path)."""
    import json
    import subprocess
    import sys
    code = (
        "import sys, json, io, contextlib\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '1', '--config', '%d']\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    bench.main()\n"
        "line = json.loads(buf.getvalue().strip().splitlines()[-1])\n"
        "maps = open('/proc/self/maps').read()\n"
        "bad = [m for m in sys.modules if m.startswith('doudizhu-rl_b200') or m == 'ddz_b200' or m == 'torch']\n"
        "print(json.dumps({'line': line, 'bad_modules': bad, 'product_so': 'libddz_b200' in maps, 'oracle_so': 'libddz_oracle' in maps}))\n")
    for cfg in (4, 5):
        out = subprocess.run([sys.executable, "-c", code % (cfg, ROOT)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        res = json.loads(out.stdout.strip().splitlines()[-1])
        assert res["bad_modules"] == [] and not res["product_so"] and res["oracle_so"]
        line = res["line"]
        assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
        assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
        assert line["unit"] == ("moves/s" if cfg == 5 else "env-steps/s")


def test_adversarial_pool_fixture_matches_the_oracle(oracle):
    """doudizhu-rl_b200/data/adversarial_pool.npz (BASELINE config 5's "previous moves"): the legal lead moves of the ten
    pool hands, as the oracle's definitional generator gives them; adversarial_pairs is deterministic and half-leading."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ddz_deals_t", os.path.join(ROOT, "doudizhu-rl_b200", "deals.py"))
    deals = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(deals)
    f = np.load(deals.POOL_FILE)
    assert np.array_equal(f["pool"], deals.ADVERSARIAL_POOL)
    z = np.zeros(15, np.int8)
    for i, h in enumerate(deals.ADVERSARIAL_POOL):
        want = oracle.pack(oracle.get_moves(h, z, fast=False))
        assert np.array_equal(f["lead_moves"][f["lead_off"][i]:f["lead_off"][i + 1]], want)
    assert f["lead_off"][1] - f["lead_off"][0] == 497                      # the worst hand of all 20-card hands
    h1, l1 = deals.adversarial_pairs(4096, seed=9)
    h2, l2 = deals.adversarial_pairs(4096, seed=9)
    assert np.array_equal(h1, h2) and np.array_equal(l1, l2) and 0.4 < (l1 == 0).mean() < 0.6
    assert set(h1.tolist()) <= set(deals.pack_counts_np(deals.ADVERSARIAL_POOL).tolist())
