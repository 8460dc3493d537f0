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
    assert D.native.lib.ddz_abi_version() == 1
    assert [D.native.lib.ddz_face_channels(v) for v in range(4)] == [4, 7, 9, 6]
    assert D.native.lib.ddz_face_channels(4) == D.native.E_ARG
    assert D.native.lib.ddz_state_bytes(1000) == 1000 * 76 and D.native.lib.ddz_state_bytes(0) == 0
    assert D.native.lib.ddz_workspace_bytes(4096) == 256 + 8 * 128


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
