"""CPU check of the flat long-list enumerator (doudizhu-rl_b200/csrc/ddz_flat.cuh): the product's descriptor walk, subset
table and decode, compiled for the host by tests/host_harness, against the oracle and the golden legal sets.  The CUDA
kernel k_legal_flat runs the same functions; its own parity tests are in test_gpu_parity.py (-m gpu)."""
import ctypes as C
import os
import subprocess
from math import comb

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "flat_host.cpp")
LIB = os.path.join(HERE, "host_harness", "libflat_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "doudizhu-rl_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    deps = [SRC, os.path.join(CSRC, "ddz_flat.cuh"), os.path.join(CSRC, "ddz_device.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    u64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    L.flat_host_tile.argtypes = [u64p, u64p, C.c_int, u64p, C.c_longlong, i32p, C.c_int, C.c_longlong, C.POINTER(C.c_int)]
    L.flat_host_tile.restype = C.c_longlong
    L.flat_host_table.argtypes = [C.POINTER(C.c_uint16)]
    return L


def run_tiles(L, hands_packed, lasts_packed, arena=512, base0=0):
    """all pairs, 32 per tile, concatenated exactly as the kernel lays them out; returns (lists, rounds)"""
    n = len(hands_packed)
    out = np.zeros(n * 600 + base0 + 8, np.uint64)
    offs = np.zeros(n + 1, np.int32)
    base, rounds = base0, 0
    u64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    for t0 in range(0, n, 32):
        nh = min(32, n - t0)
        h = np.ascontiguousarray(hands_packed[t0:t0 + nh]); l = np.ascontiguousarray(lasts_packed[t0:t0 + nh])
        o = np.zeros(nh + 1, np.int32)
        r = C.c_int(0)
        tot = L.flat_host_tile(h.ctypes.data_as(u64p), l.ctypes.data_as(u64p), nh, out.ctypes.data_as(u64p), len(out),
                               o.ctypes.data_as(i32p), arena, base, C.byref(r))
        assert tot >= 0, "harness error %d in tile %d" % (tot, t0 // 32)
        offs[t0:t0 + nh + 1] = o
        base += tot
        rounds += r.value
    return [out[offs[i]:offs[i + 1]] for i in range(n)], rounds


def _rand_hands(rng, n, lo=1, hi=21):
    deck = np.array([i // 4 for i in range(52)] + [13, 14])
    return np.stack([np.bincount(rng.permutation(deck)[:rng.integers(lo, hi)], minlength=15) for _ in range(n)]).astype(np.int8)


def test_subset_table_is_colex(harness):
    t = np.zeros(909, np.uint16)
    assert harness.flat_host_table(t.ctypes.data_as(C.POINTER(C.c_uint16))) == 909
    off = [0, 1, 16, 107, 327, 657, 909]
    nmax = [0, 15, 14, 12, 11, 10]
    assert t[0] == 0
    for k in range(1, 6):
        seg = t[off[k]:off[k + 1]].astype(np.int64)
        assert len(seg) == comb(nmax[k], k)
        assert all(bin(int(v)).count("1") == k for v in seg) and np.all(np.diff(seg) > 0) and seg[-1] < (1 << nmax[k])
        for n in range(k, nmax[k] + 1):                      # the first C(n,k) entries are the subsets of {0..n-1}
            assert seg[comb(n, k) - 1] < (1 << n) and (comb(n, k) == len(seg) or seg[comb(n, k)] >= (1 << n))


def test_flat_lists_equal_golden_legal_sets(harness, oracle, golden):
    g = golden.legal_sets
    lists, _ = run_tiles(harness, oracle.pack(g["hands"]), oracle.pack(g["lasts"]))
    off = g["legal_off"]
    for i, got in enumerate(lists):
        want = oracle.pack(g["universe"][g["legal_idx"][off[i]:off[i + 1]]])
        assert np.array_equal(got, np.atleast_1d(want)), (i, g["hands"][i], g["lasts"][i])


@pytest.mark.parametrize("arena,base0", [(512, 0), (160, 1), (192, 7)])
def test_flat_lists_random_and_adversarial_vs_oracle(harness, oracle, arena, base0):
    rng = np.random.default_rng(23 + arena)
    z = np.zeros(15, np.int8)
    pool = np.array([[1, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0], [3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0],
                     [4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1], [3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 2, 2, 1, 1],
                     [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1], [2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0],
                     [3, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 0, 0], [4, 4, 0, 0, 0, 0, 2, 2, 2, 2, 0, 0, 0, 1, 1],
                     [1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 0, 0, 0], [2, 2, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],
                     [0] * 15, [4] * 13 + [1, 1], [3] * 13 + [0, 1], [2] * 13 + [1, 0]], np.int8)
    hands = np.concatenate([_rand_hands(rng, 1500), pool[rng.integers(0, len(pool), 500)]])
    n = len(hands)
    lasts = np.zeros((n, 15), np.int8)
    for i in range(n):
        if i % 3:
            om = oracle.get_moves(pool[rng.integers(0, 10)] if i % 2 else _rand_hands(rng, 1, 2, 21)[0], z, fast=True)
            lasts[i] = om[rng.integers(0, len(om))]
    lists, rounds = run_tiles(harness, oracle.pack(hands), oracle.pack(lasts), arena=arena, base0=base0)
    assert rounds >= (n + 31) // 32
    big = 0
    for i, got in enumerate(lists):
        if hands[i].sum() > 20:                             # beyond a physical hand: the oracle's buffers end at 512 moves
            cnt = oracle.count_moves_batch(hands[i:i + 1], lasts[i:i + 1])[0]
            assert len(got) == cnt and len(set(got.tolist())) == cnt
            big += 1
            continue
        want = oracle.pack(oracle.get_moves(hands[i], lasts[i], fast=(i % 5 != 0)))
        assert np.array_equal(got, np.atleast_1d(want)), (i, hands[i], lasts[i])
    assert big > 0


def test_flat_tiny_tiles_with_odd_base(harness, oracle):
    """tiles with very few groups and an odd list offset: the first window starts one slot before the tile's first move
    and lanes look past the arena's sentinel (a regression: they once reported a group start at slot 0)"""
    rng = np.random.default_rng(99)
    z = np.zeros(15, np.int8)
    for trial in range(300):
        nh = int(rng.integers(1, 6))
        hands = _rand_hands(rng, nh, 1, 4)
        lasts = np.zeros((nh, 15), np.int8)
        if trial % 2:
            lasts[0] = np.eye(15, dtype=np.int8)[int(rng.integers(0, 13))]
        lists, _ = run_tiles(harness, oracle.pack(hands), oracle.pack(lasts), base0=int(rng.integers(0, 4)))
        for i, got in enumerate(lists):
            assert np.array_equal(got, np.atleast_1d(oracle.pack(oracle.get_moves(hands[i], lasts[i], fast=True)))), (trial, i)
