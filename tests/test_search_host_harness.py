"""CPU check of the search bot's move pruning as the product computes it (doudizhu-rl_b200/csrc/ddz_search.cuh: integer
value keys, the generated rank table, pruned_pick), compiled for the host by tests/host_harness, against the golden vectors of
the unmodified server/mcts/evaluator.py + get_moves.py and against the oracle.  The CUDA kernels run the same functions;
their parity tests are in test_gpu_parity.py (-m gpu)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "search_host.cpp")
LIB = os.path.join(HERE, "host_harness", "libsearch_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "doudizhu-rl_b200", "csrc")


@pytest.fixture(scope="module")
def harness():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("ddz_search.cuh", "ddz_device.cuh", "ddz_value_rank.inc")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.search_host_value2.argtypes = [C.c_uint64]
    L.search_host_value_key.argtypes = [C.c_uint64, C.c_int]
    L.search_host_value_key.restype = C.c_uint32
    L.search_host_pruned_list.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    return L


def test_rank_table_is_what_the_generator_writes():
    """ddz_value_rank.inc is generated: the committed file equals the generator's output"""
    out = subprocess.check_output(["python", os.path.join(CSRC, "make_value_rank.py")], text=True)
    assert out == open(os.path.join(CSRC, "ddz_value_rank.inc")).read()


def test_value2_is_twice_cards_value(harness, oracle, golden):
    """every key of the reference's cards_value table (evaluator.py:17-57): the device's integer value is exactly twice it"""
    g = golden.mcts_moves
    packed = oracle.pack(g["value_keys"])
    got = np.array([harness.search_host_value2(int(p)) for p in packed], np.float64)
    assert np.array_equal(got, 2.0 * g["value_vals"])
    assert got.min() >= -14 and got.max() <= 57


def test_value_key_orders_like_python_doubles(harness, oracle, golden):
    """value_key(move, handnum) compares exactly like  cards_value - 0.1 * (handnum - len(move))  does in Python, ties and
    one-ulp differences included: checked on every move of the table at every hand size, class by class"""
    g = golden.mcts_moves
    keys, vals = g["value_keys"], g["value_vals"]
    packed = oracle.pack(keys)
    size = keys.sum(axis=1)
    rng = np.random.default_rng(0)
    pick = rng.permutation(len(keys))[:1500]
    for handnum in (20, 17, 11, 6):
        ok = pick[size[pick] <= handnum]
        want = np.array([vals[i] - 0.1 * (handnum - int(size[i])) for i in ok])       # the reference's expression
        got = np.array([harness.search_host_value_key(int(packed[i]), handnum) for i in ok], np.int64)
        order = np.argsort(want, kind="stable")
        w, k = want[order], got[order]
        assert np.all((w[1:] > w[:-1]) == (k[1:] > k[:-1])) and np.all((w[1:] == w[:-1]) == (k[1:] == k[:-1]))
    counts, _, _, _, extra = oracle.universe()
    for c in counts[extra == 1]:
        assert harness.search_host_value_key(int(oracle.pack(c)), 20) == 0xFFFF


def test_pruned_pick_reproduces_get_moves_py(harness, oracle, golden):
    """the pruned list read entry by entry through pruned_pick == server/mcts/get_moves.py on the 769 golden positions"""
    g = golden.mcts_moves
    offs = g["offsets"]
    out = np.zeros(344, np.uint64)
    for i, (h, l) in enumerate(zip(g["hands"], g["lasts"])):
        n = harness.search_host_pruned_list(int(oracle.pack(h)), int(oracle.pack(l)), out.ctypes.data_as(C.POINTER(C.c_uint64)), 344)
        assert n == offs[i + 1] - offs[i], (i, n)
        assert np.array_equal(oracle.unpack(out[:n]), g["moves"][offs[i]:offs[i + 1]]), i


def test_pruned_pick_random_hands_against_oracle(harness, oracle):
    rng = np.random.default_rng(11)
    deck = np.array([i // 4 for i in range(52)] + [13, 14])
    out = np.zeros(344, np.uint64)
    pruned = 0
    for t in range(300):
        hand = np.bincount(rng.permutation(deck)[:rng.integers(2, 21)], minlength=15).astype(np.int8)
        last = np.zeros(15, np.int8)
        if t % 3:
            other = np.bincount(rng.permutation(deck)[:rng.integers(5, 21)], minlength=15).astype(np.int8)
            lead = oracle.get_moves(other, last, fast=True)
            last = lead[rng.integers(len(lead))]
        want = oracle.mcts_moves(hand, last)
        n = harness.search_host_pruned_list(int(oracle.pack(hand)), int(oracle.pack(last)), out.ctypes.data_as(C.POINTER(C.c_uint64)), 344)
        assert n == len(want) and np.array_equal(oracle.unpack(out[:n]), want), t
        pruned += len(oracle.get_moves(hand, last, fast=True)) > 10
    assert pruned > 50
