"""ctypes binding of the CPU oracle (oracle/ddz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (doudizhu-rl_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("DDZ_ORACLE_LIB") or os.path.join(_HERE, "libddz_oracle.so")   # DDZ_ORACLE_LIB: the form-B build
_lib = None

MAX_LEGAL = 512
NUNIVERSE = 13551
NACTIONS_CARD_PY = 13527
VARIANT_CHANNELS = (4, 7, 9, 6)
DEFAULT_REWARDS = (50, 100, 50)  # role 0 up, 1 lord, 2 down (reference game.py:13-14)


class RefEnv(C.Structure):
    _fields_ = [("hand", (C.c_int8 * 15) * 3), ("hist", (C.c_int8 * 15) * 3), ("recent", (C.c_int8 * 15) * 3),
                ("cur", C.c_int8), ("done", C.c_int8), ("winner", C.c_int8), ("err", C.c_int8),
                ("games", C.c_int32)]


ENV_DTYPE = np.dtype([("hand", np.int8, (3, 15)), ("hist", np.int8, (3, 15)), ("recent", np.int8, (3, 15)),
                      ("cur", np.int8), ("done", np.int8), ("winner", np.int8), ("err", np.int8),
                      ("games", np.int32)], align=True)
assert ENV_DTYPE.itemsize == C.sizeof(RefEnv)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ddz_oracle.c", "ddz_oracle.h")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        i8p, i32p, i64p, u64p, f32p, u8p, u32p = (C.POINTER(t) for t in (
            C.c_int8, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_uint8, C.c_uint32))
        ip = C.POINTER(C.c_int)
        envp = C.c_void_p
        L.ddz_ref_universe_size.restype = C.c_int
        L.ddz_ref_universe_get.argtypes = [C.c_int, i8p, ip, ip, ip, ip]
        L.ddz_ref_classify.argtypes = [i8p, ip, ip, ip]
        L.ddz_ref_bigger_than.argtypes = [C.c_int] * 6
        L.ddz_ref_get_moves.argtypes = [i8p, i8p, i8p, C.c_int]
        L.ddz_ref_get_moves_fast.argtypes = [i8p, i8p, i8p, C.c_int]
        L.ddz_ref_count_moves_fast.argtypes = [i8p, i8p]
        L.ddz_ref_env_clear.argtypes = [envp]
        L.ddz_ref_env_clear.restype = None
        L.ddz_ref_env_deal.argtypes = [envp, i8p, C.c_int]
        L.ddz_ref_env_last.argtypes = [envp, i8p]
        L.ddz_ref_env_last.restype = None
        L.ddz_ref_env_legal.argtypes = [envp, i8p, C.c_int, C.c_int]
        L.ddz_ref_env_step.argtypes = [envp, i8p, i32p, ip, ip, ip, f32p]
        L.ddz_ref_prob_form.restype = C.c_int
        L.ddz_ref_state_prob.argtypes = [envp, f32p]
        L.ddz_ref_state_prob.restype = None
        L.ddz_ref_state_prob_manual.argtypes = [i32p, C.c_int, C.c_int, f32p]
        L.ddz_ref_state_prob_manual.restype = None
        L.ddz_ref_env_face.argtypes = [envp, C.c_int, f32p]
        L.ddz_ref_encode_actions.argtypes = [i8p, C.c_int, f32p]
        L.ddz_ref_encode_actions.restype = None
        L.ddz_ref_pack.argtypes = [i8p]
        L.ddz_ref_pack.restype = C.c_uint64
        L.ddz_ref_unpack.argtypes = [C.c_uint64, i8p]
        L.ddz_ref_unpack.restype = None
        L.ddz_ref_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.ddz_ref_philox.restype = C.c_uint32
        L.ddz_ref_batch_observe.argtypes = [envp, C.c_int, C.c_int, C.c_int, i32p, u64p, f32p, C.c_int64, f32p]
        L.ddz_ref_batch_observe.restype = C.c_int64
        L.ddz_ref_batch_step.argtypes = [envp, C.c_int, i32p, u64p, i32p, C.c_int, C.c_uint64, C.c_uint64,
                                         C.c_uint32, i32p, i8p, u8p, i8p, f32p, i64p]
        L.ddz_ref_batch_deal.argtypes = [envp, C.c_int, i8p, i8p, C.c_int, C.c_int]
        L.ddz_ref_batch_export.argtypes = [envp, C.c_int, u64p, u32p]
        L.ddz_ref_batch_export.restype = None
        L.ddz_ref_rollout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, i8p, i8p, C.c_int, C.c_int, i64p, u64p,
                                      C.POINTER(C.c_double)]
        L.ddz_ref_rollout.restype = C.c_int64
        L.ddz_ref_rollout_export.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, i8p, i8p, C.c_int, C.c_int,
                                             i64p, u64p, u32p]
        L.ddz_ref_rollout_export.restype = C.c_int64
        L.ddz_ref_get_moves_batch.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int, i32p, u64p, C.POINTER(C.c_double)]
        L.ddz_ref_get_moves_batch.restype = C.c_int64
        L.ddz_ref_cards_value.argtypes = [i8p]
        L.ddz_ref_cards_value.restype = C.c_double
        L.ddz_ref_mcts_moves.argtypes = [i8p, i8p, i8p, C.c_int]
        L.ddz_ref_get_moves_digest.argtypes = [u64p, u64p, C.c_int, i32p, u64p, u64p]
        L.ddz_ref_get_moves_digest.restype = C.c_int64
        L.ddz_ref_count_lead_closed.argtypes = [i8p]
        L.ddz_ref_max_lead_moves_exhaustive.argtypes = [C.c_int, i8p, C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def _i8(a):
    return np.ascontiguousarray(a, dtype=np.int8)


# ---------------------------------------------------------------- rules
def universe():
    """(counts int8[U,15], cat, len, val, is_extra) in canonical order."""
    L = lib()
    n = L.ddz_ref_universe_size()
    counts = np.zeros((n, 15), np.int8)
    meta = np.zeros((n, 4), np.int32)
    c, l, v, x = (C.c_int() for _ in range(4))
    for i in range(n):
        L.ddz_ref_universe_get(i, _ptr(counts[i], C.c_int8), C.byref(c), C.byref(l), C.byref(v), C.byref(x))
        meta[i] = (c.value, l.value, v.value, x.value)
    return counts, meta[:, 0], meta[:, 1], meta[:, 2], meta[:, 3]


def classify(counts):
    c, l, v = C.c_int(), C.c_int(), C.c_int()
    a = _i8(counts)
    idx = lib().ddz_ref_classify(_ptr(a, C.c_int8), C.byref(c), C.byref(l), C.byref(v))
    return idx, c.value, l.value, v.value


def bigger_than(a, g):
    return bool(lib().ddz_ref_bigger_than(*[int(x) for x in a], *[int(x) for x in g]))


def get_moves(hand, last, fast=False):
    """r.get_moves(hand15, last15) -> int8[N,15]"""
    h, l = _i8(hand), _i8(last)
    out = np.zeros((MAX_LEGAL, 15), np.int8)
    f = lib().ddz_ref_get_moves_fast if fast else lib().ddz_ref_get_moves
    n = f(_ptr(h, C.c_int8), _ptr(l, C.c_int8), _ptr(out, C.c_int8), MAX_LEGAL)
    if n < 0 or n > MAX_LEGAL:
        raise ValueError("get_moves failed: %d" % n)
    return out[:n].copy()


def cards_value(counts):
    """cards_value[tuple(counts)] of server/mcts/evaluator.py; None for a vector that is not in that table"""
    a = _i8(counts)
    v = float(lib().ddz_ref_cards_value(_ptr(a, C.c_int8)))
    return None if v < -1e29 else v


def mcts_moves(hand, last):
    """mcts.get_moves.get_moves(hand, last) (server/mcts/get_moves.py:36-69) -> int8[M,15] in the reference's order"""
    h, l = _i8(hand), _i8(last)
    out = np.zeros((MAX_LEGAL, 15), np.int8)
    n = lib().ddz_ref_mcts_moves(_ptr(h, C.c_int8), _ptr(l, C.c_int8), _ptr(out, C.c_int8), MAX_LEGAL)
    if n < 0 or n > MAX_LEGAL:
        raise ValueError("mcts_moves failed: %d" % n)
    return out[:n].copy()


def pack(counts):
    a = _i8(counts)
    if a.ndim == 1:
        return int(lib().ddz_ref_pack(_ptr(a, C.c_int8)))
    sh = (np.arange(15, dtype=np.uint64) * np.uint64(4))
    return (a.astype(np.uint64) << sh).sum(axis=-1, dtype=np.uint64)


def unpack(packed):
    p = np.asarray(packed, dtype=np.uint64)
    sh = (np.arange(15, dtype=np.uint64) * np.uint64(4))
    return ((p[..., None] >> sh) & np.uint64(15)).astype(np.int8)


def philox(seed, env, step):
    return int(lib().ddz_ref_philox(int(seed), int(env), int(step)))


def state_prob_manual(known60, size1, size2):
    k = np.ascontiguousarray(known60, dtype=np.int32).reshape(60)
    out = np.zeros(120, np.float32)
    lib().ddz_ref_state_prob_manual(_ptr(k, C.c_int32), int(size1), int(size2), _ptr(out, C.c_float))
    return out


# ---------------------------------------------------------------- batched env
class RefBatch:
    """B oracle envs driven in lock-step; mirrors the product's batched API for parity tests."""

    def __init__(self, B, variant=2, rewards=DEFAULT_REWARDS):
        self.B, self.variant = int(B), int(variant)
        self.C = VARIANT_CHANNELS[variant]
        self.envs = np.zeros(self.B, ENV_DTYPE)
        self.envs["winner"] = -1
        self.envs["cur"] = 1
        self.rewards = np.asarray(rewards, np.int32)
        self.stats = np.zeros(16, np.int64)
        self.offsets = np.zeros(self.B + 1, np.int32)
        self.actions_u64 = np.zeros(0, np.uint64)

    @property
    def _p(self):
        return self.envs.ctypes.data_as(C.c_void_p)

    def clear(self):
        for b in range(self.B):
            lib().ddz_ref_env_clear(C.c_void_p(self.envs.ctypes.data + b * ENV_DTYPE.itemsize))

    def deal(self, perm, lord_pile=None, only_done=False, pool_games=1):
        perm = _i8(perm)
        lp = None if lord_pile is None else _i8(lord_pile)
        rc = lib().ddz_ref_batch_deal(self._p, self.B, _ptr(perm, C.c_int8), _ptr(lp, C.c_int8),
                                      int(only_done), int(pool_games))
        if rc != 0:
            raise ValueError("bad deal")

    def observe(self, fast=True, want_f32=True, want_face=True):
        cap = self.B * MAX_LEGAL
        offsets = np.zeros(self.B + 1, np.int32)
        # first pass for the count keeps the buffers tight
        total = lib().ddz_ref_batch_observe(self._p, self.B, self.variant, int(fast), _ptr(offsets, C.c_int32),
                                            None, None, 0, None)
        if total < 0:
            raise ValueError("observe failed")
        au = np.zeros(total, np.uint64)
        af = np.zeros((total, 15, 4), np.float32) if want_f32 else None
        face = np.zeros((self.B, self.C, 15, 4), np.float32) if want_face else None
        lib().ddz_ref_batch_observe(self._p, self.B, self.variant, int(fast), _ptr(offsets, C.c_int32),
                                    _ptr(au, C.c_uint64), _ptr(af, C.c_float), total,
                                    _ptr(face, C.c_float))
        self.offsets, self.actions_u64 = offsets, au
        return offsets, au, af, face

    def step(self, choice=None, mode=0, seed=0, env0=0, step=0):
        ch = None if choice is None else np.ascontiguousarray(choice).view(np.int32)
        r = np.zeros(self.B, np.int8)
        done = np.zeros(self.B, np.uint8)
        cat = np.zeros(self.B, np.int8)
        rew = np.zeros((self.B, 3), np.float32)
        lib().ddz_ref_batch_step(self._p, self.B, _ptr(self.offsets, C.c_int32), _ptr(self.actions_u64, C.c_uint64),
                                 _ptr(ch, C.c_int32), int(mode), int(seed), int(env0), int(step),
                                 _ptr(self.rewards, C.c_int32), _ptr(r, C.c_int8), _ptr(done, C.c_uint8),
                                 _ptr(cat, C.c_int8), _ptr(rew, C.c_float), _ptr(self.stats, C.c_int64))
        return r, done, cat, rew

    def export(self):
        f = np.zeros((9, self.B), np.uint64)
        meta = np.zeros(self.B, np.uint32)
        lib().ddz_ref_batch_export(self._p, self.B, _ptr(f, C.c_uint64), _ptr(meta, C.c_uint32))
        return f, meta


def rollout(B, warm_steps, steps, variant, seed, perm_pool, lord_pool, pool_games, nthreads):
    """CPU baseline: returns (timed env_steps, seconds, stats int64[16], checksum)."""
    perm_pool = _i8(perm_pool)
    lp = None if lord_pool is None else _i8(lord_pool)
    stats = np.zeros(16, np.int64)
    cs = C.c_uint64(0)
    sec = C.c_double(0)
    n = lib().ddz_ref_rollout(int(B), int(warm_steps), int(steps), int(variant), int(seed), _ptr(perm_pool, C.c_int8),
                              _ptr(lp, C.c_int8), int(pool_games), int(nthreads), _ptr(stats, C.c_int64),
                              C.byref(cs), C.byref(sec))
    return int(n), float(sec.value), stats, int(cs.value)


def rollout_export(B, steps, variant, seed, perm_pool, lord_pool, pool_games, nthreads, env0=0):
    """the rollout above from the deal, all envs (global env ids env0 .. env0 + B - 1): (env_steps, stats int64[16],
    fields uint64[9,B], meta uint32[B])"""
    perm_pool = _i8(perm_pool)
    lp = None if lord_pool is None else _i8(lord_pool)
    stats = np.zeros(16, np.int64)
    f = np.zeros((9, int(B)), np.uint64)
    meta = np.zeros(int(B), np.uint32)
    n = lib().ddz_ref_rollout_export(int(B), int(steps), int(variant), int(seed), int(env0), _ptr(perm_pool, C.c_int8),
                                     _ptr(lp, C.c_int8), int(pool_games), int(nthreads), _ptr(stats, C.c_int64),
                                     _ptr(f, C.c_uint64), _ptr(meta, C.c_uint32))
    if n < 0:
        raise ValueError("rollout_export failed")
    return int(n), stats, f, meta


def get_moves_batch(hands_packed, lasts_packed, reps=1, nthreads=1):
    """CPU baseline of config 5: (moves emitted, seconds, counts int32[n], checksum) for packed uint64 pairs"""
    h = np.ascontiguousarray(hands_packed, dtype=np.uint64)
    l = np.ascontiguousarray(lasts_packed, dtype=np.uint64)
    counts = np.zeros(len(h), np.int32)
    cs, sec = C.c_uint64(0), C.c_double(0)
    n = lib().ddz_ref_get_moves_batch(_ptr(h, C.c_uint64), _ptr(l, C.c_uint64), len(h), int(reps), int(nthreads),
                                      _ptr(counts, C.c_int32), C.byref(cs), C.byref(sec))
    if n < 0:
        raise ValueError("get_moves_batch failed")
    return int(n), float(sec.value), counts, int(cs.value)


PHI = 0x9E3779B97F4A7C15


def get_moves_digest(hands_packed, lasts_packed):
    """(moves, counts int32[n], unordered digest, ordered digest) of the lists of packed uint64 pairs; see lists_digest"""
    h = np.ascontiguousarray(hands_packed, dtype=np.uint64)
    l = np.ascontiguousarray(lasts_packed, dtype=np.uint64)
    counts = np.zeros(len(h), np.int32)
    un, od = C.c_uint64(0), C.c_uint64(0)
    n = lib().ddz_ref_get_moves_digest(_ptr(h, C.c_uint64), _ptr(l, C.c_uint64), len(h), _ptr(counts, C.c_int32),
                                       C.byref(un), C.byref(od))
    if n < 0:
        raise ValueError("get_moves_digest failed")
    return int(n), counts, int(un.value), int(od.value)


def lists_digest(actions_u64, offsets):
    """the same two digests computed from CSR lists (numpy, wrapping uint64 arithmetic)"""
    a = np.ascontiguousarray(actions_u64).view(np.uint64)
    off = np.asarray(offsets, np.int64)
    a = a[:off[-1]]
    k = np.arange(len(a), dtype=np.int64) - np.repeat(off[:-1], np.diff(off))
    with np.errstate(over="ignore"):
        p = a * np.uint64(PHI)
        return int(p.sum(dtype=np.uint64)), int((p * (2 * k + 1).astype(np.uint64)).sum(dtype=np.uint64))


def count_lead_closed(hand):
    h = _i8(hand)
    return int(lib().ddz_ref_count_lead_closed(_ptr(h, C.c_int8)))


def max_lead_moves_exhaustive(nthreads=8):
    """(max lead moves over all 20-card hands, an argmax hand, hands visited)"""
    best = np.zeros(15, np.int8)
    vis = C.c_longlong(0)
    m = lib().ddz_ref_max_lead_moves_exhaustive(int(nthreads), _ptr(best, C.c_int8), C.byref(vis))
    return int(m), best, int(vis.value)


def count_moves_batch(hands, lasts):
    """number of legal moves for each (hand, last) pair (constructive generator, counting only) -> int32 [n]"""
    h, l = _i8(hands).reshape(-1, 15), _i8(lasts).reshape(-1, 15)
    out = np.zeros(len(h), np.int32)
    f = lib().ddz_ref_count_moves_fast
    for i in range(len(h)):
        out[i] = f(_ptr(h[i], C.c_int8), _ptr(l[i], C.c_int8))
    return out
