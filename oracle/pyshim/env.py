"""Stand-in for the reference's absent native module `env` (reference envi.py:12-13), backed by the C oracle.

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's UNMODIFIED envi.py can be imported in the
build container to generate tests/golden/ fixtures (tests/golden/make_golden.py).  API inferred from the
call sites listed in SURVEY.md Appendix A; semantics frozen in SURVEY.md Appendix C2 ("parity unpinned").
"""
import ctypes as C

import numpy as np

from oracle import ddz_oracle as O

# Deal stream used by the bare `prepare()` the reference calls (game.py:171): game g of the process is
# numpy PCG64(20260101 + g).permutation(54) with lord_pile 0.
DEAL_SEED0 = 20260101
_next_game = [0]


def reset_deal_stream(g=0):
    _next_game[0] = int(g)


def next_deal():
    g = _next_game[0]
    _next_game[0] += 1
    perm = np.random.Generator(np.random.PCG64(DEAL_SEED0 + g)).permutation(54).astype(np.int8)
    return perm, 0


class Env(object):
    def __init__(self, seed=None):
        self._e = np.zeros(1, O.ENV_DTYPE)
        self._e["winner"] = -1
        self._e["cur"] = 1
        self._seed = seed

    @property
    def _p(self):
        return C.c_void_p(self._e.ctypes.data)

    def reset(self):
        O.lib().ddz_ref_env_clear(self._p)

    def prepare(self):
        perm, lord = next_deal()
        self.prepare_manual(perm, lord)

    def prepare_manual(self, perm, lord_pile):
        perm = np.ascontiguousarray(perm, np.int8)
        if O.lib().ddz_ref_env_deal(self._p, O._ptr(perm, C.c_int8), int(lord_pile)) != 0:
            raise ValueError("perm must be a permutation of 0..53 and lord_pile in 0..2")

    def get_role_ID(self):
        return int(self._e["cur"][0]) + 1

    def get_curr_handcards(self):
        cnt = self._e["hand"][0][int(self._e["cur"][0])]
        return np.array([r + 3 for r in range(15) for _ in range(int(cnt[r]))], dtype=int)

    def _cards(self, cnt):
        return [r + 3 for r in range(15) for _ in range(int(cnt[r]))]

    def get_last_two_cards(self):
        s = int(self._e["cur"][0])
        rec = self._e["recent"][0]
        return [self._cards(rec[(s + 2) % 3]), self._cards(rec[(s + 1) % 3])]

    def get_last_outcards(self):
        last = np.zeros(15, np.int8)
        O.lib().ddz_ref_env_last(self._p, O._ptr(last, C.c_int8))
        return np.array(self._cards(last), dtype=int)

    def get_state_prob(self):
        out = np.zeros(120, np.float32)
        O.lib().ddz_ref_state_prob(self._p, O._ptr(out, C.c_float))
        return out

    def get_state_prob_manual(self, known60, size1, size2):
        return O.state_prob_manual(known60, size1, size2)

    def step_manual(self, cards):
        cnt = np.zeros(15, np.int8)
        for c in np.asarray(cards, dtype=int).reshape(-1):
            cnt[int(c) - 3] += 1
        r, d, cat = C.c_int(), C.c_int(), C.c_int()
        rc = O.lib().ddz_ref_env_step(self._p, O._ptr(cnt, C.c_int8), None, C.byref(r), C.byref(d), C.byref(cat), None)
        if rc != 0:
            raise ValueError("illegal move %s" % list(cards))
        return r.value, bool(d.value), cat.value

    def step_auto(self):
        raise NotImplementedError("RHCP step_auto is out of scope (SURVEY.md 2.1)")
