"""Generate tests/golden/*.npz from the REFERENCE's own importable pure-Python code.

Run once in the build container (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py [/root/reference]

What is pinned, and by which reference code:
  card_space.npz   rule_based/utils/card.py:34-159 action_space (13 527 moves, order), Category2Range,
                   and CardGroup.to_cardgroup(...).type/len/value for every move (card.py:372-527)
  beats.npz        CardGroup.bigger_than (card.py:307-325) on sampled move pairs
  legal_sets.npz   server/rule_utils/utils.py:21-38 get_mask_onehot60 evaluated on random and adversarial
                   (hand, last) pairs over the universe "action space + the 24 rocket-kicker moves"
  rocket24.npz     the 24 vectors built by server/mcts/get_moves.py:22-34 (those source lines are exec'd)
  envi_trace.npz   the UNMODIFIED envi.py (all four Env classes, envi.py:16-217) run on top of the oracle's
                   stand-ins for the absent natives (oracle/pyshim), random play from a fixed index stream
  converters.npz   envi.py:118-157 arr2cards / cards2arr / batch_arr2onehot / onehot2arr samples
  game_transitions.npz  the UNMODIFIED game.py Game.play (lord training, all three seats agents) with a stub DQN that records
                   every perceive(s0, a0, r, s1, a1, done) call: pins the delayed-feedback transition assembly
  core_payloads.npz  server/core.py:26-67 Predictor.face / Predictor.valid_actions (unmodified, over oracle/pyshim) on the
                   example payloads of server/client.py and on payloads cut from random games

Nothing of the reference is copied into the repo: only input/output vectors are stored.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyshim"))
sys.path.insert(0, REF)

np.int = int  # envi.py:23 uses the alias NumPy removed in 1.24  
np.bool = bool

import rule_based.utils.card as card  # noqa: E402
import server.rule_utils.utils as rutils  # noqa: E402
from oracle import ddz_oracle as O  # noqa: E402

RANKS = card.Card.cards


def to_counts(chars):
    c = np.zeros(15, np.int8)
    for x in chars:
        c[RANKS.index(x)] += 1
    return c


def to_chars(counts):
    return [RANKS[r] for r in range(15) for _ in range(int(counts[r]))]


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print("wrote", name, os.path.getsize(path), "bytes")


# ------------------------------------------------------------------ card_space
A = card.action_space
space = np.array([to_counts(a) for a in A], np.int8)
attrs = np.array([(g.type, g.len, g.value) for g in (card.CardGroup.to_cardgroup(a) for a in A)], np.int32)
save("card_space.npz", counts=space, attrs=attrs, cat_range=np.array(card.Category2Range, np.int32))

# ------------------------------------------------------------------ rocket24
src = open(os.path.join(REF, "server", "mcts", "get_moves.py"), encoding="utf-8").read().splitlines()
ns = {}
exec("\n".join(src[21:34]), ns)  # lines 22-34: builds sidaihuojian (13) and sandaihuojian (11)
rocket24 = np.array(ns["sidaihuojian"] + ns["sandaihuojian"], np.int8)
assert rocket24.shape == (24, 15)
save("rocket24.npz", counts=rocket24)

# universe = action space with the 24 extras (classified by the reference's own to_cardgroup)
extra_attrs = np.array([(g.type, g.len, g.value) for g in
                        (card.CardGroup.to_cardgroup(to_chars(c)) for c in rocket24)], np.int32)

# ------------------------------------------------------------------ beats
rng = np.random.default_rng(7)
groups = [card.CardGroup.to_cardgroup(a) for a in A]
pairs = rng.integers(0, len(A), size=(40000, 2))
# make sure same-category pairs are well represented
for lo, hi in card.Category2Range:
    if hi - lo > 1:
        extra = rng.integers(lo, hi, size=(min(2000, (hi - lo) ** 2), 2))
        pairs = np.concatenate([pairs, extra])
beats = np.array([groups[i].bigger_than(groups[j]) for i, j in pairs], np.uint8)
save("beats.npz", pairs=pairs.astype(np.int32), beats=beats)

# ------------------------------------------------------------------ legal sets
ucounts, ucat, ulen, uval, uextra = O.universe()
uchars = [to_chars(c) for c in ucounts]
deck = np.array([i // 4 for i in range(52)] + [13, 14])


def rand_hand(n):
    c = np.zeros(15, np.int8)
    for x in rng.permutation(deck)[:n]:
        c[x] += 1
    return c


ADVERSARIAL = [
    [1, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0],  # 497 lead moves
    [3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0],  # 5-trio airplane + 5 singles
    [4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1],  # 4 bombs + rocket
    [3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 2, 2, 1, 1],  # 4 trios + both jokers
    [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1],  # 12-straight + pair of 2 + rocket
    [2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0, 0],  # 10 consecutive pairs
    [3, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 2, 0, 0],  # 6-trio line
    [4, 4, 0, 0, 0, 0, 2, 2, 2, 2, 0, 0, 0, 1, 1],  # bombs + pair kickers + rocket
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],  # rocket only
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0],  # single card
    [4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],  # bomb + rocket (4+rocket extra)
    [3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],  # 2-airplane + rocket extra
    [0, 0, 0, 0, 0, 0, 0, 3, 3, 3, 3, 3, 1, 1, 1],  # get_moves.py:8 example hand
]
hands, lasts = [], []
zero = np.zeros(15, np.int8)
for h in ADVERSARIAL:
    h = np.array(h, np.int8)
    hands.append(h)
    lasts.append(zero)
    # follow a sample of moves another (full-deck-minus-hand) player could have played
    rest = np.repeat(np.arange(15), np.array([4] * 13 + [1, 1]) - h)
    other = np.bincount(rng.permutation(rest)[:20], minlength=15).astype(np.int8)
    om = O.get_moves(other, zero, fast=True)
    for k in rng.choice(len(om), size=min(6, len(om)), replace=False):
        hands.append(h)
        lasts.append(om[k])
for it in range(700):
    h = rand_hand(int(rng.integers(1, 21)))
    if it % 3 == 0:
        last = zero
    else:
        oh = rand_hand(int(rng.integers(2, 21)))
        om = O.get_moves(oh, zero, fast=True)
        # bias toward multi-card moves so every category gets followed
        w = om.sum(axis=1).astype(np.float64) ** 2
        last = om[rng.choice(len(om), p=w / w.sum())]
    hands.append(h)
    lasts.append(last)
# every one of the 24 extras as the trick to beat, against a strong hand
for c in rocket24:
    hands.append(np.array([4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 2, 1, 1], np.int8))
    lasts.append(c)

legal_idx, legal_off = [], [0]
for n, (h, last) in enumerate(zip(hands, lasts)):
    lead = not last.any()
    mask = rutils.get_mask_onehot60(to_chars(h), uchars, None if lead else to_chars(last))
    valid = np.flatnonzero(mask.sum(axis=1) > 0)  # index 0 (pass) is all-zero in this form
    if not lead:
        valid = np.concatenate([[0], valid])      # follow => pass legal (server/CFR.py:152-174)
    legal_idx.append(valid.astype(np.int32))
    legal_off.append(legal_off[-1] + len(valid))
    if n % 100 == 0:
        print("legal set", n, "/", len(hands), flush=True)
save("legal_sets.npz", hands=np.array(hands, np.int8), lasts=np.array(lasts, np.int8),
     legal_idx=np.concatenate(legal_idx), legal_off=np.array(legal_off, np.int64),
     universe=ucounts, universe_attrs=np.stack([ucat, ulen, uval, uextra], 1).astype(np.int32),
     extra_attrs=extra_attrs)

# ------------------------------------------------------------------ envi trace
import envi  # noqa: E402  (unmodified reference wrapper over oracle/pyshim/{env,r}.py)
import env as shim_env  # noqa: E402

CLASSES = [envi.Env, envi.EnvComplicated, envi.EnvCooperation, envi.EnvCooperationSimplify]
NGAMES, SEED = 12, 20260101
shim_env.reset_deal_stream(0)
e = envi.EnvCooperation(seed=None)
rec = {k: [] for k in ("game", "role", "n_legal", "choice", "r", "done", "cat", "left", "taken")}
faces = [[] for _ in CLASSES]
legal_rows, action_tensors, hist_rows, recent_rows, perms = [], [], [], [], []
gstep = 0
for g in range(NGAMES):
    perms.append(np.random.Generator(np.random.PCG64(SEED + g)).permutation(54).astype(np.int8))
    e.reset()
    e.prepare()
    done, t = False, 0
    while not done:
        role = e.get_role_ID() - 1
        for ci, cls in enumerate(CLASSES):
            faces[ci].append(cls.face.fget(e).numpy().copy())
        acts = e.valid_actions(tensor=False)
        at = e.valid_actions(tensor=True).numpy()
        k = O.philox(SEED, g, t) % len(acts)
        legal_rows.extend(acts)
        action_tensors.append(at.copy())
        r, done, cat = e.step_manual(at[k])
        rec["game"].append(g); rec["role"].append(role); rec["n_legal"].append(len(acts)); rec["choice"].append(k)
        rec["r"].append(r); rec["done"].append(int(done)); rec["cat"].append(cat)
        rec["left"].append(np.array(e.left).copy()); rec["taken"].append(np.array(e.taken).copy())
        hist_rows.append(np.stack([e.history[q] for q in range(3)]))
        recent_rows.append(np.stack([e.recent_handout[q] for q in range(3)]))
        t += 1
    # the face is still queried after the terminal move (game.py:121-122)
    for ci, cls in enumerate(CLASSES):
        faces[ci].append(cls.face.fget(e).numpy().copy())
save("envi_trace.npz", perms=np.array(perms), seed=np.int64(SEED),
     legal=np.array(legal_rows, np.int8), actions_f32=np.concatenate(action_tensors).astype(np.float32),
     face_first=np.array(faces[0]), face_complicated=np.array(faces[1]),
     face_cooperation=np.array(faces[2]), face_simplify=np.array(faces[3]),
     history=np.array(hist_rows, np.int8), recent=np.array(recent_rows, np.int8),
     **{k: np.array(v) for k, v in rec.items()})

# ------------------------------------------------------------------ server/core.py payload path
import ast  # noqa: E402
import server.core as core  # noqa: E402

pred = core.Predictor.__new__(core.Predictor)      # __init__ would load absent checkpoints; face/valid_actions need only this
pred.mock_env = envi.Env(seed=0)
payloads = []
tree = ast.parse(open(os.path.join(REF, "server", "client.py"), encoding="utf-8").read())
for node in tree.body:
    if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict) and node.targets[0].id.startswith("payload"):
        payloads.append(ast.literal_eval(node.value))
n_client = len(payloads)
# positions cut from random games played on the oracle-backed envi.Env
shim_env.reset_deal_stream(1000)
e = envi.EnvCooperationSimplify()
prng = np.random.default_rng(99)
for g in range(30):
    e.reset(); e.prepare()
    done, hist_cards, last_cards = False, {0: [], 1: [], 2: []}, {0: [], 1: [], 2: []}
    while not done:
        role = e.get_role_ID() - 1
        if prng.random() < 0.35:
            payloads.append({"role_id": role, "cur_cards": [int(c) for c in e.get_curr_handcards()],
                             "history": {q: list(hist_cards[q]) for q in range(3)},
                             "last_taken": {q: list(last_cards[q]) for q in range(3)},
                             "left": {q: int(e.left[q]) for q in range(3)}})
        acts = e.valid_actions(tensor=False)
        a = acts[prng.integers(len(acts))]
        cards = [int(c) for c in envi.Env.arr2cards(a)]
        hist_cards[role] += cards
        last_cards[role] = cards
        _, done, _ = e.step_manual(envi.Env.batch_arr2onehot([a])[0])
faces, acts_rows, acts_off = [], [], [0]
for p in payloads:
    faces.append(pred.face(**p).numpy())
    _, at = pred.valid_actions(**p)
    acts_rows.append(at.numpy())
    acts_off.append(acts_off[-1] + at.shape[0])
sys.path.insert(0, os.path.join(ROOT))
def _cnt(cards):
    c = np.zeros(15, np.int8)
    for v in cards:
        c[int(v) - 3] += 1
    return c
save("core_payloads.npz", n_client=np.int64(n_client),
     role=np.array([p["role_id"] for p in payloads], np.int8),
     hand=np.array([_cnt(p["cur_cards"]) for p in payloads]),
     history=np.array([[_cnt(p["history"][q]) for q in range(3)] for p in payloads]),
     last_taken=np.array([[_cnt(p["last_taken"][q]) for q in range(3)] for p in payloads]),
     left=np.array([[p["left"][q] for q in range(3)] for p in payloads], np.int8),
     face=np.array(faces, np.float32), actions=np.concatenate(acts_rows).astype(np.float32),
     actions_off=np.array(acts_off, np.int64))

# ------------------------------------------------------------------ game.py transition assembly
import tempfile  # noqa: E402
import config as ref_config  # noqa: E402
_tmp = tempfile.mkdtemp()
ref_config.LOG_DIR, ref_config.WIN_DIR, ref_config.MODEL_DIR = (os.path.join(_tmp, d) for d in ("logs", "win", "models"))
import game as ref_game  # noqa: E402

TSEED = 4711
STEP = [0]          # env-steps played so far in the current game
GAME = [0]
RECORDS = []


class _StubNet:
    def load(self, *a): pass
    def save(self, *a): pass


class StubDQN:
    """stands in for dqn.DQNFirst: picks legal move number philox(TSEED, game, 2*step + kind) % N, records perceive()"""
    def __init__(self, net_cls):
        self.policy_net, self.target_net, self.epsilon = _StubNet(), _StubNet(), 0.0
        self.is_lord = net_cls == "lord"

    def _pick(self, actions, kind):
        return actions[O.philox(TSEED, GAME[0], 2 * STEP[0] + kind) % actions.shape[0]]

    def e_greedy_action(self, face, actions):          # the lord's move (it is the only seat in training)
        return self._pick(actions, 0)

    def greedy_action(self, face, actions):            # farmers: their move; lord: the a1 of Game.feedback (game.py:125)
        return self._pick(actions, 1 if self.is_lord else 0)

    def perceive(self, s0, a0, r, s1, a1, done):
        RECORDS.append((GAME[0], s0.numpy().copy(), a0.numpy().copy(), float(r), s1.numpy().copy(), a1.numpy().copy(), bool(done)))
        return None

    def update_epsilon(self, episode): pass
    def update_target(self, episode): pass


class CountingEnv(envi.EnvCooperation):
    def step_manual(self, onehot_cards):
        out = super().step_manual(onehot_cards)
        STEP[0] += 1
        return out


NG_T = 40
shim_env.reset_deal_stream(5000)
gm = ref_game.Game(CountingEnv, {"lord": "lord", "down": "down", "up": "up"}, {"lord": StubDQN, "down": StubDQN, "up": StubDQN},
                   train_dict={"lord": True, "down": False, "up": False})
tperms = []
for g in range(NG_T):
    GAME[0], STEP[0] = g, 0
    tperms.append(np.random.Generator(np.random.PCG64(SEED + 5000 + g)).permutation(54).astype(np.int8))
    gm.play()
save("game_transitions.npz", seed=np.int64(TSEED), perms=np.array(tperms),
     game=np.array([r[0] for r in RECORDS], np.int32), s0=np.array([r[1] for r in RECORDS], np.float32),
     a0=np.array([r[2] for r in RECORDS], np.float32), r=np.array([r[3] for r in RECORDS], np.float32),
     s1=np.array([r[4] for r in RECORDS], np.float32), a1=np.array([r[5] for r in RECORDS], np.float32),
     done=np.array([r[6] for r in RECORDS], np.uint8), lord_wins=np.int64(gm.lord_total_wins),
     farmer_wins=np.int64(gm.up_total_wins + gm.down_total_wins))

# ------------------------------------------------------------------ converters
arrs = np.array([rand_hand(int(rng.integers(0, 21))) for _ in range(64)], np.int8)
cards = [envi.Env.arr2cards(a) for a in arrs]
assert all(np.array_equal(envi.Env.cards2arr(c), a) for c, a in zip(cards, arrs))
onehot = envi.Env.batch_arr2onehot(arrs)
back = np.array([envi.Env.onehot2arr(o) for o in onehot])
assert np.array_equal(back, arrs)
save("converters.npz", arrs=arrs, cards_flat=np.concatenate(cards).astype(np.int8),
     cards_off=np.cumsum([0] + [len(c) for c in cards]).astype(np.int32), onehot=onehot.astype(np.int8))
print("done")
