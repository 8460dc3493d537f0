"""CPU: pin the oracle to the reference's own golden vectors (tests/golden/, made by make_golden.py from
rule_based/utils/card.py, server/rule_utils/utils.py, server/mcts/get_moves.py and the unmodified envi.py)."""
import numpy as np
import pytest


def test_universe_is_card_py_action_space_plus_24(oracle, golden):
    cnt, cat, ln, val, extra = oracle.universe()
    g = golden.card_space
    assert len(cnt) == 13551 and int(extra.sum()) == 24
    keep = extra == 0
    assert np.array_equal(cnt[keep], g["counts"])                      # card.py:34-159, same order
    assert np.array_equal(np.stack([cat, ln, val], 1)[keep], g["attrs"])  # card.py:372-527 type/len/value
    # Category2Range (card.py:31) == first/last index of each category in the 13 527 list
    c = cat[keep]
    for k, (lo, hi) in enumerate(g["cat_range"]):
        assert (c[lo:hi] == k).all() and (lo == 0 or c[lo - 1] != k) and (hi == len(c) or c[hi] != k)
    assert len(set(oracle.pack(cnt).tolist())) == 13551                # duplicate-free universe
    assert int(cnt.sum(axis=1).max()) == 20


def test_extras_are_the_24_rocket_kicker_moves(oracle, golden):
    cnt, cat, ln, val, extra = oracle.universe()
    mine = {tuple(r) for r in cnt[extra == 1].tolist()}
    ref = {tuple(r) for r in golden.rocket24["counts"].tolist()}      # server/mcts/get_moves.py:22-34
    assert mine == ref
    ls = golden.legal_sets
    for c, (t, l, v) in zip(golden.rocket24["counts"], ls["extra_attrs"]):
        idx, mc, ml, mv = oracle.classify(c)
        assert idx >= 0 and (mc, ml, mv) == (t, l, v)                  # reference to_cardgroup agrees


def test_bigger_than_matches_card_py(oracle, golden):
    g = golden.beats
    attrs = golden.card_space["attrs"]
    a, b = attrs[g["pairs"][:, 0]], attrs[g["pairs"][:, 1]]
    mine = np.array([oracle.bigger_than(x, y) for x, y in zip(a, b)], np.uint8)
    assert np.array_equal(mine, g["beats"])


@pytest.mark.parametrize("fast", [False, True])
def test_legal_sets_match_get_mask_onehot60(oracle, golden, fast):
    g = golden.legal_sets
    assert np.array_equal(g["universe"], oracle.universe()[0])
    off = g["legal_off"]
    for i, (h, last) in enumerate(zip(g["hands"], g["lasts"])):
        want = g["universe"][g["legal_idx"][off[i]:off[i + 1]]]
        got = oracle.get_moves(h, last, fast=fast)
        assert np.array_equal(got, want), (i, h, last)


def test_known_answers(oracle):
    z = np.zeros(15, np.int8)
    # 20-card hand 333444 5..A 22 *$ -> 167 lead moves; vs 9995 -> only rocket (+ pass)
    h = np.array([3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1], np.int8)
    assert h.sum() == 20
    # card.py space gives 167; the oracle additionally emits 333444+*$ (rocket-kicker extra) -> 168
    moves = oracle.get_moves(h, z)
    extra = np.array([3, 3] + [0] * 11 + [1, 1], np.int8)
    assert len(moves) == 168 and any((m == extra).all() for m in moves)
    last = np.zeros(15, np.int8); last[6] = 3; last[2] = 1   # 999 + 5
    got = oracle.get_moves(h, last)
    assert len(got) == 2 and not got[0].any() and got[1][13] == 1 and got[1][14] == 1
    worst = np.array([1, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0], np.int8)
    assert len(oracle.get_moves(worst, z)) == 497               # SURVEY.md App. B worst lead hand
    assert len(oracle.get_moves(z, z)) == 0                      # empty hand, lead: nothing
    assert oracle.get_moves(z, worst * 0 + np.eye(15, dtype=np.int8)[0]).shape == (1, 15)  # follow: pass only


def test_fast_equals_definitional_random(oracle):
    rng = np.random.default_rng(123)
    deck = np.array([i // 4 for i in range(52)] + [13, 14])
    z = np.zeros(15, np.int8)
    for it in range(4000):
        h = np.bincount(rng.permutation(deck)[:rng.integers(1, 21)], minlength=15).astype(np.int8)
        if it % 2:
            o = np.bincount(rng.permutation(deck)[:rng.integers(1, 21)], minlength=15).astype(np.int8)
            om = oracle.get_moves(o, z, fast=True)
            last = om[rng.integers(len(om))]
        else:
            last = z
        assert np.array_equal(oracle.get_moves(h, last), oracle.get_moves(h, last, fast=True))


def _replay(oracle, golden, variant, key):
    """Replay the recorded game (perms + choice stream) through the oracle's BATCHED driver and compare with what
    the unmodified envi.py produced step by step."""
    t = golden.envi_trace
    games = t["game"]
    faces = t[key]
    fi = li = ai = 0
    for g in range(len(t["perms"])):
        rb = oracle.RefBatch(1, variant)
        rb.deal(t["perms"][g][None], np.zeros(1, np.int8))
        for s in np.flatnonzero(games == g):
            off, au, af, face = rb.observe(fast=False)
            n = int(t["n_legal"][s])
            assert off[1] == n
            assert np.array_equal(oracle.unpack(au), t["legal"][li:li + n])
            assert np.array_equal(af, t["actions_f32"][ai:ai + n])
            assert np.array_equal(face[0], faces[fi]), (g, s)
            f, meta = rb.export()
            assert (meta[0] & 3) == t["role"][s]
            li += n; ai += n; fi += 1
            r, done, cat, rew = rb.step(np.array([t["choice"][s]], np.int32), mode=0)
            assert (r[0], done[0], cat[0]) == (t["r"][s], t["done"][s], t["cat"][s])
            f, meta = rb.export()
            hands = oracle.unpack(f[0:3, 0]); hist = oracle.unpack(f[3:6, 0]); rec = oracle.unpack(f[6:9, 0])
            assert np.array_equal(hands.sum(1), t["left"][s])            # envi.py:39 left[role]
            assert np.array_equal(hist, t["history"][s])                 # envi.py:41
            assert np.array_equal(rec, t["recent"][s])                   # envi.py:42-43
            assert np.array_equal(hist.sum(0), t["taken"][s])            # envi.py:40
        # terminal face (game.py:121-122)
        _, _, _, face = rb.observe()
        assert np.array_equal(face[0], faces[fi])
        fi += 1
    assert fi == len(faces)


@pytest.mark.parametrize("variant,key", [(0, "face_first"), (1, "face_complicated"),
                                         (2, "face_cooperation"), (3, "face_simplify")])
def test_envi_trace_replay(oracle, golden, variant, key):
    _replay(oracle, golden, variant, key)


def test_philox_stream_matches_trace(oracle, golden):
    t = golden.envi_trace
    seed = int(t["seed"])
    for g in range(len(t["perms"])):
        idx = np.flatnonzero(t["game"] == g)
        for k, s in enumerate(idx):
            assert oracle.philox(seed, g, k) % int(t["n_legal"][s]) == t["choice"][s]
    # Philox4x32-10 known-answer (Random123 kat_vectors: counter=0, key=0 -> 6627e8d5 ...)
    assert oracle.philox(0, 0, 0) == 0x6627E8D5


def test_converters_golden(golden):
    g = golden.converters
    arrs, onehot = g["arrs"], g["onehot"]
    therm = (np.arange(4)[None, None, :] < arrs[:, :, None]).astype(np.int8)   # envi.py:140-146
    assert np.array_equal(therm, onehot)
    for i, a in enumerate(arrs):
        cards = g["cards_flat"][g["cards_off"][i]:g["cards_off"][i + 1]]
        assert np.array_equal(np.repeat(np.arange(15) + 3, a), cards)           # envi.py:119-130


def test_deal_rejects_bad_perm(oracle):
    rb = oracle.RefBatch(1)
    bad = np.zeros((1, 54), np.int8)
    with pytest.raises(ValueError):
        rb.deal(bad)


def test_illegal_and_done_steps_are_noops(oracle):
    rng = np.random.default_rng(5)
    rb = oracle.RefBatch(2)
    rb.deal(np.stack([rng.permutation(54), rng.permutation(54)]).astype(np.int8))
    off, au, _, _ = rb.observe()
    before = rb.export()
    r, done, cat, rew = rb.step(np.array([off[1] - off[0], -1], np.int32))       # out of range both
    after = rb.export()
    assert (cat == -1).all() and (r == 0).all() and not done.any()
    assert np.array_equal(before[0], after[0]) and rb.stats[7] == 2
    assert ((after[1] >> 5) & 1).all()                                           # sticky error flag


def test_max_legal_bound_is_proved_exhaustively(oracle):
    """DDZ_MAX_LEGAL = 512: the closed-form lead count equals the enumerators, legal sets are monotone in the hand, and
    the maximum over ALL 153 009 740 twenty-card hands of one deck is 497 (SURVEY.md App. B asked for this proof)."""
    rng = np.random.default_rng(9)
    deck = np.array([i // 4 for i in range(52)] + [13, 14])
    z = np.zeros(15, np.int8)
    for _ in range(2000):
        h = np.bincount(rng.permutation(deck)[:rng.integers(1, 21)], minlength=15).astype(np.int8)
        n = oracle.count_lead_closed(h)
        assert n == len(oracle.get_moves(h, z, fast=True))
        bigger = h.copy()
        r = int(rng.integers(0, 15))
        if bigger[r] < (4 if r < 13 else 1):
            bigger[r] += 1
            assert oracle.count_lead_closed(bigger) >= n                      # monotone
    best, hand, visited = oracle.max_lead_moves_exhaustive(8)
    assert visited == 153009740 and best == 497 and hand.sum() == 20
    assert len(oracle.get_moves(hand, z)) == 497 <= oracle.MAX_LEGAL


def test_core_payload_path_matches_predictor(oracle, golden):
    """server/core.py:26-67 (Predictor.face / valid_actions on a JSON payload): the oracle, loaded with the payload's
    position, produces the same 6-channel face and the same legal-action tensor as the reference's unmodified code."""
    g = golden.core_payloads
    n = len(g["role"])
    assert int(g["n_client"]) >= 2 and n > 100
    rb = oracle.RefBatch(n, 3)
    rb.envs["cur"] = g["role"]
    rb.envs["hist"] = g["history"]
    rb.envs["recent"] = g["last_taken"]
    hands = np.zeros((n, 3, 15), np.int8)
    for b in range(n):
        for q in range(3):
            if q == g["role"][b]:
                hands[b, q] = g["hand"][b]
            else:                                            # unknown cards: only the SIZE enters the features
                left, c = int(g["left"][b, q]), np.zeros(15, np.int8)
                for r in range(15):
                    c[r] = min(4 if r < 13 else 1, left); left -= c[r]
                hands[b, q] = c
    rb.envs["hand"] = hands
    off, au, af, face = rb.observe(fast=False)
    assert np.array_equal(face, g["face"])
    assert np.array_equal(off.astype(np.int64), g["actions_off"])
    assert np.array_equal(af, g["actions"])


def test_config1_thousand_seeded_games_random_play(oracle):
    """BASELINE config 1 on the oracle: 1000 games from the documented deal stream (PCG64(20260101+g), lord_pile 0), all
    seats uniform random.  The dynamics must reproduce what SURVEY.md 3.3 / BASELINE.md 3 measured independently on the
    reference's own action space: about 61.8 decisions per game, about 5.6 legal moves per decision, 51 % passes."""
    import ddz_b200 as D
    n = 1000
    perm, lord = D.default_deals(0, n)
    rb = oracle.RefBatch(n, 0)
    rb.deal(perm, lord)
    moves = 0
    for t in range(200):
        off, _, _, _ = rb.observe(want_f32=False, want_face=False)
        moves += int(off[n])
        rb.step(mode=2, seed=20260101, env0=0, step=t)
        if rb.envs["done"].all():
            break
    st = rb.stats
    assert st[0] == n and st[1] + st[2] + st[3] == n
    per_game, nbar, passes = st[4] / n, moves / st[4], st[9] / st[4]
    assert 58 < per_game < 66 and 5.0 < nbar < 6.2 and 0.47 < passes < 0.55, (per_game, nbar, passes)


def test_prob_plane_forms_follow_their_definitions(golden):
    """get_state_prob_manual (server/core.py:26-33) in both layouts of oracle/SEMANTICS.md, on the known60 inputs the
    unmodified server/core.py built for 591 payloads: form A = thermometer(total - known), form B = deck60 - known60."""
    import ctypes as C
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    g = golden.core_payloads
    # server/core.py:26-31 get_prob: known = thermometer(all played cards + own hand), sizes of the next / next-next player
    role = g["role"].astype(np.int64)
    counts = g["history"].sum(1).astype(np.int64) + g["hand"]
    known = (np.arange(4)[None, None, :] < counts[:, :, None]).astype(np.int32).reshape(-1, 60)
    ar = np.arange(len(role))
    s1, s2 = g["left"][ar, (role + 1) % 3], g["left"][ar, (role + 2) % 3]
    deck = np.ones((15, 4), np.float32); deck[13:, 1:] = 0
    from oracle import ddz_oracle as O
    O.build()
    for name, form in (("libddz_oracle.so", 0), ("libddz_oracle_formB.so", 1)):
        L = C.CDLL(os.path.join(os.path.dirname(here), "oracle", name))
        assert L.ddz_ref_prob_form() == form
        for i in range(0, len(known), 7):
            k = (known[i].reshape(15, 4) != 0)
            if form == 0:
                u = np.clip(np.array([4] * 13 + [1, 1]) - k.sum(-1), 0, None)
                unk = (np.arange(4)[None, :] < u[:, None]).astype(np.float32)
            else:
                unk = np.clip(deck - k.astype(np.float32), 0, None)
            tot = int(s1[i]) + int(s2[i])
            p = [np.float32(s1[i]) / np.float32(tot), np.float32(s2[i]) / np.float32(tot)] if tot else [0, 0]
            want = np.concatenate([(unk * p[0]).reshape(60), (unk * p[1]).reshape(60)]).astype(np.float32)
            out = np.zeros(120, np.float32)
            kk = np.ascontiguousarray(known[i], dtype=np.int32)
            L.ddz_ref_state_prob_manual(kk.ctypes.data_as(C.POINTER(C.c_int32)), int(s1[i]), int(s2[i]), out.ctypes.data_as(C.POINTER(C.c_float)))
            assert np.array_equal(out, want), (name, i)
            if form == 0:      # the default form is what the unmodified Predictor.face produced over the oracle shim
                assert np.array_equal(out.reshape(2, 15, 4), g["face"][i][4:6]), i


def test_cards_value_matches_evaluator_py(oracle, golden):
    """server/mcts/evaluator.py:17-57, the whole table: every key of the reference's cards_value, exact."""
    g = golden.mcts_moves
    keys, vals = g["value_keys"], g["value_vals"]
    assert len(keys) == oracle.NACTIONS_CARD_PY
    got = np.array([oracle.cards_value(k) for k in keys], np.float64)
    assert np.array_equal(got, vals)
    counts, _, _, _, extra = oracle.universe()
    for c in counts[extra == 1]:
        assert oracle.cards_value(c) is None            # KeyError in the reference (get_moves.py:56 drops them first)


def test_mcts_moves_match_get_moves_py(oracle, golden):
    """server/mcts/get_moves.py:36-69 run unmodified (over oracle/pyshim/r.py) on 769 positions: the pruned list, in order."""
    g = golden.mcts_moves
    offs = g["offsets"]
    pruned = 0
    for i, (h, l) in enumerate(zip(g["hands"], g["lasts"])):
        want = g["moves"][offs[i]:offs[i + 1]]
        got = oracle.mcts_moves(h, l)
        assert np.array_equal(got, want), i
        n = int(g["full_counts"][i])
        assert len(got) == (n if n <= 10 else 2 * (n // 3 + 1))
        pruned += n > 10
    assert pruned > 100
