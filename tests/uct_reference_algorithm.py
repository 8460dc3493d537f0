"""The reference's UCT search (server/mcts/interface.py:37-45, tree.py, tree_policy.py, default_policy.py, backup.py,
get_bestchild.py) restated over the CPU oracle, for tests: the oracle supplies r.get_moves and plays the random playouts
(Philox index stream, the one ddz_playout uses).  With prune=True both the tree and the playouts draw from the bot's pruned
move list (oracle.mcts_moves, pinned against the unmodified server/mcts/get_moves.py), as the reference does.
TEST INFRASTRUCTURE: the product's search (doudizhu-rl_b200/search.py) must pick the same move with the same seed and
budget."""
import math

import numpy as np


class Node:
    def __init__(self, parent, state):
        self.parent, self.state = parent, state
        self.children, self.reward, self.visit = [], 0.0, 0
        self.untried = None


class State:
    def __init__(self, hands, recent, player, winner, action):
        self.hands, self.recent, self.player, self.winner, self.action = hands, recent, player, winner, action

    def last_move(self):
        prev = self.recent[(self.player + 2) % 3]
        return prev if prev.any() else self.recent[(self.player + 1) % 3]

    def play(self, move):
        hands, recent = self.hands.copy(), self.recent.copy()
        hands[self.player] -= move
        recent[self.player] = move
        winner = self.player if move.any() and not hands[self.player].any() else -1
        return State(hands, recent, (self.player + 1) % 3, winner, move)


def pruned_playout_step(oracle, rb, seed, env0, t):
    """one decision of every unfinished env of `rb`: entry  philox(seed, env0 + b, t) % len  of the bot's pruned move list
    (tree.py:83-91: get_moves(...) then np.random.choice)"""
    offs, acts, _, _ = rb.observe(want_f32=False, want_face=False)
    choice = np.zeros(rb.B, np.int32)
    for b in range(rb.B):
        e = rb.envs[b]
        if e["done"]:
            continue
        cur = int(e["cur"])
        prev, pp = e["recent"][(cur + 2) % 3], e["recent"][(cur + 1) % 3]
        lst = oracle.mcts_moves(e["hand"][cur], prev if prev.any() else pp)
        mv = lst[oracle.philox(seed, env0 + b, t) % len(lst)]
        full = oracle.unpack(acts[offs[b]:offs[b + 1]])
        choice[b] = int(np.flatnonzero((full == mv).all(axis=1))[0])
    rb.step(choice, mode=0)


def uct(oracle, role, hands, last_taken, budget, width=1, c=0.7, seed=1, prune=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    me = int(role)
    root = Node(None, State(np.asarray(hands, np.int64).copy(), np.asarray(last_taken, np.int64).copy(), me, -1, None))

    def won(winner):
        return 1.0 if (winner == 1) == (me == 1) else 0.0

    def pick(values, target):                                          # np.random.choice among ties (get_bestchild.py)
        tied = np.flatnonzero(values == target)
        return int(tied[0] if len(tied) == 1 else tied[rng.integers(len(tied))])

    def tree_policy(node):
        while node.state.winner == -1:
            if node.untried is None:
                st = node.state
                lst = oracle.mcts_moves(st.hands[st.player], st.last_move()) if prune else \
                    oracle.get_moves(st.hands[st.player], st.last_move(), fast=True)
                node.untried = [m.astype(np.int64) for m in lst]
                node.nmoves = len(node.untried)
            if len(node.children) < node.nmoves:                       # expand one untried move, chosen at random
                mv = node.untried.pop(int(rng.integers(len(node.untried))))
                sub = Node(node, node.state.play(mv))
                node.children.append(sub)
                return sub
            visit = np.array([n.visit for n in node.children], np.float64)
            reward = np.array([n.reward for n in node.children], np.float64)
            values = reward / visit + c * np.sqrt(2.0 * math.log(node.visit) / visit)
            node = node.children[pick(values, values.max() if node.state.player == me else values.min())]   # UCB1 / UCB2
        return node

    def default_policy(node, it):
        st = node.state
        if st.winner != -1:
            return width * won(st.winner)
        rb = oracle.RefBatch(width, 0)
        rb.envs["hand"][:] = st.hands
        rb.envs["recent"][:] = st.recent
        rb.envs["cur"] = st.player
        rb.envs["winner"] = -1
        t = 0
        while not rb.envs["done"].all():
            if prune:
                pruned_playout_step(oracle, rb, seed, it * width, t)
            else:
                rb.observe(want_f32=False, want_face=False)
                rb.step(mode=2, seed=seed, env0=it * width, step=t)    # finished envs do nothing
            t += 1
        return float(sum(won(int(w)) for w in rb.envs["winner"]))

    for it in range(budget):
        node = tree_policy(root)
        reward = default_policy(node, it)
        while node is not None:                                        # back-up
            node.visit += width
            node.reward += reward
            node = node.parent
    rate = np.array([n.reward / n.visit for n in root.children])
    best = root.children[pick(rate, rate.max())]
    return best.state.action, root
