"""UCT search on the batched env: the reference's search bot, `mcts(payload)`
(server/mcts/interface.py:15-45), with the device doing what costs time.

The reference algorithm, kept as it is:
  * `computation_budget` iterations of  tree policy -> default policy -> back-up  (interface.py:37-41);
  * tree policy (tree_policy.py:4-13): walk down while the node is fully expanded, choosing the child with the highest
    UCB value when the node's player is the searching player and the LOWEST otherwise (get_bestchild.py:4-33,
    value = reward/visit + 0.7 * sqrt(2 ln(parent visits) / visit)); at the first node with an untried move, expand ONE
    untried move chosen at random (tree.py:28-52) and stop there;
  * default policy (default_policy.py:4-10): uniformly random legal moves until somebody is out of cards; the reward is 1
    when the searching player's side has won (tree.py:71-81);
  * back-up (backup.py:1-5): visit += 1, reward += reward on the path to the root;
  * answer: the root child with the best reward / visit (get_bestchild_.py:35-42).

  * both the tree and the default policy take their moves from the bot's OWN move list, mcts.get_moves.get_moves
    (get_moves.py:36-69; tree.py:33,86): a list of more than 10 moves loses its rocket-kicker moves, is ranked by
    cards_value - 0.1 * cards left (evaluator.py:17-57) and keeps its lowest- and highest-valued thirds.  `prune=True`
    (the default) does the same; `prune=False` searches and plays over the full r.get_moves lists.

What runs where: a node's move list comes from ddz_mcts_moves (ddz_legal_moves without pruning), the default policy is
ddz_playout_pruned (ddz_playout) -- `width` random playouts of the expanded node in one launch (Philox stream keyed by the
iteration), backed up as `width` visits; width = 1 is the reference's sequential search.  The tree itself (a few thousand
nodes, pointer chasing) stays on the host.
Exact ties are broken at random among the tied candidates, as the reference does with np.random.choice
(get_bestchild.py:8,17,37 -- with one visit per child most of a young node's children tie), and the untried move to
expand is drawn at random (tree.py:42); both draws come from ONE seeded numpy generator, so with the same seed the search
is reproducible, and the same algorithm written over the CPU restatement of the env (tests/uct_reference_algorithm.py)
picks the same move.
"""
import math

import numpy as np
import torch

from . import _native as N
from .deals import pack_counts_np
from .env import BatchedEnv, MoveGenerator

UCB_C = 0.7                      # get_bestchild.py:4


class _Node:
    __slots__ = ("parent", "children", "reward", "visit", "hands", "recent", "cur", "winner", "action", "untried", "nmoves")

    def __init__(self, parent, hands, recent, cur, winner, action):
        self.parent, self.children, self.reward, self.visit = parent, [], 0.0, 0
        self.hands, self.recent, self.cur, self.winner, self.action = hands, recent, cur, winner, action
        self.untried, self.nmoves = None, None


def _trick(recent, cur):
    """the hand-out to beat: previous player's, else the one before (envi.py:103-110)"""
    p = recent[(cur + 2) % 3]
    return p if p.any() else recent[(cur + 1) % 3]


class UctSearch:
    def __init__(self, role, hands, last_taken, width=1, c=UCB_C, seed=1, device=None, prune=True):
        if not torch.cuda.is_available():
            raise N.DdzError("UctSearch needs a CUDA device")
        self.me = int(role)
        self.width, self.c, self.seed, self.prune = int(width), float(c), int(seed), bool(prune)
        self.rng = np.random.Generator(np.random.PCG64(self.seed))
        self.env = BatchedEnv(self.width, seed=self.seed if self.seed else None, device=device)
        self.env.seed = self.seed
        self.dev = self.env.device
        self.gen = MoveGenerator(1, device=self.dev)
        self._pair = torch.zeros(2, dtype=torch.int64).pin_memory()
        self._pair_d = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self._list_d = torch.zeros(N.MCTS_MAX_MOVES + 1, dtype=torch.int64, device=self.dev)     # pruned list | its length
        self._list_h = torch.zeros(N.MCTS_MAX_MOVES + 1, dtype=torch.int64).pin_memory()
        self._state_h = torch.zeros(10, dtype=torch.int64).pin_memory()
        self._state_d = torch.zeros(10, dtype=torch.int64, device=self.dev)
        hands = np.asarray(hands, np.int64).reshape(3, 15).copy()
        recent = np.asarray(last_taken, np.int64).reshape(3, 15).copy()
        self.root = _Node(None, hands, recent, self.me, -1, None)
        self.iterations = 0
        self.playouts = 0

    # ------------------------------------------------------------------ device work
    def _legal_moves(self, node):
        """int64 [n,15]: the move list of the player to move against the trick to beat -- the bot's pruned list in the
        reference's order, or r.get_moves in canonical order"""
        self._pair[0] = int(pack_counts_np(node.hands[node.cur]).astype(np.int64))
        self._pair[1] = int(pack_counts_np(_trick(node.recent, node.cur)).astype(np.int64))
        self._pair_d.copy_(self._pair, non_blocking=True)
        if self.prune:
            count = self._list_d[N.MCTS_MAX_MOVES:].view(torch.int32)
            with torch.cuda.device(self.dev):
                N.check(N.lib.ddz_mcts_moves(self._pair_d[0:1].data_ptr(), self._pair_d[1:2].data_ptr(),
                                             self._list_d.data_ptr(), count.data_ptr(), 1,
                                             torch.cuda.current_stream(self.dev).cuda_stream), "ddz_mcts_moves")
            self._list_h.copy_(self._list_d, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            n = int(self._list_h[N.MCTS_MAX_MOVES:].view(torch.int32)[0])
            packed = self._list_h[:n].numpy().view(np.uint64).copy()
        else:
            acts, offs = self.gen.generate(self._pair_d[0:1], self._pair_d[1:2])
            n = int(offs[1].item())
            packed = acts[:n].cpu().numpy().view(np.uint64)
        return ((packed[:, None] >> (np.arange(15, dtype=np.uint64) * np.uint64(4))) & np.uint64(15)).astype(np.int64)

    def _default_policy(self, node, it):
        """`width` random playouts of `node` in one launch; returns the number the searching side won"""
        if node.winner >= 0:
            return float(self.width) * self._won(node.winner)
        env, W = self.env, self.width
        self._state_h[0:3] = torch.as_tensor(pack_counts_np(node.hands).astype(np.int64))
        self._state_h[3:6] = 0
        self._state_h[6:9] = torch.as_tensor(pack_counts_np(node.recent).astype(np.int64))
        self._state_h[9] = node.cur
        self._state_d.copy_(self._state_h, non_blocking=True)
        f, meta = env._fields()
        f.copy_(self._state_d[0:9, None].expand(9, W))
        meta.copy_(self._state_d[9:10].expand(W).to(torch.int32))
        env._fresh = False
        env.env0, env._stepno = it * W, 0            # Philox counters of this iteration's playouts
        env.playout(max_steps=400, policy="search" if self.prune else "uniform")
        winner = (env._fields()[1] >> 3) & 3
        lord_won = int((winner == 1).sum().item())
        self.playouts += W
        return float(lord_won if self.me == 1 else W - lord_won)

    def _won(self, winner):
        return 1.0 if (winner == 1) == (self.me == 1) else 0.0          # tree.py:71-81: the farmers win together

    # ------------------------------------------------------------------ the tree (host)
    def _expand(self, node):
        i = int(self.rng.integers(len(node.untried)))
        mv = node.untried.pop(i)
        hands, recent = node.hands.copy(), node.recent.copy()
        hands[node.cur] -= mv
        recent[node.cur] = mv
        winner = node.cur if mv.any() and not hands[node.cur].any() else -1
        child = _Node(node, hands, recent, (node.cur + 1) % 3, winner, mv)
        node.children.append(child)
        return child

    def _best_child(self, node):
        visit = np.array([ch.visit for ch in node.children], np.float64)
        reward = np.array([ch.reward for ch in node.children], np.float64)
        values = reward / visit + self.c * np.sqrt(2.0 * math.log(node.visit) / visit)
        return node.children[self._pick(values, values.max() if node.cur == self.me else values.min())]

    def _pick(self, values, target):
        """index of the entry equal to `target`; a random one of them when several tie (get_bestchild.py:8-17)"""
        tied = np.flatnonzero(values == target)
        return int(tied[0] if len(tied) == 1 else tied[self.rng.integers(len(tied))])

    def _tree_policy(self):
        node = self.root
        while node.winner < 0:
            if node.untried is None:
                moves = self._legal_moves(node)
                node.untried, node.nmoves = [m for m in moves], len(moves)
            if len(node.children) < node.nmoves:
                return self._expand(node)
            node = self._best_child(node)
        return node

    def run(self, computation_budget=1000):
        for _ in range(int(computation_budget)):
            node = self._tree_policy()
            reward = self._default_policy(node, self.iterations)
            while node is not None:                      # backup.py
                node.visit += self.width
                node.reward += reward
                node = node.parent
            self.iterations += 1
        return self

    def root_table(self):
        """(moves int64 [n,15], visits [n], win rate [n]) of the root's children in expansion order"""
        ch = self.root.children
        return (np.stack([c.action for c in ch]), np.array([c.visit for c in ch]),
                np.array([c.reward / c.visit for c in ch]))

    def best_move(self):
        """get_bestchild_: the root child with the best reward / visit (a random one of them on ties, get_bestchild.py:35-42);
        counts int64 [15].  Consumes the generator when there is a tie: call it once."""
        moves, _, rate = self.root_table()
        return moves[self._pick(rate, rate.max())]
