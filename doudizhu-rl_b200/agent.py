"""Batched Q-scoring shim for the reference's agent.

Reference: DQNFirst.greedy_action / e_greedy_action (dqn.py:50-71) score the legal moves of ONE decision with
`net(face, actions)` (net.py:81-102, which accepts a pre-batched face [N,C,15,4]) and take the argmax (or, with
probability epsilon, a uniform random move).  Here the same is done for all B envs at once: the network runs over the
CSR action list in chunks, and a hand-written kernel (ddz_select_actions) does the segmented argmax / epsilon draw.
The network itself is the caller's torch module -- the consumer contract, not part of this library.  A network that
exposes `forward_state_action(x)` for the already concatenated input x [n, C+1, 15, 4] (what net.py:90 builds with
`face.repeat` + `torch.cat`) gets that tensor written directly by ddz_encode_state_actions.
"""
import torch

from . import _native as N


class BatchedGreedyPolicy:
    def __init__(self, net, epsilon=0.0, seed=0, chunk_actions=1 << 17):
        self.net, self.epsilon, self.seed, self.chunk = net, float(epsilon), int(seed), int(chunk_actions)

    @torch.no_grad()
    def q_values(self, env, env_mask=None):
        """float32 [sumN]: Q(face[env of move i], move i) for every legal move of every env.  env_mask (bool [B]): score
        only the moves of those envs (e.g. the ones where it is the learning role's turn); the others get 0."""
        if hasattr(self.net, "forward_state_action") and hasattr(env, "state_actions"):
            # the network takes the concatenated [n, C+1, 15, 4] input: one kernel writes it in place (no gather, no cat)
            x, rows = env.state_actions(env_mask)
            q = torch.zeros(env.num_actions, dtype=torch.float32, device=x.device)
            for lo in range(0, x.shape[0], self.chunk):
                out = self.net.forward_state_action(x[lo:lo + self.chunk]).reshape(-1).to(torch.float32)
                if rows is None:
                    q[lo:lo + out.numel()] = out
                else:
                    q[rows[lo:lo + out.numel()]] = out
            return q
        face = env.face
        actions, offsets = env.valid_actions()
        n = actions.shape[0]
        counts = (offsets[1:] - offsets[:-1]).to(torch.int64)
        owner = torch.repeat_interleave(torch.arange(env.B, device=face.device), counts, output_size=n)
        q = torch.zeros(n, dtype=torch.float32, device=face.device)
        rows = torch.arange(n, device=face.device) if env_mask is None else env_mask[owner].nonzero(as_tuple=True)[0]
        for lo in range(0, rows.numel(), self.chunk):
            sel = rows[lo:lo + self.chunk]
            q[sel] = self.net(face[owner[sel]], actions[sel]).reshape(-1).to(torch.float32)
        return q

    def select(self, env, q, stepno=None):
        """int32 [B] index into each env's legal list (dqn.py:60,70), -1 for finished envs."""
        choice = torch.empty(env.B, dtype=torch.int32, device=q.device)
        q = q.contiguous()
        with torch.cuda.device(q.device):
            N.check(N.lib.ddz_select_actions(q.data_ptr(), env.offsets.data_ptr(), self.epsilon, self.seed, env.env0,
                                             env._stepno if stepno is None else int(stepno), choice.data_ptr(), env.B,
                                             torch.cuda.current_stream(q.device).cuda_stream), "ddz_select_actions")
        return choice

    def act(self, env, env_mask=None):
        return self.select(env, self.q_values(env, env_mask))
