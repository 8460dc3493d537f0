"""Batched rollout / trainer driver pieces.

Reference: Game.play / Game.step / Game.feedback (game.py:90-181) assemble one transition (s0, a0, r, s1, a1, done) per
decision of a learning role with a DELAYED feedback rule, and DQNFirst.perceive (dqn.py:21-48) pushes it into a
`deque(maxlen=20000)` and does one TD step on a random minibatch.  Here the same assembly runs for all envs at once on
the device, the replay buffer is a ring of GPU tensors, and the TD step is the reference's arithmetic.

The delayed-feedback rule for a learning role R (game.py:128-168):
  * at R's turn it observes s0 = face and plays a0;
  * the NEXT time it is R's turn (nobody finished in between) the pending (s0, a0) is closed with r = 0,
    s1 = the face R sees now, a1 = the greedy action at s1, done = 0;
  * when the game ends first, the pending (s0, a0) is closed with r = +reward if R's side won else -reward,
    s1 = the face after the terminal move (still queried, game.py:121-122), a1 = zeros, done = 1.
"""
import torch


class ReplayBuffer:
    """Ring buffer of transitions on the GPU (dqn.py:14 deque(maxlen=REPLAY_SIZE), :22-29 sample)."""

    def __init__(self, capacity, face_channels, device):
        self.capacity, self.size, self.head = int(capacity), 0, 0
        dev = torch.device(device)
        self.s0 = torch.zeros((capacity, face_channels, 15, 4), dtype=torch.float32, device=dev)
        self.s1 = torch.zeros_like(self.s0)
        self.a0 = torch.zeros((capacity, 15, 4), dtype=torch.float32, device=dev)
        self.a1 = torch.zeros_like(self.a0)
        self.r = torch.zeros((capacity, 1), dtype=torch.float32, device=dev)
        self.done = torch.zeros((capacity, 1), dtype=torch.float32, device=dev)

    def __len__(self):
        return self.size

    def append(self, s0, a0, r, s1, a1, done):
        """append n transitions (oldest are overwritten, like a bounded deque)"""
        n = s0.shape[0]
        if n == 0:
            return
        if n > self.capacity:
            s0, a0, r, s1, a1, done = (x[-self.capacity:] for x in (s0, a0, r, s1, a1, done))
            n = self.capacity
        idx = (self.head + torch.arange(n, device=self.s0.device)) % self.capacity
        self.s0[idx], self.a0[idx], self.s1[idx], self.a1[idx] = s0, a0, s1, a1
        self.r[idx] = r.reshape(n, 1).to(torch.float32)
        self.done[idx] = done.reshape(n, 1).to(torch.float32)
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.s0.device, generator=generator)
        return self.s0[idx], self.a0[idx], self.r[idx], self.s1[idx], self.a1[idx], self.done[idx]


def td_step(policy_net, target_net, optimizer, batch, gamma):
    """one minibatch update with the reference's arithmetic (dqn.py:39-47): y = r + (1 - done) * gamma * Q_target(s1, a1),
    MSE against Q_policy(s0, a0)."""
    s0, a0, r, s1, a1, done = batch
    with torch.no_grad():
        y_true = r + (1 - done) * gamma * target_net(s1, a1)
    loss = torch.nn.functional.mse_loss(policy_net(s0, a0), y_true)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()


class TransitionCollector:
    """Delayed-feedback transition assembly for one learning role over a batched env (game.py:109-168)."""

    def __init__(self, env, role=1, reward=100.0):
        self.env, self.role, self.reward = env, int(role), float(reward)
        B, C, dev = env.B, env.C, env.device
        self.s0 = torch.zeros((B, C, 15, 4), dtype=torch.float32, device=dev)
        self.a0 = torch.zeros((B, 15, 4), dtype=torch.float32, device=dev)
        self.pending = torch.zeros(B, dtype=torch.bool, device=dev)

    def on_turn(self, a0, a1):
        """Call BEFORE stepping, with the env observed.  a0 [B,15,4]: the action each env is about to play (used where it is
        the role's turn); a1 [B,15,4]: the greedy action at the current state (the reference evaluates it separately,
        game.py:125).  Returns the transitions closed by this turn: (s0, a0, r, s1, a1, done, env_index)."""
        env = self.env
        mine = ((env.get_role_ID() - 1) == self.role) & ~env.is_done
        close = mine & self.pending
        idx = close.nonzero(as_tuple=True)[0]
        face = env.face
        out = (self.s0[idx].clone(), self.a0[idx].clone(), torch.zeros(idx.numel(), device=face.device),
               face[idx].clone(), a1[idx].clone(), torch.zeros(idx.numel(), device=face.device), idx)
        m = mine.nonzero(as_tuple=True)[0]
        self.s0[m], self.a0[m] = face[m], a0[m]
        self.pending |= mine
        return out

    def on_step_done(self, done_now):
        """Call AFTER stepping WITHOUT re-deal and after env.observe(): done_now bool [B] = envs that finished with this
        step.  Closes their pending transition with the terminal reward; env.face is the face after the terminal move."""
        env = self.env
        close = done_now & self.pending
        idx = close.nonzero(as_tuple=True)[0]
        face = env.face
        winner = env.winner[idx]
        won = (winner == 1) if self.role == 1 else (winner != 1)
        r = torch.where(won, torch.full_like(won, self.reward, dtype=torch.float32),
                        torch.full_like(won, -self.reward, dtype=torch.float32))
        out = (self.s0[idx].clone(), self.a0[idx].clone(), r, face[idx].clone(),
               torch.zeros((idx.numel(), 15, 4), device=face.device), torch.ones(idx.numel(), device=face.device), idx)
        self.pending &= ~done_now
        return out
