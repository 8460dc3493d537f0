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
    def __init__(self, net, epsilon=0.0, seed=0, chunk_actions=1 << 17, fused=False, precision="fp32"):
        """fused=True: score with FusedQScorer (below) -- for the NetComplicated family of net.py; call
        `policy.scorer.refresh()` after the network's weights change."""
        self.net, self.epsilon, self.seed, self.chunk = net, float(epsilon), int(seed), int(chunk_actions)
        self.fused, self.precision, self.scorer = bool(fused), precision, None

    @torch.no_grad()
    def q_values(self, env, env_mask=None):
        """float32 [sumN]: Q(face[env of move i], move i) for every legal move of every env.  env_mask (bool [B]): score
        only the moves of those envs (e.g. the ones where it is the learning role's turn); the others get 0."""
        if self.fused:
            if self.scorer is None:
                self.scorer = FusedQScorer(self.net, env.C, precision=self.precision, device=env.device)
            return self.scorer.q_values(env, env_mask)
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
        q = q.contiguous() if q.numel() else torch.zeros(1, dtype=torch.float32, device=q.device)   # every env finished: no moves
        with torch.cuda.device(q.device):
            N.check(N.lib.ddz_select_actions(q.data_ptr(), env.offsets.data_ptr(), self.epsilon, self.seed, env.env0,
                                             env._stepno if stepno is None else int(stepno), choice.data_ptr(), env.B,
                                             torch.cuda.current_stream(q.device).cuda_stream), "ddz_select_actions")
        return choice

    def act(self, env, env_mask=None):
        return self.select(env, self.q_values(env, env_mask))


# ------------------------------------------------------------------------------------------------
# Q-scoring without the [n, C+1, 15, 4] tensor (second half)
# ------------------------------------------------------------------------------------------------
def _thermometer_lut():
    """row values by nibble, as the face encoder writes them (csrc/ddz_kernels.cu lut_entry): 0..4 = count, ones
    left-aligned; 8 + count = the same ones right-aligned (probability planes of the form-B build)"""
    lut = torch.zeros(16, 4)
    for c in range(5):
        lut[c, :c] = 1.0
        lut[8 + c, 4 - c:] = 1.0
    return lut


def q_tables(rank_convs, line_conv, dtype=torch.float32):
    """The convolutions of net.py:65-139 (NetComplicated family) as lookup tables over what the env already holds.

    Every plane of the network's input is a row of four 0/1 slots per rank that is a function of ONE nibble (a count);
    conv_k is a (1,k) kernel with stride (1,4), i.e. exactly one output per rank that sees slots 0..k-1 of that rank
    (net.py:69-72), so  conv_k(x)[o, r] = bias_k[o] + sum_c scale_c * T[c][nibble_c(r)][k][o]  with
    T[c][v][o][k] = sum_{j<k} W_k[o, c, 0, j] * lut[v][j];  conv_shunzi (15,1) is  bias[o] + sum_{c,r} scale_c *
    L[c][r][o] * lut[nibble_c(r)][j]  for each slot j.  Returns (T [C+1,16,W,4], rank_bias [W,4], L [C+1,15,W], line_bias [W])."""
    dev = rank_convs[0].weight.device                          # built where the weights live (a refresh after every update)
    lut = _thermometer_lut().to(dev, torch.float64)
    W = rank_convs[0].weight.shape[0]
    cin = rank_convs[0].weight.shape[1]
    T = torch.zeros(cin, 16, W, 4, dtype=torch.float64, device=dev)
    bias = torch.zeros(W, 4, dtype=torch.float64, device=dev)
    for k, conv in enumerate(rank_convs):                      # kernel (1, k+1)
        w = conv.weight.detach().to(torch.float64)             # [W, cin, 1, k+1]
        assert w.shape[2] == 1 and w.shape[3] == k + 1 and tuple(conv.stride) == (1, 4)
        T[:, :, :, k] = torch.einsum("ocj,vj->cvo", w[:, :, 0, :], lut[:, :k + 1])
        bias[:, k] = conv.bias.detach().to(torch.float64)
    lw = line_conv.weight.detach().to(torch.float64)           # [W, cin, 15, 1]
    assert lw.shape[2] == 15 and lw.shape[3] == 1
    L = lw[:, :, :, 0].permute(1, 2, 0).contiguous()           # [cin, 15, W]
    lb = line_conv.bias.detach().to(torch.float64)
    return T.to(dtype).contiguous(), bias.to(dtype).contiguous(), L.to(dtype).contiguous(), lb.to(dtype).contiguous()


def net_parts(net):
    """(rank_convs, line_conv, fc1, fc2) of a NetComplicated-family module (net.py:65-139: conv1..conv4, conv_shunzi,
    fc1, fc2) or of anything that exposes the same layers as rank_convs / line_conv"""
    if hasattr(net, "rank_convs"):
        return list(net.rank_convs), net.line_conv, net.fc1, net.fc2
    if all(hasattr(net, a) for a in ("conv1", "conv2", "conv3", "conv4", "conv_shunzi", "fc1", "fc2")):
        return [net.conv1, net.conv2, net.conv3, net.conv4], net.conv_shunzi, net.fc1, net.fc2
    raise TypeError("not a NetComplicated-shaped network: need conv1..conv4, conv_shunzi, fc1, fc2 (net.py:65-139)")


class FusedQScorer:
    """net(face, actions) of the NetComplicated family (net.py:65-139) for every legal move of a batched env, without the
    network's input tensor: ddz_q_features computes the first layer (four rank convolutions + max-pool + the line
    convolution, net.py:91-97) from the packed state and move lists through the tables of q_tables(), as rows of the matrix
    fc1 multiplies (rank-major; fc1's columns are permuted to match once per refresh); fc1 -> ReLU -> fc2 (net.py:99-101, dropout is the identity in eval mode) are two library GEMMs.
    precision: "fp32" (default: float32 rows, TF32 off), "tf32", "bf16" (rows written as bfloat16, fc1 in bfloat16)."""

    def __init__(self, net, face_channels, precision="fp32", device=None, chunk_rows=1 << 18):
        if precision not in ("fp32", "tf32", "bf16"):
            raise ValueError("precision must be fp32, tf32 or bf16")
        if not torch.cuda.is_available():
            raise N.DdzError("FusedQScorer needs a CUDA device")
        self.net, self.C, self.precision, self.chunk_rows = net, int(face_channels), precision, int(chunk_rows)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._feat = None
        self.refresh()

    def refresh(self):
        """rebuild the tables from the network's current weights"""
        convs, line, fc1, fc2 = net_parts(self.net)
        if convs[0].weight.shape[1] != self.C + 1:
            raise ValueError("the network takes %d input planes, the env's face has %d + 1" % (convs[0].weight.shape[1], self.C))
        T, rb, L, lb = q_tables(convs, line)
        dev, dt = self.device, (torch.bfloat16 if self.precision == "bf16" else torch.float32)
        self.T, self.rank_bias, self.L, self.line_bias = T.to(dev), rb.to(dev), L.to(dev), lb.to(dev)
        self.W = int(T.shape[2])
        if self.W > 256 or self.W % 4:
            raise ValueError("ddz_q_features handles widths up to 256 that are multiples of 4 (net.py: 256)")
        if fc1.weight.shape[1] != 19 * self.W:
            raise ValueError("fc1 must take 19 * width inputs (net.py:77,138)")
        # ddz_q_features lays a row out rank-major, [15][W] | [W][4]; net.py:93 flattens [W][15] | [W][4]: permute fc1's columns
        w1 = fc1.weight.detach()
        H, W = w1.shape[0], self.W
        w1 = torch.cat((w1[:, :15 * W].reshape(H, W, 15).permute(0, 2, 1).reshape(H, 15 * W), w1[:, 15 * W:]), dim=1)
        self.w1t = w1.t().contiguous().to(dev, dt)
        self.b1 = fc1.bias.detach().to(dev, dt)
        self.w2 = fc2.weight.detach().reshape(-1).to(dev, torch.float32)
        self.b2 = fc2.bias.detach().to(dev, torch.float32)

    @torch.no_grad()
    def q_values(self, env, env_mask=None):
        """float32 [sumN]; moves of envs outside env_mask (bool [B]) get 0"""
        env._ensure()
        dev = self.device
        off = env._offsets[env._cur]
        n_all = env.num_actions
        q = torch.zeros(n_all, dtype=torch.float32, device=dev)
        if env_mask is None:
            dst, rows, mask8 = off, None, None
        else:
            mask = env_mask.to(dev, torch.bool)
            cnt = (off[1:] - off[:-1]) * mask
            dst = torch.zeros(env.B + 1, dtype=torch.int32, device=dev)
            dst[1:] = torch.cumsum(cnt, 0)
            rows = torch.repeat_interleave(mask, (off[1:] - off[:-1]).to(torch.int64), output_size=n_all).nonzero(as_tuple=True)[0]
            mask8 = mask.to(torch.uint8)
        dst_h = dst.cpu().numpy()
        total = int(dst_h[-1])
        if total == 0:
            return q
        bf16 = self.precision == "bf16"
        cap = min(total, self.chunk_rows + N.MAX_LEGAL)
        if self._feat is None or self._feat.shape[0] < cap:
            self._feat = torch.empty((cap, 19 * self.W), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.precision == "tf32"
        try:
            b_lo = 0
            while b_lo < env.B:
                lo = int(dst_h[b_lo])
                b_hi = int(min(env.B, max(b_lo + 1, dst_h.searchsorted(lo + self.chunk_rows, side="right") - 1)))
                hi = int(dst_h[b_hi])
                if hi > lo:
                    feat = self._feat[:hi - lo]
                    with torch.cuda.device(dev):
                        N.check(N.lib.ddz_q_features(env._p(env._state), env.VARIANT, env._p(off),
                                                     env._p(env._actions_u64[env._cur]), env._p(mask8),
                                                     None if env_mask is None else dst.data_ptr(), b_lo, b_hi - b_lo, lo,
                                                     self.T.data_ptr(), self.rank_bias.data_ptr(), self.L.data_ptr(),
                                                     self.line_bias.data_ptr(), self.W, feat.data_ptr(), int(bf16), env.B,
                                                     env._stream()), "ddz_q_features")
                    h = torch.addmm(self.b1, feat, self.w1t).relu_()
                    out = torch.mv(h.to(torch.float32), self.w2) + self.b2
                    if rows is None:
                        q[lo:hi] = out
                    else:
                        q[rows[lo:hi]] = out
                b_lo = b_hi
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        return q
