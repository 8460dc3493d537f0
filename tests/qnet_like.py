"""A Q-network with the reference's input/output contract (net.py:81-102: face [N,C,15,4] or [C,15,4], actions [N,15,4]
-> [N,1]) for tests of the batched Q-scoring shim.  Architecture as described in SURVEY.md section 2 (#4): the C face
channels + 1 action channel go through (1,k) convolutions with stride (1,4), k = 1..4, and a (15,1) "shunzi"
convolution, then two linear layers (width = hidden = 256 gives NetCooperation's sizes, net.py:125-139: ~3.6 MFLOP per
scored action).  Written from that description for test purposes; weights are random.  With width = hidden = 256 it
computes what the reference's NetComplicated family computes: tests/golden/net_forward.npz (make_net_golden.py) holds the
outputs of the UNMODIFIED net.py classes carrying this module's seeded weights."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class QNetLike(nn.Module):
    def __init__(self, face_channels, width=32, hidden=64, seed=None):
        super().__init__()
        if seed is not None:
            torch.manual_seed(seed)
        cin = face_channels + 1
        self.rank_convs = nn.ModuleList([nn.Conv2d(cin, width, (1, k), (1, 4)) for k in (1, 2, 3, 4)])
        self.line_conv = nn.Conv2d(cin, width, (15, 1), 1)
        self.fc1 = nn.Linear(width * (15 + 4), hidden)
        self.fc2 = nn.Linear(hidden, 1)

    def forward(self, face, actions):
        if face.dim() == 3:
            face = face.unsqueeze(0).repeat((actions.shape[0], 1, 1, 1))
        return self.forward_state_action(torch.cat((face, actions.unsqueeze(1)), dim=1))

    def forward_state_action(self, x):
        """x [N, C+1, 15, 4]: the concatenated state-action input (net.py:90)"""
        r = torch.cat([c(x) for c in self.rank_convs], -1).max(-1).values.flatten(1)      # [N, width*15]
        l = self.line_conv(x).flatten(1)                                                  # [N, width*4]
        return self.fc2(F.relu(self.fc1(torch.cat([r, l], -1))))


def load_into_reference_net(like, ref):
    """copy a QNetLike(width = hidden = 256)'s weights into one of net.py's NetComplicated-family modules"""
    with torch.no_grad():
        for k, conv in enumerate((ref.conv1, ref.conv2, ref.conv3, ref.conv4)):
            conv.weight.copy_(like.rank_convs[k].weight); conv.bias.copy_(like.rank_convs[k].bias)
        ref.conv_shunzi.weight.copy_(like.line_conv.weight); ref.conv_shunzi.bias.copy_(like.line_conv.bias)
        ref.fc1.weight.copy_(like.fc1.weight); ref.fc1.bias.copy_(like.fc1.bias)
        ref.fc2.weight.copy_(like.fc2.weight); ref.fc2.bias.copy_(like.fc2.bias)


def golden_inputs(face_channels, n, seed):
    """seeded inputs of the env's shape: 0/1 thermometer rows, the last two face planes scaled like the probability planes"""
    import numpy as np
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 5, (n, face_channels + 1, 15))
    x = (counts[..., None] > np.arange(4)).astype(np.float32)
    x[:, face_channels - 2] *= rng.random((n, 1, 1)).astype(np.float32)
    x[:, face_channels - 1] *= rng.random((n, 1, 1)).astype(np.float32)
    return torch.from_numpy(x[:, :face_channels].copy()), torch.from_numpy(x[:, face_channels].copy())
