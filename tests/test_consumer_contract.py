"""The consumer of the env's tensors: the reference's Q-networks, net.py:65-139.  The GPU box has no
reference tree, so the shim tests and bench.py --config 3 score actions with tests/qnet_like.py; here that stand-in is
pinned to the UNMODIFIED net.py classes through tests/golden/net_forward.npz (make_net_golden.py: the stand-in's seeded
weights loaded into NetComplicated / NetMoreComplicated / NetCooperation / NetCooperationSimplify, their forward stored)."""
import numpy as np
import pytest
import torch

from qnet_like import QNetLike, golden_inputs


def _cases(golden):
    g = golden.net_forward
    return g, [(str(n), int(c), k) for k, (n, c) in enumerate(zip(g["names"], g["channels"]))]


def test_qnet_like_computes_what_net_py_computes(golden):
    g, cases = _cases(golden)
    assert [n for n, _, _ in cases] == ["NetComplicated", "NetMoreComplicated", "NetCooperation", "NetCooperationSimplify"]
    torch.set_num_threads(1)
    for name, C, k in cases:
        like = QNetLike(C, 256, 256, seed=100 + k).eval()
        face, actions = golden_inputs(C, 48, seed=200 + k)
        with torch.no_grad():
            batch = like(face, actions).numpy()
            single = like(face[0], actions).numpy()
            fused = like.forward_state_action(torch.cat((face, actions.unsqueeze(1)), dim=1)).numpy()
        assert batch.shape == (48, 1)
        np.testing.assert_allclose(batch, g[name + "_batch"], rtol=0, atol=2e-6, err_msg=name)
        np.testing.assert_allclose(single, g[name + "_single"], rtol=0, atol=2e-6, err_msg=name)
        assert np.array_equal(fused, batch)              # the shim's in-place input is the same tensor net.py:90 builds


@pytest.mark.gpu
def test_qnet_like_on_the_gpu_matches_net_py_golden(golden):
    """fp32 (TF32 off), tolerance 1e-4 absolute on outputs of magnitude ~0.1-1: cuDNN's summation order differs from the CPU's"""
    g, cases = _cases(golden)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for name, C, k in cases:
            like = QNetLike(C, 256, 256, seed=100 + k).eval().cuda()
            face, actions = golden_inputs(C, 48, seed=200 + k)
            with torch.no_grad():
                batch = like(face.cuda(), actions.cuda()).cpu().numpy()
            np.testing.assert_allclose(batch, g[name + "_batch"], rtol=0, atol=1e-4, err_msg=name)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
