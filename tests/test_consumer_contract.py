"""The consumer of the env's tensors: the reference's Q-networks, net.py:65-139.  The GPU box has no
reference tree, so the shim tests and bench.py --config 3 score actions with tests/qnet_like.py; here that stand-in is
pinned to the UNMODIFIED net.py classes through tests/golden/net_forward.npz (make_net_golden.py: the stand-in's seeded
weights loaded into NetComplicated / NetMoreComplicated / NetCooperation / NetCooperationSimplify, their forward stored)."""
import numpy as np
import pytest
import torch

from qnet_like import QNetLike, golden_inputs


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ddz_b200
    return ddz_b200


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


def test_q_tables_reproduce_the_first_layer(golden):
    """agent.q_tables + the lookup formulation (tests/qscore_reference.py) == the network's convolutions, max-pool and
    concatenation (net.py:91-97), and the whole forward through fc1 / fc2 == the golden outputs of the unmodified net.py"""
    from qscore_reference import features, nibbles_of
    import ddz_b200.agent as agent
    g, cases = _cases(golden)
    for name, C, k in cases:
        like = QNetLike(C, 256, 256, seed=100 + k).eval()
        face, actions = golden_inputs(C, 48, seed=200 + k)
        x = torch.cat((face, actions.unsqueeze(1)), dim=1)
        convs, line, fc1, fc2 = agent.net_parts(like)
        T, rb, L, lb = agent.q_tables(convs, line, dtype=torch.float64)
        nib, scale = nibbles_of(x.numpy())
        feat = features(nib, scale, T.numpy(), rb.numpy(), L.numpy(), lb.numpy())
        with torch.no_grad():
            r = torch.cat([c(x) for c in like.rank_convs], -1).max(-1).values.flatten(1)
            want = torch.cat([r, like.line_conv(x).flatten(1)], -1).numpy()
        np.testing.assert_allclose(feat, want, rtol=0, atol=1e-5, err_msg=name)
        h = np.maximum(feat @ fc1.weight.detach().numpy().T.astype(np.float64) + fc1.bias.detach().numpy(), 0)
        q = h @ fc2.weight.detach().numpy().T.astype(np.float64) + fc2.bias.detach().numpy()
        np.testing.assert_allclose(q, g[name + "_batch"], rtol=0, atol=1e-5, err_msg=name)


def _midgame_env(D, cls, B, steps, seed):
    perm, lord = D.random_deals(B, seed=seed, pool_games=2)
    pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
    env = getattr(D, cls)(B, seed=seed)
    env.prepare(pd, ld, pool_games=2)
    for _ in range(steps):
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=2)
    return env


@pytest.mark.gpu
@pytest.mark.parametrize("cls,C", [("BatchedEnv", 4), ("BatchedEnvComplicated", 7), ("BatchedEnvCooperation", 9),
                                   ("BatchedEnvCooperationSimplify", 6)])
def test_fused_q_scorer_equals_the_network(D, cls, C):
    """ddz_q_features + fc1 / fc2 == net(face, actions) over the env's own tensors (net.py:81-102), fp32, every legal move of
    every env, all four env classes; the masked form (only the landlord's decisions) and the chunked form agree with it.
    Tolerance: 1e-5 absolute + 1e-4 relative on float32 Q values (different summation order, no TF32 on either side)."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        env = _midgame_env(D, cls, 3000, 23, seed=3)
        width = 256 if C == 9 else 32
        net = QNetLike(C, width, 96, seed=5).eval().cuda()
        want = D.BatchedGreedyPolicy(net).q_values(env)
        fused = D.BatchedGreedyPolicy(net, fused=True)
        got = fused.q_values(env)
        assert got.shape == want.shape and got.shape[0] == env.num_actions > 3000
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        assert (D.BatchedGreedyPolicy(net).select(env, want) == fused.select(env, got)).float().mean() > 0.999
        mask = env.get_role_ID() == 2                                  # the landlord's decisions only
        assert 0 < int(mask.sum()) < env.B
        want_m = D.BatchedGreedyPolicy(net).q_values(env, mask)
        got_m = fused.q_values(env, mask)
        torch.testing.assert_close(got_m, want_m, rtol=1e-4, atol=1e-5)
        small = D.FusedQScorer(net, C, chunk_rows=700)
        # chunks change the GEMM's row count, and with it the library's choice of kernel: equal up to float32 rounding
        torch.testing.assert_close(small.q_values(env, mask), got_m, rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(small.q_values(env), got, rtol=1e-5, atol=2e-6)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.gpu
def test_fused_q_scorer_reduced_precisions(D):
    """tf32 / bf16 rows: Q values within 1e-2 of the float32 ones (outputs of magnitude ~0.1-1) and the same greedy move in
    at least 97 % of the decisions of a random-weight network"""
    env = _midgame_env(D, "BatchedEnvCooperation", 4096, 17, seed=9)
    net = QNetLike(9, 256, 256, seed=11).eval().cuda()
    exact = D.BatchedGreedyPolicy(net, fused=True)
    q32 = exact.q_values(env)
    for precision in ("tf32", "bf16"):
        pol = D.BatchedGreedyPolicy(net, fused=True, precision=precision)
        q = pol.q_values(env)
        assert float((q - q32).abs().max()) < 1e-2, precision
        assert (pol.select(env, q) == exact.select(env, q32)).float().mean() > 0.97, precision


@pytest.mark.gpu
def test_fused_scoring_follows_the_weights_during_training(D):
    """BatchedDQN(fused=True): after TD updates the scorer's tables are rebuilt, so its Q values stay those of the module"""
    import functools
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        net_cls = functools.partial(QNetLike, 9, 256, 64)
        game = D.BatchedGame(D.BatchedEnvCooperation, {"lord": net_cls},
                             {"lord": functools.partial(D.BatchedDQN, fused=True, batch_size=64)}, num_envs=512, seed=3)
        before = [p.detach().clone() for p in game.lord.policy_net.parameters()]
        game.train(episodes=2)
        assert any(not torch.equal(a, b) for a, b in zip(before, game.lord.policy_net.parameters()))    # it did learn
        env = game.env
        env.observe()
        got = game.lord.q_values(env)
        want = D.BatchedGreedyPolicy(game.lord.policy_net.eval()).q_values(env)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        assert game.lord._policy.scorer is not None
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.gpu
def test_fused_q_scorer_with_finished_envs(D):
    """envs whose game is over have no legal moves: the scorer skips them (mixed batch) and returns an empty vector when
    every env is finished"""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        perm, lord = D.random_deals(600, seed=4)
        env = D.BatchedEnvCooperation(600, seed=4)
        env.prepare(torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda())
        net = QNetLike(9, 256, 64, seed=2).eval().cuda()
        fused = D.BatchedGreedyPolicy(net, fused=True)
        for _ in range(40):                                   # no re-deal: games end one by one
            env.rollout_step()
        done = env.is_done
        assert 0 < int(done.sum()) < env.B
        env.observe()
        got, want = fused.q_values(env), D.BatchedGreedyPolicy(net).q_values(env)
        assert got.shape[0] == env.num_actions
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        env.playout(max_steps=400)
        env.observe()
        assert bool(env.is_done.all()) and env.num_actions == 0
        assert fused.q_values(env).numel() == 0
        assert (fused.select(env, fused.q_values(env)) == -1).all()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
