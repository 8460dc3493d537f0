"""Run by test_gpu_parity.py::test_prob_form_b_build_matches_form_b_oracle in a subprocess with DDZ_LIB / DDZ_ORACLE_LIB set
to the -DDDZ_PROB_FORM_B builds: the fused rollout of all four face variants, bit-exact against the form-B oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddz_b200 as D
from oracle import ddz_oracle as O

assert D.native.lib.ddz_prob_form() == 1 and O.lib().ddz_ref_prob_form() == 1, "not the form-B builds"
B, G = 384, 2
perm, lord = D.random_deals(B, seed=5, pool_games=G)
pd, ld = torch.as_tensor(perm).cuda(), torch.as_tensor(lord).cuda()
right_aligned = 0
for cls, variant in (("BatchedEnv", 0), ("BatchedEnvComplicated", 1), ("BatchedEnvCooperation", 2), ("BatchedEnvCooperationSimplify", 3)):
    env = getattr(D, cls)(B, seed=11)
    env.prepare(pd, ld, pool_games=G)
    ref = O.RefBatch(B, variant)
    ref.deal(perm, lord, pool_games=G)
    for t in range(70):
        off, au, af, face = ref.observe()
        assert np.array_equal(env.face.cpu().numpy(), face), (cls, t)
        assert np.array_equal(env.offsets.cpu().numpy(), off) and np.array_equal(env.valid_actions()[0].cpu().numpy(), af)
        standalone = torch.empty_like(env.face)
        D.native.check(D.native.lib.ddz_encode_face(env._state.data_ptr(), variant, standalone.data_ptr(), B,
                                                    torch.cuda.current_stream().cuda_stream), "ddz_encode_face")
        assert torch.equal(standalone, env.face)
        prob = face[:, -2:]                                   # form B: ones right-aligned in ranks 3..2
        right_aligned += int(((prob[:, :, :13, 0] == 0) & (prob[:, :, :13, 3] > 0)).sum())
        env.rollout_step(perm=pd, lord_pile=ld, pool_games=G)
        ref.step(mode=2, seed=11, env0=0, step=t)
        ref.deal(perm, lord, only_done=True, pool_games=G)
    x, _ = env.state_actions()
    want = np.concatenate([np.repeat(ref.observe()[3], np.diff(ref.observe()[0]), axis=0), ref.observe()[2][:, None]], 1)
    assert np.array_equal(x.cpu().numpy(), want), cls
    # the fused Q-scorer reads the form-B nibbles (reversed rows for ranks 3..2 of the probability planes, jokers unchanged)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from qnet_like import QNetLike
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    for width in (32, 256):
        net = QNetLike(env.C, width, 64, seed=3).eval().cuda()
        torch.testing.assert_close(D.BatchedGreedyPolicy(net, fused=True).q_values(env), D.BatchedGreedyPolicy(net).q_values(env),
                                   rtol=1e-4, atol=1e-5)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    k60 = (np.random.default_rng(1).random((5, 60)) < 0.5).astype(np.int32)
    k60 = np.sort(k60.reshape(5, 15, 4), -1)[:, :, ::-1].reshape(5, 60).copy()   # thermometer rows
    k60.reshape(5, 15, 4)[:, 13:, 1:] = 0
    got = D.BatchedEnv.get_state_prob_manual(k60, [3, 5, 7, 1, 2], [4, 4, 1, 9, 2], device="cuda").cpu().numpy()
    for i in range(5):
        assert np.array_equal(got[i], O.state_prob_manual(k60[i], [3, 5, 7, 1, 2][i], [4, 4, 1, 9, 2][i]))
assert right_aligned > 0
print("form B ok")
