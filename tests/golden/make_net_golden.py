"""Generate tests/golden/net_forward.npz from the REFERENCE's Q-networks (net.py:65-139), run unmodified:

    python tests/golden/make_net_golden.py [/root/reference]

The reference's networks are the CONSUMER of the env's tensors (contract only).  tests/qnet_like.py is the
stand-in the shim tests and bench.py --config 3 use on the GPU box, where the reference tree does not exist; this fixture
pins it: for NetComplicated / NetMoreComplicated / NetCooperation / NetCooperationSimplify the stand-in's weights (seeded
init) are loaded into the reference class, and the reference's forward(face, actions) on seeded inputs is stored -- both
call forms of net.py:81-90 (a face per action, and one face [C,15,4] repeated over the actions).
Only input/output vectors are stored; nothing of the reference is copied.
"""
import os
import sys

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyshim"))        # pytz stand-in for config.py
sys.path.insert(0, REF)

import net as refnet  # noqa: E402
from qnet_like import QNetLike, load_into_reference_net, golden_inputs  # noqa: E402

NETS = (("NetComplicated", 4), ("NetMoreComplicated", 7), ("NetCooperation", 9), ("NetCooperationSimplify", 6))


def main():
    out = {}
    torch.set_num_threads(1)
    for k, (name, C) in enumerate(NETS):
        like = QNetLike(C, 256, 256, seed=100 + k)
        ref = getattr(refnet, name)()
        load_into_reference_net(like, ref)
        ref.eval()
        face, actions = golden_inputs(C, 48, seed=200 + k)
        with torch.no_grad():
            out[name + "_batch"] = ref(face, actions).numpy()                     # net.py:87: face.dim() == 4
            out[name + "_single"] = ref(face[0], actions).numpy()                 # one face, repeated (net.py:88)
    path = os.path.join(HERE, "net_forward.npz")
    np.savez_compressed(path, names=np.array([n for n, _ in NETS]), channels=np.array([c for _, c in NETS]), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
