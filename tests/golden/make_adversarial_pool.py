"""Writes doudizhu-rl_b200/data/adversarial_pool.npz: the legal LEAD moves of the ten adversarial hands of BASELINE
config 5, from the oracle's definitional generator (filter of the card.py action universe).  They are
the "previous moves" the following half of the config-5 pairs has to beat."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ddz_oracle as O

spec = importlib.util.spec_from_file_location("ddz_deals", os.path.join(ROOT, "doudizhu-rl_b200", "deals.py"))
deals = importlib.util.module_from_spec(spec)
spec.loader.exec_module(deals)

z = np.zeros(15, np.int8)
lists = [O.pack(O.get_moves(h, z, fast=False)) for h in deals.ADVERSARIAL_POOL]
moves = np.concatenate(lists).astype(np.uint64)
np.savez_compressed(deals.POOL_FILE, pool=deals.ADVERSARIAL_POOL, lead_moves=moves,
                    lead_off=np.cumsum([0] + [len(l) for l in lists]).astype(np.int32))
print("wrote %s: %d lead moves of %d hands" % (deals.POOL_FILE, len(moves), len(lists)))
