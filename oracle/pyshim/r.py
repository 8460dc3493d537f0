"""Stand-in for the reference's absent native module `r` (reference envi.py:11,111), backed by the C oracle's
definitional legal-move generator.  TEST INFRASTRUCTURE ONLY (see env.py in this directory)."""
from oracle import ddz_oracle as O


def get_moves(hand, last):
    return [[int(x) for x in row] for row in O.get_moves(list(hand), list(last), fast=False)]
