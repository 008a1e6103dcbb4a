"""Mirror of the reference's pssgp/kalman/base.py:3."""
from collections import namedtuple

LGSSM = namedtuple("LGSSM", ["P0", "Fs", "Qs", "H", "R"])
