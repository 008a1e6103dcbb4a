"""Generates tests/golden/*.npz from the importable part of the reference (pssgp.toymodels is pure numpy;
the TF-dependent modules cannot be imported in this image).  Run in the build container only:

    PYTHONPATH=/root/reference python scripts/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from pssgp.toymodels import obs_noise, sinu, comp_sinu, rect  # noqa: E402

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)
# experiments/toy_models/common.py:28-46 (get_data) for N = 1000, seed 0, noise 0.1 = BASELINE configs[0]
t = np.linspace(0, 4, 1000)
np.savez_compressed(os.path.join(out, "toy_sinusoid_n1000.npz"), t=t, ft=sinu(t), y=obs_noise(sinu(t), 0.1, 0),
                    y_seed666=obs_noise(sinu(t), 0.5, 666), comp=comp_sinu(t), rect=rect(t))
print("written", out)
