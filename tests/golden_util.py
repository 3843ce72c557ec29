"""Load golden fixtures (reference outputs) and rebuild their seeded inputs/parameters."""
import os

import numpy as np

from cases import CASES, inputs  # tests/golden/cases.py (on sys.path via conftest)
from oracle.glom_oracle import synth_params

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    case = CASES[name]
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        outs = {k: z[k] for k in z.files}
    params = synth_params(case["dim"], case["levels"], case["image_size"], case["patch_size"],
                          seed=case.get("param_seed", 0))
    return case, params, outs


def model_kwargs(case):
    return dict(dim=case["dim"], levels=case["levels"], image_size=case["image_size"],
                patch_size=case["patch_size"], consensus_self=case.get("consensus_self", False),
                local_consensus_radius=case.get("local_consensus_radius", 0))


__all__ = ["CASES", "GOLDEN_DIR", "inputs", "load", "model_kwargs"]
