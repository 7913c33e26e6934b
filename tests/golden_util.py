"""Helpers shared by the golden-vector tests: load a fixture and rebuild its inputs from the seed."""
import ast
import os

import numpy as np
import torch

from oracle import gscan_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = ["tiny_aux", "tiny_nocond", "demo", "demo_f32", "comp_small", "comp_aux", "tlen_small"]
# (plus "greedy_long": predict() outputs only - sequences and attention weights at max_decoding_steps = 120)


def load_case(name, dtype=None):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    meta = ast.literal_eval(str(z["__meta__"]))
    cfg = dict(O.CONFIGS[meta["cfg_name"]])
    cfg.update(ast.literal_eval(meta["overrides"]))
    batch_kw = ast.literal_eval(meta["batch_kw"])
    dtype = dtype or getattr(torch, meta["dtype"])
    params = O.synthetic_params(cfg, meta["seed"], scale=meta["param_scale"], dtype=dtype)
    batch = O.synthetic_batch(cfg, seed=meta["seed"] + 1, **batch_kw)
    return cfg, meta, params, batch, z
