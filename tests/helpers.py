"""Shared builders for tests: identical weights/inputs for the reference, the oracle and the CUDA path."""
import os

import torch

from tests.golden.make_golden import CASES
from toc3d_b200 import EVA_ViT, ToC3DEVAViT
from toc3d_b200.synthetic import make_gumbel, make_inputs, randomize_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_model(kind, cfg):
    torch.manual_seed(0)
    return (ToC3DEVAViT if kind == "toc3d" else EVA_ViT)(**cfg).eval()


def case_setup(name):
    """-> (fixture, kind, cfg, model (CPU), state_dict, inputs, gumbel) rebuilt exactly as make_golden did."""
    kind, cfg, hw, views, bias_std, prev, seed, sub = CASES[name]
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"))
    model = build_model(kind, cfg)
    sd = randomize_state_dict(model.state_dict(), seed=seed, bias_std=bias_std)
    model.load_state_dict(sd)
    inp = make_inputs(1, views, hw, seed=seed, pose="random")
    inp["prev_exists"] = prev
    gn = make_gumbel(views, (hw[0] // 16) * (hw[1] // 16), seed=seed + 100)
    return fx, kind, cfg, model, sd, inp, gn


def run_oracle(kind, cfg, sd, inp, gn, tap=None):
    from oracle import toc3d_oracle as O
    with torch.no_grad():
        if kind == "dense":
            return O.forward_dense(sd, cfg, inp["x"], tap=tap)
        return O.forward_toc3d(sd, cfg, inp["x"], inp["temp_queries"], inp["temp_ref_points"], inp["temp_vel"],
                               inp["temp_timestamp"], inp["temp_ego_pose"], inp["ego_pose_inv"], inp["prev_exists"],
                               gn, tap=tap)


def to_cuda(inp):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
