"""Shared helpers for the -m gpu parity tests."""
import numpy as np
import torch

import multimodal_seq2seq_gscan_b200 as pkg
from oracle import gscan_oracle as O

DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def full_state_dict(params):
    sd = {k: v.detach().clone().float() for k, v in params.items()}
    for att in ("textual_attention", "visual_attention"):
        for layer in ("key_layer", "query_layer", "energy_layer"):
            sd[f"attention_decoder.{att}.{layer}.weight"] = sd[f"{att}.{layer}.weight"]
    return sd


def build_model(cfg, params, train=True):
    model = pkg.Model(**O.model_kwargs(cfg)).to(DEV)
    model.load_state_dict(full_state_dict(params), strict=True)
    model.train(train)
    return model


def to_dev(batch):
    return {
        "commands": torch.tensor(batch["commands"], device=DEV),
        "situations": torch.tensor(batch["situations"], device=DEV),
        "targets": torch.tensor(batch["targets"], device=DEV),
        "positions": torch.tensor(batch["target_positions"], device=DEV),
    }


def oracle_run(cfg, params, batch, weight_target_loss=0.3, dropout=None, dtype=torch.float64):
    """Oracle forward + loss + gradients on CPU in float64."""
    p = {k: v.detach().to(dtype).requires_grad_(True) for k, v in params.items()}
    drop = None if dropout is None else {k: v.to(dtype) for k, v in dropout.items()}
    logp, aux = O.model_forward(p, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                                torch.tensor(batch["situations"]).to(dtype), torch.tensor(batch["targets"]),
                                cfg["conditional_attention"], cfg["auxiliary_task"], dropout=drop)
    loss = O.nll_loss(logp, torch.tensor(batch["targets"]))
    if cfg["auxiliary_task"]:
        loss = loss + weight_target_loss * O.aux_nll_loss(aux, torch.tensor(batch["target_positions"]))
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items()}
    return logp.detach(), None if aux is None else aux.detach(), loss.detach(), grads


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return (a - b).norm().item() / max(b.norm().item(), 1e-30)
