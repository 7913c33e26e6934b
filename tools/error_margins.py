"""How far inside the 1e-4 parity bar the shipping kernels are: full-size (B=200, Tt=121, aux on) forward + backward
against the float64 oracle; prints the max-abs log-prob error, the loss error and the worst / median gradient rel-L2."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gscan_oracle as O
from tests.gpu_util import build_model, to_dev, oracle_run, rel_l2

cfg = dict(O.CONFIGS["comp"]); cfg["auxiliary_task"] = True
params = O.synthetic_params(cfg, 1234, scale=2.0)
batch = O.synthetic_batch(cfg, batch_size=200, seed=99)
model = build_model(cfg, params, train=True)
d = to_dev(batch)
logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"], situations_input=d["situations"],
                  target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
loss.backward()
logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch)
named = dict(model.named_parameters())
errs = {n: rel_l2(named[n].grad, grads_o[n]) for n, _ in O.param_shapes(cfg)}
worst = max(errs, key=errs.get)
print(json.dumps({"shape": "comp, B=200, Tt=121, aux on, fp64 oracle", "logp_max_abs": float((logp.detach().cpu().double() - logp_o).abs().max()),
                  "aux_logp_max_abs": float((aux.detach().cpu().double() - aux_o).abs().max()),
                  "loss_rel": abs(loss.item() - loss_o.item()) / abs(loss_o.item()),
                  "grad_rel_l2_worst": errs[worst], "grad_worst_tensor": worst,
                  "grad_rel_l2_median": float(np.median(list(errs.values()))), "tolerance": 1e-4}))
