"""The hand-derived backward (oracle/manual_backward.py - the recipe the CUDA kernels transliterate)
must agree with the reference's golden gradients and with autograd of the oracle, incl. dropout masks."""
import numpy as np
import pytest
import torch

from oracle import gscan_oracle as O
from oracle import manual_backward as MB
from tests.golden_util import load_case


@pytest.mark.parametrize("name", ["tiny_aux", "tiny_nocond", "demo", "comp_small", "comp_aux", "tlen_small"])
def test_manual_matches_golden(name):
    cfg, meta, params, batch, z = load_case(name)
    out, S = MB.forward(params, batch, cfg)
    np.testing.assert_allclose(out["logp"], z["logp"], rtol=1e-9, atol=1e-9)
    loss, dlogp, daux = MB.loss_and_upstream(out, batch, cfg["auxiliary_task"], meta["weight_target_loss"])
    np.testing.assert_allclose(loss, z["loss"], rtol=1e-10)
    grads = MB.backward(params, batch, cfg, S, dlogp, daux)
    for pname, _ in O.param_shapes(cfg):
        ref = z["grad." + pname].astype(np.float64)
        tol = 1e-9 if z["grad." + pname].dtype == np.float64 else 3e-6
        assert np.abs(grads[pname] - ref).max() <= tol * max(1.0, np.abs(ref).max()), pname


def test_manual_matches_autograd_with_dropout():
    cfg = dict(O.CONFIGS["tiny"])
    params = O.synthetic_params(cfg, 7, scale=2.0, dtype=torch.float64)
    batch = O.synthetic_batch(cfg, batch_size=4, seed=8, max_cmd_len=6, min_cmd_len=3, max_tgt_len=8)
    rng = np.random.default_rng(3)
    B, Ti, Tt = 4, batch["commands"].shape[1], batch["targets"].shape[1]
    M, D, E, H = cfg["grid_size"] ** 2, 3 * cfg["cnn_hidden_num_channels"], cfg["embedding_dimension"], cfg["decoder_hidden_size"]
    drop = {"cnn": (rng.random((B, M, D)) > 0.3) / 0.7, "enc": (rng.random((B, Ti, E)) > 0.3) / 0.7,
            "dec": (rng.random((B, Tt, H)) > 0.3) / 0.7}
    for v in params.values():
        v.requires_grad_(True)
    tdrop = {k: torch.tensor(v) for k, v in drop.items()}
    logp, aux = O.model_forward(params, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                                torch.tensor(batch["situations"], dtype=torch.float64),
                                torch.tensor(batch["targets"]), True, True, dropout=tdrop)
    loss = O.nll_loss(logp, torch.tensor(batch["targets"])) + 0.3 * O.aux_nll_loss(
        aux, torch.tensor(batch["target_positions"]))
    loss.backward()
    out, S = MB.forward(params_np(params), batch, cfg, dropout=drop)
    np.testing.assert_allclose(out["logp"], logp.detach().numpy(), rtol=1e-10, atol=1e-10)
    l2, dlogp, daux = MB.loss_and_upstream(out, batch, True, 0.3)
    assert l2 == pytest.approx(loss.item(), rel=1e-12)
    grads = MB.backward(params_np(params), batch, cfg, S, dlogp, daux, dropout=drop)
    for pname, _ in O.param_shapes(cfg):
        np.testing.assert_allclose(grads[pname], params[pname].grad.numpy(), rtol=1e-8, atol=1e-11, err_msg=pname)


def params_np(params):
    return {k: v.detach().numpy() for k, v in params.items()}
