"""Pin the CPU oracle (oracle/gscan_oracle.py) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import gscan_oracle as O
from tests.golden_util import CASE_NAMES, load_case


def _tol(meta):
    return (1e-9, 1e-9) if meta["dtype"] == "float64" else (2e-5, 2e-5)


def _close(a, b, rtol, atol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.dtype != a.dtype:
        b = b.astype(a.dtype)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_forward_loss_grads(name):
    cfg, meta, params, batch, z = load_case(name)
    rtol, atol = _tol(meta)
    for v in params.values():
        v.requires_grad_(True)
    dtype = next(iter(params.values())).dtype
    commands = torch.tensor(batch["commands"])
    situations = torch.tensor(batch["situations"], dtype=dtype)
    targets = torch.tensor(batch["targets"])
    logp, aux = O.model_forward(params, commands, batch["cmd_lengths"], situations, targets,
                                conditional_attention=cfg["conditional_attention"],
                                auxiliary_task=cfg["auxiliary_task"])
    _close(logp.detach().numpy(), z["logp"], max(rtol, 1e-6), max(atol, 1e-6) if z["logp"].dtype == np.float32 else atol)
    loss = O.nll_loss(logp, targets)
    _close(loss.item(), z["nll"], rtol, atol)
    if cfg["auxiliary_task"]:
        _close(aux.detach().numpy(), z["aux_logp"], rtol, atol)
        aux_loss = O.aux_nll_loss(aux, torch.tensor(batch["target_positions"]))
        _close(aux_loss.item(), z["aux_nll"], rtol, atol)
        loss = loss + meta["weight_target_loss"] * aux_loss
    _close(loss.item(), z["loss"], rtol, atol)
    acc, exact = O.metrics(logp.detach(), targets)
    assert acc == pytest.approx(float(z["accuracy"])) and exact == pytest.approx(float(z["exact_match"]))
    loss.backward()
    for pname, _ in O.param_shapes(cfg):
        g_ref = z["grad." + pname]
        g = params[pname].grad.numpy()
        # big fixtures were stored as float32
        tol = 1e-9 if g_ref.dtype == np.float64 and meta["dtype"] == "float64" else 3e-6
        scale = max(1e-12, float(np.abs(g_ref).max()))
        assert np.abs(g - g_ref).max() <= tol * max(1.0, scale) + tol, pname


@pytest.mark.parametrize("name", CASE_NAMES)
def test_encode_input_parts(name):
    cfg, meta, params, batch, z = load_case(name)
    rtol, atol = _tol(meta)
    dtype = next(iter(params.values())).dtype
    with torch.no_grad():
        feat = O.cnn_forward(params, torch.tensor(batch["situations"], dtype=dtype))
        hidden, enc_out = O.encoder_forward(params, torch.tensor(batch["commands"]), batch["cmd_lengths"])
    _close(feat.numpy(), z["encoded_situations"], max(rtol, 2e-7), max(atol, 2e-7))
    _close(enc_out.numpy(), z["encoder_outputs"], max(rtol, 2e-7), max(atol, 2e-7))
    _close(hidden.numpy(), z["hidden_states"], max(rtol, 2e-7), max(atol, 2e-7))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_greedy_decode(name):
    cfg, meta, params, batch, z = load_case(name)
    dtype = next(iter(params.values())).dtype
    commands = torch.tensor(batch["commands"])
    situations = torch.tensor(batch["situations"], dtype=dtype)
    seqs, alphas, betas, beta_sum = O.greedy_decode(
        params, commands, batch["cmd_lengths"], situations, int(z["greedy_max_steps"]),
        conditional_attention=cfg["conditional_attention"])
    assert [len(s) for s in seqs] == z["greedy_lengths"].tolist()
    for b, s in enumerate(seqs):
        assert s == z["greedy_sequences"][b, :len(s)].tolist()
        tgt = batch["targets"][b, :int(batch["tgt_lengths"][b])].tolist()[1:-1]
        assert O.sequence_accuracy(s, tgt) == pytest.approx(float(z["greedy_accuracy"][b]))
    if cfg["auxiliary_task"]:
        pred = beta_sum.argmax(dim=1).numpy()
        acc = 100.0 * (pred == batch["target_positions"]).astype(np.float64)
        np.testing.assert_allclose(acc, z["greedy_aux_accuracy"])
    # N+1 cap with an unreachable EOS
    seqs2, *_ = O.greedy_decode(params, commands, batch["cmd_lengths"], situations, 5, eos_idx=-1,
                                conditional_attention=cfg["conditional_attention"])
    assert np.array(seqs2).tolist() == z["greedy_noeos_sequences"].tolist()
    assert all(len(s) == 6 for s in seqs2)


def test_sequence_accuracy_edge_cases():
    # helpers.py:44-64 semantics
    assert O.sequence_accuracy([], []) == 0.0
    assert O.sequence_accuracy([], [3, 4]) == 0.0
    assert O.sequence_accuracy([3, 4], [3, 4]) == 100.0
    assert O.sequence_accuracy([3, 4, 5], [3, 4]) == pytest.approx(200.0 / 3)
    assert O.sequence_accuracy([0], [3]) == 0.0


@pytest.mark.parametrize("tag,eos", [("eos", 2), ("noeos", -1)])
def test_greedy_decode_full_length_with_attention(tag, eos):
    """Round 2 golden: the reference's predict() at max_decoding_steps = 120 (float32), sequences AND the per-step
    attention weights it returns (predict.py:108-109); lengths 13 / 3 / 0 / 121 with the real EOS, 121 everywhere
    without."""
    cfg, meta, params, batch, z = load_case("greedy_long")
    N = int(meta["max_decoding_steps"])
    seqs, alphas, betas, beta_sum = O.greedy_decode(params, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                                                    torch.tensor(batch["situations"]), N, eos_idx=eos)
    assert [len(s) for s in seqs] == z[f"{tag}_lengths"].tolist()
    for b, s in enumerate(seqs):
        n = len(s)
        assert s == z[f"{tag}_sequences"][b, :n].tolist()
        n_in = int(batch["cmd_lengths"][b])
        if n:
            np.testing.assert_allclose(np.stack(alphas[b]), z[f"{tag}_alphas"][b, :n, :n_in], rtol=0, atol=2e-6)
            np.testing.assert_allclose(np.stack(betas[b]), z[f"{tag}_betas"][b, :n], rtol=0, atol=2e-6)
    pred = beta_sum.argmax(dim=1).numpy()
    np.testing.assert_allclose(100.0 * (pred == batch["target_positions"]), z[f"{tag}_aux_accuracy"])
