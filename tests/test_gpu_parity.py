"""Parity of the sm_100a kernels (through the drop-in Model -> C ABI) against the CPU oracle and the
reference's golden vectors.  Tolerances: the north star asks for fp32 agreement within 1e-4
relative; the reference's own fp32-vs-fp64 noise is ~5e-7 on log-probs and ~4e-6 rel-L2 on
gradients, so we require:
    log-probs     max-abs <= 2e-5     loss  rel <= 1e-5     gradients  rel-L2 <= 1e-4
Greedy sequences, lengths and exact-match must be identical."""
import numpy as np
import pytest
import torch

import multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import ops
from oracle import gscan_oracle as O
from tests.golden_util import CASE_NAMES, load_case
from tests.gpu_util import DEV, build_model, full_state_dict, oracle_run, rel_l2, to_dev

pytestmark = pytest.mark.gpu

LOGP_ATOL = 2e-5
GRAD_RTOL = 1e-4


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (128, 64, 16), (129, 65, 17), (300, 400, 100), (1000, 9, 100),
                                   (9, 100, 2000)])
def test_sgemm_forms(M, N, K):
    lib = pkg.load()
    g = torch.Generator().manual_seed(M * 1000 + N * 10 + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    Bm = torch.randn(K, N, generator=g, dtype=torch.float64)
    bias = torch.randn(N, generator=g, dtype=torch.float64)
    ref = A @ Bm
    Ad, Bd = A.float().to(DEV), Bm.float().to(DEV)
    st = torch.cuda.current_stream().cuda_stream

    def run(a, a_rs, a_cs, b, b_rs, b_cs, bias_t=None, act=0, acc=0, C=None):
        C = torch.zeros(M, N, device=DEV) if C is None else C
        rc = lib.gscan_sgemm(a.data_ptr(), a_rs, a_cs, b.data_ptr(), b_rs, b_cs, C.data_ptr(), N, M, N, K,
                             None if bias_t is None else bias_t.data_ptr(), act, acc, st)
        assert rc == 0
        return C.cpu().double()

    tol = 1e-5 * max(1.0, ref.abs().max().item())
    assert (run(Ad, K, 1, Bd, N, 1) - ref).abs().max() < tol                                  # NN
    Bt = Bd.t().contiguous()
    assert (run(Ad, K, 1, Bt, 1, K) - ref).abs().max() < tol                                  # NT
    At = Ad.t().contiguous()
    assert (run(At, 1, M, Bd, N, 1) - ref).abs().max() < tol                                  # TN
    out = run(Ad, K, 1, Bt, 1, K, bias.float().to(DEV), act=1)
    assert (out - torch.tanh(ref + bias)).abs().max() < tol + 1e-6     # |tanh'| <= 1
    out = run(Ad, K, 1, Bt, 1, K, bias.float().to(DEV), act=2)
    assert (out - torch.relu(ref + bias)).abs().max() < tol
    C0 = torch.ones(M, N, device=DEV)
    assert (run(Ad, K, 1, Bd, N, 1, acc=1, C=C0) - (ref + 1)).abs().max() < tol


@pytest.mark.parametrize("M,N,K,ksplit", [(24200, 100, 400, 1), (24200, 400, 100, 1), (200, 100, 100, 1),
                                          (7200, 100, 152, 1), (400, 100, 24200, 37), (400, 200, 24200, 18),
                                          (100, 100, 24200, 148), (300, 260, 1000, 1), (64, 32, 32, 1),
                                          (129, 132, 40, 3)])
def test_sgemm_tcgen05_path(M, N, K, ksplit):
    """The tcgen05 / TMA GEMM (csrc/gemm_tc.cuh) in all four operand-major combinations, against fp64 and against
    the mma.sync kernel.  Tolerance: 2e-5 of the typical |sum| (both kernels are 3xTF32; fp32 itself is ~1e-6)."""
    lib = pkg.load()
    g = torch.Generator().manual_seed(M + 7 * N + 13 * K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    Bm = 0.1 * torch.randn(K, N, generator=g, dtype=torch.float64)
    bias = torch.randn(N, generator=g, dtype=torch.float64)
    ref = A @ Bm
    scale = (A.abs() @ Bm.abs()).max().item() / K ** 0.5
    tol = 2e-5 * scale
    st = torch.cuda.current_stream().cuda_stream
    # leading dimensions padded to multiples of 4 floats (TMA needs 16-byte row strides)
    Kp, Np, Mp = (K + 3) // 4 * 4, (N + 3) // 4 * 4, (M + 3) // 4 * 4
    A_k = torch.zeros(M, Kp, device=DEV); A_k[:, :K] = A.float().to(DEV)          # A(i,k) K-contiguous
    A_m = torch.zeros(K, Mp, device=DEV); A_m[:, :M] = A.t().float().to(DEV)      # A(i,k) M-contiguous
    B_k = torch.zeros(N, Kp, device=DEV); B_k[:, :K] = Bm.t().float().to(DEV)     # B(k,j) K-contiguous
    B_n = torch.zeros(K, Np, device=DEV); B_n[:, :N] = Bm.float().to(DEV)         # B(k,j) N-contiguous

    def run(path, a, a_rs, a_cs, b, b_rs, b_cs, bias_t=None, act=0, acc=0, init=0.0, ks=ksplit):
        C = torch.full((M, Np), init, device=DEV)
        rc = lib.gscan_sgemm_path(a.data_ptr(), a_rs, a_cs, b.data_ptr(), b_rs, b_cs, C.data_ptr(), Np, M, N, K,
                                  None if bias_t is None else bias_t.data_ptr(), act, acc, ks, path, st)
        assert rc == 0, rc
        torch.cuda.synchronize()
        return C[:, :N].cpu().double()

    forms = {"NT": (A_k, Kp, 1, B_k, 1, Kp), "NN": (A_k, Kp, 1, B_n, Np, 1),
             "TN": (A_m, 1, Mp, B_n, Np, 1), "TT": (A_m, 1, Mp, B_k, 1, Kp)}
    for name, f in forms.items():
        out = run(1, *f)
        assert (out - ref).abs().max() < tol, (name, (out - ref).abs().max().item(), tol)
        old = run(0, *f)
        assert (out - old).abs().max() < tol, name
    if ksplit == 1:
        bt = bias.float().to(DEV)
        out = run(1, *forms["NT"], bias_t=bt, act=1)
        assert (out - torch.tanh(ref + bias)).abs().max() < tol + 1e-6
        out = run(1, *forms["NN"], bias_t=bt, act=2)
        assert (out - torch.relu(ref + bias)).abs().max() < tol
        out = run(1, *forms["NT"], acc=1, init=1.0)
        assert (out - (ref + 1)).abs().max() < tol
    else:   # split-K adds onto what C holds
        out = run(1, *forms["TN"], init=2.0)
        assert (out - (ref + 2)).abs().max() < tol


def test_sgemm_tcgen05_rejects_unaligned():
    lib = pkg.load()
    st = torch.cuda.current_stream().cuda_stream
    A = torch.zeros(256, 150, device=DEV)     # row stride 600 B: not a multiple of 16
    Bm = torch.zeros(128, 150, device=DEV)
    C = torch.zeros(256, 128, device=DEV)
    rc = lib.gscan_sgemm_path(A.data_ptr(), 150, 1, Bm.data_ptr(), 1, 150, C.data_ptr(), 128, 256, 128, 150, None, 0, 0,
                              1, 1, st)
    assert rc == -2


@pytest.mark.parametrize("name", CASE_NAMES)
def test_encode_input_matches_golden(name):
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    enc = model.encode_input(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                             situations_input=d["situations"])
    assert enc["encoded_situations"].shape == z["encoded_situations"].shape
    np.testing.assert_allclose(enc["encoded_situations"].cpu().numpy(), z["encoded_situations"], atol=5e-6, rtol=1e-5)
    np.testing.assert_allclose(enc["encoded_commands"]["encoder_outputs"].cpu().numpy(), z["encoder_outputs"],
                               atol=5e-6, rtol=1e-5)
    np.testing.assert_allclose(enc["hidden_states"].cpu().numpy(), z["hidden_states"], atol=5e-6, rtol=1e-5)
    assert enc["encoded_commands"]["sequence_lengths"] == [int(l) for l in batch["cmd_lengths"]]
    # the situation encoder is callable on its own, like the reference sub-module
    feat = model.situation_encoder(d["situations"])
    np.testing.assert_allclose(feat.cpu().numpy(), z["encoded_situations"], atol=5e-6, rtol=1e-5)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_forward_loss_gradients_match_golden(name):
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                      situations_input=d["situations"], target_batch=d["targets"],
                      target_lengths=batch["tgt_lengths"])
    assert logp.shape == z["logp"].shape
    assert np.abs(logp.detach().cpu().numpy().astype(np.float64) - z["logp"]).max() <= LOGP_ATOL
    loss = model.get_loss(logp, d["targets"])
    assert abs(loss.item() - float(z["nll"])) <= 1e-5 * max(1.0, abs(float(z["nll"])))
    if cfg["auxiliary_task"]:
        assert np.abs(aux.detach().cpu().numpy().astype(np.float64) - z["aux_logp"]).max() <= LOGP_ATOL
        aux_loss = model.get_auxiliary_loss(aux, d["positions"])
        assert abs(aux_loss.item() - float(z["aux_nll"])) <= 1e-5 * max(1.0, abs(float(z["aux_nll"])))
        loss += meta["weight_target_loss"] * aux_loss       # in-place, exactly as train.py:107 does
        assert model.get_auxiliary_accuracy(aux, d["positions"]) == pytest.approx(float(z["aux_accuracy"]))
    else:
        assert isinstance(aux, tuple) and len(aux) == 2      # model.py:217
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * max(1.0, abs(float(z["loss"])))
    acc, exact = model.get_metrics(logp, d["targets"])
    assert acc == pytest.approx(float(z["accuracy"])) and exact == pytest.approx(float(z["exact_match"]))
    loss.backward()
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        ref = torch.tensor(z["grad." + pname].astype(np.float64))
        err = rel_l2(named[pname].grad, ref)
        assert err <= GRAD_RTOL, f"{pname}: rel-L2 {err:.3e}"


@pytest.mark.parametrize("name", CASE_NAMES)
def test_greedy_decode_matches_reference_predict(name):
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    N = int(z["greedy_max_steps"])
    out = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], N, 1, 2, return_attention=True)
    toks, lens, steps = out["tokens"].cpu().numpy(), out["lengths"].cpu().numpy(), out["steps"].cpu().numpy()
    assert lens.tolist() == z["greedy_lengths"].tolist()
    for b in range(len(lens)):
        assert toks[b, :lens[b]].tolist() == z["greedy_sequences"][b, :lens[b]].tolist()
        assert (toks[b, lens[b]:] == -1).all()
        assert steps[b] == min(lens[b] + 1, N + 1)
        tgt = batch["targets"][b, :int(batch["tgt_lengths"][b])].tolist()[1:-1]
        assert O.sequence_accuracy(toks[b, :lens[b]].tolist(), tgt) == pytest.approx(float(z["greedy_accuracy"][b]))
    # attention rows of kept steps are probability distributions over the valid keys
    al = out["attention_weights_commands"].cpu().numpy()
    be = out["attention_weights_situations"].cpu().numpy()
    for b in range(len(lens)):
        if lens[b]:
            np.testing.assert_allclose(al[b, :lens[b]].sum(-1), 1.0, atol=1e-5)
            np.testing.assert_allclose(be[b, :lens[b]].sum(-1), 1.0, atol=1e-5)
            assert (al[b, :, int(batch["cmd_lengths"][b]):] == 0).all()
    if cfg["auxiliary_task"]:
        pred = out["aux_logp"].argmax(dim=1).cpu().numpy()
        acc = 100.0 * (pred == batch["target_positions"]).astype(np.float64)
        np.testing.assert_allclose(acc, z["greedy_aux_accuracy"])
    # N+1 cap when EOS is unreachable (predict.py:101 uses <=)
    out2 = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], 5, 1, -1)
    assert out2["lengths"].cpu().tolist() == [6] * len(lens)
    assert out2["tokens"].cpu().numpy().tolist() == z["greedy_noeos_sequences"].tolist()


@pytest.mark.parametrize("tag,eos", [("eos", 2), ("noeos", -1)])
def test_greedy_decode_full_length_with_attention_matches_reference(tag, eos):
    """120 decoding steps against the reference's own predict() (golden `greedy_long`, float32): token sequences,
    lengths, and the attention weights alpha / beta of every kept step (predict.py:108-109) to 2e-5."""
    cfg, meta, params, batch, z = load_case("greedy_long", dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    N = int(meta["max_decoding_steps"])
    out = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], N, 1, eos, return_attention=True)
    toks, lens = out["tokens"].cpu().numpy(), out["lengths"].cpu().numpy()
    assert lens.tolist() == z[f"{tag}_lengths"].tolist()
    al = out["attention_weights_commands"].cpu().numpy()
    be = out["attention_weights_situations"].cpu().numpy()
    for b in range(len(lens)):
        n, n_in = int(lens[b]), int(batch["cmd_lengths"][b])
        assert toks[b, :n].tolist() == z[f"{tag}_sequences"][b, :n].tolist()
        if n:
            np.testing.assert_allclose(al[b, :n, :n_in], z[f"{tag}_alphas"][b, :n, :n_in], rtol=0, atol=2e-5)
            np.testing.assert_allclose(be[b, :n], z[f"{tag}_betas"][b, :n], rtol=0, atol=2e-5)
    pred = out["aux_logp"].argmax(dim=1).cpu().numpy()
    np.testing.assert_allclose(100.0 * (pred == batch["target_positions"]), z[f"{tag}_aux_accuracy"])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_batched_predict_and_evaluate_match_reference(name):
    """predict() / evaluate() of the package over ragged batches give, per example, what the reference's
    batch-1 predict() / evaluate() gave for the golden case (sequences, accuracies, exact match, aux accuracy)."""
    from multimodal_seq2seq_gscan_b200 import predict as P
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    N = int(z["greedy_max_steps"])
    B = d["commands"].shape[0]
    cuts = [0, max(1, B // 3), B]

    def iterator():
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            Ti = int(batch["cmd_lengths"][lo:hi].max())
            Tt = int(batch["tgt_lengths"][lo:hi].max())
            yield (d["commands"][lo:hi, :Ti], batch["cmd_lengths"][lo:hi], [f"deriv{b}" for b in range(lo, hi)],
                   d["situations"][lo:hi], [{"id": b} for b in range(lo, hi)], d["targets"][lo:hi, :Tt],
                   batch["tgt_lengths"][lo:hi], torch.zeros(hi - lo, dtype=torch.long, device=DEV),
                   torch.tensor(batch["target_positions"][lo:hi], device=DEV))

    recs = list(P.predict(iterator(), model, N, 0, 1, 2))
    assert len(recs) == B
    for b, (inp, deriv, sit, out, tgt, att_c, att_s, aux) in enumerate(recs):
        n = int(z["greedy_lengths"][b])
        assert out == z["greedy_sequences"][b, :n].tolist()
        assert deriv == [f"deriv{b}"] and sit == [{"id": b}]
        assert inp.shape == (1, int(batch["cmd_lengths"][b])) and tgt.shape == (1, int(batch["tgt_lengths"][b]))
        assert P.sequence_accuracy(out, tgt[0].tolist()[1:-1]) == pytest.approx(float(z["greedy_accuracy"][b]))
        assert len(att_c) == n and len(att_s) == n
        if n:
            assert len(att_c[0][0]) == int(batch["cmd_lengths"][b]) and len(att_s[0][0]) == cfg["grid_size"] ** 2
            np.testing.assert_allclose(np.sum(att_c[0][0]), 1.0, atol=1e-5)
        if cfg["auxiliary_task"]:
            assert aux == float(z["greedy_aux_accuracy"][b])
    acc, em, aux_acc = P.evaluate(iterator(), model, N, 0, 1, 2)
    assert acc == pytest.approx(float(np.mean(z["greedy_accuracy"])))
    assert em == pytest.approx(100.0 * float(np.mean(z["greedy_accuracy"] == 100)))
    if cfg["auxiliary_task"]:
        assert aux_acc == pytest.approx(float(np.mean(z["greedy_aux_accuracy"])))


@pytest.mark.parametrize("name", ["tiny_aux", "demo", "comp_small"])
def test_decode_input_step_api(name):
    """predict.py's calling sequence: encode_input, key_layer x2, initialize_hidden, decode_input loop."""
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    enc = model.encode_input(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                             situations_input=d["situations"])
    keys_vis = model.visual_attention.key_layer(enc["encoded_situations"])
    keys_txt = model.textual_attention.key_layer(enc["encoded_commands"]["encoder_outputs"])
    hidden = model.attention_decoder.initialize_hidden(model.tanh(model.enc_hidden_to_dec_hidden(enc["hidden_states"])))
    # oracle, float64
    p64 = {k: v.double() for k, v in params.items()}
    feat = O.cnn_forward(p64, torch.tensor(batch["situations"]).double())
    hid, enc_out = O.encoder_forward(p64, torch.tensor(batch["commands"]), batch["cmd_lengths"])
    kt, kv = O.project_keys(p64, enc_out, feat)
    h, c = O.initial_state(p64, hid)
    assert (keys_vis.cpu().double() - kv).abs().max() < 1e-5
    assert (keys_txt.cpu().double() - kt.transpose(0, 1)).abs().max() < 1e-5
    assert (hidden[0][0].cpu().double() - h).abs().max() < 1e-5
    lens = torch.tensor([int(l) for l in batch["cmd_lengths"]])
    for t in range(3):
        tok = torch.tensor(batch["targets"][:, t])
        logits, hidden, beta, alpha, beta2 = model.decode_input(
            target_token=tok.to(DEV), hidden=hidden, encoder_outputs=keys_txt,
            input_lengths=batch["cmd_lengths"], encoded_situations=keys_vis)
        lo, h, c, al, be = O.decoder_step(p64, O.embed(p64["attention_decoder.embedding.weight"], tok), h, c, kt,
                                          lens, kv, cfg["conditional_attention"])
        assert hidden[0].shape == (1, len(lens), cfg["decoder_hidden_size"])
        assert (logits.cpu().double() - lo).abs().max() < 2e-5
        assert (hidden[0][0].cpu().double() - h).abs().max() < 1e-5
        assert (hidden[1][0].cpu().double() - c).abs().max() < 1e-5
        assert (alpha.cpu().double() - al).abs().max() < 1e-5
        assert (beta.cpu().double() - be).abs().max() < 1e-5 and torch.equal(beta, beta2)


def _random_masks(cfg, B, Ti, Tt, seed, p=0.3):
    g = torch.Generator().manual_seed(seed)
    M, D, E, H = cfg["grid_size"] ** 2, 3 * cfg["cnn_hidden_num_channels"], cfg["embedding_dimension"], cfg["decoder_hidden_size"]
    mk = lambda *s: (torch.rand(*s, generator=g) > p).float() / (1 - p)
    return {"cnn": mk(B, M, D), "enc": mk(B, Ti, E), "dec": mk(B, Tt, H)}


@pytest.mark.parametrize("cfg_name,B,tgt", [("tiny", 5, 9), ("comp", 7, 25)])
def test_dropout_masks_forward_backward(cfg_name, B, tgt):
    """Train-mode arithmetic with explicit dropout masks equals the oracle with the same masks."""
    cfg = dict(O.CONFIGS[cfg_name])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 21, scale=2.0)
    batch = O.synthetic_batch(cfg, batch_size=B, seed=22, max_cmd_len=8, min_cmd_len=3, max_tgt_len=tgt)
    Ti, Tt = batch["commands"].shape[1], batch["targets"].shape[1]
    masks = _random_masks(cfg, B, Ti, Tt, 5)
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    cmd_len = ops.lengths_to_device(batch["cmd_lengths"], DEV)
    logp, aux = ops.ModelForward.apply(model._cfg(cfg["grid_size"]), d["commands"], cmd_len, Ti, d["situations"],
                                       d["targets"], tuple(masks[k].to(DEV) for k in ("cnn", "enc", "dec")),
                                       *model._param_list())
    loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
    loss.backward()
    logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch, dropout=masks)
    assert (logp.detach().cpu().double() - logp_o).abs().max() <= LOGP_ATOL
    assert (aux.detach().cpu().double() - aux_o).abs().max() <= LOGP_ATOL
    assert abs(loss.item() - loss_o.item()) <= 1e-5 * max(1.0, abs(loss_o.item()))
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        err = rel_l2(named[pname].grad, grads_o[pname])
        assert err <= GRAD_RTOL, f"{pname}: rel-L2 {err:.3e}"


def test_train_mode_dropout_statistics():
    """With p > 0 and model.train() the masks are drawn internally: outputs differ between calls,
    stay normalised, and the kept fraction matches 1-p."""
    cfg = dict(O.CONFIGS["comp"])
    cfg.update(encoder_dropout_p=0.3, decoder_dropout_p=0.3, cnn_dropout_p=0.1)
    params = O.synthetic_params(cfg, 3)
    batch = O.synthetic_batch(cfg, batch_size=16, seed=4, max_tgt_len=10)
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    args = dict(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"], situations_input=d["situations"],
                target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
    a, _ = model(**args)
    b, _ = model(**args)
    assert not torch.equal(a, b)
    assert torch.allclose(a.exp().sum(-1), torch.ones_like(a[..., 0]), atol=1e-5)
    m = ops._dropout_mask((64, 1000), 0.3, DEV)
    assert abs((m > 0).float().mean().item() - 0.7) < 0.01 and abs(m.max().item() - 1 / 0.7) < 1e-6
    model.eval()
    a, _ = model(**args)
    b, _ = model(**args)
    assert torch.equal(a, b)


@pytest.mark.parametrize("cfg_name,aux,lo", [("comp", False, 3), ("comp", True, 3), ("tlen", False, 17)])
def test_full_size_against_oracle(cfg_name, aux, lo):
    """BASELINE.json configs 2, 3 and 5 at full size (B=200, Tt=121): forward + gradients vs the
    float64 oracle, plus size-independent properties (normalisation, batch-split invariance)."""
    cfg = dict(O.CONFIGS[cfg_name])
    cfg["auxiliary_task"] = aux
    params = O.synthetic_params(cfg, 1234)
    batch = O.synthetic_batch(cfg, batch_size=200, seed=1235, min_tgt_len=lo)
    assert batch["targets"].shape == (200, 121) and batch["commands"].shape == (200, 10)
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    logp, auxo = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                       situations_input=d["situations"], target_batch=d["targets"],
                       target_lengths=batch["tgt_lengths"])
    loss = model.get_loss(logp, d["targets"])
    if aux:
        loss = loss + 0.3 * model.get_auxiliary_loss(auxo, d["positions"])
    loss.backward()
    assert torch.isfinite(logp).all()
    assert torch.allclose(logp.exp().sum(-1), torch.ones(200, 121, device=DEV), atol=1e-5)
    logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch)
    assert (logp.detach().cpu().double() - logp_o).abs().max() <= LOGP_ATOL
    if aux:
        assert (auxo.detach().cpu().double() - aux_o).abs().max() <= LOGP_ATOL
    assert abs(loss.item() - loss_o.item()) <= 1e-5 * abs(loss_o.item())
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        err = rel_l2(named[pname].grad, grads_o[pname])
        assert err <= GRAD_RTOL, f"{pname}: rel-L2 {err:.3e}"
    # batch-split invariance: examples are independent, so a ragged sub-batch reproduces its rows
    sub = slice(37, 37 + 51)
    sub_len = batch["cmd_lengths"][sub]
    Ti_sub = int(sub_len.max())
    with torch.no_grad():
        logp_sub, _ = model(commands_input=d["commands"][sub, :Ti_sub].contiguous(), commands_lengths=sub_len,
                            situations_input=d["situations"][sub], target_batch=d["targets"][sub],
                            target_lengths=batch["tgt_lengths"][sub])
    assert (logp_sub - logp[sub]).abs().max() <= 1e-5


def test_full_size_greedy_decode_properties():
    """Config 4: B=200, max_decoding_steps=120.  Batched result == per-example result; EOS-free run
    emits exactly 121 tokens per sequence; prefix property in max_decoding_steps."""
    cfg = dict(O.CONFIGS["comp"])
    params = O.synthetic_params(cfg, 1234, scale=2.0)
    batch = O.synthetic_batch(cfg, batch_size=200, seed=99)
    model = build_model(cfg, params, train=False)
    d = to_dev(batch)
    out = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], 120, 1, 2)
    lens = out["lengths"].cpu().numpy()
    toks = out["tokens"].cpu().numpy()
    assert ((lens >= 0) & (lens <= 121)).all()
    for b in range(200):
        assert (toks[b, :lens[b]] != 2).all() and (toks[b, :lens[b]] >= 0).all()
    # the oracle on ALL 200 examples (float64: a near-tie in the argmax would show up as a token difference)
    p64 = {k: v.double() for k, v in params.items()}
    seqs, *_ = O.greedy_decode(p64, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                               torch.tensor(batch["situations"]).double(), 120)
    assert lens.tolist() == [len(s) for s in seqs]
    for b in range(200):
        assert toks[b, :lens[b]].tolist() == seqs[b], b
    # singles == batched
    for b in (3, 150):
        n = int(batch["cmd_lengths"][b])
        o1 = model.greedy_decode(d["commands"][b:b + 1, :n].contiguous(), batch["cmd_lengths"][b:b + 1],
                                 d["situations"][b:b + 1], 120, 1, 2)
        assert o1["tokens"][0, :lens[b]].cpu().tolist() == toks[b, :lens[b]].tolist()
        assert int(o1["lengths"][0]) == lens[b]
    noeos = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], 120, 1, -1)
    assert (noeos["lengths"].cpu().numpy() == 121).all() and (noeos["steps"].cpu().numpy() == 121).all()
    # ... and the worst case of the bench (EOS unreachable, 200 x 121 steps) token for token against the oracle
    seqs_noeos, *_ = O.greedy_decode(p64, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                                     torch.tensor(batch["situations"]).double(), 120, eos_idx=-1)
    assert noeos["tokens"].cpu().numpy().tolist() == seqs_noeos
    short = model.greedy_decode(d["commands"], batch["cmd_lengths"], d["situations"], 9, 1, -1)
    assert torch.equal(short["tokens"], noeos["tokens"][:, :10])


def test_ragged_batch_sizes_and_lengths():
    """B not a multiple of the per-CTA tile, Tt = 1..3, Ti = 1..3, single-example batches."""
    cfg = dict(O.CONFIGS["demo"])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 5, scale=2.0)
    for B, tmax, cmax in [(1, 3, 3), (3, 3, 4), (5, 4, 3), (9, 6, 8)]:
        batch = O.synthetic_batch(cfg, batch_size=B, seed=B, max_cmd_len=cmax, min_cmd_len=3, max_tgt_len=tmax,
                                  min_tgt_len=3)
        model = build_model(cfg, params, train=True)
        d = to_dev(batch)
        logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                          situations_input=d["situations"], target_batch=d["targets"],
                          target_lengths=batch["tgt_lengths"])
        loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
        loss.backward()
        logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch)
        assert (logp.detach().cpu().double() - logp_o).abs().max() <= LOGP_ATOL
        named = dict(model.named_parameters())
        for pname, _ in O.param_shapes(cfg):
            assert rel_l2(named[pname].grad, grads_o[pname]) <= GRAD_RTOL, (B, pname)


@pytest.mark.parametrize("aux", [False, True])
def test_fused_trainer_gradient_equals_autograd_path(aux):
    """FusedTrainer forms d(loss)/d(logp) from the targets before the forward pass and lets gscan_forward_train run the
    output-head backward inside the forward call (no auxiliary task), or takes the plain autograd path (auxiliary task):
    either way its flat gradient must be the gradient of `get_loss (+ 0.3 get_auxiliary_loss)` through Model, the path
    the golden tests pin against the reference.  Dropout off so that both passes see the same network."""
    from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = aux
    params = O.synthetic_params(cfg, 11, scale=1.5)
    batch = O.synthetic_batch(cfg, batch_size=24, seed=12, max_tgt_len=33)
    d = to_dev(batch)
    kw = O.model_kwargs(cfg)
    for k in list(kw):
        if "dropout" in k:
            kw[k] = 0.0

    def make():
        m = pkg.Model(**kw).to(DEV)
        m.load_state_dict(full_state_dict(params), strict=True)
        m.train(True)
        return m

    ref = make()
    logp, a = ref(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"], situations_input=d["situations"],
                  target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
    loss = ref.get_loss(logp, d["targets"])
    if aux:
        loss = loss + 0.3 * ref.get_auxiliary_loss(a, d["positions"])
    loss.backward()
    model = make()
    trainer = FusedTrainer(model, weight_target_loss=0.3)
    out = trainer.train_step(d["commands"], batch["cmd_lengths"], d["situations"], d["targets"], batch["tgt_lengths"],
                             d["positions"] if aux else None)
    assert abs(out.item() - loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))
    views = trainer._views(trainer.last_flat_grad)
    for (name, pr), g in zip(ref.named_parameters(), views):
        assert rel_l2(g, pr.grad) <= 2e-6, name
    # and a second step on the trainer's persistent buffers gives the same gradient from the same parameters
    model2 = make()
    t2 = FusedTrainer(model2, weight_target_loss=0.3, learning_rate=0.0)
    for _ in range(2):
        t2.train_step(d["commands"], batch["cmd_lengths"], d["situations"], d["targets"], batch["tgt_lengths"],
                      d["positions"] if aux else None)
    for (name, pr), g in zip(ref.named_parameters(), t2._views(t2.last_flat_grad)):
        assert rel_l2(g, pr.grad) <= 2e-6, name


def test_early_loss_gradient_is_bit_identical_and_head_fallback_is_safe():
    """(1) d(loss)/d(logp) formed from the targets alone (gscan_nll_count + gscan_nll_backward) is bit-identical to what
    autograd returns through NLLLoss, in mean form and in the SUM form of the data-parallel step.  (2) A forward pass
    that was given an early d_logp (gscan_forward_train: the output-head backward runs inside the forward call) followed
    by a backward pass with ANOTHER gradient recomputes the head: gradients equal the plain path's."""
    cfg = dict(O.CONFIGS["comp"])
    params = O.synthetic_params(cfg, 21, scale=1.5)
    batch = O.synthetic_batch(cfg, batch_size=16, seed=22, max_tgt_len=27)
    d = to_dev(batch)
    model = build_model(cfg, params, train=True)
    logp, _ = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"], situations_input=d["situations"],
                    target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
    V = logp.shape[2]
    lp = logp.detach().clone().requires_grad_(True)
    nll, n_tok = ops.NLLLoss.apply(lp, d["targets"], 0, 1)
    (g_mean,) = torch.autograd.grad(nll, lp, retain_graph=True)
    (g_sum,) = torch.autograd.grad(nll * n_tok.detach(), lp)
    e_mean, out = ops.nll_grad_from_targets(d["targets"], V, 0, 1, sum_form=False)
    e_sum, _ = ops.nll_grad_from_targets(d["targets"], V, 0, 1, sum_form=True)
    assert torch.equal(e_mean, g_mean) and torch.equal(e_sum, g_sum)
    assert out[1].item() == n_tok.item()

    def grads_with(early, w):
        m = build_model(cfg, params, train=True)
        if early is not None:
            ops.set_early_dlogp(early)
        try:
            lg, _ = m(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"], situations_input=d["situations"],
                      target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
        finally:
            ops.set_early_dlogp(None)
        gs = torch.autograd.grad([lg], list(m.parameters()), grad_outputs=[w])
        return lg.detach(), [g.clone() for g in gs]

    torch.manual_seed(3)
    w = torch.randn_like(logp) * 0.01
    lp0, g0 = grads_with(None, w)                       # plain path
    lp1, g1 = grads_with(e_mean, w)                     # early head for e_mean, then a backward pass with w: recomputed
    lp2, g2 = grads_with(w, w)                          # early head for w, backward with the same tensor: skipped there
    assert torch.equal(lp0, lp1) and torch.equal(lp0, lp2)
    for a, b_, c in zip(g0, g1, g2):
        assert rel_l2(b_, a) <= 1e-6
        assert rel_l2(c, a) <= 2e-6


def test_adam_step_matches_torch():
    torch.manual_seed(0)
    n = 100003
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 4):
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
        assert (p - ref.detach()).abs().max() < 1e-6


def test_fork_join_streams_keep_stream_semantics():
    """gscan_forward / gscan_backward fork onto library-owned helper streams and join back with events; seen from the
    caller the call must still behave like work on the ONE stream it was given: results on a side torch stream, with
    consumers enqueued right behind on that stream and no synchronisation in between, equal the default-stream results."""
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 5, scale=1.5)
    batch = O.synthetic_batch(cfg, batch_size=48, seed=6, max_tgt_len=40)
    d = to_dev(batch)
    pos = torch.tensor(batch["target_positions"], device=DEV)

    def run():
        model = build_model(cfg, params, train=True)
        logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                          situations_input=d["situations"], target_batch=d["targets"],
                          target_lengths=batch["tgt_lengths"])
        loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, pos)
        loss.backward()
        grads = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        return logp.detach().clone(), grads.clone(), loss.detach().clone()

    ref = run()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    outs = []
    with torch.cuda.stream(side):
        for _ in range(3):              # back to back: the helper streams are re-forked while earlier work is in flight
            outs.append(run())
    side.synchronize()
    for logp, grads, loss in outs:
        # split-K atomics make the weight gradients run-to-run reproducible only to fp32 rounding
        assert torch.equal(logp, ref[0])
        assert (grads - ref[1]).norm() <= 1e-5 * ref[1].norm()
        assert abs(float(loss) - float(ref[2])) <= 1e-6 * abs(float(ref[2]))



@pytest.mark.parametrize("B,tmax,lo", [(100, 60, 20), (136, 40, 10), (33, 121, 60), (200, 30, 5)])
def test_medium_shapes_against_oracle(B, tmax, lo):
    """Shapes between the tiny golden cases and the full-size one: the number of clusters of the sweeps (and with it
    the SMs left idle for the shadow launches), the rows per shadow chunk and the K-splits of the grouped
    weight-gradient GEMM all depend on B and Tt; the cut-off below which the shadow schedule is not used at all
    (1,024 rows per chunk) falls inside this set."""
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 77, scale=1.5)
    batch = O.synthetic_batch(cfg, batch_size=B, seed=78, max_tgt_len=tmax, min_tgt_len=lo)
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                      situations_input=d["situations"], target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
    loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
    loss.backward()
    logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch)
    assert (logp.detach().cpu().double() - logp_o).abs().max() <= LOGP_ATOL
    assert (aux.detach().cpu().double() - aux_o).abs().max() <= LOGP_ATOL
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        err = rel_l2(named[pname].grad, grads_o[pname])
        assert err <= GRAD_RTOL, f"B={B} Tt={tmax} {pname}: rel-L2 {err:.3e}"


FALLBACK_ENVS = [
    {"GSCAN_SHADOW": "0"},                                   # grouped weight-gradient GEMM after the sweep only
    {"GSCAN_SHADOW_Z": "1"},                                 # value-path Z kernel chunked into the shadow too (off by default)
    {"GSCAN_SHADOW_CUTS": "45"},                             # one progress signal instead of three
    {"GSCAN_SHADOW_CUTS": "80,60,40,20"},                    # four
    {"GSCAN_NO_GROUP_GEMM": "1", "GSCAN_SHADOW": "0"},       # one split-K launch per weight gradient
    {"GSCAN_ENC_STREAMING": "1"},                            # encoder sweeps that stream W_hh from L2
    {"GSCAN_HEAD_UNFUSED": "1"},                             # log-softmax backward + two products for the output head
    {"GSCAN_CAP_PRELUDE": "0", "GSCAN_CAP_POST": "0", "GSCAN_MAIN_WAITS_VALUE": "0"},
]


@pytest.mark.parametrize("env", FALLBACK_ENVS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_alternative_schedules_keep_parity(env):
    """The library reads its schedule switches from the environment once per process, so every alternative path
    (fallback kernels, other stream / SM-budget arrangements) reruns the full-size oracle comparison in a child
    process: same tolerances, same inputs."""
    import os, subprocess, sys
    child_env = dict(os.environ, **env)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_full_size_against_oracle and comp-True"], cwd=root, env=child_env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "1 passed" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("cfg_name,B", [("tiny", 5), ("comp", 4)])
def test_dense_non_binary_situations(cfg_name, B):
    """The situation CNN is a zero-skipping gather (gSCAN grids are {0,1} and ~3 % dense) but it must stay EXACT for any
    input: dense, non-binary, negative situations through forward, loss and all gradients against the oracle."""
    cfg = dict(O.CONFIGS[cfg_name])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 41, scale=2.0)
    batch = O.synthetic_batch(cfg, batch_size=B, seed=42, max_cmd_len=8, min_cmd_len=3, max_tgt_len=9)
    rng = np.random.default_rng(43)
    batch["situations"] = rng.normal(size=batch["situations"].shape).astype(np.float32)
    batch["situations"][0] = 0.0          # and one all-zero grid
    model = build_model(cfg, params, train=True)
    d = to_dev(batch)
    logp, aux = model(commands_input=d["commands"], commands_lengths=batch["cmd_lengths"],
                      situations_input=d["situations"], target_batch=d["targets"], target_lengths=batch["tgt_lengths"])
    loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
    loss.backward()
    logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch)
    assert (logp.detach().cpu().double() - logp_o).abs().max() <= LOGP_ATOL
    assert (aux.detach().cpu().double() - aux_o).abs().max() <= LOGP_ATOL
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        err = rel_l2(named[pname].grad, grads_o[pname])
        assert err <= GRAD_RTOL, f"{pname}: rel-L2 {err:.3e}"


def test_in_kernel_dropout_matches_materialised_masks():
    """Training-mode dropout is drawn inside the kernels (Philox stream per site, gscan_forward_rng / _backward_rng).
    The masks such a stream stands for can be materialised (gscan_dropout_mask): passing THEM explicitly gives the same
    log-probs and gradients, and the oracle with those masks agrees - so the in-kernel draw is covered by the same
    parity bar as everything else.  Also: kept fraction, scale, independence of the call counter."""
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = True
    params = O.synthetic_params(cfg, 21, scale=2.0)
    B = 7
    batch = O.synthetic_batch(cfg, batch_size=B, seed=22, max_cmd_len=8, min_cmd_len=3, max_tgt_len=25)
    Ti, Tt = batch["commands"].shape[1], batch["targets"].shape[1]
    M, D, E, H = cfg["grid_size"] ** 2, 3 * cfg["cnn_hidden_num_channels"], cfg["embedding_dimension"], cfg["decoder_hidden_size"]
    rng = ops.Dropout(0.1, 0.3, 0.3, 1234, 7)
    masks = ops.dropout_masks_of(rng, ((B, M, D), (B, Ti, E), (B, Tt, H)), DEV)
    for m, p in zip(masks, (0.1, 0.3, 0.3)):
        vals = torch.unique(m).cpu().tolist()
        assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1 / (1 - p)) < 1e-6
        assert abs((m > 0).float().mean().item() - (1 - p)) < 0.03
    other = ops.dropout_masks_of(ops.Dropout(0.1, 0.3, 0.3, 1234, 8), ((B, M, D), (B, Ti, E), (B, Tt, H)), DEV)
    assert not torch.equal(other[2], masks[2])
    d = to_dev(batch)
    cmd_len = ops.lengths_to_device(batch["cmd_lengths"], DEV)
    results = []
    for how in (rng, tuple(masks)):
        model = build_model(cfg, params, train=True)
        logp, aux = ops.ModelForward.apply(model._cfg(cfg["grid_size"]), d["commands"], cmd_len, Ti, d["situations"],
                                           d["targets"], how, *model._param_list())
        loss = model.get_loss(logp, d["targets"]) + 0.3 * model.get_auxiliary_loss(aux, d["positions"])
        loss.backward()
        results.append((logp.detach(), aux.detach(), {n: p.grad.detach() for n, p in model.named_parameters()}))
    (lp_a, aux_a, g_a), (lp_b, aux_b, g_b) = results
    assert torch.equal(lp_a, lp_b) and torch.equal(aux_a, aux_b)
    for name in g_a:      # (split-K atomics: equal up to summation order)
        assert rel_l2(g_a[name], g_b[name]) <= 1e-6, name
    logp_o, aux_o, loss_o, grads_o = oracle_run(cfg, params, batch,
                                                dropout={"cnn": masks[0].cpu(), "enc": masks[1].cpu(), "dec": masks[2].cpu()})
    assert (lp_a.cpu().double() - logp_o).abs().max() <= LOGP_ATOL
    for pname, _ in O.param_shapes(cfg):
        assert rel_l2(g_a[pname], grads_o[pname]) <= GRAD_RTOL, pname
