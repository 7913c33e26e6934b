"""CPU oracle for the gSCAN multimodal seq2seq hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, as explicit arithmetic, of the algorithm the reference
implements with ``nn.Conv2d`` / ``nn.LSTM`` / ``nn.Linear`` modules.  It never runs on the
product path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product (``multimodal_seq2seq_gscan_b200``)
fails loudly when its CUDA library is missing; it never falls back to this file.

Parity pinning: the reference's own tests hold NO golden vectors for this path
(``seq2seq/seq2seq_test.py:1-35`` is a stub).  The oracle is therefore pinned against outputs
of the reference itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/seq2seq`` unchanged) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against those fixtures.

All functions take a flat ``dict`` of parameters keyed by the reference's ``state_dict`` names
 and work in whatever dtype the parameters carry (float32 or float64).
Gradients come from autograd over these explicit formulas.

Reference lines followed (relative to /root/reference/seq2seq):
  cnn_forward            cnn_model.py:22-36
  encoder_forward        seq2seq_model.py:47-89
  attention              seq2seq_model.py:105-139, helpers.py:11-32
  decoder_step           seq2seq_model.py:359-428
  decoder_forward        seq2seq_model.py:433-490, 494-504
  model_forward          model.py:172-219
  nll_loss / metrics     model.py:108-160
  greedy_decode          predict.py:57-128
  sequence_accuracy      helpers.py:44-64
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------
# Situation CNN  (cnn_model.py:22-36)
# ----------------------------------------------------------------------------------------
def _same_conv_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Stride-1 'same' cross-correlation on an NHWC grid, written tap by tap.

    x [B,G,G,C] indexed [b,row,col,ch]; weight [F,C,k,k] in the reference's layout.  The
    reference feeds ``input.transpose(1,3)`` (cnn_model.py:28) to Conv2d, i.e. the conv's
    "height" axis is the grid *column* and its "width" axis is the grid *row*.  Weight tap
    [f,c,i,j] therefore multiplies x[b, row + j - p, col + i - p, c].
    """
    B, G, _, C = x.shape
    F, _, k, _ = weight.shape
    p = k // 2
    xp = torch.zeros(B, G + 2 * p, G + 2 * p, C, dtype=x.dtype)
    xp[:, p:p + G, p:p + G, :] = x
    out = bias.view(1, 1, 1, F).expand(B, G, G, F).clone()
    for i in range(k):          # conv "height" tap  -> grid column offset
        for j in range(k):      # conv "width" tap   -> grid row offset
            patch = xp[:, j:j + G, i:i + G, :]                    # [B,G,G,C]
            out = out + torch.einsum("brcq,fq->brcf", patch, weight[:, :, i, j])
    return out


def cnn_forward(p: Params, situations: torch.Tensor,
                dropout_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu(cat[conv1, conv5, conv_k3]) (optionally * dropout_mask) -> [B, G*G, 3F].

    ``dropout_mask`` (already scaled by 1/(1-p)) has the output's shape [B, G*G, 3F].
    Output cell index = row*G + col (cnn_model.py:34-36 transposes back before reshaping).
    """
    outs = []
    for name in ("conv_1", "conv_2", "conv_3"):
        outs.append(_same_conv_nhwc(situations, p[f"situation_encoder.{name}.weight"],
                                    p[f"situation_encoder.{name}.bias"]))
    feat = torch.relu(torch.cat(outs, dim=-1))
    B, G, _, D = feat.shape
    feat = feat.reshape(B, G * G, D)
    if dropout_mask is not None:
        feat = feat * dropout_mask
    return feat


# ----------------------------------------------------------------------------------------
# Embedding with a padding row  (nn.Embedding(padding_idx=...), seq2seq_model.py:42,351)
# ----------------------------------------------------------------------------------------
def embed(weight: torch.Tensor, idx: torch.Tensor, pad_idx: int = 0) -> torch.Tensor:
    """Row lookup; the padding row is read like any other but receives no gradient."""
    keep = (idx != pad_idx).unsqueeze(-1)
    return torch.where(keep, weight[idx], weight.detach()[idx])


# ----------------------------------------------------------------------------------------
# LSTM cell, gate order i,f,g,o  (torch.nn.LSTM semantics used at seq2seq_model.py:44,353)
# ----------------------------------------------------------------------------------------
def lstm_cell(x: torch.Tensor, h: torch.Tensor, c: torch.Tensor, w_ih: torch.Tensor,
              w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor):
    H = h.shape[-1]
    a = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i = torch.sigmoid(a[..., 0:H])
    f = torch.sigmoid(a[..., H:2 * H])
    g = torch.tanh(a[..., 2 * H:3 * H])
    o = torch.sigmoid(a[..., 3 * H:4 * H])
    c_new = f * c + i * g
    h_new = o * torch.tanh(c_new)
    return h_new, c_new


# ----------------------------------------------------------------------------------------
# Command encoder  (seq2seq_model.py:47-89)
# ----------------------------------------------------------------------------------------
def encoder_forward(p: Params, commands: torch.Tensor, lengths: Sequence[int],
                    dropout_mask: Optional[torch.Tensor] = None):
    """Bidirectional LSTM over the valid tokens of each command.

    Returns (hidden [B,H], encoder_outputs [Ti,B,H]) with Ti = max(lengths);
    outputs are h_fwd + h_bwd at valid positions and exactly zero at padded ones
    (pad_packed_sequence, seq2seq_model.py:73); hidden = h_fwd(len-1) + h_bwd(0) (80-82).
    The reference's sort / pack / unsort (64-69, 85-88) is a per-row no-op and is skipped.
    """
    lengths = [int(l) for l in lengths]
    B = commands.shape[0]
    Ti = max(lengths)
    emb = embed(p["encoder.embedding.weight"], commands[:, :Ti])     # [B,Ti,E]
    if dropout_mask is not None:
        emb = emb * dropout_mask[:, :Ti]
    H = p["encoder.lstm.weight_hh_l0"].shape[1]
    dtype = emb.dtype
    len_t = torch.tensor(lengths)
    outs = torch.zeros(Ti, B, H, dtype=dtype)
    finals = []
    for suffix, reverse in (("", False), ("_reverse", True)):
        w_ih = p[f"encoder.lstm.weight_ih_l0{suffix}"]
        w_hh = p[f"encoder.lstm.weight_hh_l0{suffix}"]
        b_ih = p[f"encoder.lstm.bias_ih_l0{suffix}"]
        b_hh = p[f"encoder.lstm.bias_hh_l0{suffix}"]
        h = torch.zeros(B, H, dtype=dtype)
        c = torch.zeros(B, H, dtype=dtype)
        steps = range(Ti - 1, -1, -1) if reverse else range(Ti)
        dir_out = [None] * Ti
        for t in steps:
            valid = (t < len_t).to(dtype).unsqueeze(1)               # [B,1]
            h_new, c_new = lstm_cell(emb[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            # A padded position leaves the state untouched (the packed sequence simply does not
            # contain it) and contributes a zero output row.
            h = valid * h_new + (1 - valid) * h
            c = valid * c_new + (1 - valid) * c
            dir_out[t] = valid * h_new
        outs = outs + torch.stack(dir_out, dim=0)
        finals.append(h)
    hidden = finals[0] + finals[1]
    return hidden, outs


# ----------------------------------------------------------------------------------------
# Attention  (seq2seq_model.py:105-139; mask helpers.py:11-32)
# ----------------------------------------------------------------------------------------
def attention(query: torch.Tensor, keys: torch.Tensor, w_q: torch.Tensor, v: torch.Tensor,
              lengths: Optional[torch.Tensor]):
    """query [B,Q]; keys = values = projected keys [B,N,H]; returns (context [B,H], weights [B,N])."""
    q = query @ w_q.t()                                               # [B,H]
    scores = torch.tanh(q.unsqueeze(1) + keys) @ v.view(-1)           # [B,N]
    if lengths is not None:
        N = keys.shape[1]
        mask = torch.arange(N).unsqueeze(0) < lengths.unsqueeze(1)
        scores = scores.masked_fill(~mask, float("-inf"))
    weights = torch.softmax(scores, dim=1)
    context = torch.einsum("bn,bnh->bh", weights, keys)
    return context, weights


# ----------------------------------------------------------------------------------------
# Decoder  (seq2seq_model.py:359-428, 433-490)
# ----------------------------------------------------------------------------------------
def decoder_step(p: Params, emb: torch.Tensor, h: torch.Tensor, c: torch.Tensor,
                 keys_text: torch.Tensor, cmd_lengths: torch.Tensor, keys_vis: torch.Tensor,
                 conditional_attention: bool):
    """One step of SURVEY.md 3.3.  emb [B,H] is the (already dropped-out) token embedding;
    keys_text [B,Ti,H] and keys_vis [B,M,H] are the PROJECTED keys, which also serve as values
    (seq2seq_model.py:388-390, 400-402).  Returns (logits, h, c, alpha, beta)."""
    c_t, alpha = attention(h, keys_text, p["textual_attention.query_layer.weight"],
                           p["textual_attention.energy_layer.weight"], cmd_lengths)
    if conditional_attention:
        q = torch.tanh(torch.cat([h, c_t], dim=1) @ p["attention_decoder.queries_to_keys.weight"].t()
                       + p["attention_decoder.queries_to_keys.bias"])
    else:
        q = h
    c_v, beta = attention(q, keys_vis, p["visual_attention.query_layer.weight"],
                          p["visual_attention.energy_layer.weight"], None)
    x = torch.cat([emb, c_t, c_v], dim=1)
    h_new, c_new = lstm_cell(x, h, c, p["attention_decoder.lstm.weight_ih_l0"],
                             p["attention_decoder.lstm.weight_hh_l0"],
                             p["attention_decoder.lstm.bias_ih_l0"],
                             p["attention_decoder.lstm.bias_hh_l0"])
    u = torch.cat([emb, h_new, c_t, c_v], dim=1)
    pre = u @ p["attention_decoder.output_to_hidden.weight"].t()
    logits = pre @ p["attention_decoder.hidden_to_output.weight"].t()
    return logits, h_new, c_new, alpha, beta


def project_keys(p: Params, enc_out: torch.Tensor, feat: torch.Tensor):
    """Key projections done once per sequence (seq2seq_model.py:466-469).
    enc_out [Ti,B,H] -> keys_text [B,Ti,H];  feat [B,M,D] -> keys_vis [B,M,H]."""
    keys_text = (enc_out @ p["textual_attention.key_layer.weight"].t()).transpose(0, 1)
    keys_vis = feat @ p["visual_attention.key_layer.weight"].t()
    return keys_text, keys_vis


def initial_state(p: Params, hidden: torch.Tensor):
    """h0 = c0 = tanh(W_e2d h_enc + b)  (model.py:195-196; seq2seq_model.py:494-504)."""
    h0 = torch.tanh(hidden @ p["enc_hidden_to_dec_hidden.weight"].t()
                    + p["enc_hidden_to_dec_hidden.bias"])
    return h0, h0.clone()


def model_forward(p: Params, commands: torch.Tensor, cmd_lengths: Sequence[int],
                  situations: torch.Tensor, targets: torch.Tensor,
                  conditional_attention: bool = True, auxiliary_task: bool = False,
                  dropout: Optional[Dict[str, torch.Tensor]] = None):
    """Teacher-forced forward of the whole model (model.py:206-219).

    Returns (logp [B,Tt,V], aux_logp [B,M] or None).  The decoder runs over ALL padded target
    steps for every example and the aux scores sum beta over all of them.
    ``dropout`` optionally carries pre-scaled masks 'cnn' [B,M,D], 'enc' [B,Ti,E], 'dec' [B,Tt,H].
    """
    dropout = dropout or {}
    feat = cnn_forward(p, situations, dropout.get("cnn"))
    hidden, enc_out = encoder_forward(p, commands, cmd_lengths, dropout.get("enc"))
    keys_text, keys_vis = project_keys(p, enc_out, feat)
    h, c = initial_state(p, hidden)
    len_t = torch.tensor([int(l) for l in cmd_lengths])
    emb_all = embed(p["attention_decoder.embedding.weight"], targets)  # [B,Tt,H]
    if dropout.get("dec") is not None:
        emb_all = emb_all * dropout["dec"]
    logits, betas = [], []
    for t in range(targets.shape[1]):
        lg, h, c, _alpha, beta = decoder_step(p, emb_all[:, t], h, c, keys_text, len_t, keys_vis,
                                              conditional_attention)
        logits.append(lg)
        betas.append(beta)
    logits = torch.stack(logits, dim=1)                               # [B,Tt,V]
    logp = torch.log_softmax(logits, dim=-1)
    aux = torch.log_softmax(torch.stack(betas, 0).sum(0), dim=-1) if auxiliary_task else None
    return logp, aux


# ----------------------------------------------------------------------------------------
# Loss and metrics  (model.py:108-160)
# ----------------------------------------------------------------------------------------
def shift_targets(targets: torch.Tensor) -> torch.Tensor:
    """Drop SOS, append one pad column (model.py:108-115)."""
    return torch.cat([targets[:, 1:], torch.zeros(targets.shape[0], 1, dtype=targets.dtype)], dim=1)


def nll_loss(logp: torch.Tensor, targets: torch.Tensor, pad_idx: int = 0) -> torch.Tensor:
    """Mean over non-pad shifted targets of -logp (NLLLoss(ignore_index), model.py:100,147-160)."""
    tgt = shift_targets(targets)
    picked = torch.gather(logp, 2, tgt.unsqueeze(-1)).squeeze(-1)
    mask = (tgt != pad_idx).to(logp.dtype)
    return -(picked * mask).sum() / mask.sum()


def aux_nll_loss(aux_logp: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    """Plain mean NLL over the batch (model.py:59,162-164)."""
    return -torch.gather(aux_logp, 1, positions.view(-1, 1)).mean()


def metrics(logp: torch.Tensor, targets: torch.Tensor, pad_idx: int = 0) -> Tuple[float, float]:
    """(token accuracy %, exact match %) as model.py:117-137."""
    tgt = shift_targets(targets)
    mask = tgt != pad_idx
    pred = logp.max(dim=2)[1]
    match = (pred == tgt) & mask
    total = int(mask.sum())
    exact = 100.0 * int((match.sum(1) == mask.sum(1)).sum()) / tgt.shape[0]
    return 100.0 * int(match.sum()) / total, exact


# ----------------------------------------------------------------------------------------
# Greedy decoding  (predict.py:57-128) and sequence accuracy (helpers.py:44-64)
# ----------------------------------------------------------------------------------------
def greedy_decode(p: Params, commands: torch.Tensor, cmd_lengths: Sequence[int],
                  situations: torch.Tensor, max_decoding_steps: int, sos_idx: int = 1,
                  eos_idx: int = 2, conditional_attention: bool = True):
    """Batched restatement of the reference's batch-size-1 loop, one example at a time in effect.

    Per example: start from SOS, take argmax(log_softmax(logits)) (first max on ties), stop
    after producing EOS or after max_decoding_steps + 1 tokens (``<=`` at predict.py:101); a
    trailing EOS and its attention rows are dropped (114-117).  Returns per-example lists
    ``sequences`` (token ids), ``alphas`` [n_i, len_i], ``betas`` [n_i, M] and ``beta_sum`` [B,M]:
    the sum of beta over every step that was executed INCLUDING the EOS-producing one
    (predict.py:111,119).
    Note the batch-1 reference masks text attention at the example's own length, with
    Ti = that length; extra padded key columns carry exactly zero weight, so batching is exact.
    """
    with torch.no_grad():
        feat = cnn_forward(p, situations)
        hidden, enc_out = encoder_forward(p, commands, cmd_lengths)
        keys_text, keys_vis = project_keys(p, enc_out, feat)
        h, c = initial_state(p, hidden)
        B = commands.shape[0]
        len_t = torch.tensor([int(l) for l in cmd_lengths])
        token = torch.full((B,), sos_idx, dtype=torch.long)
        alive = torch.ones(B, dtype=torch.bool)
        seqs: List[List[int]] = [[] for _ in range(B)]
        alphas: List[List[np.ndarray]] = [[] for _ in range(B)]
        betas: List[List[np.ndarray]] = [[] for _ in range(B)]
        beta_sum = torch.zeros(B, keys_vis.shape[1], dtype=keys_vis.dtype)
        emb_w = p["attention_decoder.embedding.weight"]
        for _ in range(max_decoding_steps + 1):
            if not bool(alive.any()):
                break
            logits, h, c, alpha, beta = decoder_step(p, embed(emb_w, token), h, c, keys_text, len_t,
                                                     keys_vis, conditional_attention)
            nxt = torch.log_softmax(logits, dim=-1).max(dim=-1)[1]
            for b in range(B):
                if alive[b]:
                    beta_sum[b] += beta[b]
                    tok = int(nxt[b])
                    if tok == eos_idx:
                        alive[b] = False
                    else:
                        seqs[b].append(tok)
                        alphas[b].append(alpha[b, :int(len_t[b])].numpy().copy())
                        betas[b].append(beta[b].numpy().copy())
            token = nxt
    return seqs, alphas, betas, beta_sum


def sequence_accuracy(prediction: List[int], target: List[int]) -> float:
    """helpers.py:44-64: pad the shorter list (prediction with 0, target with -1), compare."""
    n = max(len(prediction), len(target))
    if n == 0:
        return 0.0
    pred = list(prediction) + [0] * (n - len(prediction))
    tgt = list(target) + [-1] * (n - len(target))
    return 100.0 * sum(int(a == b) for a, b in zip(pred, tgt)) / n


# ----------------------------------------------------------------------------------------
# Parameter registry, named configurations and deterministic synthetic data live with the product
# (they are also what bench.py feeds the kernels); re-exported here for the tests.
# ----------------------------------------------------------------------------------------
from multimodal_seq2seq_gscan_b200.synthetic import (  # noqa: E402,F401
    CONFIGS, model_kwargs, param_shapes, synthetic_batch, synthetic_params)
