"""Hand-derived forward + backward of the whole path in numpy float64.  TEST INFRASTRUCTURE ONLY.

This mirrors, buffer for buffer, the plan the CUDA kernels in
``multimodal_seq2seq_gscan_b200/csrc`` follow (DESIGN.md "Kernel plan"): the same saved
activations, the same split into batched GEMMs and recurrent sweeps, the same BPTT recipe
(SURVEY.md A.6).  It exists so that (1) the derivation is checked against autograd of
``gscan_oracle`` on CPU (tests/test_manual_backward.py) before it is transliterated to CUDA, and
(2) GPU tests can compare individual workspace buffers, not just final outputs.

Nothing on the product path imports this file.
"""
from __future__ import annotations

import numpy as np


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def conv_weight_nhwc(w):
    """[F,C,k,k] reference conv weight -> [k(row off),k(col off),C,F] taps for an NHWC(row,col)
    grid; the reference convolves the transposed grid (cnn_model.py:28), so tap [f,c,i,j]
    applies to row offset j, col offset i."""
    return np.transpose(w, (3, 2, 1, 0)).copy()


def forward(p, batch, cfg, dropout=None):
    """Returns (outputs, saved) with saved activations named as in the CUDA workspace."""
    dropout = dropout or {}
    P = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}
    cond = cfg["conditional_attention"]
    x = np.asarray(batch["situations"], dtype=np.float64)
    cmds = np.asarray(batch["commands"])
    tgts = np.asarray(batch["targets"])
    cl = np.asarray(batch["cmd_lengths"]).astype(np.int64)
    B, G, _, C = x.shape
    M = G * G
    Ti = int(cl.max())
    Tt = tgts.shape[1]
    H = P["attention_decoder.lstm.weight_hh_l0"].shape[1]
    S = {"B": B, "M": M, "Ti": Ti, "Tt": Tt, "H": H, "G": G, "cl": cl}

    # F1 situation CNN -> feat [B,M,D]
    convs = []
    for name in ("conv_1", "conv_2", "conv_3"):
        w = conv_weight_nhwc(P[f"situation_encoder.{name}.weight"])
        k = w.shape[0]
        pd = k // 2
        xp = np.zeros((B, G + 2 * pd, G + 2 * pd, C))
        xp[:, pd:pd + G, pd:pd + G] = x
        o = np.zeros((B, G, G, w.shape[3])) + P[f"situation_encoder.{name}.bias"]
        for dr in range(k):
            for dc in range(k):
                o += xp[:, dr:dr + G, dc:dc + G, :] @ w[dr, dc]
        convs.append(o)
    feat = np.maximum(np.concatenate(convs, -1), 0.0).reshape(B, M, -1)
    if dropout.get("cnn") is not None:
        feat = feat * dropout["cnn"]
    S["feat"] = feat
    # F2 visual keys
    KV = feat @ P["visual_attention.key_layer.weight"].T                   # [B,M,H]
    S["KV"] = KV

    # F3/F4 encoder
    emb_w = P["encoder.embedding.weight"]
    enc_x = emb_w[cmds[:, :Ti]]                                              # [B,Ti,E]
    if dropout.get("enc") is not None:
        enc_x = enc_x * dropout["enc"][:, :Ti]
    S["enc_x"] = enc_x
    enc_out = np.zeros((Ti, B, H))
    enc_h = np.zeros((2, Ti, B, H))
    enc_c = np.zeros((2, Ti, B, H))
    enc_g = np.zeros((2, Ti, B, 4 * H))
    h_enc = np.zeros((B, H))
    for d, suffix in enumerate(("", "_reverse")):
        w_ih = P[f"encoder.lstm.weight_ih_l0{suffix}"]
        w_hh = P[f"encoder.lstm.weight_hh_l0{suffix}"]
        bias = P[f"encoder.lstm.bias_ih_l0{suffix}"] + P[f"encoder.lstm.bias_hh_l0{suffix}"]
        xg = enc_x @ w_ih.T + bias                                           # [B,Ti,4H]
        h = np.zeros((B, H))
        c = np.zeros((B, H))
        order = range(Ti) if d == 0 else range(Ti - 1, -1, -1)
        for t in order:
            valid = (t < cl)[:, None]
            a = xg[:, t] + h @ w_hh.T
            i, f, g, o = _sig(a[:, :H]), _sig(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), _sig(a[:, 3 * H:])
            cn = f * c + i * g
            hn = o * np.tanh(cn)
            enc_g[d, t] = np.concatenate([i, f, g, o], 1)
            h = np.where(valid, hn, h)
            c = np.where(valid, cn, c)
            enc_h[d, t] = h       # state AFTER position t (carried through pads)
            enc_c[d, t] = c
            enc_out[t] += np.where(valid, hn, 0.0)
        h_enc += h
    S.update(enc_out=enc_out, enc_h=enc_h, enc_c=enc_c, enc_g=enc_g, h_enc=h_enc)

    # F5 textual keys + initial state
    KT = enc_out @ P["textual_attention.key_layer.weight"].T               # [Ti,B,H]
    h0 = np.tanh(h_enc @ P["enc_hidden_to_dec_hidden.weight"].T + P["enc_hidden_to_dec_hidden.bias"])
    S.update(KT=KT, h0=h0)

    # F6 decoder embeddings (time-major) and their input-gate contribution
    E_all = P["attention_decoder.embedding.weight"][tgts]                  # [B,Tt,H]
    if dropout.get("dec") is not None:
        E_all = E_all * dropout["dec"]
    E_all = np.transpose(E_all, (1, 0, 2))                                   # [Tt,B,H]
    w_ih = P["attention_decoder.lstm.weight_ih_l0"]
    w_hh = P["attention_decoder.lstm.weight_hh_l0"]
    Xe = E_all @ w_ih[:, :H].T + P["attention_decoder.lstm.bias_ih_l0"] + P["attention_decoder.lstm.bias_hh_l0"]

    # F7 recurrent sweep
    U = np.zeros((Tt + 1, B, 4 * H))          # rows [e | h | cT | cV]; group 0 holds h0 in the h block
    U[0, :, H:2 * H] = h0
    U[1:, :, 0:H] = E_all
    Cs = np.zeros((Tt + 1, B, H))
    Cs[0] = h0
    gates = np.zeros((Tt, B, 4 * H))
    alpha = np.zeros((Tt, B, Ti))
    beta = np.zeros((Tt, B, M))
    Qp = np.zeros((Tt, B, H))
    qT = np.zeros((Tt, B, H))
    qV = np.zeros((Tt, B, H))
    vT = P["textual_attention.energy_layer.weight"].reshape(-1)
    vV = P["visual_attention.energy_layer.weight"].reshape(-1)
    WqT = P["textual_attention.query_layer.weight"]
    WqV = P["visual_attention.query_layer.weight"]
    if cond:
        Wc = P["attention_decoder.queries_to_keys.weight"]
        bc = P["attention_decoder.queries_to_keys.bias"]
    KTb = np.transpose(KT, (1, 0, 2))                                        # [B,Ti,H]
    tmask = np.arange(Ti)[None, :] < cl[:, None]
    for t in range(Tt):
        h = U[t, :, H:2 * H]
        c = Cs[t]
        qT[t] = h @ WqT.T
        s = np.tanh(qT[t][:, None, :] + KTb) @ vT
        s = np.where(tmask, s, -np.inf)
        s = np.exp(s - s.max(1, keepdims=True))
        alpha[t] = s / s.sum(1, keepdims=True)
        cT = np.einsum("bj,bjh->bh", alpha[t], KTb)
        if cond:
            Qp[t] = np.tanh(np.concatenate([h, cT], 1) @ Wc.T + bc)
        else:
            Qp[t] = h
        qV[t] = Qp[t] @ WqV.T
        r = np.tanh(qV[t][:, None, :] + KV) @ vV
        r = np.exp(r - r.max(1, keepdims=True))
        beta[t] = r / r.sum(1, keepdims=True)
        cV = np.einsum("bm,bmh->bh", beta[t], KV)
        a = Xe[t] + np.concatenate([cT, cV], 1) @ w_ih[:, H:].T + h @ w_hh.T
        i, f, g, o = _sig(a[:, :H]), _sig(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), _sig(a[:, 3 * H:])
        Cs[t + 1] = f * c + i * g
        hn = o * np.tanh(Cs[t + 1])
        gates[t] = np.concatenate([i, f, g, o], 1)
        U[t + 1, :, H:2 * H] = hn
        U[t + 1, :, 2 * H:3 * H] = cT
        U[t + 1, :, 3 * H:] = cV
    S.update(U=U, Cs=Cs, gates=gates, alpha=alpha, beta=beta, Qp=Qp, qT=qT, qV=qV, Xe=Xe)

    # F8/F9 output projection + log-softmax
    pre = U[1:] @ P["attention_decoder.output_to_hidden.weight"].T         # [Tt,B,H]
    logits = pre @ P["attention_decoder.hidden_to_output.weight"].T        # [Tt,B,V]
    z = logits - logits.max(-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(-1, keepdims=True))
    S.update(pre=pre, logp_tm=logp)
    out = {"logp": np.transpose(logp, (1, 0, 2))}
    # F10 aux
    bsum = beta.sum(0)
    z = bsum - bsum.max(-1, keepdims=True)
    out["aux_logp"] = z - np.log(np.exp(z).sum(-1, keepdims=True))
    S["beta_sum"] = bsum
    return out, S


def backward(p, batch, cfg, S, dlogp, daux=None, dropout=None):
    """dlogp [B,Tt,V], daux [B,M] (or None): upstream gradients.  Returns dict of param grads."""
    dropout = dropout or {}
    P = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}
    cond = cfg["conditional_attention"]
    B, M, Ti, Tt, H, G, cl = S["B"], S["M"], S["Ti"], S["Tt"], S["H"], S["G"], S["cl"]
    x = np.asarray(batch["situations"], dtype=np.float64)
    cmds = np.asarray(batch["commands"])
    tgts = np.asarray(batch["targets"])
    g = {k: np.zeros_like(v) for k, v in P.items()}

    # B1 log-softmax backward, hidden_to_output
    dlogp_tm = np.transpose(dlogp, (1, 0, 2))
    dlogits = dlogp_tm - np.exp(S["logp_tm"]) * dlogp_tm.sum(-1, keepdims=True)
    Wh2o = P["attention_decoder.hidden_to_output.weight"]
    Wo2h = P["attention_decoder.output_to_hidden.weight"]
    dpre = dlogits @ Wh2o
    g["attention_decoder.hidden_to_output.weight"] = np.einsum("tbv,tbh->vh", dlogits, S["pre"])
    # B2 output_to_hidden
    U = S["U"]
    dU = dpre @ Wo2h                                                        # [Tt,B,4H]
    g["attention_decoder.output_to_hidden.weight"] = np.einsum("tbh,tbk->hk", dpre, U[1:])
    # B3 aux log-softmax backward -> the same d(beta) at every step
    if daux is not None:
        bs = S["beta_sum"]
        sm = np.exp(bs - bs.max(-1, keepdims=True))
        sm /= sm.sum(-1, keepdims=True)
        dbeta_aux = daux - sm * daux.sum(-1, keepdims=True)
    else:
        dbeta_aux = np.zeros((B, M))

    # B4 reverse-time sweep
    w_ih = P["attention_decoder.lstm.weight_ih_l0"]
    w_hh = P["attention_decoder.lstm.weight_hh_l0"]
    vT = P["textual_attention.energy_layer.weight"].reshape(-1)
    vV = P["visual_attention.energy_layer.weight"].reshape(-1)
    WqT = P["textual_attention.query_layer.weight"]
    WqV = P["visual_attention.query_layer.weight"]
    if cond:
        Wc = P["attention_decoder.queries_to_keys.weight"]
    KV = S["KV"]
    KTb = np.transpose(S["KT"], (1, 0, 2))
    tmask = np.arange(Ti)[None, :] < cl[:, None]
    dgates = np.zeros((Tt, B, 4 * H))
    dd = np.zeros((Tt, B, H))
    dqV = np.zeros((Tt, B, H))
    dqT = np.zeros((Tt, B, H))
    dKV = np.zeros((B, M, H))
    dKTb = np.zeros((B, Ti, H))
    dvT = np.zeros(H)
    dvV = np.zeros(H)
    dh = np.zeros((B, H))
    dc = np.zeros((B, H))
    for t in range(Tt - 1, -1, -1):
        i, f, gg, o = (S["gates"][t][:, :H], S["gates"][t][:, H:2 * H], S["gates"][t][:, 2 * H:3 * H],
                       S["gates"][t][:, 3 * H:])
        c_prev, c_new = S["Cs"][t], S["Cs"][t + 1]
        h_prev = U[t, :, H:2 * H]
        cT, cV = U[t + 1, :, 2 * H:3 * H], U[t + 1, :, 3 * H:]
        # LSTM cell
        dh_t = dh + dU[t, :, H:2 * H]
        tc = np.tanh(c_new)
        do = dh_t * tc
        dc_t = dc + dh_t * o * (1 - tc * tc)
        da = np.concatenate([dc_t * gg * i * (1 - i), dc_t * c_prev * f * (1 - f),
                             dc_t * i * (1 - gg * gg), do * o * (1 - o)], 1)
        dgates[t] = da
        dc = dc_t * f
        dx = da @ w_ih[:, H:]                                               # [B,2H] -> cT, cV
        dh = da @ w_hh
        dcT = dx[:, :H] + dU[t, :, 2 * H:3 * H]
        dcV = dx[:, H:] + dU[t, :, 3 * H:]
        # visual attention
        b_ = S["beta"][t]
        dbeta = np.einsum("bh,bmh->bm", dcV, KV) + dbeta_aux
        dKV += b_[:, :, None] * dcV[:, None, :]
        dr = b_ * (dbeta - (dbeta * b_).sum(1, keepdims=True))
        zz = np.tanh(S["qV"][t][:, None, :] + KV)
        dvV += np.einsum("bm,bmh->h", dr, zz)
        dzpre = dr[:, :, None] * vV[None, None, :] * (1 - zz * zz)
        dKV += dzpre
        dqV[t] = dzpre.sum(1)
        dQp = dqV[t] @ WqV
        # conditional query
        if cond:
            dd[t] = dQp * (1 - S["Qp"][t] ** 2)
            dhc = dd[t] @ Wc
            dh = dh + dhc[:, :H]
            dcT = dcT + dhc[:, H:]
        else:
            dh = dh + dQp
        # textual attention
        a_ = S["alpha"][t]
        dalpha = np.einsum("bh,bjh->bj", dcT, KTb)
        dKTb += a_[:, :, None] * dcT[:, None, :]
        ds = a_ * (dalpha - (dalpha * a_).sum(1, keepdims=True))
        ds = np.where(tmask, ds, 0.0)
        zz = np.tanh(S["qT"][t][:, None, :] + KTb)
        dvT += np.einsum("bj,bjh->h", ds, zz)
        dzpre = ds[:, :, None] * vT[None, None, :] * (1 - zz * zz)
        dKTb += dzpre
        dqT[t] = dzpre.sum(1)
        dh = dh + dqT[t] @ WqT
    dh0 = dh + dc

    # B5 decoder weight gradients as batched GEMMs over the saved per-step quantities
    Hprev = U[:-1, :, H:2 * H]
    X_c = U[1:, :, 2 * H:]                                                   # [cT | cV]
    E_all = U[1:, :, :H]
    gw = np.zeros_like(w_ih)
    gw[:, :H] = np.einsum("tbg,tbh->gh", dgates, E_all)
    gw[:, H:] = np.einsum("tbg,tbk->gk", dgates, X_c)
    g["attention_decoder.lstm.weight_ih_l0"] = gw
    g["attention_decoder.lstm.weight_hh_l0"] = np.einsum("tbg,tbh->gh", dgates, Hprev)
    g["attention_decoder.lstm.bias_ih_l0"] = dgates.sum((0, 1))
    g["attention_decoder.lstm.bias_hh_l0"] = dgates.sum((0, 1))
    g["textual_attention.query_layer.weight"] = np.einsum("tbo,tbh->oh", dqT, Hprev)
    g["visual_attention.query_layer.weight"] = np.einsum("tbo,tbh->oh", dqV, S["Qp"])
    g["textual_attention.energy_layer.weight"] = dvT.reshape(1, -1)
    g["visual_attention.energy_layer.weight"] = dvV.reshape(1, -1)
    if cond:
        gc = np.zeros_like(Wc)
        gc[:, :H] = np.einsum("tbo,tbh->oh", dd, Hprev)
        gc[:, H:] = np.einsum("tbo,tbh->oh", dd, U[1:, :, 2 * H:3 * H])
        g["attention_decoder.queries_to_keys.weight"] = gc
        g["attention_decoder.queries_to_keys.bias"] = dd.sum((0, 1))
    # decoder embedding: dE = dU[:, :, :H] + dgates @ W_ih[:, :H]; scatter by token (skip pad)
    dE = dU[:, :, :H] + dgates @ w_ih[:, :H]                                 # [Tt,B,H]
    dE = np.transpose(dE, (1, 0, 2))
    if dropout.get("dec") is not None:
        dE = dE * dropout["dec"]
    ge = np.zeros_like(P["attention_decoder.embedding.weight"])
    np.add.at(ge, tgts.reshape(-1), dE.reshape(-1, H))
    ge[0] = 0.0
    g["attention_decoder.embedding.weight"] = ge

    # B6 visual keys -> CNN
    feat = S["feat"]
    g["visual_attention.key_layer.weight"] = np.einsum("bmh,bmd->hd", dKV, feat)
    dfeat = dKV @ P["visual_attention.key_layer.weight"]                   # [B,M,D]
    if dropout.get("cnn") is not None:
        dfeat = dfeat * dropout["cnn"]
    dconv = (dfeat * (feat > 0)).reshape(B, G, G, -1)
    C = x.shape[3]
    off = 0
    for name in ("conv_1", "conv_2", "conv_3"):
        w = P[f"situation_encoder.{name}.weight"]
        F_, _, k, _ = w.shape
        pd = k // 2
        dy = dconv[..., off:off + F_]
        off += F_
        xp = np.zeros((B, G + 2 * pd, G + 2 * pd, C))
        xp[:, pd:pd + G, pd:pd + G] = x
        gw = np.zeros((k, k, C, F_))
        for dr_ in range(k):
            for dc_ in range(k):
                gw[dr_, dc_] = np.einsum("brcq,brcf->qf", xp[:, dr_:dr_ + G, dc_:dc_ + G, :], dy)
        g[f"situation_encoder.{name}.weight"] = np.transpose(gw, (3, 2, 1, 0))
        g[f"situation_encoder.{name}.bias"] = dy.sum((0, 1, 2))

    # B7 textual keys, initial state
    dKT = np.transpose(dKTb, (1, 0, 2))                                      # [Ti,B,H]
    g["textual_attention.key_layer.weight"] = np.einsum("tbo,tbh->oh", dKT, S["enc_out"])
    denc_out = dKT @ P["textual_attention.key_layer.weight"]
    dpre0 = dh0 * (1 - S["h0"] ** 2)
    g["enc_hidden_to_dec_hidden.weight"] = dpre0.T @ S["h_enc"]
    g["enc_hidden_to_dec_hidden.bias"] = dpre0.sum(0)
    dh_enc = dpre0 @ P["enc_hidden_to_dec_hidden.weight"]

    # B8 encoder BPTT (both directions receive denc_out at valid positions and dh_enc at their end)
    enc_x = S["enc_x"]
    denc_x = np.zeros_like(enc_x)
    for d, suffix in enumerate(("", "_reverse")):
        w_ih_e = P[f"encoder.lstm.weight_ih_l0{suffix}"]
        w_hh_e = P[f"encoder.lstm.weight_hh_l0{suffix}"]
        dga = np.zeros((Ti, B, 4 * H))
        hprev_all = np.zeros((Ti, B, H))
        dh = dh_enc.copy()
        dc = np.zeros((B, H))
        order = list(range(Ti)) if d == 0 else list(range(Ti - 1, -1, -1))
        for idx in range(Ti - 1, -1, -1):
            t = order[idx]
            valid = (t < cl)[:, None]
            if idx > 0:
                h_prev, c_prev = S["enc_h"][d, order[idx - 1]], S["enc_c"][d, order[idx - 1]]
            else:
                h_prev, c_prev = np.zeros((B, H)), np.zeros((B, H))
            hprev_all[t] = h_prev
            gt = S["enc_g"][d, t]
            i, f, gg, o = gt[:, :H], gt[:, H:2 * H], gt[:, 2 * H:3 * H], gt[:, 3 * H:]
            c_new = f * c_prev + i * gg        # recomputed; equals enc_c at valid positions
            tc = np.tanh(c_new)
            dh_t = dh + denc_out[t]
            do = dh_t * tc
            dc_t = dc + dh_t * o * (1 - tc * tc)
            da = np.concatenate([dc_t * gg * i * (1 - i), dc_t * c_prev * f * (1 - f),
                                 dc_t * i * (1 - gg * gg), do * o * (1 - o)], 1)
            da = np.where(valid, da, 0.0)
            dga[t] = da
            # padded position: state passes straight through
            dh = np.where(valid, da @ w_hh_e, dh)
            dc = np.where(valid, dc_t * f, dc)
        g[f"encoder.lstm.weight_ih_l0{suffix}"] = np.einsum("tbg,bte->ge", dga, enc_x)
        g[f"encoder.lstm.weight_hh_l0{suffix}"] = np.einsum("tbg,tbh->gh", dga, hprev_all)
        g[f"encoder.lstm.bias_ih_l0{suffix}"] = dga.sum((0, 1))
        g[f"encoder.lstm.bias_hh_l0{suffix}"] = dga.sum((0, 1))
        denc_x += np.transpose(dga @ w_ih_e, (1, 0, 2))
    if dropout.get("enc") is not None:
        denc_x = denc_x * dropout["enc"][:, :Ti]
    gemb = np.zeros_like(P["encoder.embedding.weight"])
    np.add.at(gemb, cmds[:, :Ti].reshape(-1), denc_x.reshape(-1, enc_x.shape[2]))
    gemb[0] = 0.0
    g["encoder.embedding.weight"] = gemb
    return g


def loss_and_upstream(out, batch, auxiliary_task, weight_target_loss=0.3, pad_idx=0):
    """Reference loss (model.py:147-164; train.py:102-107) and its gradient w.r.t. logp / aux_logp."""
    tg = np.asarray(batch["targets"])
    B, Tt = tg.shape
    sh = np.concatenate([tg[:, 1:], np.zeros((B, 1), dtype=tg.dtype)], 1)
    mask = sh != pad_idx
    n = mask.sum()
    logp = out["logp"]
    picked = np.take_along_axis(logp, sh[..., None], 2)[..., 0]
    loss = -(picked * mask).sum() / n
    dlogp = np.zeros_like(logp)
    np.put_along_axis(dlogp, sh[..., None], (-(mask / n))[..., None], 2)
    daux = None
    if auxiliary_task:
        pos = np.asarray(batch["target_positions"])
        aux = out["aux_logp"]
        loss = loss + weight_target_loss * (-aux[np.arange(B), pos].mean())
        daux = np.zeros_like(aux)
        daux[np.arange(B), pos] = -weight_target_loss / B
    return loss, dlogp, daux
