"""Deterministic synthetic gSCAN-shaped parameters and batches and the named model
configurations of BASELINE.json.  Used by bench.py, the tests and the CPU oracle; the real data
blobs of the reference are not available offline (reference .MISSING_LARGE_BLOBS)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]


def param_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names and shapes in ``model.parameters()`` order (Adam index order)."""
    C, F, k3 = cfg["num_cnn_channels"], cfg["cnn_hidden_num_channels"], cfg["cnn_kernel_size"]
    E, H = cfg["embedding_dimension"], cfg["decoder_hidden_size"]
    Vi, V = cfg["input_vocabulary_size"], cfg["target_vocabulary_size"]
    assert cfg["encoder_hidden_size"] == H
    D = 3 * F
    shapes = [
        ("situation_encoder.conv_1.weight", (F, C, 1, 1)), ("situation_encoder.conv_1.bias", (F,)),
        ("situation_encoder.conv_2.weight", (F, C, 5, 5)), ("situation_encoder.conv_2.bias", (F,)),
        ("situation_encoder.conv_3.weight", (F, C, k3, k3)), ("situation_encoder.conv_3.bias", (F,)),
        ("visual_attention.key_layer.weight", (H, D)),
        ("visual_attention.query_layer.weight", (H, H)),
        ("visual_attention.energy_layer.weight", (1, H)),
        ("encoder.embedding.weight", (Vi, E)),
    ]
    for suffix in ("", "_reverse"):
        shapes += [(f"encoder.lstm.weight_ih_l0{suffix}", (4 * H, E)),
                   (f"encoder.lstm.weight_hh_l0{suffix}", (4 * H, H)),
                   (f"encoder.lstm.bias_ih_l0{suffix}", (4 * H,)),
                   (f"encoder.lstm.bias_hh_l0{suffix}", (4 * H,))]
    shapes += [
        ("enc_hidden_to_dec_hidden.weight", (H, H)), ("enc_hidden_to_dec_hidden.bias", (H,)),
        ("textual_attention.key_layer.weight", (H, H)),
        ("textual_attention.query_layer.weight", (H, H)),
        ("textual_attention.energy_layer.weight", (1, H)),
    ]
    if cfg.get("conditional_attention", True):
        shapes += [("attention_decoder.queries_to_keys.weight", (H, 2 * H)),
                   ("attention_decoder.queries_to_keys.bias", (H,))]
    shapes += [
        ("attention_decoder.embedding.weight", (V, H)),
        ("attention_decoder.lstm.weight_ih_l0", (4 * H, 3 * H)),
        ("attention_decoder.lstm.weight_hh_l0", (4 * H, H)),
        ("attention_decoder.lstm.bias_ih_l0", (4 * H,)),
        ("attention_decoder.lstm.bias_hh_l0", (4 * H,)),
        ("attention_decoder.output_to_hidden.weight", (H, 4 * H)),
        ("attention_decoder.hidden_to_output.weight", (V, H)),
    ]
    return shapes


def synthetic_params(cfg: dict, seed: int, scale: float = 1.0, dtype=torch.float32) -> Params:
    """Deterministic parameters from numpy's PCG64 (identical on every machine, unlike
    torch's default init).  Uniform(-s, s) with s = scale / sqrt(fan_in); embeddings N(0,1)*0.5
    with the padding row (index 0) zeroed as nn.Embedding(padding_idx=0) does."""
    rng = np.random.default_rng(seed)
    out: Params = {}
    for name, shape in param_shapes(cfg):
        if "embedding" in name:
            w = rng.standard_normal(shape) * 0.5
            w[0] = 0.0
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            s = scale / np.sqrt(fan_in)
            w = rng.uniform(-s, s, size=shape)
        out[name] = torch.tensor(w, dtype=dtype)
    return out


def synthetic_batch(cfg: dict, batch_size: int, seed: int, max_cmd_len: int = 10, min_cmd_len: int = 5,
                    max_tgt_len: int = 121, min_tgt_len: int = 3, force_max: bool = True,
                    bernoulli_situations: bool = False) -> dict:
    """Synthetic gSCAN-shaped batch as specified in SURVEY.md 8(d).  Lengths include SOS/EOS;
    example 0 is forced to the maximum lengths so Ti/Tt are fixed; PAD 0, SOS 1, EOS 2."""
    rng = np.random.default_rng(seed)
    G, C = cfg["grid_size"], cfg["num_cnn_channels"]
    Vi, V = cfg["input_vocabulary_size"], cfg["target_vocabulary_size"]
    B = batch_size

    def seqs(lo, hi, vocab):
        lens = rng.integers(lo, hi + 1, size=B)
        if force_max:
            lens[0] = hi
        T = int(lens.max())
        arr = np.zeros((B, T), dtype=np.int64)
        for b in range(B):
            n = int(lens[b])
            arr[b, 0] = 1
            if n > 2:
                arr[b, 1:n - 1] = rng.integers(3, vocab, size=n - 2)
            arr[b, n - 1] = 2
        return arr, lens.astype(np.float64)

    commands, cmd_len = seqs(min_cmd_len, max_cmd_len, Vi)
    targets, tgt_len = seqs(min_tgt_len, max_tgt_len, V)
    if bernoulli_situations or C < 12:
        sit = (rng.random((B, G, G, C)) < 0.1).astype(np.float32)
    else:
        sit = np.zeros((B, G, G, C), dtype=np.float32)
        n_dir = 4
        agent_ch = C - n_dir - 1
        for b in range(B):
            cells = rng.permutation(G * G)
            r, c = divmod(int(cells[0]), G)
            sit[b, r, c, agent_ch] = 1.0
            sit[b, r, c, agent_ch + 1 + int(rng.integers(0, n_dir))] = 1.0
            n_obj = int(rng.integers(1, min(8, G * G - 1) + 1))
            for cell in cells[1:1 + n_obj]:
                r, c = divmod(int(cell), G)
                sit[b, r, c, int(rng.integers(0, 4))] = 1.0
                attrs = rng.choice(np.arange(4, agent_ch), size=2, replace=False)
                sit[b, r, c, attrs] = 1.0
    positions = rng.integers(0, G * G, size=B).astype(np.int64)
    return {"commands": commands, "cmd_lengths": cmd_len, "situations": sit, "targets": targets,
            "tgt_lengths": tgt_len, "target_positions": positions}


CONFIGS = {
    # README.md:177,265-296 demo model (BASELINE.json configs[0])
    "demo": dict(input_vocabulary_size=14, embedding_dimension=5, encoder_hidden_size=20,
                 num_encoder_layers=1, target_vocabulary_size=6, encoder_dropout_p=0.0,
                 encoder_bidirectional=True, num_decoder_layers=1, decoder_dropout_p=0.0,
                 decoder_hidden_size=20, num_cnn_channels=15, cnn_kernel_size=7, cnn_dropout_p=0.0,
                 cnn_hidden_num_channels=50, input_padding_idx=0, target_pad_idx=0, target_eos_idx=2,
                 output_directory="/tmp", conditional_attention=True, auxiliary_task=False,
                 simple_situation_representation=True, attention_type="bahdanau", grid_size=4),
    # compositional_splits shape (configs[1], configs[2] with auxiliary_task=True)
    "comp": dict(input_vocabulary_size=21, embedding_dimension=25, encoder_hidden_size=100,
                 num_encoder_layers=1, target_vocabulary_size=9, encoder_dropout_p=0.0,
                 encoder_bidirectional=True, num_decoder_layers=1, decoder_dropout_p=0.0,
                 decoder_hidden_size=100, num_cnn_channels=16, cnn_kernel_size=7, cnn_dropout_p=0.0,
                 cnn_hidden_num_channels=50, input_padding_idx=0, target_pad_idx=0, target_eos_idx=2,
                 output_directory="/tmp", conditional_attention=True, auxiliary_task=False,
                 simple_situation_representation=True, attention_type="bahdanau", grid_size=6),
    # target_length_split shape (configs[4])
    "tlen": dict(input_vocabulary_size=17, embedding_dimension=25, encoder_hidden_size=100,
                 num_encoder_layers=1, target_vocabulary_size=8, encoder_dropout_p=0.0,
                 encoder_bidirectional=True, num_decoder_layers=1, decoder_dropout_p=0.0,
                 decoder_hidden_size=100, num_cnn_channels=16, cnn_kernel_size=13, cnn_dropout_p=0.0,
                 cnn_hidden_num_channels=50, input_padding_idx=0, target_pad_idx=0, target_eos_idx=2,
                 output_directory="/tmp", conditional_attention=True, auxiliary_task=False,
                 simple_situation_representation=True, attention_type="bahdanau", grid_size=6),
    # a deliberately odd small shape for ragged / non-multiple-of-tile coverage
    "tiny": dict(input_vocabulary_size=11, embedding_dimension=7, encoder_hidden_size=12,
                 num_encoder_layers=1, target_vocabulary_size=7, encoder_dropout_p=0.0,
                 encoder_bidirectional=True, num_decoder_layers=1, decoder_dropout_p=0.0,
                 decoder_hidden_size=12, num_cnn_channels=5, cnn_kernel_size=3, cnn_dropout_p=0.0,
                 cnn_hidden_num_channels=6, input_padding_idx=0, target_pad_idx=0, target_eos_idx=2,
                 output_directory="/tmp", conditional_attention=True, auxiliary_task=True,
                 simple_situation_representation=True, attention_type="bahdanau", grid_size=3),
}


def model_kwargs(cfg: dict) -> dict:
    """Strip the oracle-only keys so the dict can be splatted into a ``Model`` constructor."""
    return {k: v for k, v in cfg.items() if k != "grid_size"}


def full_state_dict(params: Params) -> Params:
    """The 38-key state_dict of the reference Model: the 32 tensors plus the six aliases under
    ``attention_decoder.{textual,visual}_attention``."""
    sd = {k: v.detach().clone().float() for k, v in params.items()}
    for att in ("textual_attention", "visual_attention"):
        for layer in ("key_layer", "query_layer", "energy_layer"):
            sd[f"attention_decoder.{att}.{layer}.weight"] = sd[f"{att}.{layer}.weight"]
    return sd
