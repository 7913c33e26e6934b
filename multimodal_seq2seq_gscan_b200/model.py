"""Drop-in replacement for the reference's ``seq2seq.model.Model`` (reference seq2seq/model.py:24-261).

Same constructor, same attribute / sub-module names (``predict.py:87-96`` reaches into
``visual_attention.key_layer``, ``textual_attention.key_layer``, ``attention_decoder.initialize_hidden``,
``tanh`` and ``enc_hidden_to_dec_hidden``), same 38-key ``state_dict`` and the same
``model.parameters()`` order (Adam state is positional), so the reference's ``train.py`` and
``predict.py`` run against it unchanged.  All arithmetic happens in libgscan_b200.so
(include/gscan_b200.h); the ``nn.Conv2d`` / ``nn.LSTM`` / ``nn.Linear`` objects below are used only as
parameter containers with PyTorch's default initialisation, never called.

Supported configuration (the one the paper uses): ``attention_type="bahdanau"``,
``simple_situation_representation=True``, one encoder and one decoder layer, bidirectional
encoder, conditional attention on or off, auxiliary task on or off.
"""
from __future__ import annotations

import os
import shutil
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops


class _KernelLinear(nn.Linear):
    """nn.Linear whose forward runs on the library GEMM (inference only; used by predict.py's
    direct calls to ``key_layer`` / ``enc_hidden_to_dec_hidden``)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # noqa: D401
        return ops.sgemm_nt(x, self.weight, self.bias)


class SituationEncoder(nn.Module):
    """Parameter container + callable for the three parallel 'same' convolutions
    (reference cnn_model.py:5-36).  Output [B, G*G, 3F]."""

    def __init__(self, num_channels: int, cnn_kernel_size: int, num_conv_channels: int, dropout_probability: float):
        super().__init__()
        self.conv_1 = nn.Conv2d(num_channels, num_conv_channels, kernel_size=1, padding=0)
        self.conv_2 = nn.Conv2d(num_channels, num_conv_channels, kernel_size=5, padding=2)
        self.conv_3 = nn.Conv2d(num_channels, num_conv_channels, kernel_size=cnn_kernel_size,
                                padding=cnn_kernel_size // 2)
        self.dropout_probability = dropout_probability
        self.output_dimension = num_conv_channels * 3
        self._owner = None  # set by Model

    def forward(self, input_images: torch.Tensor) -> torch.Tensor:
        model = self._owner()
        mask = None
        if self.training and self.dropout_probability > 0:
            B, G = input_images.shape[0], input_images.shape[1]
            mask = ops._dropout_mask((B, G * G, self.output_dimension), self.dropout_probability, input_images.device)
        return ops.cnn_forward(model._cfg(input_images.shape[1]), model._param_list(), input_images, mask)


class CommandEncoder(nn.Module):
    """Parameter container for the bidirectional LSTM command encoder (reference seq2seq_model.py:19-94)."""

    def __init__(self, input_size: int, embedding_dim: int, hidden_size: int, num_layers: int,
                 dropout_probability: float, bidirectional: bool, padding_idx: int):
        super().__init__()
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.input_size = input_size
        self.embedding_dim = embedding_dim
        self.dropout_probability = dropout_probability
        self.bidirectional = bidirectional
        self.embedding = nn.Embedding(input_size, embedding_dim, padding_idx=padding_idx)
        self.lstm = nn.LSTM(input_size=embedding_dim, hidden_size=hidden_size, num_layers=num_layers,
                            bidirectional=bidirectional)


class BahdanauAttention(nn.Module):
    """Parameter container for additive attention (reference seq2seq_model.py:97-139)."""

    def __init__(self, key_size: int, query_size: int, hidden_size: int):
        super().__init__()
        self.key_layer = _KernelLinear(key_size, hidden_size, bias=False)
        self.query_layer = _KernelLinear(query_size, hidden_size, bias=False)
        self.energy_layer = _KernelLinear(hidden_size, 1, bias=False)


class AttentionDecoder(nn.Module):
    """Parameter container for the attention decoder (reference seq2seq_model.py:330-509)."""

    def __init__(self, hidden_size: int, output_size: int, num_layers: int, textual_attention: BahdanauAttention,
                 visual_attention: BahdanauAttention, dropout_probability: float, padding_idx: int,
                 conditional_attention: bool):
        super().__init__()
        self.num_layers = num_layers
        self.conditional_attention = conditional_attention
        if conditional_attention:
            self.queries_to_keys = nn.Linear(hidden_size * 2, hidden_size)
        self.hidden_size = hidden_size
        self.output_size = output_size
        self.dropout_probability = dropout_probability
        self.embedding = nn.Embedding(output_size, hidden_size, padding_idx=padding_idx)
        self.lstm = nn.LSTM(hidden_size * 3, hidden_size, num_layers=num_layers)
        self.textual_attention = textual_attention
        self.visual_attention = visual_attention
        self.output_to_hidden = nn.Linear(hidden_size * 4, hidden_size, bias=False)
        self.hidden_to_output = nn.Linear(hidden_size, output_size, bias=False)

    def initialize_hidden(self, encoder_message: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Both the hidden and the cell state of every layer start from the encoder message
        (reference seq2seq_model.py:494-504)."""
        msg = encoder_message.unsqueeze(0).expand(self.num_layers, -1, -1).contiguous()
        return msg.clone(), msg.clone()


class Model(nn.Module):

    def __init__(self, input_vocabulary_size: int, embedding_dimension: int, encoder_hidden_size: int,
                 num_encoder_layers: int, target_vocabulary_size: int, encoder_dropout_p: float,
                 encoder_bidirectional: bool, num_decoder_layers: int, decoder_dropout_p: float,
                 decoder_hidden_size: int, num_cnn_channels: int, cnn_kernel_size: int,
                 cnn_dropout_p: float, cnn_hidden_num_channels: int, input_padding_idx: int, target_pad_idx: int,
                 target_eos_idx: int, output_directory: str, conditional_attention: bool, auxiliary_task: bool,
                 simple_situation_representation: bool, attention_type: str, **kwargs):
        super().__init__()
        if attention_type != "bahdanau":
            if attention_type == "luong":
                raise NotImplementedError("Luong attention is marked 'not correctly implemented' in the reference "
                                          "(model.py:89) and is not provided by the B200 kernels.")
            raise ValueError("Unknown attention type {} specified.".format(attention_type))
        if not simple_situation_representation:
            raise NotImplementedError("Only simple_situation_representation=True is supported (the reference "
                                      "raises for image input too, __main__.py:112-114).")
        if num_encoder_layers != 1 or num_decoder_layers != 1 or not encoder_bidirectional:
            raise NotImplementedError("The B200 kernels implement the paper configuration: one bidirectional "
                                      "encoder layer and one decoder layer.")
        if encoder_hidden_size != decoder_hidden_size:
            raise NotImplementedError("encoder_hidden_size must equal decoder_hidden_size.")
        self.simple_situation_representation = simple_situation_representation
        self.situation_encoder = SituationEncoder(num_channels=num_cnn_channels, cnn_kernel_size=cnn_kernel_size,
                                                  num_conv_channels=cnn_hidden_num_channels,
                                                  dropout_probability=cnn_dropout_p)
        self.visual_attention = BahdanauAttention(key_size=cnn_hidden_num_channels * 3,
                                                  query_size=decoder_hidden_size, hidden_size=decoder_hidden_size)
        self.auxiliary_task = auxiliary_task
        if auxiliary_task:
            self.auxiliary_loss_criterion = nn.NLLLoss()
        self.encoder = CommandEncoder(input_size=input_vocabulary_size, embedding_dim=embedding_dimension,
                                      hidden_size=encoder_hidden_size, num_layers=num_encoder_layers,
                                      dropout_probability=encoder_dropout_p, bidirectional=encoder_bidirectional,
                                      padding_idx=input_padding_idx)
        self.enc_hidden_to_dec_hidden = _KernelLinear(encoder_hidden_size, decoder_hidden_size)
        self.textual_attention = BahdanauAttention(key_size=encoder_hidden_size, query_size=decoder_hidden_size,
                                                   hidden_size=decoder_hidden_size)
        self.attention_type = attention_type
        self.attention_decoder = AttentionDecoder(hidden_size=decoder_hidden_size, output_size=target_vocabulary_size,
                                                  num_layers=num_decoder_layers,
                                                  textual_attention=self.textual_attention,
                                                  visual_attention=self.visual_attention,
                                                  dropout_probability=decoder_dropout_p, padding_idx=target_pad_idx,
                                                  conditional_attention=conditional_attention)
        self.target_eos_idx = target_eos_idx
        self.target_pad_idx = target_pad_idx
        self.input_padding_idx = input_padding_idx
        self.loss_criterion = nn.NLLLoss(ignore_index=target_pad_idx)
        self.tanh = nn.Tanh()
        self.output_directory = output_directory
        self.trained_iterations = 0
        self.best_iteration = 0
        self.best_exact_match = 0
        self.best_accuracy = 0
        self.conditional_attention = conditional_attention
        self._static_cfg = dict(C=num_cnn_channels, F=cnn_hidden_num_channels, K3=cnn_kernel_size,
                                E=embedding_dimension, H=decoder_hidden_size, Vi=input_vocabulary_size,
                                V=target_vocabulary_size, conditional_attention=int(conditional_attention),
                                auxiliary_task=int(auxiliary_task), pad_idx_in=input_padding_idx,
                                pad_idx_out=target_pad_idx)
        self._dropout_p = (cnn_dropout_p, encoder_dropout_p, decoder_dropout_p)
        import weakref
        self.situation_encoder._owner = weakref.ref(self)

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _cfg(self, grid_size: int) -> dict:
        cfg = dict(self._static_cfg)
        cfg["G"] = int(grid_size)
        return cfg

    def _param_list(self) -> List[Optional[torch.Tensor]]:
        """The 32 tensors in gscan_param order (include/gscan_b200.h) = model.parameters() order."""
        se, dec = self.situation_encoder, self.attention_decoder
        lstm_e, lstm_d = self.encoder.lstm, dec.lstm
        cond = dec.conditional_attention
        return [
            se.conv_1.weight, se.conv_1.bias, se.conv_2.weight, se.conv_2.bias, se.conv_3.weight, se.conv_3.bias,
            self.visual_attention.key_layer.weight, self.visual_attention.query_layer.weight,
            self.visual_attention.energy_layer.weight,
            self.encoder.embedding.weight,
            lstm_e.weight_ih_l0, lstm_e.weight_hh_l0, lstm_e.bias_ih_l0, lstm_e.bias_hh_l0,
            lstm_e.weight_ih_l0_reverse, lstm_e.weight_hh_l0_reverse, lstm_e.bias_ih_l0_reverse,
            lstm_e.bias_hh_l0_reverse,
            self.enc_hidden_to_dec_hidden.weight, self.enc_hidden_to_dec_hidden.bias,
            self.textual_attention.key_layer.weight, self.textual_attention.query_layer.weight,
            self.textual_attention.energy_layer.weight,
            dec.queries_to_keys.weight if cond else None, dec.queries_to_keys.bias if cond else None,
            dec.embedding.weight,
            lstm_d.weight_ih_l0, lstm_d.weight_hh_l0, lstm_d.bias_ih_l0, lstm_d.bias_hh_l0,
            dec.output_to_hidden.weight, dec.hidden_to_output.weight,
        ]

    def _device(self) -> torch.device:
        return self.enc_hidden_to_dec_hidden.weight.device

    def _masks(self, B: int, G: int, Ti: int, Tt: int, device):
        if not self.training:
            return (None, None, None)
        p_cnn, p_enc, p_dec = self._dropout_p
        cfg = self._static_cfg
        # same order as the reference draws them: CNN features, command embeddings, target embeddings
        return (ops._dropout_mask((B, G * G, 3 * cfg["F"]), p_cnn, device),
                ops._dropout_mask((B, Ti, cfg["E"]), p_enc, device),
                ops._dropout_mask((B, Tt, cfg["H"]), p_dec, device) if Tt > 0 else None)

    # ------------------------------------------------------------------------------------------
    # reference API: losses and metrics (model.py:108-170)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def remove_start_of_sequence(input_tensor: torch.Tensor) -> torch.Tensor:
        """Drop the SOS column and append a padding (0) column."""
        pad = torch.zeros(input_tensor.size(0), 1, dtype=input_tensor.dtype, device=input_tensor.device)
        return torch.cat([input_tensor[:, 1:], pad], dim=1)

    def get_metrics(self, target_scores: torch.Tensor, targets: torch.Tensor) -> Tuple[float, float]:
        match, total, exact = ops.metrics_counts(target_scores, targets, self.target_pad_idx).tolist()
        return 100. * match / total, 100. * exact / targets.size(0)

    @staticmethod
    def get_auxiliary_accuracy(target_scores: torch.Tensor, targets: torch.Tensor) -> float:
        with torch.no_grad():
            predicted = target_scores.max(dim=1)[1]
            equal = torch.eq(targets.view(-1), predicted).long().sum().item()
        return 100. * equal / len(targets)

    def get_loss(self, target_scores: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """Mean NLL of [B,Tt,V] log-probabilities against targets shifted left by one, ignoring pad.  The position
        behind the last target holds the literal 0 the reference appends (model.py:108-115): ignored when
        target_pad_idx == 0 (the gSCAN vocabularies), scored as class 0 otherwise - the kernels do the same."""
        return ops.NLLLoss.apply(target_scores, targets, self.target_pad_idx, 1)[0]

    def get_auxiliary_loss(self, auxiliary_scores_target: torch.Tensor, target_target_positions: torch.Tensor):
        return ops.NLLLoss.apply(auxiliary_scores_target.unsqueeze(1), target_target_positions.view(-1, 1),
                                 -100, 0)[0]

    def auxiliary_task_forward(self, output_scores_target_pos: torch.Tensor) -> torch.Tensor:
        assert self.auxiliary_task, "Please set auxiliary_task to True if using it."
        return torch.log_softmax(output_scores_target_pos, dim=-1)

    # ------------------------------------------------------------------------------------------
    # reference API: step-wise interface used by predict.py (model.py:172-188)
    # ------------------------------------------------------------------------------------------
    def encode_input(self, commands_input: torch.LongTensor, commands_lengths, situations_input: torch.Tensor
                     ) -> Dict[str, torch.Tensor]:
        B, G = situations_input.shape[0], situations_input.shape[1]
        Ti = ops.max_length(commands_lengths)
        masks = self._masks(B, G, Ti, 0, situations_input.device)
        feat, enc_out, hidden = ops.encode(self._cfg(G), self._param_list(), commands_input, commands_lengths,
                                           situations_input, masks[:2])
        lengths = [int(l) for l in np.asarray(commands_lengths).reshape(-1)] \
            if not isinstance(commands_lengths, torch.Tensor) else commands_lengths.tolist()
        return {"encoded_situations": feat,
                "encoded_commands": {"encoder_outputs": enc_out, "sequence_lengths": lengths},
                "hidden_states": hidden}

    def decode_input(self, target_token: torch.LongTensor, hidden: Tuple[torch.Tensor, torch.Tensor],
                     encoder_outputs: torch.Tensor, input_lengths, encoded_situations: torch.Tensor):
        """One decoder step on PROJECTED keys; returns (logits, (h, c), beta, alpha, beta) exactly as
        ``forward_step`` does (reference seq2seq_model.py:427-428)."""
        h, c = hidden
        G = int(round(encoded_situations.shape[1] ** 0.5))
        drop = None
        if self.training and self._dropout_p[2] > 0:
            drop = ops._dropout_mask((h.shape[-2], h.shape[-1]), self._dropout_p[2], h.device)
        logits, h_new, c_new, alpha, beta = ops.decoder_step(self._cfg(G), self._param_list(), target_token, h, c,
                                                             encoder_outputs, input_lengths, encoded_situations, drop)
        return logits, (h_new.unsqueeze(0), c_new.unsqueeze(0)), beta, alpha, beta

    # ------------------------------------------------------------------------------------------
    # reference API: training forward (model.py:190-219)
    # ------------------------------------------------------------------------------------------
    def forward(self, commands_input: torch.LongTensor, commands_lengths, situations_input: torch.Tensor,
                target_batch: torch.LongTensor, target_lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        assert commands_input.size(0) == len(commands_lengths), "Wrong amount of lengths passed to .forward()"
        dev = situations_input.device
        B, G = situations_input.shape[0], situations_input.shape[1]
        Ti = ops.max_length(commands_lengths)
        Tt = target_batch.shape[1]
        cmd_len = ops.lengths_to_device(commands_lengths, dev)
        # training: the dropout of the three sites is drawn inside the kernels (no mask tensors, no ATen launches)
        masks = (ops.dropout_rng(*self._dropout_p) if self.training else None) or (None, None, None)
        logp, aux = ops.ModelForward.apply(self._cfg(G), commands_input, cmd_len, Ti, situations_input, target_batch,
                                           masks, *self._param_list())
        if not self.auxiliary_task:
            aux = (torch.zeros(1), torch.zeros(1))     # what the reference returns (model.py:217)
        return logp, aux

    # ------------------------------------------------------------------------------------------
    # new, additive: batched greedy decoding (replaces predict.py's batch-size-1 loop)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def greedy_decode(self, commands_input: torch.LongTensor, commands_lengths, situations_input: torch.Tensor,
                      max_decoding_steps: int, sos_idx: int, eos_idx: int, return_attention: bool = False) -> dict:
        """Decode a whole batch at once with the per-sequence semantics of ``predict.py:97-117``.
        Returns device tensors: ``tokens`` [B, max_decoding_steps+1] (-1 beyond ``lengths``), ``lengths``,
        ``steps``, ``beta_sum`` [B, G*G], ``aux_logp`` (if the auxiliary task is on) and, on request,
        the per-step attention weights."""
        G = situations_input.shape[1]
        if self._static_cfg["V"] > 32:
            raise NotImplementedError(
                "batched greedy decoding keeps the logits of one example in the registers of one lane: target "
                "vocabularies above 32 entries (gSCAN has 8-9) are not supported by gscan_greedy_decode; "
                "use the step-wise encode_input / decode_input loop of predict.py for such a model")
        return ops.greedy_decode(self._cfg(G), self._param_list(), commands_input, commands_lengths,
                                 situations_input, max_decoding_steps, sos_idx, eos_idx,
                                 return_attention=return_attention, want_aux=bool(self.auxiliary_task))

    # ------------------------------------------------------------------------------------------
    # reference API: bookkeeping and checkpoints (model.py:221-261)
    # ------------------------------------------------------------------------------------------
    def update_state(self, is_best: bool, accuracy=None, exact_match=None) -> None:
        self.trained_iterations += 1
        if is_best:
            self.best_exact_match = exact_match
            self.best_accuracy = accuracy
            self.best_iteration = self.trained_iterations

    def load_model(self, path_to_checkpoint: str) -> dict:
        checkpoint = torch.load(path_to_checkpoint, map_location=self._device())
        self.trained_iterations = checkpoint["iteration"]
        self.best_iteration = checkpoint["best_iteration"]
        self.load_state_dict(checkpoint["state_dict"])
        self.best_exact_match = checkpoint["best_exact_match"]
        self.best_accuracy = checkpoint["best_accuracy"]
        return checkpoint["optimizer_state_dict"]

    def get_current_state(self) -> dict:
        return {"iteration": self.trained_iterations, "state_dict": self.state_dict(),
                "best_iteration": self.best_iteration, "best_accuracy": self.best_accuracy,
                "best_exact_match": self.best_exact_match}

    def save_checkpoint(self, file_name: str, is_best: bool, optimizer_state_dict: dict) -> str:
        path = os.path.join(self.output_directory, file_name)
        state = self.get_current_state()
        state["optimizer_state_dict"] = optimizer_state_dict
        torch.save(state, path)
        if is_best:
            shutil.copyfile(path, os.path.join(self.output_directory, "model_best.pth.tar"))
        return path
