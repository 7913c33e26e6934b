"""Host-side glue between torch tensors and the C ABI (include/gscan_b200.h).

torch owns memory, streams and autograd bookkeeping; every arithmetic operation of the path runs
in libgscan_b200.so.  The ``autograd.Function`` classes here play the role autograd's recorded
graph plays in the reference: ``ModelForward`` saves the workspace written by ``gscan_forward``
and hands it to ``gscan_backward``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Dims, Dropout, ParamArray

# launches issued by the library since import, for bench.py's "gpu_launches" claim
_call_counts = {"forward": 0, "backward": 0, "greedy": 0, "other": 0}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("multimodal_seq2seq_gscan_b200 runs only on CUDA tensors (sm_100a kernels); "
                               "there is no CPU fallback. Move the model and its inputs to a CUDA device.")


def _param_array(tensors: Sequence[Optional[torch.Tensor]]) -> ParamArray:
    arr = ParamArray()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _PinnedRing:
    """A few pinned staging buffers of one size, reused round-robin.  A host->device copy from PAGEABLE memory makes the
    CUDA runtime synchronise the stream before it starts (the host then waits for everything enqueued so far: measured
    as one host/GPU lock-step per training step, with every launch after it exposed at the host's launch cadence);
    from pinned memory the copy is truly stream-ordered.  A slot is rewritten only after the copy that last read it
    has completed (its event), which with several slots is long past."""

    SLOTS = 8

    def __init__(self, n: int):
        self.bufs = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(self.SLOTS)]
        self.events = [None] * self.SLOTS
        self.i = 0

    def stage(self, arr: np.ndarray, device) -> torch.Tensor:
        k = self.i
        self.i = (k + 1) % self.SLOTS
        if self.events[k] is not None:
            self.events[k].synchronize()
        np.copyto(self.bufs[k].numpy(), arr, casting="unsafe")
        out = torch.empty(arr.size, dtype=torch.int32, device=device)
        out.copy_(self.bufs[k], non_blocking=True)
        if self.events[k] is None:
            self.events[k] = torch.cuda.Event()
        self.events[k].record(torch.cuda.current_stream(device))
        return out


_RINGS: dict = {}


def lengths_to_device(lengths, device) -> torch.Tensor:
    """The reference passes lengths as host lists / numpy float64 arrays (gSCAN_dataset.py:198-199);
    the kernels want int32 on the device.  No sync: the copy is stream-ordered (pinned staging ring)."""
    if isinstance(lengths, torch.Tensor):
        return lengths.to(device=device, dtype=torch.int32)
    arr = np.asarray(lengths).reshape(-1)
    device = torch.device(device)
    if device.type != "cuda":
        return torch.from_numpy(np.ascontiguousarray(arr.astype(np.int32))).to(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), arr.size)
    ring = _RINGS.get(key)
    if ring is None:
        ring = _RINGS[key] = _PinnedRing(arr.size)
    return ring.stage(arr, device)


def max_length(lengths) -> int:
    if isinstance(lengths, torch.Tensor):
        return int(lengths.max().item())
    return int(np.max(np.asarray(lengths)))


def make_dims(cfg: dict, B: int, Ti: int, Tt: int, Ti_stride: int) -> Dims:
    d = Dims()
    d.B, d.Ti, d.Tt, d.Ti_stride = B, Ti, Tt, Ti_stride
    for k in ("G", "C", "F", "K3", "E", "H", "Vi", "V", "conditional_attention", "auxiliary_task", "pad_idx_in",
              "pad_idx_out"):
        setattr(d, k, int(cfg[k]))
    return d


_ONES: dict = {}


def _dropout_mask(shape, p: float, device, generator=None) -> Optional[torch.Tensor]:
    """Inverted-dropout mask (0 or 1/(1-p)).  Without an explicit generator it is ONE fused kernel - dropout of a cached
    tensor of ones - instead of bernoulli_ + mul_ (the three masks of a step are drawn at its very start, on the
    critical path: 6 launches, 28 us on B200, became 3)."""
    if p <= 0.0:
        return None
    if generator is None:
        key = (tuple(shape), str(device))
        ones = _ONES.get(key)
        if ones is None:
            if len(_ONES) > 64:
                _ONES.clear()
            ones = _ONES[key] = torch.ones(shape, dtype=torch.float32, device=device)
        return torch.nn.functional.dropout(ones, p, True)
    mask = torch.empty(shape, dtype=torch.float32, device=device)
    mask.bernoulli_(1.0 - p, generator=generator)
    return mask.mul_(1.0 / (1.0 - p))


_rng_calls = 0
_ZEROS1: dict = {}


def _zeros1(dev) -> torch.Tensor:
    """A fresh view of a per-device constant zeros(1) (the placeholder auxiliary output): no fill kernel per step."""
    z = _ZEROS1.get(dev)
    if z is None:
        z = _ZEROS1[dev] = torch.zeros(1, device=dev)
    return z.view(1)



def dropout_rng(p_cnn: float, p_enc: float, p_dec: float) -> Optional[Dropout]:
    """The dropout of ONE training step, drawn inside the kernels: a Philox stream keyed by torch's seed (mixed with
    the rank under torch.distributed, so that replicas draw different masks) and a per-process step counter.  No
    device work, no host sync.  None when all probabilities are zero."""
    global _rng_calls
    if p_cnn <= 0 and p_enc <= 0 and p_dec <= 0:
        return None
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        seed = (seed + 0x9E3779B97F4A7C15 * (torch.distributed.get_rank() + 1)) & 0xFFFFFFFFFFFFFFFF
    _rng_calls += 1
    return Dropout(float(p_cnn), float(p_enc), float(p_dec), seed, _rng_calls)


def dropout_masks_of(rng: Dropout, shapes, device) -> list:
    """The three masks a ``Dropout`` stands for, materialised (tests): shapes = ((B, M, 3F), (B, Ti, E), (B, Tt, H))."""
    lib = _lib.load()
    out = []
    for which, shape in enumerate(shapes):
        m = torch.empty(shape, dtype=torch.float32, device=device)
        _lib.check(lib.gscan_dropout_mask(rng, which, m.numel(), _ptr(m), _stream(device)), "gscan_dropout_mask")
        out.append(m)
    return out


class ModelForward(torch.autograd.Function):
    """(logp [B,Tt,V], aux_logp [B,M]) = gscan_forward(...); backward = gscan_backward(...).  ``masks`` is a triple of
    explicit dropout masks (or Nones), or a ``Dropout`` (masks drawn inside the kernels: gscan_forward_rng)."""

    @staticmethod
    def forward(ctx, cfg, commands, cmd_len_dev, Ti, situations, targets, masks, *params):
        lib = _lib.load()
        _require_cuda(commands, situations, targets, *[p for p in params if p is not None])
        dev = situations.device
        B, Tt = targets.shape
        commands = commands.contiguous()
        situations = situations.contiguous().float()
        targets = targets.contiguous()
        dims = make_dims(cfg, B, Ti, Tt, commands.shape[1])
        _lib.check(lib.gscan_check_dims(dims), "gscan_check_dims")
        ctx.set_materialize_grads(False)     # no zero-filled gradients for outputs nobody differentiates
        M = dims.G * dims.G
        n_ws = lib.gscan_workspace_floats(dims)
        ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
        logp = torch.empty(B, Tt, dims.V, dtype=torch.float32, device=dev)
        aux = torch.empty(B, M, dtype=torch.float32, device=dev) if dims.auxiliary_task else None
        parr = _param_array([None if p is None else p.detach() for p in params])
        ctx.rng = masks if isinstance(masks, Dropout) else None
        early = _take_early_dlogp(logp)
        if early is not None:
            # training step with d(loss)/d(logp) known in advance: the output-head backward runs inside the forward call
            d_early, ready = early
            rng = ctx.rng
            if rng is not None:
                masks = (None, None, None)
            rc = lib.gscan_forward_train(dims, parr, _ptr(commands), _ptr(cmd_len_dev), _ptr(situations), _ptr(targets),
                                         _ptr(masks[0]), _ptr(masks[1]), _ptr(masks[2]), rng, _ptr(ws), n_ws, _ptr(logp),
                                         _ptr(aux), _ptr(d_early), ready, _stream(dev))
        elif ctx.rng is not None:
            rc = lib.gscan_forward_rng(dims, parr, _ptr(commands), _ptr(cmd_len_dev), _ptr(situations), _ptr(targets),
                                       ctx.rng, _ptr(ws), n_ws, _ptr(logp), _ptr(aux), _stream(dev))
            masks = (None, None, None)
        else:
            drop_cnn, drop_enc, drop_dec = masks
            rc = lib.gscan_forward(dims, parr, _ptr(commands), _ptr(cmd_len_dev), _ptr(situations), _ptr(targets),
                                   _ptr(drop_cnn), _ptr(drop_enc), _ptr(drop_dec), _ptr(ws), n_ws, _ptr(logp),
                                   _ptr(aux), _stream(dev))
        _lib.check(rc, "gscan_forward")
        _call_counts["forward"] += 1
        ctx.dims = dims
        ctx.n_ws = n_ws
        ctx.param_shapes = [None if p is None else p.shape for p in params]
        ctx.save_for_backward(commands, cmd_len_dev, situations, targets, ws, *[m for m in masks if m is not None],
                              *[p for p in params if p is not None])
        ctx.mask_present = [m is not None for m in masks]
        ctx.param_present = [p is not None for p in params]
        if aux is None:
            aux = _zeros1(dev)
            ctx.mark_non_differentiable(aux)
        return logp, aux

    @staticmethod
    def backward(ctx, d_logp, d_aux):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        commands, cmd_len_dev, situations, targets, ws = saved[:5]
        rest = saved[5:]
        masks = []
        for present in ctx.mask_present:
            masks.append(rest.pop(0) if present else None)
        params = []
        for present in ctx.param_present:
            params.append(rest.pop(0) if present else None)
        dims = ctx.dims
        dev = situations.device
        if d_logp is None:
            d_logp = torch.zeros(dims.B, dims.Tt, dims.V, dtype=torch.float32, device=dev)
        d_logp = d_logp.contiguous().float()
        if not dims.auxiliary_task:
            d_aux = None
        elif d_aux is not None:
            d_aux = d_aux.contiguous().float()
        # one flat fp32 gradient buffer in model.parameters() order; the per-parameter gradients
        # returned to autograd are views into it (this is also the data-parallel all-reduce buffer)
        sizes, offsets = flat_layout(ctx.param_shapes)
        n_flat = int(offsets[-1])
        prezeroed = _flat_grad_prezeroed
        flat = _take_flat_grad_target(n_flat, dev)
        if flat is None:
            flat = torch.zeros(n_flat + FLAT_TAIL, dtype=torch.float32, device=dev)
        elif not prezeroed:
            flat[:n_flat].zero_()      # the tail (data-parallel counts) belongs to the caller
        views = [None if s is None else flat[int(o):int(o) + n].view(s)
                 for s, o, n in zip(ctx.param_shapes, offsets[:-1], sizes)]
        if ctx.rng is not None:
            rc = lib.gscan_backward_rng(dims, _param_array(params), _ptr(commands), _ptr(cmd_len_dev), _ptr(situations),
                                        _ptr(targets), ctx.rng, _ptr(ws), ctx.n_ws, _ptr(d_logp), _ptr(d_aux),
                                        _param_array(views), _stream(dev))
        else:
            rc = lib.gscan_backward(dims, _param_array(params), _ptr(commands), _ptr(cmd_len_dev), _ptr(situations),
                                    _ptr(targets), _ptr(masks[0]), _ptr(masks[1]), _ptr(masks[2]), _ptr(ws), ctx.n_ws,
                                    _ptr(d_logp), _ptr(d_aux), _param_array(views), _stream(dev))
        _lib.check(rc, "gscan_backward")
        _call_counts["backward"] += 1
        return (None, None, None, None, None, None, None, *views)


# floats behind the last gradient of the flat buffer: the data-parallel step appends [n_tok, n_examples] there so
# that gradients and counts travel in ONE all-reduce (dp.COUNT_SLOTS)
FLAT_TAIL = 4
_flat_grad_target: Optional[torch.Tensor] = None
_flat_grad_prezeroed = False


def set_flat_grad_target(buf: Optional[torch.Tensor], prezeroed: bool = False) -> None:
    """The NEXT ``ModelForward.backward`` writes its flat gradient into ``buf`` (fp32, at least
    ``flat_layout(...)[1][-1] + FLAT_TAIL`` elements; only the gradient part is zeroed) instead of a fresh
    tensor.  ``FusedTrainer`` hands in its persistent buffer: no allocation per step, and the counts it has
    already written into the tail survive.  ``prezeroed``: the caller has zeroed the gradient part already
    (the trainer does it at the start of the step, off the path between the two sweeps)."""
    global _flat_grad_target, _flat_grad_prezeroed
    _flat_grad_target = buf
    _flat_grad_prezeroed = bool(prezeroed) and buf is not None


_early_dlogp = None


def set_early_dlogp(d_logp: Optional[torch.Tensor], ready_event: Optional["torch.cuda.Event"] = None) -> None:
    """The NEXT ``ModelForward.forward`` is the forward pass of a training step whose d(loss)/d(logp) is ``d_logp``
    ([B, Tt, V] fp32, contiguous; ``nll_grad_from_targets``): it goes through gscan_forward_train, which also runs the
    output-head backward pass.  The backward call must then be given the SAME tensor (``torch.autograd.grad(...,
    grad_outputs=[d_logp])``); anything else is correct too, the head is then recomputed.  ``ready_event``: recorded
    after the work that fills ``d_logp`` when that work runs on another stream."""
    global _early_dlogp
    _early_dlogp = None if d_logp is None else (d_logp, ready_event)


def _take_early_dlogp(logp: torch.Tensor):
    global _early_dlogp
    e, _early_dlogp = _early_dlogp, None
    if e is None:
        return None
    d, ev = e
    if d.shape != logp.shape or d.dtype != torch.float32 or d.device != logp.device or not d.is_contiguous():
        return None
    return d, (None if ev is None else ev.cuda_event)


def _take_flat_grad_target(n_flat: int, device) -> Optional[torch.Tensor]:
    global _flat_grad_target
    buf, _flat_grad_target = _flat_grad_target, None
    if buf is None or buf.device != device or buf.dtype != torch.float32 or buf.numel() < n_flat + FLAT_TAIL \
            or not buf.is_contiguous():
        return None
    return buf


def flat_layout(shapes):
    """Offsets (in floats, each padded to a multiple of 4 = 16 bytes) of the tensors of a flat fp32
    buffer in model.parameters() order; absent tensors (shape None) take no room.  Used for the flat
    gradient buffer produced by ``ModelForward.backward`` and for ``FusedTrainer``'s flat parameter /
    Adam-state buffers, so the two line up element for element."""
    key = tuple(None if s is None else tuple(s) for s in shapes)
    hit = _FLAT_LAYOUTS.get(key)          # computed once per model: this runs in every backward pass
    if hit is None:
        sizes = [0 if s is None else int(np.prod(s)) for s in shapes]
        offsets = np.concatenate([[0], np.cumsum([(n + 3) // 4 * 4 for n in sizes])]).astype(np.int64)
        hit = _FLAT_LAYOUTS[key] = (sizes, offsets)
    return hit


_FLAT_LAYOUTS: dict = {}


class NLLLoss(torch.autograd.Function):
    """(mean NLL over non-ignored targets shifted by ``shift``, number of scored targets)
    (gscan_nll_forward / _backward)."""

    @staticmethod
    def forward(ctx, logp, targets, pad_idx, shift, clone_out=True, out=None):
        lib = _lib.load()
        _require_cuda(logp, targets)
        ctx.set_materialize_grads(False)
        logp = logp.contiguous().float()
        targets = targets.contiguous()
        B, T, V = logp.shape
        if out is None:
            out = torch.empty(68, dtype=torch.float32, device=logp.device)   # GSCAN_NLL_OUT_FLOATS: [mean, count | scratch]
        _lib.check(lib.gscan_nll_forward(_ptr(logp), _ptr(targets), B, T, V, int(pad_idx), int(shift), _ptr(out),
                                         _stream(logp.device)), "gscan_nll_forward")
        _call_counts["other"] += 1
        ctx.save_for_backward(targets, out)
        ctx.meta = (B, T, V, int(pad_idx), int(shift))
        # The reference's driver updates the loss in place (train.py:107), which autograd forbids on an output that is
        # a view of a buffer made here: by default the mean is a copy.  The fused trainer never does and takes views
        # (clone_out=False): no copy kernels between the head and the backward pass.
        count = out[1]
        ctx.mark_non_differentiable(count)
        return (out[0].clone() if clone_out else out[0]), count

    @staticmethod
    def backward(ctx, d_loss, _d_count):
        lib = _lib.load()
        if d_loss is None:
            return None, None, None, None, None, None
        targets, out = ctx.saved_tensors
        B, T, V, pad_idx, shift = ctx.meta
        d_loss = d_loss.contiguous().float().reshape(1)
        d_logp = torch.empty(B, T, V, dtype=torch.float32, device=out.device)
        _lib.check(lib.gscan_nll_backward(_ptr(targets), B, T, V, pad_idx, shift, _ptr(out), _ptr(d_loss), _ptr(d_logp),
                                          _stream(out.device)), "gscan_nll_backward")
        _call_counts["other"] += 1
        return d_logp, None, None, None, None, None


def nll_grad_from_targets(targets: torch.Tensor, V: int, pad_idx: int, shift: int, sum_form: bool,
                          d_logp: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                          d_loss: Optional[torch.Tensor] = None):
    """d(loss)/d(logp) of the mean (``sum_form`` False) or the sum of the NLL terms ``NLLLoss`` scores, formed from the
    targets alone (gscan_nll_count + gscan_nll_backward) - bit-identical to what ``NLLLoss.backward`` returns for
    ``loss = mean`` resp. ``loss = mean * count``.  Returns (d_logp [B, T, V], result buffer with the count at [1]);
    ``d_logp`` / ``out`` may be handed in (buffers a caller reuses from step to step); ``d_loss`` overrides the factor
    (the weight of an auxiliary loss term)."""
    lib = _lib.load()
    _require_cuda(targets)
    targets = targets.contiguous()
    B, T = targets.shape
    dev = targets.device
    if out is None:
        out = torch.empty(68, dtype=torch.float32, device=dev)
    _lib.check(lib.gscan_nll_count(_ptr(targets), B, T, int(pad_idx), int(shift), _ptr(out), _stream(dev)), "gscan_nll_count")
    if d_logp is None or d_logp.shape != (B, T, V):
        d_logp = torch.empty(B, T, V, dtype=torch.float32, device=dev)
    if d_loss is None:                                  # (a 1-element device tensor: d loss / d mean of these terms)
        d_loss = out[1:2] if sum_form else _ones1(dev)  # d loss / d mean = count, resp. 1
    _lib.check(lib.gscan_nll_backward(_ptr(targets), B, T, V, int(pad_idx), int(shift), _ptr(out), _ptr(d_loss), _ptr(d_logp),
                                      _stream(dev)), "gscan_nll_backward")
    _call_counts["other"] += 2
    return d_logp, out


_ONES1: dict = {}


def _ones1(dev) -> torch.Tensor:
    t = _ONES1.get(dev)
    if t is None:
        t = _ONES1[dev] = torch.ones(1, dtype=torch.float32, device=dev)
    return t


def metrics_counts(logp: torch.Tensor, targets: torch.Tensor, pad_idx: int) -> torch.Tensor:
    """int32[3] on the device: matching tokens, non-pad tokens, exactly matching sequences."""
    lib = _lib.load()
    _require_cuda(logp, targets)
    logp = logp.detach().contiguous().float()
    targets = targets.contiguous()
    B, T, V = logp.shape
    counts = torch.empty(3, dtype=torch.int32, device=logp.device)
    _lib.check(lib.gscan_metrics(_ptr(logp), _ptr(targets), B, T, V, int(pad_idx), _ptr(counts),
                                 _stream(logp.device)), "gscan_metrics")
    _call_counts["other"] += 1
    return counts


def sgemm_nt(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """y = act(x @ weight.T + bias) on the library's GEMM (no autograd)."""
    lib = _lib.load()
    _require_cuda(x, weight)
    shape = x.shape
    x2 = x.detach().reshape(-1, shape[-1]).contiguous().float()
    w = weight.detach().contiguous()
    M, K = x2.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.gscan_sgemm(_ptr(x2), K, 1, _ptr(w), 1, K, _ptr(y), N, M, N, K,
                               None if bias is None else bias.detach().data_ptr(), act, 0, _stream(x.device)),
               "gscan_sgemm")
    _call_counts["other"] += 1
    return y.view(*shape[:-1], N)


def encode(cfg, params, commands, lengths, situations, masks=(None, None)):
    lib = _lib.load()
    _require_cuda(commands, situations)
    dev = situations.device
    commands = commands.contiguous()
    situations = situations.contiguous().float()
    B = commands.shape[0]
    Ti = max_length(lengths)
    cmd_len = lengths_to_device(lengths, dev)
    dims = make_dims(cfg, B, Ti, 1, commands.shape[1])
    _lib.check(lib.gscan_check_dims(dims), "gscan_check_dims")
    n_ws = lib.gscan_encode_workspace_floats(dims)
    ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
    M, D = dims.G * dims.G, 3 * dims.F
    feat = torch.empty(B, M, D, dtype=torch.float32, device=dev)
    enc_out = torch.empty(Ti, B, dims.H, dtype=torch.float32, device=dev)
    hidden = torch.empty(B, dims.H, dtype=torch.float32, device=dev)
    rc = lib.gscan_encode(dims, _param_array([None if p is None else p.detach() for p in params]), _ptr(commands),
                          _ptr(cmd_len), _ptr(situations), _ptr(masks[0]), _ptr(masks[1]), _ptr(ws), n_ws, _ptr(feat),
                          _ptr(enc_out), _ptr(hidden), _stream(dev))
    _lib.check(rc, "gscan_encode")
    _call_counts["other"] += 1
    return feat, enc_out, hidden


def decoder_step(cfg, params, tokens, h, c, keys_text, lengths, keys_vis, drop_dec=None):
    lib = _lib.load()
    _require_cuda(tokens, h, c, keys_text, keys_vis)
    dev = h.device
    tokens = tokens.contiguous()
    h2 = h.detach().reshape(-1, h.shape[-1]).contiguous().float()
    c2 = c.detach().reshape(-1, c.shape[-1]).contiguous().float()
    keys_text = keys_text.detach().contiguous().float()
    keys_vis = keys_vis.detach().contiguous().float()
    Ti, B, H = keys_text.shape
    cmd_len = lengths_to_device(lengths, dev)
    dims = make_dims(cfg, B, Ti, 1, Ti)
    _lib.check(lib.gscan_check_dims(dims), "gscan_check_dims")
    n_ws = lib.gscan_step_workspace_floats(dims)
    ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
    M = dims.G * dims.G
    logits = torch.empty(B, dims.V, dtype=torch.float32, device=dev)
    h_out = torch.empty(B, H, dtype=torch.float32, device=dev)
    c_out = torch.empty(B, H, dtype=torch.float32, device=dev)
    alpha = torch.empty(B, Ti, dtype=torch.float32, device=dev)
    beta = torch.empty(B, M, dtype=torch.float32, device=dev)
    rc = lib.gscan_decoder_step(dims, _param_array([None if p is None else p.detach() for p in params]), _ptr(tokens),
                                _ptr(h2), _ptr(c2), _ptr(keys_text), _ptr(cmd_len), _ptr(keys_vis), _ptr(drop_dec),
                                _ptr(ws), n_ws, _ptr(logits), _ptr(h_out), _ptr(c_out), _ptr(alpha), _ptr(beta),
                                _stream(dev))
    _lib.check(rc, "gscan_decoder_step")
    _call_counts["other"] += 1
    return logits, h_out, c_out, alpha, beta


def greedy_decode(cfg, params, commands, lengths, situations, max_decoding_steps, sos_idx, eos_idx,
                  return_attention=False, want_aux=False):
    lib = _lib.load()
    _require_cuda(commands, situations)
    dev = situations.device
    commands = commands.contiguous()
    situations = situations.contiguous().float()
    B = commands.shape[0]
    Ti = max_length(lengths)
    cmd_len = lengths_to_device(lengths, dev)
    dims = make_dims(cfg, B, Ti, 1, commands.shape[1])
    _lib.check(lib.gscan_check_dims(dims), "gscan_check_dims")
    n_ws = lib.gscan_greedy_workspace_floats(dims)
    ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
    T = int(max_decoding_steps) + 1
    M = dims.G * dims.G
    tokens = torch.empty(B, T, dtype=torch.int64, device=dev)
    out_len = torch.empty(B, dtype=torch.int32, device=dev)
    out_steps = torch.empty(B, dtype=torch.int32, device=dev)
    beta_sum = torch.empty(B, M, dtype=torch.float32, device=dev)
    aux = torch.empty(B, M, dtype=torch.float32, device=dev) if want_aux else None
    alphas = torch.zeros(B, T, Ti, dtype=torch.float32, device=dev) if return_attention else None
    betas = torch.zeros(B, T, M, dtype=torch.float32, device=dev) if return_attention else None
    rc = lib.gscan_greedy_decode(dims, _param_array([None if p is None else p.detach() for p in params]),
                                 _ptr(commands), _ptr(cmd_len), _ptr(situations), int(max_decoding_steps), int(sos_idx),
                                 int(eos_idx), _ptr(ws), n_ws, _ptr(tokens), _ptr(out_len), _ptr(out_steps),
                                 _ptr(beta_sum), _ptr(aux), _ptr(alphas), _ptr(betas), _stream(dev))
    _lib.check(rc, "gscan_greedy_decode")
    _call_counts["greedy"] += 1
    return {"tokens": tokens, "lengths": out_len, "steps": out_steps, "beta_sum": beta_sum, "aux_logp": aux,
            "attention_weights_commands": alphas, "attention_weights_situations": betas}


def cnn_forward(cfg, params, situations, drop_cnn=None):
    lib = _lib.load()
    _require_cuda(situations)
    dev = situations.device
    situations = situations.contiguous().float()
    B = situations.shape[0]
    dims = make_dims(cfg, B, 1, 1, 1)
    M, D = dims.G * dims.G, 3 * dims.F
    n_ws = dims.C * dims.F * (1 + 25 + dims.K3 * dims.K3) + 16
    ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
    feat = torch.empty(B, M, D, dtype=torch.float32, device=dev)
    rc = lib.gscan_cnn_forward(dims, _param_array([None if p is None else p.detach() for p in params]),
                               _ptr(situations), _ptr(drop_cnn), _ptr(ws), n_ws, _ptr(feat), _stream(dev))
    _lib.check(rc, "gscan_cnn_forward")
    _call_counts["other"] += 1
    return feat


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0, grad_denom=None):
    """One fused Adam launch over the flat buffers; ``grad_denom`` (a 1-element CUDA tensor, e.g. the all-reduced
    token count inside the flat gradient buffer) divides the gradient on the device (gscan_adam_step_dev)."""
    lib = _lib.load()
    _require_cuda(param, grad, exp_avg, exp_avg_sq, grad_denom)
    n = param.numel()
    _lib.check(lib.gscan_adam_step_dev(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n, float(lr),
                                       float(beta1), float(beta2), float(eps), int(step), float(grad_scale),
                                       _ptr(grad_denom), _stream(param.device)), "gscan_adam_step_dev")
    _call_counts["other"] += 1
