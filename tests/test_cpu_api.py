"""CPU-only checks of the boundary: the C-ABI library loads and exports every declared symbol, the
drop-in Model has the reference's parameter registry / state_dict layout / checkpoint format, and the
product path refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import _lib
from oracle import gscan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return pkg.build()


def test_library_exports_every_declared_symbol(libpath):
    header = open(os.path.join(ROOT, "include", "gscan_b200.h")).read()
    declared = set(re.findall(r"\b(gscan_[a-z_0-9]+)\s*\(", header))
    declared -= {"gscan_check_dims)"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(libpath)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.gscan_abi_version.restype = ctypes.c_int32
    assert lib.gscan_abi_version() == 1


def test_workspace_queries_and_dim_checks(libpath):
    lib = _lib.load()
    cfg = O.CONFIGS["comp"]
    d = _lib.Dims(B=200, Ti=10, Tt=121, G=6, C=16, F=50, K3=7, E=25, H=100, Vi=21, V=9, conditional_attention=1,
                  auxiliary_task=0, pad_idx_in=0, pad_idx_out=0, Ti_stride=10)
    n = lib.gscan_workspace_floats(d)
    assert 40e6 < n < 120e6          # ~ 300 MB of fp32 saved activations + scratch at the headline shape
    assert lib.gscan_encode_workspace_floats(d) < n
    d.H = 102                        # not a multiple of 4
    assert lib.gscan_check_dims(d) == -2
    d.H = 100
    d.K3 = 6                         # even kernels have no 'same' padding in the reference either
    assert lib.gscan_check_dims(d) == -2
    d.K3 = 7
    d.Ti_stride = 5
    assert lib.gscan_check_dims(d) == -1


@pytest.mark.parametrize("cfg_name,cond", [("comp", True), ("comp", False), ("demo", True), ("tlen", True)])
def test_parameter_registry_matches_reference_order(cfg_name, cond):
    cfg = dict(O.CONFIGS[cfg_name])
    cfg["conditional_attention"] = cond
    model = pkg.Model(**O.model_kwargs(cfg))
    names = [n for n, _ in model.named_parameters()]
    expected = O.param_shapes(cfg)
    assert names == [n for n, _ in expected]
    for (n, p), (_, shape) in zip(model.named_parameters(), expected):
        assert tuple(p.shape) == tuple(shape), n
    sd = model.state_dict()
    aliases = [f"attention_decoder.{a}.{l}.weight" for a in ("textual_attention", "visual_attention")
               for l in ("key_layer", "query_layer", "energy_layer")]
    assert set(sd) == set(names) | set(aliases)
    assert len(sd) == len(names) + 6
    # the 32 tensors handed to the C ABI are the parameters, in order (None for absent conditional layer)
    plist = [p for p in model._param_list() if p is not None]
    assert [id(p) for p in plist] == [id(p) for p in model.parameters()]
    if cfg_name == "comp" and cond:
        assert sum(p.numel() for p in model.parameters()) == 440275   # adverb_run_1.txt:58
    if cfg_name == "tlen":
        assert sum(p.numel() for p in model.parameters()) == 535975   # target_lengths_run_1.txt:79


def test_constructor_errors_and_extra_kwargs():
    cfg = O.model_kwargs(O.CONFIGS["demo"])
    pkg.Model(**cfg, mode="train", data_path="x", k=0)       # unknown flags are swallowed like the reference
    with pytest.raises(ValueError):
        pkg.Model(**{**cfg, "attention_type": "nope"})
    with pytest.raises(NotImplementedError):
        pkg.Model(**{**cfg, "simple_situation_representation": False})


def test_no_cpu_fallback():
    cfg = dict(O.CONFIGS["tiny"])
    model = pkg.Model(**O.model_kwargs(cfg))
    batch = O.synthetic_batch(cfg, batch_size=2, seed=1, max_cmd_len=5, min_cmd_len=3, max_tgt_len=5)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(commands_input=torch.tensor(batch["commands"]), commands_lengths=batch["cmd_lengths"],
              situations_input=torch.tensor(batch["situations"]), target_batch=torch.tensor(batch["targets"]),
              target_lengths=batch["tgt_lengths"])
    with pytest.raises(RuntimeError, match="CUDA"):
        model.get_loss(torch.zeros(2, 5, 7), torch.tensor(batch["targets"]))


def test_checkpoint_layout_round_trip(tmp_path):
    cfg = dict(O.CONFIGS["demo"])
    cfg["output_directory"] = str(tmp_path)
    model = pkg.Model(**O.model_kwargs(cfg))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.update_state(is_best=False)
    model.update_state(is_best=True, accuracy=91.0, exact_match=42.0)
    path = model.save_checkpoint("checkpoint.pth.tar", is_best=True, optimizer_state_dict=opt.state_dict())
    assert os.path.exists(os.path.join(str(tmp_path), "model_best.pth.tar"))
    ckpt = torch.load(path)
    assert set(ckpt) == {"iteration", "state_dict", "best_iteration", "best_accuracy", "best_exact_match",
                         "optimizer_state_dict"}
    other = pkg.Model(**O.model_kwargs(cfg))
    opt_state = other.load_model(path)
    assert other.trained_iterations == 2 and other.best_iteration == 2
    assert other.best_exact_match == 42.0 and other.best_accuracy == 91.0
    for (n, a), (_, b) in zip(model.named_parameters(), other.named_parameters()):
        assert torch.equal(a, b), n
    torch.optim.Adam(other.parameters(), lr=1e-3).load_state_dict(opt_state)


@pytest.mark.skipif(not os.path.isdir("/root/reference/seq2seq"), reason="reference checkout not present")
def test_state_dict_interchangeable_with_reference():
    from tests.golden.make_golden import import_reference
    RefModel, _, _ = import_reference()
    cfg = dict(O.CONFIGS["comp"])
    ref = RefModel(**O.model_kwargs(cfg))
    ours = pkg.Model(**O.model_kwargs(cfg))
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
