"""The UNMODIFIED reference drivers run against `multimodal_seq2seq_gscan_b200.Model` on the GPU (north_star:
"drops into seq2seq/train.py and predict.py unchanged"; SURVEY.md section 7 test (v)).

The reference sources come from `oracle/_ref/reference_seq2seq.zip` (vendored byte for byte by `oracle/make_ref.py`
in the build container, git-ignored, shipped with the snapshot); only module-level NAMES are rebound before calling
them - `seq2seq.train.Model` / `seq2seq.train.GroundedScanDataset` point at this package's drop-ins - not a line
of the reference code is edited.

  * `seq2seq.predict.predict` (predict.py:57-128: encode_input, key_layer x 2, initialize_hidden, tanh,
    enc_hidden_to_dec_hidden, the decode_input loop) and `seq2seq.evaluate.evaluate` on the package Model give the
    sequences / accuracies the reference Model gave (goldens, incl. the 120-step one);
  * `seq2seq.train.train` (train.py:14-149: Model(**cfg), torch.optim.Adam over model.parameters(), LambdaLR,
    get_loss / get_auxiliary_loss / `loss +=` / backward / update_state / get_metrics) on the package Model gives the
    loss curve `FusedTrainer` gives through the package's own train() on the same batches, to 1e-5.
"""
import json
import logging
import os

import numpy as np
import pytest
import torch

from oracle import gscan_oracle as O
from oracle import ref_loader
from tests.golden_util import CASE_NAMES, load_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not vendored (python oracle/make_ref.py)")]

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _iterator(batch, dev):
    commands = torch.tensor(batch["commands"], device=dev)
    situations = torch.tensor(batch["situations"], device=dev)
    targets = torch.tensor(batch["targets"], device=dev)
    positions = torch.tensor(batch["target_positions"], device=dev)
    for b in range(commands.shape[0]):
        n_in, n_tg = int(batch["cmd_lengths"][b]), int(batch["tgt_lengths"][b])
        yield (commands[b:b + 1, :n_in], [n_in], [""], situations[b:b + 1], [{}], targets[b:b + 1, :n_tg], [n_tg],
               torch.zeros(1, dtype=torch.long, device=dev), positions[b:b + 1])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_reference_predict_and_evaluate_run_on_package_model(name):
    from tests.gpu_util import DEV, build_model
    ref = ref_loader.load()
    assert ref.predict.device.type == "cuda"        # the reference picked the GPU at import (predict.py:13)
    cfg, meta, params, batch, z = load_case(name, dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    N = int(z["greedy_max_steps"])
    with torch.no_grad():
        recs = list(ref.predict.predict(_iterator(batch, DEV), model=model, max_decoding_steps=N, pad_idx=0,
                                        sos_idx=1, eos_idx=2))
    assert len(recs) == batch["commands"].shape[0]
    for b, (_i, _d, _s, output_sequence, target_sequence, att_cmd, att_sit, aux_acc) in enumerate(recs):
        n = int(z["greedy_lengths"][b])
        assert output_sequence == z["greedy_sequences"][b, :n].tolist()
        assert len(att_cmd) == n and len(att_sit) == n
        acc = ref.helpers.sequence_accuracy(output_sequence, target_sequence[0].tolist()[1:-1])
        assert acc == pytest.approx(float(z["greedy_accuracy"][b]))
        if cfg["auxiliary_task"]:
            assert float(aux_acc) == float(z["greedy_aux_accuracy"][b])
    with torch.no_grad():
        acc, em, aux = ref.evaluate.evaluate(_iterator(batch, DEV), model=model, max_decoding_steps=N, pad_idx=0,
                                             sos_idx=1, eos_idx=2)
    assert acc == pytest.approx(float(np.mean(z["greedy_accuracy"])))
    assert em == pytest.approx(100.0 * float(np.mean(z["greedy_accuracy"] == 100)))


def test_reference_predict_full_length_on_package_model():
    from tests.gpu_util import DEV, build_model
    ref = ref_loader.load()
    cfg, meta, params, batch, z = load_case("greedy_long", dtype=torch.float32)
    model = build_model(cfg, params, train=False)
    with torch.no_grad():
        recs = list(ref.predict.predict(_iterator(batch, DEV), model=model, max_decoding_steps=120, pad_idx=0,
                                        sos_idx=1, eos_idx=2))
    for b, (_i, _d, _s, seq, _t, att_cmd, att_sit, aux_acc) in enumerate(recs):
        n, n_in = int(z["eos_lengths"][b]), int(batch["cmd_lengths"][b])
        assert seq == z["eos_sequences"][b, :n].tolist()
        for t in range(n):
            np.testing.assert_allclose(np.asarray(att_cmd[t]).reshape(-1), z["eos_alphas"][b, t, :n_in], atol=2e-5)
            np.testing.assert_allclose(np.asarray(att_sit[t]).reshape(-1), z["eos_betas"][b, t], atol=2e-5)
        assert float(aux_acc) == float(z["eos_aux_accuracy"][b])


def _train_args(tmp, **over):
    a = dict(data_path=os.path.join(tmp, "dataset.txt"), data_directory=tmp, generate_vocabularies=False,
             input_vocab_path="training_input_vocab.txt", target_vocab_path="training_target_vocab.txt",
             embedding_dimension=25, num_encoder_layers=1, encoder_dropout_p=0.0, encoder_bidirectional=True,
             training_batch_size=8, test_batch_size=1, max_decoding_steps=30, num_decoder_layers=1,
             decoder_dropout_p=0.0, cnn_kernel_size=7, cnn_dropout_p=0.0, cnn_hidden_num_channels=50,
             simple_situation_representation=True, decoder_hidden_size=100, encoder_hidden_size=100,
             learning_rate=2e-3, adam_beta_1=0.9, adam_beta_2=0.999, lr_decay=0.9, lr_decay_steps=20000,
             resume_from_file="", max_training_iterations=20, output_directory=tmp, print_every=5,
             evaluate_every=10, conditional_attention=True, auxiliary_task=True, weight_target_loss=0.3,
             attention_type="bahdanau", k=0, max_training_examples=None, seed=3, max_testing_examples=4)
    a.update(over)
    return a


def test_reference_train_loop_runs_on_package_model(tmp_path, caplog):
    import multimodal_seq2seq_gscan_b200 as pkg
    from multimodal_seq2seq_gscan_b200 import train as our_train
    from multimodal_seq2seq_gscan_b200.dataset import GroundedScanDataset
    from multimodal_seq2seq_gscan_b200.synthetic import full_state_dict
    from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
    ref = ref_loader.load()
    tmp = str(tmp_path)
    data = json.load(open(os.path.join(GOLD, "dataset_small.txt")))
    data["examples"]["dev"] = data["examples"].pop("test")
    json.dump(data, open(os.path.join(tmp, "dataset.txt"), "w"))
    ds = GroundedScanDataset(os.path.join(tmp, "dataset.txt"), tmp, split="train",
                             input_vocabulary_file="training_input_vocab.txt",
                             target_vocabulary_file="training_target_vocab.txt", generate_vocabulary=True)
    ds.save_vocabularies("training_input_vocab.txt", "training_target_vocab.txt")

    losses = {"ref": [], "ours": []}
    state = {}

    class RecordingModel(pkg.Model):
        """pkg.Model with a fixed initialisation; get_loss keeps the tensor the driver then adds the auxiliary loss
        to IN PLACE (train.py:107), so that its final value is the step's total loss at full precision."""

        def __init__(self, **kw):
            super().__init__(**kw)
            if "sd" not in state:
                cfg = dict(O.CONFIGS["comp"], input_vocabulary_size=kw["input_vocabulary_size"],
                           target_vocabulary_size=kw["target_vocabulary_size"],
                           num_cnn_channels=kw["num_cnn_channels"], auxiliary_task=True)
                state["sd"] = full_state_dict(O.synthetic_params(cfg, 77, scale=1.0))
            self.load_state_dict(state["sd"], strict=True)

        def get_loss(self, target_scores, targets):
            loss = super().get_loss(target_scores, targets)
            losses["ref"].append(loss)
            return loss

    # ---- the reference's train(), its names rebound to the drop-ins --------------------------------------------
    ref.train.Model = RecordingModel
    ref.train.GroundedScanDataset = GroundedScanDataset
    caplog.set_level(logging.INFO)
    np.random.seed(3)          # the reference seeds torch only (train.py:27); the shuffles draw from numpy
    ref.train.train(**_train_args(tmp))
    assert "Finished training." in caplog.text and "Evaluation Accuracy" in caplog.text
    ref_curve = [float(l) for l in losses["ref"]]
    assert len(ref_curve) == 20

    # ---- the package's train() on FusedTrainer, same initialisation, same batches ---------------------------------
    class SameInitModel(RecordingModel):
        def get_loss(self, target_scores, targets):
            return pkg.Model.get_loss(self, target_scores, targets)

    orig_step = FusedTrainer.train_step

    def recording_step(self, *a, **kw):
        loss = orig_step(self, *a, **kw)
        losses["ours"].append(loss)
        return loss

    our_train.Model = SameInitModel
    FusedTrainer.train_step = recording_step
    try:
        our_train.train(**_train_args(tmp))
    finally:
        FusedTrainer.train_step = orig_step
        our_train.Model = pkg.Model
    our_curve = [float(l) for l in losses["ours"]]
    assert len(our_curve) == 20
    np.testing.assert_allclose(our_curve, ref_curve, rtol=0, atol=1e-5)
    assert our_curve[-1] < our_curve[0]
