"""End-to-end check of the widened path on the GPU: gym-free loader -> Model -> FusedTrainer -> batched in-training
evaluation -> checkpoint -> resume -> predict_and_save, through the reference-shaped `train()` driver."""
import json
import logging
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _args(tmp, **over):
    a = dict(data_path=os.path.join(tmp, "dataset.txt"), data_directory=tmp, generate_vocabularies=True,
             input_vocab_path="training_input_vocab.txt", target_vocab_path="training_target_vocab.txt",
             embedding_dimension=25, num_encoder_layers=1, encoder_dropout_p=0.0, encoder_bidirectional=True,
             training_batch_size=8, test_batch_size=5, max_decoding_steps=30, num_decoder_layers=1,
             decoder_dropout_p=0.0, cnn_kernel_size=7, cnn_dropout_p=0.0, cnn_hidden_num_channels=50,
             simple_situation_representation=True, decoder_hidden_size=100, encoder_hidden_size=100,
             learning_rate=2e-3, adam_beta_1=0.9, adam_beta_2=0.999, lr_decay=0.9, lr_decay_steps=20000,
             resume_from_file="", max_training_iterations=30, output_directory=tmp, print_every=10, evaluate_every=15,
             conditional_attention=True, auxiliary_task=True, weight_target_loss=0.3, attention_type="bahdanau", k=0,
             max_training_examples=None, seed=3, max_testing_examples=None)
    a.update(over)
    return a


def test_train_driver_end_to_end(tmp_path, caplog):
    from multimodal_seq2seq_gscan_b200 import train as T, predict as P
    from multimodal_seq2seq_gscan_b200.dataset import GroundedScanDataset
    tmp = str(tmp_path)
    data = json.load(open(os.path.join(GOLD, "dataset_small.txt")))
    data["examples"]["dev"] = data["examples"].pop("test")          # the driver evaluates on the dev split
    json.dump(data, open(os.path.join(tmp, "dataset.txt"), "w"))
    caplog.set_level(logging.INFO)
    out = T.train(**_args(tmp))
    assert out["iterations"] == 30
    text = caplog.text
    assert "Iteration 00000010, loss" in text and "Evaluation Accuracy" in text and "Finished training." in text
    model = out["model"]
    assert model.trained_iterations >= 30
    # the loss goes down on 23 examples in 30 steps
    first = float(text.split("Iteration 00000010, loss")[1].split(",")[0])
    last = float(text.split("Iteration 00000030, loss")[1].split(",")[0])
    assert last < first, (first, last)
    # vocabularies were written in the reference's format and a best checkpoint, if any, is loadable
    assert os.path.exists(os.path.join(tmp, "training_input_vocab.txt"))
    ckpt = os.path.join(tmp, "checkpoint.pth.tar")
    if os.path.exists(ckpt):
        state = torch.load(ckpt, map_location="cpu", weights_only=False)
        assert set(state) == {"iteration", "state_dict", "best_iteration", "best_accuracy", "best_exact_match",
                              "optimizer_state_dict"}
        assert len(state["state_dict"]) == 38
        # resume: picks up the iteration count and the Adam moments, runs on
        out2 = T.train(**_args(tmp, resume_from_file=ckpt, generate_vocabularies=False, max_training_iterations=40))
        assert out2["iterations"] == 40 and out2["trainer"].step_count > 10
    # predictions JSON over the dev split, batched
    dev = GroundedScanDataset(os.path.join(tmp, "dataset.txt"), tmp, split="dev",
                              input_vocabulary_file="training_input_vocab.txt",
                              target_vocabulary_file="training_target_vocab.txt", generate_vocabulary=False)
    dev.read_dataset()
    path = P.predict_and_save(dev, model, os.path.join(tmp, "predict.json"), max_decoding_steps=30, batch_size=3)
    preds = json.load(open(path))
    assert len(preds) == dev.num_examples == 8
    assert all(len(p["attention_weights_situation"]) == len(p["prediction"]) for p in preds)
    # batch composition does not change the decoded sequences
    path1 = P.predict_and_save(dev, model, os.path.join(tmp, "predict1.json"), max_decoding_steps=30, batch_size=1)
    preds1 = json.load(open(path1))
    assert [p["prediction"] for p in preds] == [p["prediction"] for p in preds1]
    acc = float(np.mean([p["accuracy"] for p in preds]))
    ev = P.evaluate(dev.get_data_iterator(batch_size=4), model, 30, 0, 1, 2)
    assert ev[0] == pytest.approx(acc)
