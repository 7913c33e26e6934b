"""The gym-free loader (multimodal_seq2seq_gscan_b200/dataset.py) against what the UNMODIFIED reference loader
(seq2seq/gSCAN_dataset.py on top of GroundedScan) produced for the same dataset file - fixture made by
tests/golden/make_dataset_golden.py."""
import json
import os

import numpy as np
import pytest
import torch

from multimodal_seq2seq_gscan_b200.dataset import GroundedScanDataset, Vocabulary, situation_grid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA = os.path.join(GOLD, "dataset_small.txt")
CPU = torch.device("cpu")


@pytest.fixture(scope="module")
def expected():
    return np.load(os.path.join(GOLD, "dataset_small_expected.npz"), allow_pickle=False)


def test_train_split_matches_reference_loader(expected, tmp_path):
    ds = GroundedScanDataset(DATA, str(tmp_path), k=0, split="train", input_vocabulary_file="in.txt",
                             target_vocabulary_file="tg.txt", generate_vocabulary=True, device=CPU)
    ds.read_dataset(max_examples=None, simple_situation_representation=True)
    assert ds.input_vocabulary._idx_to_word == expected["input_vocab"].tolist()
    assert ds.target_vocabulary._idx_to_word == expected["target_vocab"].tolist()
    assert ds.num_examples == int(expected["num_examples"])
    assert ds.image_dimensions == int(expected["image_dimensions"])
    assert ds.image_channels == int(expected["image_channels"]) == 16
    assert ds.input_vocabulary_size == len(expected["input_vocab"])
    batches = list(ds.get_data_iterator(batch_size=10))
    assert len(batches) == int(expected["num_batches"]) == 3
    for bi, (inp, inp_len, deriv, sit, sit_repr, tgt, tgt_len, agent_pos, target_pos) in enumerate(batches):
        assert inp.dtype == torch.int64 and tgt.dtype == torch.int64 and sit.dtype == torch.float32
        assert isinstance(inp_len, np.ndarray) and inp_len.dtype == np.float64      # trap A.4-13
        np.testing.assert_array_equal(inp.numpy(), expected[f"b{bi}_input"])
        np.testing.assert_array_equal(inp_len, expected[f"b{bi}_input_lengths"])
        np.testing.assert_array_equal(sit.numpy(), expected[f"b{bi}_situation"])
        np.testing.assert_array_equal(tgt.numpy(), expected[f"b{bi}_target"])
        np.testing.assert_array_equal(tgt_len, expected[f"b{bi}_target_lengths"])
        np.testing.assert_array_equal(agent_pos.numpy(), expected[f"b{bi}_agent_positions"])
        np.testing.assert_array_equal(target_pos.numpy(), expected[f"b{bi}_target_positions"])
        assert list(deriv) == expected[f"b{bi}_derivations"].tolist()
        assert len(sit_repr) == inp.shape[0] and "placed_objects" in sit_repr[0]
    assert batches[-1][0].shape[0] == 3          # 23 = 10 + 10 + 3: the last batch is smaller
    # vocabulary files have the reference's JSON layout
    ds.save_vocabularies("in.txt", "tg.txt")
    ref_in = json.load(open(os.path.join(GOLD, "dataset_small_input_vocab.json")))
    mine = json.load(open(tmp_path / "in.txt"))
    assert mine["idx_to_word"] == ref_in["idx_to_word"] and mine["word_frequencies"] == ref_in["word_frequencies"]
    assert {k: v for k, v in ref_in["word_to_idx"].items() if v} == {k: v for k, v in mine["word_to_idx"].items() if v}


def test_saved_vocabularies_and_max_examples_rule(expected, tmp_path):
    for name in ("input", "target"):
        src = os.path.join(GOLD, f"dataset_small_{name}_vocab.json")
        (tmp_path / f"{name}.txt").write_text(open(src).read())
    ds = GroundedScanDataset(DATA, str(tmp_path), k=0, split="test", input_vocabulary_file="input.txt",
                             target_vocabulary_file="target.txt", generate_vocabulary=False, device=CPU)
    ds.read_dataset(max_examples=3)
    assert ds.num_examples == int(expected["test_num_examples"]) == 4      # the reference reads max_examples + 1
    b = next(ds.get_data_iterator(batch_size=50))
    np.testing.assert_array_equal(b[0].numpy(), expected["test_input"])
    np.testing.assert_array_equal(b[5].numpy(), expected["test_target"])
    np.testing.assert_array_equal(b[3].numpy(), expected["test_situation"])
    assert ds.array_to_sentence(b[0][0].tolist(), "input")[0] == "<SOS>"
    assert ds.sentence_to_array(["walk", "never-seen-word"], "input")[2] == 0       # unknown -> <PAD>, as the reference


def test_grid_superimposes_agent_and_object():
    sit = {"grid_size": 3, "agent_position": {"row": "1", "column": "2"}, "agent_direction": 3,
           "target_object": {"vector": "0100101", "position": {"row": "1", "column": "2"}, "object": {}},
           "placed_objects": {"0": {"vector": "0100101", "position": {"row": "1", "column": "2"}, "object": {}},
                              "1": {"vector": "1000010", "position": {"row": "0", "column": "0"}, "object": {}}}}
    g = situation_grid(sit)
    assert g.shape == (3, 3, 12) and g.dtype == np.uint8
    assert g[1, 2].tolist() == [0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1]      # object bits AND agent bit + direction 3
    assert g[0, 0].tolist() == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    assert g.sum() == 3 + 2 + 2


def test_shuffle_keeps_examples_aligned(tmp_path):
    ds = GroundedScanDataset(DATA, str(tmp_path), split="train", generate_vocabulary=True, device=CPU)
    ds.read_dataset()
    before = {tuple(c.tolist()) + tuple(t.tolist()) + (int(p),)
              for b in ds.get_data_iterator(5) for c, t, p in zip(b[0], b[5], b[8])}
    np.random.seed(0)
    ds.shuffle_data()
    after = set()
    for b in ds.get_data_iterator(7):
        for i, (c, t, p) in enumerate(zip(b[0], b[5], b[8])):
            assert int(b[1][i]) == int((c != 0).sum()) and int(b[6][i]) == int((t != 0).sum())
            row = int(b[4][i]["target_object"]["position"]["row"]); col = int(b[4][i]["target_object"]["position"]["column"])
            assert int(p) == row * 6 + col
    assert ds.num_examples == 23


def test_vocabulary_roundtrip(tmp_path):
    v = Vocabulary()
    v.add_sentence(["walk", "to", "a", "walk"])
    assert (v.pad_idx, v.sos_idx, v.eos_idx, v.size) == (0, 1, 2, 6)
    assert v.word_to_idx("walk") == 3 and v.word_to_idx("nope") == 0 and not v.contains_word("nope")
    v.save(str(tmp_path / "v.json"))
    w = Vocabulary.load(str(tmp_path / "v.json"))
    assert w._idx_to_word == v._idx_to_word and w.most_common(1) == [("walk", 2)]
