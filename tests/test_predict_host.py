"""Host logic of the batched predict / evaluate / predict_and_save drivers (no GPU: the decode kernel is
replaced by a stub that returns canned tokens), checked against the oracle's restatement of the reference
helpers and against the reference's record layout (predict.py:44-51, 118-120)."""
import json
import random

import numpy as np
import torch

from multimodal_seq2seq_gscan_b200 import predict as P
from oracle import gscan_oracle as O


def test_sequence_accuracy_matches_reference_helper():
    rng = random.Random(3)
    for _ in range(500):
        a = [rng.randrange(0, 5) for _ in range(rng.randrange(0, 7))]
        b = [rng.randrange(0, 5) for _ in range(rng.randrange(0, 7))]
        assert P.sequence_accuracy(a, b) == O.sequence_accuracy(a, b)
    assert P.sequence_accuracy([], []) == 0.0
    assert P.sequence_accuracy([], [3, 4]) == 0.0          # empty prediction is padded with 0s (trap A.4-6)
    assert P.sequence_accuracy([3, 4], [3, 4]) == 100.0


class StubModel:
    """greedy_decode returns what the kernel would: tokens [B, N+1] (-1 beyond length), lengths, attention."""
    auxiliary_task = True

    def __init__(self, table):
        self.table = table      # command first-word -> generated tokens (EOS already removed)
        self.calls = 0

    def eval(self):
        return self

    def greedy_decode(self, commands, lengths, situations, max_steps, sos, eos, return_attention=False):
        self.calls += 1
        B, Ti = commands.shape
        T = max_steps + 1
        toks = torch.full((B, T), -1, dtype=torch.int64)
        lens = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            seq = self.table[int(commands[b, 1])][:T]
            toks[b, :len(seq)] = torch.tensor(seq, dtype=torch.int64)
            lens[b] = len(seq)
        al = torch.rand(B, T, Ti)
        be = torch.rand(B, T, 4)
        aux = torch.zeros(B, 4)
        aux[:, 1] = 1.0          # always predicts position 1
        return {"tokens": toks, "lengths": lens, "steps": lens + 1, "beta_sum": be.sum(1), "aux_logp": aux,
                "attention_weights_commands": al if return_attention else None,
                "attention_weights_situations": be if return_attention else None}


def make_batches():
    # two ragged batches (3 + 2 examples); PAD 0, SOS 1, EOS 2
    cmds1 = torch.tensor([[1, 3, 4, 2], [1, 4, 2, 0], [1, 5, 3, 2]])
    cmds2 = torch.tensor([[1, 3, 2], [1, 5, 2]])
    tg1 = torch.tensor([[1, 3, 3, 2, 0], [1, 4, 2, 0, 0], [1, 3, 4, 5, 2]])
    tg2 = torch.tensor([[1, 3, 3, 2], [1, 4, 4, 2]])
    sit = lambda n: torch.zeros(n, 2, 2, 5)
    return [
        (cmds1, np.array([4., 3., 4.]), ["d0", "d1", "d2"], sit(3), [{"s": 0}, {"s": 1}, {"s": 2}], tg1,
         np.array([4., 3., 5.]), torch.zeros(3, dtype=torch.long), torch.tensor([1, 0, 1])),
        (cmds2, np.array([3., 3.]), ["d3", "d4"], sit(2), [{"s": 3}, {"s": 4}], tg2, np.array([4., 4.]),
         torch.zeros(2, dtype=torch.long), torch.tensor([1, 1])),
    ]


TABLE = {3: [3, 3], 4: [], 5: [3, 4, 4, 4, 4, 4, 4]}   # exact match, empty output, over-long output


def test_predict_yields_reference_records_per_example():
    model = StubModel(TABLE)
    recs = list(P.predict(iter(make_batches()), model, max_decoding_steps=5, pad_idx=0, sos_idx=1, eos_idx=2))
    assert model.calls == 2 and len(recs) == 5          # one decode per BATCH, one record per EXAMPLE
    inp, deriv, sit, out, tgt, att_c, att_s, aux = recs[0]
    assert inp.shape == (1, 4) and tgt.shape == (1, 4) and deriv == ["d0"] and sit == [{"s": 0}]
    assert out == [3, 3] and aux == 100.0
    assert len(att_c) == 2 and len(att_c[0]) == 1 and len(att_c[0][0]) == 4      # steps x [1][Ti_b]
    assert len(att_s) == 2 and len(att_s[0][0]) == 4                               # steps x [1][G*G]
    # padded command: the attention row is cut to the example's own length, as a batch-1 run would give
    assert recs[1][0].shape == (1, 3) and recs[1][3] == [] and recs[1][5] == [] and recs[1][7] == 0.0
    assert recs[2][3] == [3, 4, 4, 4, 4, 4]            # N + 1 = 6 tokens at most
    assert [r[1] for r in recs] == [["d0"], ["d1"], ["d2"], ["d3"], ["d4"]]
    # max_examples_to_evaluate counts examples and stops inside a batch
    few = list(P.predict(iter(make_batches()), StubModel(TABLE), 5, 0, 1, 2, max_examples_to_evaluate=4))
    assert len(few) == 4


def test_evaluate_matches_per_example_formula():
    acc, em, aux = P.evaluate(iter(make_batches()), StubModel(TABLE), 5, 0, 1, 2)
    targets = [[3, 3], [4], [3, 4, 5], [3, 3], [4, 4]]
    outs = [[3, 3], [], [3, 4, 4, 4, 4, 4], [3, 3], [3, 4, 4, 4, 4, 4]]
    accs = [O.sequence_accuracy(o, t) for o, t in zip(outs, targets)]
    assert acc == float(np.mean(accs))
    assert em == 100.0 * sum(a == 100 for a in accs) / 5
    assert aux == float(np.mean([100.0, 0.0, 100.0, 100.0, 100.0]))


class StubVocab:
    pad_idx, sos_idx, eos_idx = 0, 1, 2


class StubDataset:
    target_vocabulary = StubVocab()
    words = {0: "<PAD>", 1: "<SOS>", 2: "<EOS>", 3: "walk", 4: "turn", 5: "push"}

    def get_data_iterator(self, batch_size):
        assert batch_size == 3
        return iter(make_batches())

    def array_to_sentence(self, arr, vocabulary):
        return [self.words[int(i)] for i in arr]


def test_predict_and_save_schema(tmp_path):
    path = P.predict_and_save(StubDataset(), StubModel(TABLE), str(tmp_path / "predict.json"), max_decoding_steps=5,
                              batch_size=3)
    data = json.load(open(path))
    assert len(data) == 5
    assert set(data[0]) == {"input", "prediction", "derivation", "target", "situation", "attention_weights_input",
                            "attention_weights_situation", "accuracy", "exact_match", "position_accuracy"}
    assert data[0]["input"] == ["walk", "turn"] and data[0]["target"] == ["walk", "walk"]
    assert data[0]["prediction"] == ["walk", "walk"] and data[0]["exact_match"] is True
    assert data[1]["prediction"] == [] and data[1]["accuracy"] == 0.0 and data[1]["exact_match"] is False
