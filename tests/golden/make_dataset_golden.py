"""Golden fixture for the data loader (tests/test_dataset_loader.py).

Runs the UNMODIFIED reference in this container (PYTHONPATH=/root/reference; it does not exist on the GPU box):
  1. `GroundedScan(...).get_data_pairs(...)` + `save_dataset` generate a small dataset.txt (uniform split,
     grid 6, default vocabulary: the compositional_splits channel count C = 16);
  2. `seq2seq.gSCAN_dataset.GroundedScanDataset` (vocabulary generation, read_dataset, get_data_iterator)
     tensorises it.
The modules the generator imports only for rendering / plotting (gym, PyQt5, matplotlib, cv2, imageio, xlwt,
pronounceable) are absent here and are replaced by empty stubs; none of them takes part in the arithmetic.
Outputs (committed): tests/golden/dataset_small.txt (trimmed to the examples used) and
tests/golden/dataset_small_expected.npz.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_dataset_golden.py
"""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Meta(type):
    def __getattr__(cls, k): return _Any()


class _Any(metaclass=_Meta):
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, k): return _Any()


gym = stub("gym", Env=object)
stub("gym.spaces", Discrete=_Any, Box=_Any, Dict=_Any)
gym.spaces = sys.modules["gym.spaces"]
stub("gym.utils")
stub("gym.utils.seeding", np_random=lambda seed=None: (np.random.RandomState(seed), seed))
sys.modules["gym.utils"].seeding = sys.modules["gym.utils.seeding"]
for name in ("imageio", "cv2", "pronounceable"):
    stub(name)
stub("xlwt", Workbook=_Any)
stub("matplotlib")
stub("matplotlib.pyplot")
stub("PyQt5")
for sub, names in {"QtCore": ["Qt", "QPoint", "QRect"], "QtGui": ["QImage", "QPixmap", "QPainter", "QColor", "QPolygon"],
                   "QtWidgets": ["QApplication", "QMainWindow", "QWidget", "QTextEdit", "QHBoxLayout", "QVBoxLayout",
                                 "QLabel", "QFrame"]}.items():
    stub("PyQt5." + sub, **{n: _Any for n in names})

import torch  # noqa: E402
from GroundedScan.dataset import GroundedScan  # noqa: E402
import GroundedScan.gym_minigrid.rendering as _rendering  # noqa: E402


class _NoRenderer:
    """The reference renders an RGB image per example even to collect the vocabulary (dataset.py:155-158 with
    the default simple_situation_representation=False); the pixels are never used on this path."""
    def __init__(self, width=0, height=0, *a, **k): self.window, self.width, self.height = None, width, height
    def getArray(self): return np.zeros((1, 1, 3), dtype=np.uint8)
    def __getattr__(self, k): return lambda *a, **kw: None


_rendering.Renderer = _NoRenderer

random.seed(7)
np.random.seed(7)
tmp = tempfile.mkdtemp()
gs = GroundedScan(intransitive_verbs=["walk"], transitive_verbs=["pull", "push"],
                  adverbs=["cautiously", "while spinning", "hesitantly", "while zigzagging"],
                  nouns=["square", "cylinder", "circle"], color_adjectives=["red", "green", "yellow", "blue"],
                  size_adjectives=["big", "small"], min_object_size=1, max_object_size=4, percentage_train=0.7,
                  percentage_dev=0.05, sample_vocabulary="default", save_directory=tmp, grid_size=6,
                  type_grammar="adverb")
gs.get_data_pairs(max_examples=60, num_resampling=1, other_objects_sample_percentage=0.5, split_type="uniform",
                  train_percentage=0.7, min_other_objects=0, k_shot_generalization=0, make_dev_set=True,
                  cut_off_target_length=None)
path = gs.save_dataset("dataset.txt")
all_data = json.load(open(path))
print({k: len(v) for k, v in all_data["examples"].items()})
# keep the file small: 23 train / 8 test examples, nothing else
all_data["examples"] = {"train": all_data["examples"]["train"][:23], "test": all_data["examples"]["test"][:8]}
small = os.path.join(HERE, "dataset_small.txt")
json.dump(all_data, open(small, "w"))

from seq2seq.gSCAN_dataset import GroundedScanDataset  # noqa: E402

out = {}
ds = GroundedScanDataset(small, tmp, k=0, split="train", input_vocabulary_file="in.txt", target_vocabulary_file="tg.txt",
                         generate_vocabulary=True)
ds.read_dataset(max_examples=None, simple_situation_representation=True)
ds.save_vocabularies("in.txt", "tg.txt")
out["input_vocab"] = np.array(ds.input_vocabulary._idx_to_word)
out["target_vocab"] = np.array(ds.target_vocabulary._idx_to_word)
out["num_examples"] = ds.num_examples
out["image_dimensions"], out["image_channels"] = ds.image_dimensions, ds.image_channels
for bi, batch in enumerate(ds.get_data_iterator(batch_size=10)):
    (inp, inp_len, deriv, sit, sit_repr, tgt, tgt_len, agent_pos, target_pos) = batch
    out[f"b{bi}_input"], out[f"b{bi}_input_lengths"] = inp.numpy(), np.asarray(inp_len)
    out[f"b{bi}_situation"] = sit.numpy()
    out[f"b{bi}_target"], out[f"b{bi}_target_lengths"] = tgt.numpy(), np.asarray(tgt_len)
    out[f"b{bi}_agent_positions"], out[f"b{bi}_target_positions"] = agent_pos.numpy(), target_pos.numpy()
    out[f"b{bi}_derivations"] = np.array(deriv)
out["num_batches"] = bi + 1
# the test split through the SAVED vocabularies, and max_examples (the reference reads max_examples + 1)
dv = GroundedScanDataset(small, tmp, k=0, split="test", input_vocabulary_file="in.txt", target_vocabulary_file="tg.txt",
                         generate_vocabulary=False)
dv.read_dataset(max_examples=3, simple_situation_representation=True)
out["test_num_examples"] = dv.num_examples
b = next(dv.get_data_iterator(batch_size=50))
out["test_input"], out["test_target"], out["test_situation"] = b[0].numpy(), b[5].numpy(), b[3].numpy()
json.dump(json.load(open(os.path.join(tmp, "in.txt"))), open(os.path.join(HERE, "dataset_small_input_vocab.json"), "w"))
json.dump(json.load(open(os.path.join(tmp, "tg.txt"))), open(os.path.join(HERE, "dataset_small_target_vocab.json"), "w"))
np.savez_compressed(os.path.join(HERE, "dataset_small_expected.npz"), **out)
print("wrote", small, os.path.getsize(small), "bytes;", out["num_examples"], "train examples in", out["num_batches"],
      "batches; dtype lengths", out["b0_input_lengths"].dtype, "situation", out["b0_situation"].shape,
      out["b0_situation"].dtype)
