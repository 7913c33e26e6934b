"""gSCAN data tensorisation and batching without the GroundedScan / gym stack.

Replaces ``seq2seq/gSCAN_dataset.py`` (Vocabulary :17-102, GroundedScanDataset :105-315) for the simple
situation representation the paper runs use.  Same constructor, attributes and methods as far as
``train.py`` / ``predict.py`` use them, same vocabulary JSON files, and ``get_data_iterator`` yields the
same 9-tuple

    (input_batch [B,Ti] i64, input_lengths np.f64[B], derivation_repr list, situation_batch [B,G,G,C] f32,
     situation_repr list[dict], target_batch [B,Tt] i64, target_lengths np.f64[B],
     agent_positions [B] i64, target_positions [B] i64)

What is different is how it gets there:

* the dataset file (``dataset.txt``, JSON) is parsed directly: commands are comma-separated strings
  (GroundedScan/dataset.py:1235-1236) and the grid is built from the ``situation`` record exactly as
  ``Grid.encode`` does (gym_minigrid/minigrid.py:380-399: object vector in the leading channels, agent bit
  and one-hot direction in the last five, superimposed when agent and object share a cell) - no world
  simulation, no rendering;
* the reference keeps one dict of three tiny device tensors per example and pads every batch with a
  ``torch.cat`` per example (gSCAN_dataset.py:200-231; the logs show 57 minutes of tensorisation for one
  split).  Here the whole split lives in five pinned host arrays (tokens pre-padded, grids as uint8); a
  batch is five slices, copied to the device asynchronously, and the grid is widened to fp32 on the device.
"""
from __future__ import annotations

import json
import logging
import os
import random
from collections import Counter, defaultdict
from typing import Iterator, List, Optional

import numpy as np
import torch

logger = logging.getLogger(__name__)


class Vocabulary(object):
    """Word <-> index map with <PAD> = 0, <SOS> = 1, <EOS> = 2 and words numbered from 3 in order of first
    appearance (gSCAN_dataset.py:17-102; same JSON file format, so vocabulary files interoperate)."""

    def __init__(self, sos_token="<SOS>", eos_token="<EOS>", pad_token="<PAD>"):
        self.sos_token, self.eos_token, self.pad_token = sos_token, eos_token, pad_token
        self._idx_to_word = [pad_token, sos_token, eos_token]
        self._word_to_idx = {pad_token: 0, sos_token: 1, eos_token: 2}
        self._word_frequencies = Counter()

    def word_to_idx(self, word: str) -> int:
        return self._word_to_idx.get(word, 0)       # unknown words map to <PAD>, as the reference's defaultdict

    def idx_to_word(self, idx: int) -> str:
        return self._idx_to_word[idx]

    def contains_word(self, word: str) -> bool:
        return self.word_to_idx(word) != 0

    def add_sentence(self, sentence: List[str]) -> None:
        for word in sentence:
            if word not in self._word_to_idx:
                self._word_to_idx[word] = self.size
                self._idx_to_word.append(word)
            self._word_frequencies[word] += 1

    def most_common(self, n=10):
        return self._word_frequencies.most_common(n=n)

    @property
    def pad_idx(self) -> int:
        return self.word_to_idx(self.pad_token)

    @property
    def sos_idx(self) -> int:
        return self.word_to_idx(self.sos_token)

    @property
    def eos_idx(self) -> int:
        return self.word_to_idx(self.eos_token)

    @property
    def size(self) -> int:
        return len(self._idx_to_word)

    @classmethod
    def load(cls, path: str) -> "Vocabulary":
        assert os.path.exists(path), "Trying to load a vocabulary from a non-existing file {}".format(path)
        with open(path, "r") as infile:
            all_data = json.load(infile)
        vocab = cls(sos_token=all_data["sos_token"], eos_token=all_data["eos_token"], pad_token=all_data["pad_token"])
        vocab._idx_to_word = all_data["idx_to_word"]
        vocab._word_to_idx = {word: int(idx) for word, idx in all_data["word_to_idx"].items()}
        vocab._word_frequencies = Counter(all_data["word_frequencies"])
        return vocab

    def to_dict(self) -> dict:
        return {"sos_token": self.sos_token, "eos_token": self.eos_token, "pad_token": self.pad_token,
                "idx_to_word": self._idx_to_word, "word_to_idx": self._word_to_idx,
                "word_frequencies": self._word_frequencies}

    def save(self, path: str) -> str:
        with open(path, "w") as outfile:
            json.dump(self.to_dict(), outfile, indent=4)
        return path


def situation_grid(situation: dict) -> np.ndarray:
    """uint8 [G, G, C] grid of one ``situation`` record, indexed [row, column, channel], C = object attributes
    + 1 agent bit + 4 direction bits - the array ``Grid.encode`` returns (gym_minigrid/minigrid.py:380-399)."""
    G = int(situation["grid_size"])
    some_vector = (situation.get("target_object") or next(iter(situation["placed_objects"].values())))["vector"]
    n_attr = len(some_vector)
    grid = np.zeros((G, G, n_attr + 5), dtype=np.uint8)
    for placed in situation["placed_objects"].values():
        row, col = int(placed["position"]["row"]), int(placed["position"]["column"])
        grid[row, col, :n_attr] = np.frombuffer(placed["vector"].encode("ascii"), dtype=np.uint8) - ord("0")
    a_row, a_col = int(situation["agent_position"]["row"]), int(situation["agent_position"]["column"])
    grid[a_row, a_col, n_attr] = 1
    grid[a_row, a_col, n_attr + 1:] = 0
    grid[a_row, a_col, n_attr + 1 + int(situation["agent_direction"])] = 1
    return grid


class GroundedScanDataset(object):
    """Drop-in for ``seq2seq.gSCAN_dataset.GroundedScanDataset`` (simple situation representation only)."""

    def __init__(self, path_to_data: str, save_directory: str, k: int = 0, split="train", input_vocabulary_file="",
                 target_vocabulary_file="", generate_vocabulary=False, device: Optional[torch.device] = None):
        assert os.path.exists(path_to_data), "Trying to read a gSCAN dataset from a non-existing file {}.".format(
            path_to_data)
        if not generate_vocabulary:
            assert os.path.exists(os.path.join(save_directory, input_vocabulary_file)) and os.path.exists(
                os.path.join(save_directory, target_vocabulary_file)), \
                "Trying to load vocabularies from non-existing files."
        if split == "test" and generate_vocabulary:
            logger.warning("WARNING: generating a vocabulary from the test set.")
        with open(path_to_data, "r") as infile:
            all_data = json.load(infile)
        self.grid_size = int(all_data["grid_size"])
        # k-shot: k random adverb_1 examples move to train AND dev (GroundedScan/dataset.py:499-512; same
        # `random.sample` call, so a seeded `random` gives the reference's selection)
        self._data_pairs = defaultdict(list)
        for split_name, examples in all_data["examples"].items():
            chosen = set(random.sample(range(0, len(examples)), k=k)) if split_name == "adverb_1" else set()
            for i, example in enumerate(examples):
                if i in chosen:
                    self._data_pairs["train"].append(example)
                    self._data_pairs["dev"].append(example)
                else:
                    self._data_pairs[split_name].append(example)
        self.device = device if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.image_dimensions = None
        self.image_channels = 3
        self.split = split
        self.directory = save_directory
        self._n = 0
        self._commands = self._targets = self._grids = None          # pinned host tensors after read_dataset
        self._input_lengths = np.array([])
        self._target_lengths = np.array([])
        self._agent_positions = self._target_positions = None
        self._situation_reprs: List[dict] = []
        self._derivation_reprs: List[Optional[str]] = []
        if generate_vocabulary:
            logger.info("Generating vocabularies...")
            self.input_vocabulary = Vocabulary()
            self.target_vocabulary = Vocabulary()
            self.read_vocabularies()
            logger.info("Done generating vocabularies.")
        else:
            logger.info("Loading vocabularies...")
            self.input_vocabulary = Vocabulary.load(os.path.join(save_directory, input_vocabulary_file))
            self.target_vocabulary = Vocabulary.load(os.path.join(save_directory, target_vocabulary_file))
            logger.info("Done loading vocabularies.")

    # ---- vocabularies ---------------------------------------------------------------------------------
    def read_vocabularies(self) -> None:
        for example in self._data_pairs[self.split]:
            self.input_vocabulary.add_sentence(example["command"].split(","))
            self.target_vocabulary.add_sentence(example["target_commands"].split(","))

    def save_vocabularies(self, input_vocabulary_file: str, target_vocabulary_file: str) -> None:
        self.input_vocabulary.save(os.path.join(self.directory, input_vocabulary_file))
        self.target_vocabulary.save(os.path.join(self.directory, target_vocabulary_file))

    def get_vocabulary(self, vocabulary: str) -> Vocabulary:
        if vocabulary == "input":
            return self.input_vocabulary
        if vocabulary == "target":
            return self.target_vocabulary
        raise ValueError("Specified unknown vocabulary in sentence_to_array: {}".format(vocabulary))

    def sentence_to_array(self, sentence: List[str], vocabulary: str) -> List[int]:
        vocab = self.get_vocabulary(vocabulary)
        return [vocab.sos_idx] + [vocab.word_to_idx(word) for word in sentence] + [vocab.eos_idx]

    def array_to_sentence(self, sentence_array: List[int], vocabulary: str) -> List[str]:
        vocab = self.get_vocabulary(vocabulary)
        return [vocab.idx_to_word(int(word_idx)) for word_idx in sentence_array]

    # ---- tensorisation --------------------------------------------------------------------------------
    def read_dataset(self, max_examples=None, simple_situation_representation=True) -> None:
        """Tensorise the split once into pinned host arrays.  ``max_examples`` keeps the reference's stopping
        rule (gSCAN_dataset.py:245-247 tests ``len > max_examples`` BEFORE appending, so max_examples + 1
        examples are read)."""
        if not simple_situation_representation:
            raise NotImplementedError("only the simple (grid) situation representation is supported; the "
                                      "reference model refuses image input too (__main__.py:112-113)")
        logger.info("Converting dataset to tensors...")
        examples = self._data_pairs[self.split]
        if max_examples:
            examples = examples[:max_examples + 1]
        n = len(examples)
        inputs = [self.sentence_to_array(e["command"].split(","), "input") for e in examples]
        targets = [self.sentence_to_array(e["target_commands"].split(","), "target") for e in examples]
        self._input_lengths = np.array([len(a) for a in inputs], dtype=np.float64)
        self._target_lengths = np.array([len(a) for a in targets], dtype=np.float64)
        Ti = int(self._input_lengths.max()) if n else 0
        Tt = int(self._target_lengths.max()) if n else 0
        cmd = np.zeros((n, Ti), dtype=np.int64)
        tgt = np.zeros((n, Tt), dtype=np.int64)
        for i, (a, b) in enumerate(zip(inputs, targets)):
            cmd[i, :len(a)] = a
            tgt[i, :len(b)] = b
        grids = [situation_grid(e["situation"]) for e in examples]
        grid = np.stack(grids) if n else np.zeros((0, self.grid_size, self.grid_size, 0), dtype=np.uint8)
        G = self.grid_size
        agent = np.array([int(e["situation"]["agent_position"]["row"]) * int(e["situation"]["grid_size"])
                          + int(e["situation"]["agent_position"]["column"]) for e in examples], dtype=np.int64)
        target = np.array([int(e["situation"]["target_object"]["position"]["row"]) * int(e["situation"]["grid_size"])
                           + int(e["situation"]["target_object"]["position"]["column"]) for e in examples],
                          dtype=np.int64)
        pin = (lambda a: torch.from_numpy(a).pin_memory()) if torch.cuda.is_available() else torch.from_numpy
        self._commands, self._targets, self._grids = pin(cmd), pin(tgt), pin(grid)
        self._agent_positions, self._target_positions = pin(agent), pin(target)
        self._situation_reprs = [e["situation"] for e in examples]
        self._derivation_reprs = [e.get("derivation") for e in examples]
        self._n = n
        if n:
            self.image_dimensions = G
            self.image_channels = int(grid.shape[-1])

    def shuffle_data(self) -> None:
        """Reorder the examples with ``np.random.permutation`` (gSCAN_dataset.py:174-181)."""
        perm = np.random.permutation(self._n)
        tperm = torch.from_numpy(perm)
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        self._commands, self._targets, self._grids = (pin(self._commands[tperm]), pin(self._targets[tperm]),
                                                      pin(self._grids[tperm]))
        self._agent_positions, self._target_positions = pin(self._agent_positions[tperm]), pin(self._target_positions[tperm])
        self._input_lengths, self._target_lengths = self._input_lengths[perm], self._target_lengths[perm]
        self._situation_reprs = [self._situation_reprs[i] for i in perm]
        self._derivation_reprs = [self._derivation_reprs[i] for i in perm]

    def get_data_iterator(self, batch_size=10) -> Iterator[tuple]:
        """Batches in dataset order, each padded to ITS OWN maximal lengths (gSCAN_dataset.py:193-231); the last
        batch may be smaller."""
        dev = self.device
        for lo in range(0, self._n, batch_size):
            hi = min(lo + batch_size, self._n)
            input_lengths = self._input_lengths[lo:hi]
            target_lengths = self._target_lengths[lo:hi]
            Ti, Tt = int(np.max(input_lengths)), int(np.max(target_lengths))
            yield (self._commands[lo:hi, :Ti].to(dev, non_blocking=True), input_lengths,
                   self._derivation_reprs[lo:hi],
                   self._grids[lo:hi].to(dev, non_blocking=True).float(), self._situation_reprs[lo:hi],
                   self._targets[lo:hi, :Tt].to(dev, non_blocking=True), target_lengths,
                   self._agent_positions[lo:hi].to(dev, non_blocking=True),
                   self._target_positions[lo:hi].to(dev, non_blocking=True))

    @property
    def num_examples(self) -> int:
        return self._n

    @property
    def input_vocabulary_size(self) -> int:
        return self.input_vocabulary.size

    @property
    def target_vocabulary_size(self) -> int:
        return self.target_vocabulary.size
