"""Import the vendored, UNMODIFIED reference (``oracle/_ref/reference_seq2seq.zip``, see make_ref.py; zipimport).  TEST / BASELINE
INFRASTRUCTURE ONLY - the product package never imports this.

The reference's data layer imports ``GroundedScan`` (the gym/MiniGrid dataset generator, absent here and out of
scope): a stub module stands in for it, exactly as tests/golden/make_golden.py does.  The reference picks its
device at import time (``seq2seq/predict.py:13``, ``train.py:12``: CUDA when visible); to run it on the host cores of
a GPU box, import it in a process started with CUDA_VISIBLE_DEVICES="" (bench.py does).
"""
import os
import sys
import types

REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_seq2seq.zip")


def available() -> bool:
    return os.path.exists(REF_ROOT)


def load():
    """Returns the imported reference package (``ref.model.Model``, ``ref.predict.predict``, ``ref.train.train``,
    ``ref.evaluate.evaluate``, ``ref.helpers``).  Raises FileNotFoundError when it has not been vendored."""
    if not available():
        raise FileNotFoundError(f"{REF_ROOT} is missing: run `python oracle/make_ref.py` in the build container")
    if "GroundedScan" not in sys.modules:
        stub = types.ModuleType("GroundedScan")
        stub_ds = types.ModuleType("GroundedScan.dataset")

        class GroundedScan:  # placeholder for the gym-dependent generator (never instantiated)
            pass

        stub_ds.GroundedScan = GroundedScan
        stub.dataset = stub_ds
        sys.modules["GroundedScan"] = stub
        sys.modules["GroundedScan.dataset"] = stub_ds
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    import seq2seq  # noqa: F401
    import seq2seq.model, seq2seq.predict, seq2seq.evaluate, seq2seq.train, seq2seq.helpers  # noqa: E401,F401
    if not os.path.abspath(seq2seq.__file__).startswith(REF_ROOT):
        raise RuntimeError(f"a different `seq2seq` package is already imported from {seq2seq.__file__}")
    return seq2seq
