"""Vendor the UNMODIFIED reference Python sources of the hot path into ``oracle/_ref/`` (git-ignored, but shipped
to the GPU box with the repo snapshot).  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # needs /root/reference (the build container); no-op elsewhere

The reference is ~1.1 kLoC of pure Python on PyTorch: nothing to compile, it only has to be
present next to the tests and the bench.  Nothing is copied into the repository history: ``oracle/_ref/`` is listed in
``.gitignore``.  Users: ``oracle/ref_loader.py`` (tests/test_gpu_reference_drivers.py, tests/test_ref_pinning.py and
``bench.py``'s reference arm / ``cpu_baseline`` / ``reference_gpu_eager`` legs).  The product package never imports it.
"""
import hashlib
import os
import sys
import zipfile

SRC = "/root/reference/seq2seq"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference_seq2seq.zip")   # one archive, imported through zipimport
FILES = ["__init__.py", "model.py", "seq2seq_model.py", "cnn_model.py", "helpers.py", "predict.py", "evaluate.py",
         "train.py", "gSCAN_dataset.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: nothing vendored (oracle/_ref stays as shipped)")
        return 0
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    lines = []
    with zipfile.ZipFile(DST, "w", zipfile.ZIP_DEFLATED) as z:
        for name in FILES:
            with open(os.path.join(SRC, name), "rb") as f:
                data = f.read()
            info = zipfile.ZipInfo(f"seq2seq/{name}", date_time=(2020, 1, 1, 0, 0, 0))   # reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
            lines.append(f"{hashlib.sha256(data).hexdigest()}  seq2seq/{name}")
    with open(os.path.join(HERE, "_ref", "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"vendored {len(FILES)} unmodified reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
