"""ctypes binding of libgscan_b200.so (include/gscan_b200.h).

There is no CPU fallback: importing the package works without the library (so that a CPU-only
box can inspect checkpoints and parameter layouts), but any compute call raises ``RuntimeError``
when the shared library has not been built or no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_float, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
# GSCAN_LIB=<path> loads another build of the library (A/B experiments with tools/exp_build.sh); the default is the
# in-tree lib/libgscan_b200.so that build() produces
LIB_PATH = os.environ.get("GSCAN_LIB") or os.path.join(LIB_DIR, "libgscan_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NUM_PARAMS = 32
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class Dims(ctypes.Structure):
    """Mirror of ``gscan_dims`` (include/gscan_b200.h)."""
    _fields_ = [(n, c_int32) for n in (
        "B", "Ti", "Tt", "G", "C", "F", "K3", "E", "H", "Vi", "V", "conditional_attention",
        "auxiliary_task", "pad_idx_in", "pad_idx_out", "Ti_stride")]


class Dropout(ctypes.Structure):
    """Mirror of ``gscan_dropout``: dropout drawn inside the kernels from a Philox stream keyed by (seed, offset, site)."""
    _fields_ = [("p_cnn", c_float), ("p_enc", c_float), ("p_dec", c_float), ("seed", ctypes.c_uint64),
                ("offset", ctypes.c_uint64)]


ParamArray = c_void_p * NUM_PARAMS

# name -> (restype, argtypes); exactly the symbols include/gscan_b200.h declares
SIGNATURES = {
    "gscan_abi_version": (c_int32, []),
    "gscan_launch_count": (ctypes.c_ulonglong, []),
    "gscan_profile": (c_int32, [c_int32]),
    "gscan_profile_read": (c_int32, [c_void_p]),
    "gscan_check_dims": (c_int32, [POINTER(Dims)]),
    "gscan_workspace_floats": (c_size_t, [POINTER(Dims)]),
    "gscan_encode_workspace_floats": (c_size_t, [POINTER(Dims)]),
    "gscan_step_workspace_floats": (c_size_t, [POINTER(Dims)]),
    "gscan_greedy_workspace_floats": (c_size_t, [POINTER(Dims)]),
    "gscan_forward": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "gscan_backward": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                 POINTER(ParamArray), c_void_p]),
    "gscan_forward_rng": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                    POINTER(Dropout), c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "gscan_forward_train": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, POINTER(Dropout), c_void_p, c_size_t, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "gscan_backward_rng": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                     POINTER(Dropout), c_void_p, c_size_t, c_void_p, c_void_p, POINTER(ParamArray),
                                     c_void_p]),
    "gscan_dropout_mask": (c_int32, [POINTER(Dropout), c_int32, c_size_t, c_void_p, c_void_p]),
    "gscan_encode": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gscan_decoder_step": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "gscan_greedy_decode": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_int32,
                                      c_int32, c_int32, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "gscan_nll_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                    c_void_p]),
    "gscan_nll_count": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "gscan_nll_backward": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "gscan_metrics": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "gscan_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float,
                                  c_float, c_int32, c_float, c_void_p]),
    "gscan_adam_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float,
                                      c_float, c_int32, c_float, c_void_p, c_void_p]),
    "gscan_sgemm": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32,
                              c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p]),
    "gscan_sgemm_path": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32,
                                   c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "gscan_cnn_forward": (c_int32, [POINTER(Dims), POINTER(ParamArray), c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p, c_void_p]),
}

_ERRORS = {-1: "GSCAN_E_BADARG (null pointer or inconsistent dims)",
           -2: "GSCAN_E_UNSUPPORTED (shape outside what the kernels handle)",
           -3: "GSCAN_E_WORKSPACE (workspace too small)"}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/gscan_api.cu into lib/libgscan_b200.so for sm_100a (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libgscan_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR))]
    sources.append(os.path.join(INCLUDE_DIR, "gscan_b200.h"))
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in sources):
        return LIB_PATH
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        os.path.join(CSRC_DIR, "gscan_api.cu"), "-o", LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). multimodal_seq2seq_gscan_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gscan_abi_version() != 1:
        raise RuntimeError("libgscan_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what} failed: {_ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what} failed with cudaError_t {rc}")
