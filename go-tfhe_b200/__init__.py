"""go-tfhe_b200 — B200-native engine for the TFHE gate-bootstrap hot path of thedonutfactory/go-tfhe.

The package holds only what the path needs: csrc/ (CUDA kernels + the C ABI of include/tfhe_b200.h, and
the host-side client helpers) and a thin host-side mirror of the reference's packages for that path:
params, key, tlwe, lut, cloudkey, evaluator, gates.  All hot-path compute runs in lib/libtfhe_b200.so
(hand-written sm_100a CUDA); there is no Python or CPU fallback.

The directory name contains a hyphen; import it with importlib.import_module("go-tfhe_b200").
"""
from . import _native, params, engine, key, tlwe, lut, poly, cloudkey, evaluator, gates, sharding, circuit, wire  # noqa: F401
from .build import build  # noqa: F401
from .engine import Context, TfheError, OPCODES  # noqa: F401

__all__ = ["params", "engine", "key", "tlwe", "lut", "poly", "cloudkey", "evaluator", "gates", "sharding", "circuit", "wire", "build", "Context",
           "TfheError", "OPCODES"]
