"""ctypes loader for the C ABI (include/tfhe_b200.h).  Fails loudly when the CUDA library is
missing: there is no Python or CPU fallback for the hot path."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_PATH = os.environ.get("TFHE_B200_LIB") or os.path.join(HERE, "lib", "libtfhe_b200.so")  # override: tuning experiments only
CLIENT_PATH = os.path.join(HERE, "lib", "libtfhe_b200_client.so")


class GateDesc(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint8), ("in0", ctypes.c_int32), ("in1", ctypes.c_int32), ("in2", ctypes.c_int32),
                ("out", ctypes.c_int32)]


class WireSection(ctypes.Structure):
    _fields_ = [("tag", ctypes.c_char * 4), ("dtype", ctypes.c_uint32), ("count", ctypes.c_uint64), ("data", ctypes.c_void_p)]


class TfheParams(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("N", ctypes.c_int32), ("L", ctypes.c_int32), ("bgbit", ctypes.c_int32),
                ("basebit", ctypes.c_int32), ("iks_t", ctypes.c_int32)]


ENGINE_SYMBOLS = [
    "tfhe_ctx_create", "tfhe_ctx_create_multi", "tfhe_ctx_device_count", "tfhe_ctx_set_pipeline_chunk", "tfhe_ctx_destroy", "tfhe_last_error", "tfhe_ctx_load_cloudkey",
    "tfhe_ctx_load_cloudkey_device", "tfhe_bootstrap_batch", "tfhe_bootstrap_batch_indexed", "tfhe_bootstrap_multi_lut_batch", "tfhe_ctx_load_reencryption_key", "tfhe_reencrypt_batch", "tfhe_ctx_set_mux_mode", "tfhe_ctx_set_circuit_graph", "tfhe_ctx_circuit_graph_replays", "tfhe_gate_batch", "tfhe_blind_rotate_batch",
    "tfhe_cmux_batch", "tfhe_sample_extract_batch", "tfhe_key_switch_batch", "tfhe_bootstrap_batch_device",
    "tfhe_gate_batch_device", "tfhe_circuit_run", "tfhe_to_fourier_batch", "tfhe_to_poly_batch", "tfhe_mul_poly_batch", "tfhe_ctx_kernel_launches", "tfhe_ctx_set_timing", "tfhe_ctx_set_blind_rotate_variant", "tfhe_ctx_set_blind_rotate_chunk_steps", "tfhe_ctx_set_key_switch_variant", "tfhe_ctx_generate_cloudkey", "tfhe_ctx_collect_timing", "tfhe_ctx_algorithmic_bytes_per_bootstrap", "tfhe_fp64_peak_probe", "tfhe_version",
]
CLIENT_SYMBOLS = [
    "tfhe_client_secret_key", "tfhe_client_encrypt_bool", "tfhe_client_decrypt_bool", "tfhe_client_encrypt_message",
    "tfhe_client_decrypt_message", "tfhe_client_gen_lut", "tfhe_client_cloud_key", "tfhe_client_chacha20_block", "tfhe_wire_pack", "tfhe_wire_unpack",
]

_engine = None
_client = None


def engine():
    """The CUDA engine library.  Raises if it has not been built (python -m go-tfhe_b200.build / __graft_entry__.build)."""
    global _engine
    if _engine is None:
        if not os.path.exists(ENGINE_PATH):
            raise RuntimeError("CUDA engine %s is missing: run __graft_entry__.build(); there is no CPU fallback"
                               % ENGINE_PATH)
        lib = ctypes.CDLL(ENGINE_PATH)
        vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32
        lib.tfhe_ctx_create.argtypes = [ctypes.POINTER(TfheParams), ctypes.c_int, ctypes.POINTER(vp)]
        if hasattr(lib, "tfhe_ctx_create_multi"):
            lib.tfhe_ctx_create_multi.argtypes = [ctypes.POINTER(TfheParams), ctypes.c_int, vp, ctypes.POINTER(vp)]
            lib.tfhe_ctx_device_count.argtypes = [vp]
            lib.tfhe_ctx_set_pipeline_chunk.argtypes = [vp, i64]
        lib.tfhe_ctx_destroy.argtypes = [vp]
        lib.tfhe_ctx_destroy.restype = None
        lib.tfhe_last_error.argtypes = [vp]
        lib.tfhe_last_error.restype = ctypes.c_char_p
        lib.tfhe_ctx_load_cloudkey.argtypes = [vp, u32, vp, vp, vp]
        lib.tfhe_ctx_load_cloudkey_device.argtypes = [vp, u32, vp, vp, vp, vp]
        lib.tfhe_bootstrap_batch.argtypes = [vp, i64, vp, vp, i64, vp]
        lib.tfhe_gate_batch.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp]
        if hasattr(lib, "tfhe_bootstrap_batch_indexed"):
            lib.tfhe_bootstrap_batch_indexed.argtypes = [vp, i64, vp, vp, i64, vp, vp]
            lib.tfhe_bootstrap_multi_lut_batch.argtypes = [vp, i64, vp, vp, i64, i32, vp]
            lib.tfhe_ctx_load_reencryption_key.argtypes = [vp, vp, i32, i32]
            lib.tfhe_reencrypt_batch.argtypes = [vp, i64, vp, vp]
            lib.tfhe_ctx_set_mux_mode.argtypes = [vp, ctypes.c_int]
        if hasattr(lib, "tfhe_ctx_set_circuit_graph"):
            lib.tfhe_ctx_set_circuit_graph.argtypes = [vp, ctypes.c_int]
            lib.tfhe_ctx_circuit_graph_replays.argtypes = [vp]
            lib.tfhe_ctx_circuit_graph_replays.restype = i64
        lib.tfhe_blind_rotate_batch.argtypes = [vp, i64, vp, vp, i64, vp]
        lib.tfhe_cmux_batch.argtypes = [vp, i64, i32, vp, vp, vp]
        lib.tfhe_sample_extract_batch.argtypes = [vp, i64, vp, vp]
        lib.tfhe_key_switch_batch.argtypes = [vp, i64, vp, vp]
        lib.tfhe_bootstrap_batch_device.argtypes = [vp, i64, vp, vp, i64, vp, vp]
        lib.tfhe_gate_batch_device.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, vp]
        lib.tfhe_to_fourier_batch.argtypes = [vp, i64, vp, vp]
        lib.tfhe_to_poly_batch.argtypes = [vp, i64, vp, vp]
        lib.tfhe_mul_poly_batch.argtypes = [vp, i64, vp, vp, vp]
        lib.tfhe_circuit_run.argtypes = [vp, i64, i32, i32, vp, vp, i32, vp, vp]
        lib.tfhe_ctx_set_timing.argtypes = [vp, ctypes.c_int]
        lib.tfhe_ctx_set_blind_rotate_variant.argtypes = [vp, ctypes.c_int]
        if hasattr(lib, "tfhe_ctx_set_blind_rotate_chunk_steps"):  # absent from pre-round-2 builds used in A/B runs
            lib.tfhe_ctx_set_blind_rotate_chunk_steps.argtypes = [vp, ctypes.c_int]
        lib.tfhe_ctx_set_key_switch_variant.argtypes = [vp, ctypes.c_int]
        lib.tfhe_ctx_generate_cloudkey.argtypes = [vp, vp, vp, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_int, vp, vp, vp, vp]
        lib.tfhe_ctx_collect_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_double * 4)]
        lib.tfhe_ctx_kernel_launches.argtypes = [vp]
        lib.tfhe_ctx_kernel_launches.restype = i64
        lib.tfhe_ctx_algorithmic_bytes_per_bootstrap.argtypes = [vp]
        lib.tfhe_ctx_algorithmic_bytes_per_bootstrap.restype = i64
        if hasattr(lib, "tfhe_fp64_peak_probe"):
            lib.tfhe_fp64_peak_probe.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double * 3)]
        lib.tfhe_version.restype = ctypes.c_char_p
        _engine = lib
    return _engine


def client():
    global _client
    if _client is None:
        if not os.path.exists(CLIENT_PATH):
            raise RuntimeError("client library %s is missing: run __graft_entry__.build()" % CLIENT_PATH)
        lib = ctypes.CDLL(CLIENT_PATH)
        vp, i64, i32, u64, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_double
        PP = ctypes.POINTER(TfheParams)
        lib.tfhe_client_secret_key.argtypes = [PP, u64, vp, vp]
        lib.tfhe_client_encrypt_bool.argtypes = [PP, dbl, vp, u64, i64, vp, vp]
        lib.tfhe_client_decrypt_bool.argtypes = [PP, vp, i64, vp, vp]
        lib.tfhe_client_encrypt_message.argtypes = [PP, dbl, vp, u64, i64, vp, i32, vp]
        lib.tfhe_client_decrypt_message.argtypes = [PP, vp, i64, vp, i32, vp]
        lib.tfhe_client_gen_lut.argtypes = [PP, i32, vp, vp]
        lib.tfhe_client_cloud_key.argtypes = [PP, dbl, dbl, vp, vp, u64, ctypes.c_int, vp, vp, vp, vp]
        lib.tfhe_client_chacha20_block.argtypes = [vp, ctypes.c_uint32, vp, vp]
        for s in CLIENT_SYMBOLS:
            getattr(lib, s).restype = None
        lib.tfhe_wire_pack.argtypes = [ctypes.c_uint32, PP, ctypes.POINTER(WireSection), ctypes.c_uint32, vp, i64]
        lib.tfhe_wire_pack.restype = i64
        lib.tfhe_wire_unpack.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_uint32), PP, ctypes.POINTER(WireSection),
                                         ctypes.POINTER(ctypes.c_uint32)]
        lib.tfhe_wire_unpack.restype = ctypes.c_int
        _client = lib
    return _client
