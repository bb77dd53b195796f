"""ctypes binding of libb200msm.so -- the C ABI declared in include/b200msm.h.

There is deliberately no fallback: if the shared library is missing, importing this module raises,
and if no GPU is usable every compute call returns B200MSM_E_CUDA which is raised as B200MsmError.
"""
import ctypes, os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MSM_LIB") or os.path.join(_HERE, "libb200msm.so")     # B200MSM_LIB: alternative build of the same ABI (A/B experiments)

OK, E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4
BLS12_381_G1, BN254_G1, BLS12_381_G2, BN254_G2 = 0, 1, 2, 3
N8 = {BLS12_381_G1: 48, BN254_G1: 32, BLS12_381_G2: 96, BN254_G2: 64}     # bytes per coordinate-field element (Fq / Fq2)

EXPORTS = [  # every symbol include/b200msm.h and include/b200msm_probes.h declare (checked by tests/test_abi.py)
    "b200msm_create", "b200msm_create_multi", "b200msm_device_count", "b200msm_destroy", "b200msm_strerror", "b200msm_last_error", "b200msm_version",
    "b200msm_set_stream", "b200msm_synchronize", "b200msm_g1_multiexp_affine", "b200msm_g1_multiexp_affine_chunk",
    "b200msm_g1_multiexp", "b200msm_g1_multiexp_chunk",
    "b200msm_upload_bases", "b200msm_free_bases", "b200msm_g1_multiexp_resident", "b200msm_g1_normalize",
    "b200msm_g1_sum", "b200msm_g1_generate_bases", "b200msm_fq_op", "b200msm_probe_imad", "b200msm_probe_imad32", "b200msm_probe_fqmul",
    "b200msm_set_option", "b200msm_constants", "b200msm_get_counter", "b200msm_g1_batch_convert",
    "b200msm_glv_decompose_scalars", "b200msm_g1_glv_preprocess", "b200msm_upload_bases_windowed", "b200msm_g1_multiexp_batch", "b200msm_fr_fft", "b200msm_fr_fft_last_phases",
    "b200msm_fq_batch_inverse", "b200msm_debug_schedule",
]
EXPERIMENT_EXPORTS = ["b200msm_probe_dfma", "b200msm_probe_dualpipe"]      # only in -DB200_EXPERIMENTS builds (include/b200msm_probes.h)


class Stats(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint32), ("window_bits", ctypes.c_uint32), ("windows", ctypes.c_uint32),
                ("buckets_per_window", ctypes.c_uint32), ("tree_rounds", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("pairs", ctypes.c_uint64), ("affine_adds", ctypes.c_uint64),
                ("ms_total", ctypes.c_float), ("ms_h2d", ctypes.c_float), ("ms_digits_sort", ctypes.c_float),
                ("ms_accumulate", ctypes.c_float), ("ms_bucket_reduce", ctypes.c_float), ("ms_window_combine", ctypes.c_float),
                ("ms_d2h", ctypes.c_float),
                ("ms_k_sort", ctypes.c_float), ("ms_k_plan", ctypes.c_float), ("ms_k_tree_fwd", ctypes.c_float), ("ms_k_inv_tree", ctypes.c_float),
                ("ms_k_tree_bwd", ctypes.c_float), ("ms_k_finish", ctypes.c_float), ("ms_k_fold", ctypes.c_float), ("ms_k_wsum", ctypes.c_float),
                ("ms_k_horner", ctypes.c_float), ("ms_host_combine", ctypes.c_float), ("launches", ctypes.c_uint64),
                ("affine_adds_round0", ctypes.c_uint64), ("ms_k_tree_bwd_round0", ctypes.c_float), ("reserved3", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class B200MsmError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        super().__init__("b200msm status %d (%s)%s" % (status, strerror(status), (": " + detail) if detail else ""))


if not os.path.exists(LIB_PATH):
    raise ImportError("libb200msm.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` -- "
                      "there is no CPU fallback" % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)
_vp, _u32, _u64, _i = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
lib.b200msm_create.argtypes = [ctypes.POINTER(_vp), _i]
lib.b200msm_create_multi.argtypes = [ctypes.POINTER(_vp), ctypes.POINTER(_i), _i]
lib.b200msm_device_count.argtypes = [_vp]
lib.b200msm_destroy.argtypes = [_vp]; lib.b200msm_destroy.restype = None
lib.b200msm_strerror.argtypes = [_i]; lib.b200msm_strerror.restype = ctypes.c_char_p
lib.b200msm_last_error.argtypes = [_vp]; lib.b200msm_last_error.restype = ctypes.c_char_p
lib.b200msm_version.restype = ctypes.c_char_p
lib.b200msm_set_stream.argtypes = [_vp, _vp]
lib.b200msm_synchronize.argtypes = [_vp]
lib.b200msm_g1_multiexp_affine.argtypes = [_vp, _i, _vp, _vp, _u32, _u64, _vp]
lib.b200msm_g1_multiexp_affine_chunk.argtypes = [_vp, _i, _vp, _vp, _u32, _u64, _u32, _u32, _vp]
lib.b200msm_g1_multiexp.argtypes = [_vp, _i, _vp, _vp, _u32, _u64, _vp]
lib.b200msm_g1_multiexp_chunk.argtypes = [_vp, _i, _vp, _vp, _u32, _u64, _u32, _u32, _vp]
lib.b200msm_upload_bases.argtypes = [_vp, _i, _vp, _u64, ctypes.POINTER(_u64)]
lib.b200msm_upload_bases_windowed.argtypes = [_vp, _i, _vp, _u64, _u32, _u32, ctypes.POINTER(_u64)]
lib.b200msm_free_bases.argtypes = [_vp, _u64]
lib.b200msm_g1_multiexp_resident.argtypes = [_vp, _u64, _vp, _u32, _u64, _vp, ctypes.POINTER(Stats)]
lib.b200msm_g1_multiexp_batch.argtypes = [_vp, _u64, _vp, _u32, _u64, _u32, _vp]
lib.b200msm_fr_fft.argtypes = [_vp, _i, _vp, _u32, _i, _vp]
lib.b200msm_fr_fft_last_phases.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_u32)]
lib.b200msm_g1_normalize.argtypes = [_vp, _i, _vp, _u64, _vp]
lib.b200msm_g1_sum.argtypes = [_vp, _i, _vp, _u64, _vp]
lib.b200msm_g1_generate_bases.argtypes = [_vp, _i, _u64, _u64, _u64, _vp]
lib.b200msm_fq_op.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _u64]
lib.b200msm_fq_batch_inverse.argtypes = [_vp, _i, _vp, _u64, _vp]
lib.b200msm_debug_schedule.argtypes = [_vp, _vp, _u32, _u64, _u32, ctypes.POINTER(_u32), _vp, _u64, _vp, _u64]
lib.b200msm_probe_imad.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
lib.b200msm_probe_imad32.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
lib.b200msm_probe_fqmul.argtypes = [_vp, _i, ctypes.POINTER(ctypes.c_double)]
if hasattr(lib, "b200msm_probe_dfma"):
    lib.b200msm_probe_dfma.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
    lib.b200msm_probe_dualpipe.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
lib.b200msm_set_option.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int64]
lib.b200msm_g1_batch_convert.argtypes = [_vp, _i, _i, _vp, _u64, _vp]
lib.b200msm_glv_decompose_scalars.argtypes = [_vp, _i, _vp, _u64, _vp, _vp]
lib.b200msm_g1_glv_preprocess.argtypes = [_vp, _i, _vp, _vp, _u64, _vp, _vp]
lib.b200msm_get_counter.argtypes = [_vp, ctypes.c_char_p, ctypes.POINTER(_u64)]
lib.b200msm_constants.argtypes = [_i, ctypes.POINTER(_u32), _vp, _vp, _vp, ctypes.POINTER(_u32)]


def strerror(status):
    return lib.b200msm_strerror(status).decode()


def constants(curve):
    """host-only: (n8, q, R mod q, R^2 mod q, np32) as the engine uses them"""
    n8 = _u32(); np32 = _u32()
    q = ctypes.create_string_buffer(48); one = ctypes.create_string_buffer(48); r2 = ctypes.create_string_buffer(48)
    rc = lib.b200msm_constants(curve, ctypes.byref(n8), q, one, r2, ctypes.byref(np32))
    if rc: raise B200MsmError(rc)
    k = n8.value
    return k, int.from_bytes(q.raw[:k], "little"), int.from_bytes(one.raw[:k], "little"), int.from_bytes(r2.raw[:k], "little"), np32.value
