"""Engine: thin Python owner of one b200msm_ctx (one GPU, one stream).

Inputs may be `bytes`/`bytearray`/numpy arrays (host) or torch CUDA tensors / raw integer device
pointers (device); they are passed to the C ABI as plain pointers -- no torch types cross the boundary.
"""
import ctypes
from . import _lib
from ._lib import lib, B200MsmError, Stats, N8


def _ptr(x):
    """(pointer, keepalive) for bytes / bytearray / memoryview / numpy / torch tensor / int"""
    if x is None: return None, None
    if isinstance(x, int): return ctypes.c_void_p(x), None
    if isinstance(x, (bytes, bytearray)):
        buf = (ctypes.c_char * len(x)).from_buffer_copy(x) if isinstance(x, bytes) else (ctypes.c_char * len(x)).from_buffer(x)
        return ctypes.cast(buf, ctypes.c_void_p), buf
    if hasattr(x, "data_ptr"):  # torch tensor (host or device)
        assert x.is_contiguous()
        return ctypes.c_void_p(x.data_ptr()), x
    if hasattr(x, "ctypes"):    # numpy
        return ctypes.c_void_p(x.ctypes.data), x
    raise TypeError("unsupported buffer type %r" % type(x))


class Engine:
    def __init__(self, device=-1, devices=None):
        """device: one GPU ordinal (-1 = current).  devices: a list of ordinals -> ONE multi-device context (b200msm_create_multi):
        every MSM call on it shards the points over those GPUs behind the same entry points."""
        self._ctx = ctypes.c_void_p()
        if devices is not None and len(devices) > 1:
            ids = (ctypes.c_int * len(devices))(*devices)
            rc = lib.b200msm_create_multi(ctypes.byref(self._ctx), ids, len(devices))
        else:
            rc = lib.b200msm_create(ctypes.byref(self._ctx), device if devices is None else devices[0])
        if rc: raise B200MsmError(rc, "b200msm_create failed: is a B200 visible? (no CPU fallback)")

    @property
    def device_count(self): return lib.b200msm_device_count(self._ctx)

    def close(self):
        if self._ctx: lib.b200msm_destroy(self._ctx); self._ctx = ctypes.c_void_p()

    def __del__(self):
        try: self.close()
        except Exception: pass

    def _ck(self, rc):
        if rc: raise B200MsmError(rc, lib.b200msm_last_error(self._ctx).decode())

    # ---- configuration
    def set_stream(self, cuda_stream_handle): self._ck(lib.b200msm_set_stream(self._ctx, ctypes.c_void_p(cuda_stream_handle)))
    def synchronize(self): self._ck(lib.b200msm_synchronize(self._ctx))
    def set_option(self, key, value): self._ck(lib.b200msm_set_option(self._ctx, key.encode(), int(value)))

    # ---- the reference entry points
    def multiexp_affine(self, curve, bases, scalars, scalar_size, n, out=None):
        """g1m_multiexpAffine: returns 3*n8 bytes (Jacobian Montgomery) unless `out` (host/device buffer) is given."""
        pb, kb = _ptr(bases); ps, ks = _ptr(scalars)
        if out is None:
            o = ctypes.create_string_buffer(3 * N8[curve])
            self._ck(lib.b200msm_g1_multiexp_affine(self._ctx, curve, pb, ps, scalar_size, n, o)); return o.raw
        po, ko = _ptr(out)
        self._ck(lib.b200msm_g1_multiexp_affine(self._ctx, curve, pb, ps, scalar_size, n, po)); return out

    def multiexp_affine_chunk(self, curve, bases, scalars, scalar_size, n, start_bit, chunk_bits):
        pb, kb = _ptr(bases); ps, ks = _ptr(scalars)
        o = ctypes.create_string_buffer(3 * N8[curve])
        self._ck(lib.b200msm_g1_multiexp_affine_chunk(self._ctx, curve, pb, ps, scalar_size, n, start_bit, chunk_bits, o)); return o.raw

    def multiexp_jacobian(self, curve, bases_jac, scalars, scalar_size, n, chunk=None):
        """g1m_multiexp / g1m_multiexp_chunk: Jacobian bases (3*n8 bytes each); chunk = (start_bit, chunk_bits) or None"""
        pb, kb = _ptr(bases_jac); ps, ks = _ptr(scalars)
        o = ctypes.create_string_buffer(3 * N8[curve])
        if chunk is None: self._ck(lib.b200msm_g1_multiexp(self._ctx, curve, pb, ps, scalar_size, n, o))
        else: self._ck(lib.b200msm_g1_multiexp_chunk(self._ctx, curve, pb, ps, scalar_size, n, chunk[0], chunk[1], o))
        return o.raw

    def upload_bases(self, curve, bases, n):
        pb, kb = _ptr(bases); h = ctypes.c_uint64()
        self._ck(lib.b200msm_upload_bases(self._ctx, curve, pb, n, ctypes.byref(h))); return h.value

    def upload_bases_windowed(self, curve, bases, n, scalar_size=32, window_bits=0):
        """resident bases + precomputed window table (rows 2^(c*w) * P_i); use the handle with multiexp_resident"""
        pb, kb = _ptr(bases); h = ctypes.c_uint64()
        self._ck(lib.b200msm_upload_bases_windowed(self._ctx, curve, pb, n, scalar_size, window_bits, ctypes.byref(h))); return h.value

    def free_bases(self, handle): self._ck(lib.b200msm_free_bases(self._ctx, handle))

    def multiexp_resident(self, handle, scalars, scalar_size, n, curve, out=None, want_stats=False):
        ps, ks = _ptr(scalars)
        st = Stats() if want_stats else None
        if out is None:
            o = ctypes.create_string_buffer(3 * N8[curve]); po = o
        else:
            po, ko = _ptr(out)
        self._ck(lib.b200msm_g1_multiexp_resident(self._ctx, handle, ps, scalar_size, n, po, ctypes.byref(st) if st is not None else None))
        res = o.raw if out is None else out
        return (res, st.as_dict()) if want_stats else res

    def multiexp_batch(self, handle, scalars, scalar_size, n, count, curve, out=None):
        """count independent MSMs over the same resident bases; scalars = count blocks of n scalars; -> count * 3*n8 bytes"""
        ps, ks = _ptr(scalars)
        if out is None:
            o = ctypes.create_string_buffer(3 * N8[curve] * count)
            self._ck(lib.b200msm_g1_multiexp_batch(self._ctx, handle, ps, scalar_size, n, count, o)); return o.raw
        po, ko = _ptr(out)
        self._ck(lib.b200msm_g1_multiexp_batch(self._ctx, handle, ps, scalar_size, n, count, po)); return out

    def fr_fft(self, curve, data, log2n, inverse=False, out=None):
        """frm_fft / frm_ifft over 2^log2n Montgomery Fr elements (32 bytes each) -> bytes, or into `out` (host/device buffer)"""
        pi, ki = _ptr(data)
        if out is None:
            o = ctypes.create_string_buffer(32 << log2n)
            self._ck(lib.b200msm_fr_fft(self._ctx, curve, pi, log2n, 1 if inverse else 0, o)); return o.raw
        po, ko = _ptr(out)
        self._ck(lib.b200msm_fr_fft(self._ctx, curve, pi, log2n, 1 if inverse else 0, po)); return out

    def fr_fft_last_phases(self):
        ms = (ctypes.c_float * 4)(); ps = (ctypes.c_uint32 * 2)()
        self._ck(lib.b200msm_fr_fft_last_phases(self._ctx, ms, ps))
        return {"ms_bitrev": ms[0], "ms_tile": ms[1], "ms_passes": ms[2], "ms_final": ms[3], "radix4_passes": ps[0], "radix2_passes": ps[1]}

    def normalize(self, curve, jac, count=1):
        """g1m_normalize + fromMontgomery: canonical x||y bytes (plain LE ints; infinity = zeros)"""
        pj, kj = _ptr(jac); o = ctypes.create_string_buffer(2 * N8[curve] * count)
        self._ck(lib.b200msm_g1_normalize(self._ctx, curve, pj, count, o)); return o.raw

    def sum_points(self, curve, jac_points, count):
        pj, kj = _ptr(jac_points); o = ctypes.create_string_buffer(3 * N8[curve])
        self._ck(lib.b200msm_g1_sum(self._ctx, curve, pj, count, o)); return o.raw

    def generate_bases(self, curve, seed, first, n, device_out):
        po, ko = _ptr(device_out)
        self._ck(lib.b200msm_g1_generate_bases(self._ctx, curve, seed, first, n, po))

    CONVERT = {"LEMtoU": 0, "LEMtoC": 1, "UtoLEM": 2, "CtoLEM": 3, "toAffine": 4, "toJacobian": 5}

    def batch_convert(self, curve, op, data, n):
        """g1m_batchLEMtoU / LEMtoC / UtoLEM / CtoLEM / batchToAffine / batchToJacobian on n points -> bytes"""
        k = self.CONVERT[op] if isinstance(op, str) else op
        n8 = N8[curve]; out_sz = [2 * n8, n8, 2 * n8, 2 * n8, 2 * n8, 3 * n8][k]
        pi, ki = _ptr(data); o = ctypes.create_string_buffer(max(1, out_sz * n))
        self._ck(lib.b200msm_g1_batch_convert(self._ctx, curve, k, pi, n, o)); return o.raw[:out_sz * n]

    def glv_decompose_scalars(self, curve, scalars, n):
        """g1m_glv_decomposeScalar over n 32-byte scalars -> (n x 64 bytes [|k1| .. |k2| ..], [sign, ...])"""
        ps, ks = _ptr(scalars); o = ctypes.create_string_buffer(max(1, 64 * n)); sg = (ctypes.c_uint32 * max(1, n))()
        self._ck(lib.b200msm_glv_decompose_scalars(self._ctx, curve, ps, n, o, sg)); return o.raw[:64 * n], list(sg)[:n]

    def glv_preprocess(self, curve, points, scalars, n, out_points=None, out_scalars=None):
        """g1m_glv_preprocessEndomorphism: -> (2n points, 2n 32-byte scalars); bytes unless output buffers are given"""
        pp, kp = _ptr(points); ps, ks = _ptr(scalars)
        if out_points is None:
            op_ = ctypes.create_string_buffer(max(1, 192 * n)); os_ = ctypes.create_string_buffer(max(1, 64 * n))
            self._ck(lib.b200msm_g1_glv_preprocess(self._ctx, curve, pp, ps, n, op_, os_)); return op_.raw[:192 * n], os_.raw[:64 * n]
        po, ko = _ptr(out_points); pso, kso = _ptr(out_scalars)
        self._ck(lib.b200msm_g1_glv_preprocess(self._ctx, curve, pp, ps, n, po, pso)); return out_points, out_scalars

    def fq_op(self, curve, op, a, b=None):
        n = len(a) // N8[curve]
        pa, ka = _ptr(a); pb_, kb = _ptr(b)
        o = ctypes.create_string_buffer(max(1, len(a)))
        self._ck(lib.b200msm_fq_op(self._ctx, curve, op, pa, pb_, o, n)); return o.raw[:len(a)]

    def fq_batch_inverse(self, curve, a):
        """f1m_batchInverse on Montgomery elements (zeros stay zero)"""
        n = len(a) // N8[curve]
        pa, ka = _ptr(a); o = ctypes.create_string_buffer(max(1, len(a)))
        self._ck(lib.b200msm_fq_batch_inverse(self._ctx, curve, pa, n, o)); return o.raw[:len(a)]

    def debug_schedule(self, scalars, scalar_size, n, window_bits=0):
        """the engine's digit / sort phase alone: (plan dict, offsets list, sorted entries list)"""
        plan = (ctypes.c_uint32 * 6)(); ps, ks = _ptr(scalars)
        self._ck(lib.b200msm_debug_schedule(self._ctx, ps, scalar_size, n, window_bits, plan, None, 0, None, 0))
        Wd, W, B, c0, rem, nbits = list(plan)
        offs = (ctypes.c_uint32 * (W * B + 1))(); srt = (ctypes.c_uint32 * max(1, n * W))()
        self._ck(lib.b200msm_debug_schedule(self._ctx, ps, scalar_size, n, window_bits, plan, offs, W * B + 1, srt, n * W))
        offs = list(offs)
        return {"Wd": Wd, "W": W, "B": B, "c0": c0, "rem": rem, "nbits": nbits}, offs, list(srt)[:offs[-1]]

    def counter(self, key):
        v = ctypes.c_uint64(); self._ck(lib.b200msm_get_counter(self._ctx, key.encode(), ctypes.byref(v))); return v.value

    def probe_imad(self):
        v = ctypes.c_double(); self._ck(lib.b200msm_probe_imad(self._ctx, ctypes.byref(v))); return v.value

    def probe_imad32(self):
        v = ctypes.c_double(); self._ck(lib.b200msm_probe_imad32(self._ctx, ctypes.byref(v))); return v.value

    def probe_dfma(self):
        v = ctypes.c_double(); self._ck(lib.b200msm_probe_dfma(self._ctx, ctypes.byref(v))); return v.value

    def probe_dualpipe(self):
        v = (ctypes.c_double * 5)(); self._ck(lib.b200msm_probe_dualpipe(self._ctx, v))
        return {"ms_imad_only": v[0], "ms_fp64_only": v[1], "ms_both": v[2], "mults_per_thread": int(v[3]), "ms_warp_specialised_half_each": v[4]}

    def probe_fqmul(self, curve):
        v = ctypes.c_double(); self._ck(lib.b200msm_probe_fqmul(self._ctx, curve, ctypes.byref(v))); return v.value
