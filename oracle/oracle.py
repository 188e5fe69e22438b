"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle.

Two libraries with the same Python face:

* ``liboracle.so``      -- this repo's plain-C restatement (``rspt_oracle.c``), kind ``"port"``.
* ``_ref/libref.so``    -- the UNMODIFIED reference compiled from ``/root/reference`` by
  ``oracle/Makefile`` (kind ``"reference"``); present in the build container and, because
  ``oracle/_ref/`` is git-ignored but not gpurun-ignored, on the GPU box as a prebuilt file.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package ``rspt_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = {"xdelta_hzr": 0, "hzr": 1, "hadamard": 2, "dct": 3}
NSYM = 261

_sz = C.c_size_t
_vp = C.c_void_p
_u8p = C.c_void_p


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/libref.so (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/lib_rspt"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


_oracle = None
_ref = None


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_new.restype = _vp
        L.oracle_new.argtypes = [C.c_int, _sz, _sz, _sz, _sz]
        L.oracle_delete.argtypes = [_vp]
        L.oracle_compress.restype = C.c_int
        L.oracle_compress.argtypes = [_vp, _u8p, _u8p, _sz, C.POINTER(_sz)]
        L.oracle_decompress.restype = C.c_int
        L.oracle_decompress.argtypes = [_vp, _u8p, C.POINTER(_sz), _u8p]
        L.oracle_nb.restype = C.c_uint
        L.oracle_nb.argtypes = [_vp]
        L.oracle_escalations.restype = C.c_uint
        L.oracle_escalations.argtypes = [_vp]
        L.oracle_max_compressed_size.restype = _sz
        L.oracle_max_compressed_size.argtypes = [_vp]
        L.oracle_header_bytes.restype = _sz
        L.oracle_header_bytes.argtypes = [_vp]
        L.oracle_transform.restype = C.c_int
        L.oracle_transform.argtypes = [_vp, _u8p, _vp, _u8p]
        L.oracle_inverse.restype = C.c_int
        L.oracle_inverse.argtypes = [_vp, _vp, _u8p, _u8p]
        L.oracle_compress_many.restype = _sz
        L.oracle_compress_many.argtypes = [_vp, _u8p, _sz, _sz, _u8p, _sz, _vp]
        L.oracle_decompress_many.restype = _sz
        L.oracle_decompress_many.argtypes = [_vp, _u8p, _sz, _sz, _u8p, _sz]
        L.oracle_crc32c.restype = C.c_uint32
        L.oracle_crc32c.argtypes = [_u8p, _sz]
        L.oracle_hzr_max_compressed_size.restype = _sz
        L.oracle_hzr_max_compressed_size.argtypes = [_sz]
        L.oracle_hzr_encode.restype = C.c_int
        L.oracle_hzr_encode.argtypes = [_u8p, _sz, _u8p, _sz, C.POINTER(_sz)]
        L.oracle_hzr_decode.restype = C.c_int
        L.oracle_hzr_decode.argtypes = [_u8p, _sz, _u8p, _sz]
        L.oracle_hzr_verify.restype = C.c_int
        L.oracle_hzr_verify.argtypes = [_u8p, _sz, C.POINTER(_sz)]
        L.oracle_hzr_histogram.argtypes = [_u8p, _sz, _vp]
        L.oracle_hzr_build_codes.restype = C.c_int
        L.oracle_hzr_build_codes.argtypes = [_vp, _vp, _vp, _vp, C.POINTER(C.c_uint32)]
        L.oracle_hzr_block_plan.restype = C.c_int
        L.oracle_hzr_block_plan.argtypes = [_u8p, _sz, C.POINTER(C.c_uint32)]
        L.oracle_average_32.restype = C.c_int32
        L.oracle_average_32.argtypes = [_vp, _sz]
        L.oracle_fwht.argtypes = [C.c_int, _vp, _vp]
        L.oracle_prdn.restype = C.c_double
        L.oracle_prdn.argtypes = [_u8p, _u8p, _sz, _sz, _sz]
        _dp = C.POINTER(C.c_double)
        L.oracle_prefilter_iir.restype = C.c_int
        L.oracle_prefilter_iir.argtypes = [_u8p, _sz, _sz, _sz, _dp, _dp, C.c_int, C.c_int]
        L.oracle_prefilter_fir.restype = C.c_int
        L.oracle_prefilter_fir.argtypes = [_u8p, _sz, _sz, _sz, _dp, C.c_int]
        L.oracle_synth_ecg.argtypes = [_u8p, C.c_uint64, _sz, C.c_int, C.c_int, C.c_int,
                                       C.c_uint64, C.c_int32, C.c_int32]
        _oracle = L
    return _oracle


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libref.so"))
        L.ref_new.restype = _vp
        L.ref_new.argtypes = [C.c_int, _sz, _sz, _sz, _sz]
        L.ref_delete.argtypes = [C.c_int, _vp]
        L.ref_compress.restype = _sz
        L.ref_compress.argtypes = [_vp, _u8p, _u8p, _sz]
        L.ref_decompress.restype = _sz
        L.ref_decompress.argtypes = [_vp, _u8p, _u8p]
        L.ref_compress_many.restype = _sz
        L.ref_compress_many.argtypes = [_vp, _u8p, _sz, _sz, _u8p, _sz, _vp]
        L.ref_decompress_many.restype = _sz
        L.ref_decompress_many.argtypes = [_vp, _u8p, _sz, _sz, _u8p, _sz]
        L.ref_hzr_max_compressed_size.restype = _sz
        L.ref_hzr_max_compressed_size.argtypes = [_sz]
        L.ref_hzr_encode.restype = C.c_int
        L.ref_hzr_encode.argtypes = [_u8p, _sz, _u8p, _sz, C.POINTER(_sz)]
        L.ref_hzr_decode.restype = C.c_int
        L.ref_hzr_decode.argtypes = [_u8p, _sz, _u8p, _sz]
        L.ref_hzr_verify.restype = C.c_int
        L.ref_hzr_verify.argtypes = [_u8p, _sz, C.POINTER(_sz)]
        _dp = C.POINTER(C.c_double)
        if hasattr(L, "ref_prefilter_iir"):
            L.ref_prefilter_iir.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int]
            L.ref_prefilter_fir.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        L.ref_crc32c.restype = C.c_uint32
        L.ref_crc32c.argtypes = [_u8p, _sz]
        _ref = L
    return _ref


def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a.view(np.uint8).reshape(-1))


class _PackerBase:
    kind: str
    bps: int
    ch: int
    ns: int

    @property
    def frame_bytes(self) -> int:
        return self.bps * self.ch * self.ns

    def roundtrip(self, src):
        comp = self.compress(src)
        dec, used = self.decompress(comp)
        return comp, dec, used


class OraclePacker(_PackerBase):
    """The C restatement behind the i_signal_packer face (signal_packer.h:29-73)."""

    impl = "port"

    def __init__(self, kind: str, bps: int, ch: int, ns: int, nb: int = 3):
        self.L = oracle_lib()
        self.kind, self.bps, self.ch, self.ns = kind, bps, ch, ns
        self.h = self.L.oracle_new(KINDS[kind], bps, ch, ns, nb)
        if not self.h:
            raise ValueError("bad packer arguments")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_delete(self.h)
            self.h = None

    @property
    def nb(self) -> int:
        return self.L.oracle_nb(self.h)

    @property
    def escalations(self) -> int:
        return self.L.oracle_escalations(self.h)

    @property
    def max_compressed_size(self) -> int:
        return self.L.oracle_max_compressed_size(self.h)

    @property
    def header_bytes(self) -> int:
        return self.L.oracle_header_bytes(self.h)

    def compress(self, src) -> bytes:
        s = _as_u8(src)
        assert s.size == self.frame_bytes
        # capacity for the worst case after any escalation to 4 planes
        cap = 1 + self.header_bytes + 4 * (4 + self.L.oracle_hzr_max_compressed_size(self.ch * self.ns))
        dst = np.empty(cap, np.uint8)
        n = _sz(0)
        rc = self.L.oracle_compress(self.h, _ptr(s), _ptr(dst), cap, C.byref(n))
        if rc:
            raise RuntimeError(f"oracle_compress rc={rc}")
        return dst[: n.value].tobytes()

    def decompress(self, comp):
        c = _as_u8(comp)
        c = np.concatenate([c, np.zeros(16, np.uint8)])
        out = np.empty(self.frame_bytes, np.uint8)
        n = _sz(0)
        self.L.oracle_decompress(self.h, _ptr(c), C.byref(n), _ptr(out))
        return out.tobytes(), n.value

    def transform(self, src):
        """(words int32[ch*ns] handed to the plane split, header bytes)."""
        s = _as_u8(src)
        w = np.empty(self.ch * self.ns, np.int32)
        hdr = np.zeros(max(1, self.header_bytes), np.uint8)
        rc = self.L.oracle_transform(self.h, _ptr(s), _ptr(w), _ptr(hdr))
        if rc:
            raise RuntimeError("oracle_transform failed")
        return w, hdr[: self.header_bytes]

    def inverse(self, words, header):
        w = np.ascontiguousarray(words, np.int32)
        hdr = np.ascontiguousarray(np.concatenate([_as_u8(header), np.zeros(1, np.uint8)]))
        out = np.empty(self.frame_bytes, np.uint8)
        self.L.oracle_inverse(self.h, _ptr(w), _ptr(hdr), _ptr(out))
        return out

    def compress_many(self, frames: np.ndarray):
        """frames uint8 [n, frame_bytes] -> (dst [n, stride], sizes uint32[n])."""
        f = np.ascontiguousarray(frames, np.uint8).reshape(-1, self.frame_bytes)
        n = f.shape[0]
        stride = 1 + self.header_bytes + 4 * (4 + self.L.oracle_hzr_max_compressed_size(self.ch * self.ns))
        dst = np.zeros((n, stride), np.uint8)
        sizes = np.zeros(n, np.uint32)
        self.L.oracle_compress_many(self.h, _ptr(f), self.frame_bytes, n, _ptr(dst), stride, _ptr(sizes))
        return dst, sizes

    def decompress_many(self, dst: np.ndarray, n: int) -> np.ndarray:
        out = np.empty((n, self.frame_bytes), np.uint8)
        self.L.oracle_decompress_many(self.h, _ptr(dst), dst.shape[1], n, _ptr(out), self.frame_bytes)
        return out


class RefPacker(_PackerBase):
    """The unmodified reference (oracle/_ref/libref.so)."""

    impl = "reference"

    def __init__(self, kind: str, bps: int, ch: int, ns: int, nb: int = 3):
        self.L = ref_lib()
        self.kind, self.bps, self.ch, self.ns = kind, bps, ch, ns
        self.k = KINDS[kind]
        self.h = self.L.ref_new(self.k, bps, ch, ns, nb)
        hb = 3 * ch if kind in ("hadamard", "dct") else 0
        self.stride = 1 + hb + 4 * (4 + self.L.ref_hzr_max_compressed_size(ch * ns)) + 64

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_delete(self.k, self.h)
            self.h = None

    def compress(self, src) -> bytes:
        s = _as_u8(src)
        assert s.size == self.frame_bytes
        # the reference over-reads <= 3 bytes past the last sample for bps < 4 (utils.cpp:141)
        s = np.concatenate([s, np.zeros(8, np.uint8)])
        dst = np.zeros(self.stride, np.uint8)
        n = self.L.ref_compress(self.h, _ptr(s), _ptr(dst), self.stride)
        return dst[:n].tobytes()

    def decompress(self, comp):
        c = np.concatenate([_as_u8(comp), np.zeros(16, np.uint8)])
        out = np.empty(self.frame_bytes + 8, np.uint8)
        n = self.L.ref_decompress(self.h, _ptr(c), _ptr(out))
        return out[: self.frame_bytes].tobytes(), n

    def compress_many(self, frames: np.ndarray):
        f = np.ascontiguousarray(frames, np.uint8).reshape(-1, self.frame_bytes)
        n = f.shape[0]
        f = np.concatenate([f.reshape(-1), np.zeros(8, np.uint8)])
        dst = np.zeros((n, self.stride), np.uint8)
        sizes = np.zeros(n, np.uint32)
        self.L.ref_compress_many(self.h, _ptr(f), self.frame_bytes, n, _ptr(dst), self.stride, _ptr(sizes))
        return dst, sizes

    def decompress_many(self, dst: np.ndarray, n: int) -> np.ndarray:
        out = np.empty(n * self.frame_bytes + 8, np.uint8)
        self.L.ref_decompress_many(self.h, _ptr(dst), dst.shape[1], n, _ptr(out), self.frame_bytes)
        return out[: n * self.frame_bytes].reshape(n, self.frame_bytes)


# ---- hzr stage helpers -----------------------------------------------------------------------
def crc32c(data, impl: str = "port") -> int:
    a = _as_u8(data)
    a = a if a.size else np.zeros(1, np.uint8)
    n = len(_as_u8(data))
    return (oracle_lib().oracle_crc32c if impl == "port" else ref_lib().ref_crc32c)(_ptr(a), n)


def hzr_encode(data, impl: str = "port") -> bytes:
    a = _as_u8(data)
    L = oracle_lib() if impl == "port" else ref_lib()
    cap = (L.oracle_hzr_max_compressed_size if impl == "port" else L.ref_hzr_max_compressed_size)(a.size)
    out = np.zeros(cap + 16, np.uint8)
    n = _sz(0)
    src = a if a.size else np.zeros(1, np.uint8)
    if impl == "port":
        rc = L.oracle_hzr_encode(_ptr(src), a.size, _ptr(out), cap, C.byref(n))
        assert rc == 0
    else:
        rc = L.ref_hzr_encode(_ptr(src), a.size, _ptr(out), cap, C.byref(n))
        assert rc == 1
    return out[: n.value].tobytes()


def hzr_decode(comp, out_size: int, impl: str = "port"):
    c = np.concatenate([_as_u8(comp), np.zeros(16, np.uint8)])
    n = len(_as_u8(comp))
    out = np.zeros(max(out_size, 1), np.uint8)
    if impl == "port":
        rc = oracle_lib().oracle_hzr_decode(_ptr(c), n, _ptr(out), out_size)
        ok = rc == 0
    else:
        rc = ref_lib().ref_hzr_decode(_ptr(c), n, _ptr(out), out_size)
        ok = rc == 1
    return out[:out_size].tobytes(), ok


def prefilter_iir(frame, bps: int, ch: int, ns: int, n, d, init_nr_samples: int, impl: str = "port") -> np.ndarray:
    """The pre-filter step of rspt_test.cpp:116-136 (IIR, iir_filter.cpp) on one native frame; returns the filtered frame."""
    out = np.array(_as_u8(frame), dtype=np.uint8, copy=True)
    na, da = np.ascontiguousarray(n, np.float64), np.ascontiguousarray(d, np.float64)
    dp = C.POINTER(C.c_double)
    if impl == "port":
        rc = oracle_lib().oracle_prefilter_iir(_ptr(out), bps, ch, ns, na.ctypes.data_as(dp), da.ctypes.data_as(dp), len(na), init_nr_samples)
        assert rc == 0
    else:
        ref_lib().ref_prefilter_iir(_ptr(out), bps, ch, ns, na.ctypes.data_as(dp), da.ctypes.data_as(dp), len(na), init_nr_samples)
    return out


def prefilter_fir(frame, bps: int, ch: int, ns: int, kernel, impl: str = "port") -> np.ndarray:
    """The same step with a FIR filter (fir_filter.cpp)."""
    out = np.array(_as_u8(frame), dtype=np.uint8, copy=True)
    ka = np.ascontiguousarray(kernel, np.float64)
    dp = C.POINTER(C.c_double)
    if impl == "port":
        rc = oracle_lib().oracle_prefilter_fir(_ptr(out), bps, ch, ns, ka.ctypes.data_as(dp), len(ka))
        assert rc == 0
    else:
        ref_lib().ref_prefilter_fir(_ptr(out), bps, ch, ns, ka.ctypes.data_as(dp), len(ka))
    return out


def hzr_verify(comp, impl: str = "port") -> bool:
    """hzr_verify (hzr_decode.c:569-624): header walk + CRC-32C of every block."""
    c = np.concatenate([_as_u8(comp), np.zeros(16, np.uint8)])
    n = len(_as_u8(comp))
    dec = _sz(0)
    if impl == "port":
        return oracle_lib().oracle_hzr_verify(_ptr(c), n, C.byref(dec)) == 0
    return ref_lib().ref_hzr_verify(_ptr(c), n, C.byref(dec)) == 1


def hzr_histogram(block) -> np.ndarray:
    a = _as_u8(block)
    h = np.zeros(NSYM, np.uint32)
    oracle_lib().oracle_hzr_histogram(_ptr(a), a.size, _ptr(h))
    return h


def hzr_build_codes(hist: np.ndarray):
    """-> (n_used, code uint32[261], len uint8[261], tree bytes, tree_nbits)."""
    h = np.ascontiguousarray(hist, np.uint32)
    code = np.zeros(NSYM, np.uint32)
    ln = np.zeros(NSYM, np.uint8)
    tree = np.zeros(360, np.uint8)
    nb = C.c_uint32(0)
    n = oracle_lib().oracle_hzr_build_codes(_ptr(h), _ptr(code), _ptr(ln), _ptr(tree), C.byref(nb))
    return n, code, ln, tree, nb.value


def hzr_block_plan(block):
    a = _as_u8(block)
    pl = C.c_uint32(0)
    mode = oracle_lib().oracle_hzr_block_plan(_ptr(a), a.size, C.byref(pl))
    return mode, pl.value


def prdn(orig, dec, bps: int, ch: int, ns: int) -> float:
    a, b = _as_u8(orig), _as_u8(dec)
    return oracle_lib().oracle_prdn(_ptr(a), _ptr(b), bps, ch, ns)


def synth_ecg(first_frame: int, n: int, bps: int, ch: int, ns: int, seed: int = 42,
              amplitude: int = 20000, sigma: int = 3) -> np.ndarray:
    """CPU side of the workload generator (include/rspt_synth.h) -> uint8 [n, bps*ch*ns]."""
    out = np.empty((n, bps * ch * ns), np.uint8)
    oracle_lib().oracle_synth_ecg(_ptr(out), first_frame, n, bps, ch, ns, seed, amplitude, sigma)
    return out


def make_packer(kind: str, bps: int, ch: int, ns: int, nb: int = 3, impl: str = "port"):
    return (OraclePacker if impl == "port" else RefPacker)(kind, bps, ch, ns, nb)
