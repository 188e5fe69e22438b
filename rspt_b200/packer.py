"""Host-side mirror of the reference's packer interface (lib_rspt/signal_packer.h:29-73) over the
C ABI in include/rspt_gpu.h.

`SignalPacker.new_xdelta_hzr / new_hzr / new_hadamard / new_dct` take the reference factories'
arguments; `compress(src) -> bytes` and `decompress(src) -> (bytes, consumed)` have the reference
methods' meaning for ONE frame in host memory (signal_packer.h:44,57).  `compress_batch` /
`decompress_batch` are the B200 path proper: many independent frames resident in HBM per call.
torch is used only to own device memory and to name the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import KINDS, Counters, RsptError, check


@dataclass
class CompressedBatch:
    """The concatenated frames of one batch plus the out-of-band facts a decoder needs
    (the reference stream carries neither a frame index nor the plane count)."""
    stream: torch.Tensor      # uint8, capacity-sized; valid bytes = offsets[-1]
    offsets: torch.Tensor     # uint64-as-int64 [n + 1], device
    frame_nb: torch.Tensor    # uint8 [n], device
    sidecar: torch.Tensor | None
    n_frames: int

    def total_bytes(self) -> int:
        return int(self.offsets[-1].item())

    def frame(self, i: int) -> bytes:
        off = self.offsets[i:i + 2].cpu().tolist()
        return bytes(self.stream[off[0]:off[1]].cpu().numpy())


class SignalPacker:
    """One reference packer instance (= one `i_signal_packer::new_*` call) on one GPU."""

    def __init__(self, kind: str, bytes_per_sample: int, nr_channels: int, nr_samples: int,
                 nr_bytes_to_encode: int = 3, device: int | None = None, max_batch_frames: int = 1):
        if kind not in KINDS:
            raise ValueError(f"unknown packer kind {kind!r}")
        if not torch.cuda.is_available():
            raise RsptError("no CUDA device: rspt_b200 has no CPU fallback")
        self.L = _lib.lib()
        self.kind, self.bps, self.ch, self.ns = kind, bytes_per_sample, nr_channels, nr_samples
        self.device = torch.cuda.current_device() if device is None else device
        self.max_batch = max_batch_frames
        self.stream_ptr = torch.cuda.current_stream(self.device).cuda_stream
        h = C.c_void_p()
        rc = self.L.rspt_gpu_create(KINDS[kind], bytes_per_sample, nr_channels, nr_samples, nr_bytes_to_encode,
                                    self.device, self.stream_ptr, max_batch_frames, C.byref(h))
        check(rc, None, "rspt_gpu_create")
        self.h = h
        self.frame_bytes = self.L.rspt_gpu_frame_bytes(h)
        self.header_bytes = self.L.rspt_gpu_header_bytes(h)
        self.max_compressed_size = self.L.rspt_gpu_max_compressed_size(h)

    # -- reference factories (signal_packer.h:59-69) ------------------------------------------
    @classmethod
    def new_xdelta_hzr(cls, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, nr_bytes_to_encode, **kw):
        return cls("xdelta_hzr", bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, nr_bytes_to_encode, **kw)

    @classmethod
    def new_hzr(cls, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, **kw):
        return cls("hzr", bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 4, **kw)

    @classmethod
    def new_hadamard(cls, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, **kw):
        return cls("hadamard", bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 3, **kw)

    @classmethod
    def new_dct(cls, bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, **kw):
        return cls("dct", bytes_per_channel, nr_of_channels, nr_of_samples_in_each_channel, 2, **kw)

    def close(self):
        if getattr(self, "h", None):
            self.L.rspt_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------------
    @property
    def nb(self) -> int:
        v = C.c_uint(0)
        check(self.L.rspt_gpu_nb(self.h, C.byref(v)), self.h, "rspt_gpu_nb")
        return v.value

    def counters(self) -> dict:
        c = Counters()
        check(self.L.rspt_gpu_get_counters(self.h, C.byref(c)), self.h, "rspt_gpu_get_counters")
        return {n: getattr(c, n) for n, _ in Counters._fields_}

    def sync(self):
        check(self.L.rspt_gpu_sync(self.h), self.h, "rspt_gpu_sync")

    def set_stage_timing(self, enable: bool = True):
        check(self.L.rspt_gpu_set_stage_timing(self.h, int(enable)), self.h, "set_stage_timing")

    def stage_times(self, reset: bool = True) -> dict:
        """{stage: (milliseconds, calls)} accumulated since the last reset (device time, CUDA events)."""
        ms = (C.c_double * len(_lib.STAGES))()
        calls = (C.c_uint64 * len(_lib.STAGES))()
        check(self.L.rspt_gpu_get_stage_times(self.h, ms, calls, int(reset)), self.h, "get_stage_times")
        return {n: (ms[i], calls[i]) for i, n in enumerate(_lib.STAGES)}

    # -- single frame, host buffers: the reference call shapes -----------------------------------
    def compress(self, src) -> bytes:
        s = np.ascontiguousarray(np.frombuffer(src, np.uint8) if not isinstance(src, np.ndarray) else src.view(np.uint8).reshape(-1))
        if s.size != self.frame_bytes:
            raise ValueError(f"expected {self.frame_bytes} bytes, got {s.size}")
        dst = np.empty(self.max_compressed_size, np.uint8)
        n = C.c_size_t(0)
        check(self.L.rspt_gpu_compress_host(self.h, s.ctypes.data, dst.ctypes.data, dst.size, C.byref(n)), self.h, "compress")
        return dst[: n.value].tobytes()

    def decompress(self, src):
        c = np.frombuffer(src, np.uint8) if not isinstance(src, np.ndarray) else src.view(np.uint8).reshape(-1)
        c = np.ascontiguousarray(np.concatenate([c, np.zeros(8, np.uint8)]))
        out = np.empty(self.frame_bytes, np.uint8)
        n = C.c_size_t(0)
        check(self.L.rspt_gpu_decompress_host(self.h, c.ctypes.data, C.byref(n), out.ctypes.data), self.h, "decompress")
        return out.tobytes(), n.value

    # -- batches, device buffers ------------------------------------------------------------
    def _dev(self):
        return torch.device("cuda", self.device)

    def _bind_stream(self):
        """Order the handle's work on torch's CURRENT stream (the handle was created on the stream that was
        current then).  The switch itself is ordered: the new stream waits for what the old one still holds."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream_ptr:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.ExternalStream(self.stream_ptr, device=self._dev()) if self.stream_ptr else torch.cuda.default_stream(self.device))
            cur.wait_event(ev)
            check(self.L.rspt_gpu_set_stream(self.h, cur.cuda_stream), self.h, "rspt_gpu_set_stream")
            self.stream_ptr = cur.cuda_stream

    def sidecar_used_bytes(self, n_frames: int, stream_bytes: int) -> int:
        """Bytes of the decode index that belong to a batch whose stream has `stream_bytes` bytes (a prefix of
        the sidecar buffer): what has to be kept next to the stream."""
        return int(self.L.rspt_gpu_sidecar_used_bytes(self.h, n_frames, stream_bytes))

    # -- multi-GPU placement (the path's only collective), through the C ABI -------------------------
    def place_offsets_async(self, comm, batch: "CompressedBatch", rank: int, world: int) -> None:
        """All-gather of the ranks' byte totals + rebase of this batch's offsets into the global stream, on the
        handle's side stream (ordered behind the compress that produced `batch`)."""
        check(self.L.rspt_gpu_place_offsets_async(self.h, comm, batch.offsets.data_ptr(), batch.n_frames, rank, world), self.h,
              "rspt_gpu_place_offsets_async")

    def place_join(self) -> None:
        check(self.L.rspt_gpu_place_join(self.h), self.h, "rspt_gpu_place_join")

    def set_dct_exact(self, exact: bool) -> None:
        check(self.L.rspt_gpu_set_dct_exact(self.h, int(exact)), self.h, "rspt_gpu_set_dct_exact")

    def alloc_output(self, n_frames: int, sidecar: bool = True) -> CompressedBatch:
        dev = self._dev()
        stream = torch.empty(n_frames * self.max_compressed_size, dtype=torch.uint8, device=dev)
        offsets = torch.empty(n_frames + 1, dtype=torch.int64, device=dev)
        frame_nb = torch.empty(n_frames, dtype=torch.uint8, device=dev)
        sc = torch.empty(self.L.rspt_gpu_sidecar_bytes(self.h, n_frames), dtype=torch.uint8, device=dev) if sidecar else None
        return CompressedBatch(stream, offsets, frame_nb, sc, n_frames)

    def compress_batch(self, frames: torch.Tensor, out: CompressedBatch | None = None, sidecar: bool = True) -> CompressedBatch:
        """frames: uint8 CUDA tensor of n * frame_bytes bytes.  Asynchronous on torch's current stream (the
        handle is re-bound to it when it changed since the last call)."""
        if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous()):
            raise ValueError("frames must be a contiguous uint8 CUDA tensor")
        n = frames.numel() // self.frame_bytes
        if n * self.frame_bytes != frames.numel():
            raise ValueError("frames is not a whole number of frames")
        if out is None:
            out = self.alloc_output(n, sidecar)
        if out.offsets.numel() < n + 1 or out.frame_nb.numel() < n or out.stream.numel() < n * self.max_compressed_size:
            raise ValueError("the output batch is too small for these frames")
        if out.sidecar is not None and out.sidecar.numel() < self.L.rspt_gpu_sidecar_bytes(self.h, n):
            raise ValueError("the output batch's sidecar is too small for these frames")
        self._bind_stream()
        sc = out.sidecar.data_ptr() if out.sidecar is not None else None
        rc = self.L.rspt_gpu_compress_batch(self.h, frames.data_ptr(), n, out.stream.data_ptr(), out.stream.numel(),
                                            out.offsets.data_ptr(), out.frame_nb.data_ptr(), sc)
        check(rc, self.h, "rspt_gpu_compress_batch")
        out.n_frames = n
        return out

    def decompress_batch(self, batch: CompressedBatch, out: torch.Tensor | None = None, status: torch.Tensor | None = None,
                         use_sidecar: bool = True) -> torch.Tensor:
        n = batch.n_frames
        if out is None:
            out = torch.empty(n * self.frame_bytes, dtype=torch.uint8, device=self._dev())
        sc = batch.sidecar.data_ptr() if (use_sidecar and batch.sidecar is not None) else None
        nbp = batch.frame_nb.data_ptr() if batch.frame_nb is not None else None
        self._bind_stream()
        rc = self.L.rspt_gpu_decompress_batch(self.h, batch.stream.data_ptr(), batch.offsets.data_ptr(), n, nbp, sc,
                                              out.data_ptr(), status.data_ptr() if status is not None else None)
        check(rc, self.h, "rspt_gpu_decompress_batch")
        return out

    def prefilter_iir(self, frames: torch.Tensor, n, d, init_nr_samples: int) -> torch.Tensor:
        """In place on device frames: the IIR pre-filter step of rspt_test.cpp:116-136 (i_filter::new_iir)."""
        na, da = np.ascontiguousarray(n, np.float64), np.ascontiguousarray(d, np.float64)
        dp = C.POINTER(C.c_double)
        self._bind_stream()
        rc = self.L.rspt_gpu_prefilter_iir(self.h, frames.data_ptr(), frames.numel() // self.frame_bytes,
                                           na.ctypes.data_as(dp), da.ctypes.data_as(dp), len(na), init_nr_samples)
        check(rc, self.h, "rspt_gpu_prefilter_iir")
        return frames

    def prefilter_fir(self, frames: torch.Tensor, kernel) -> torch.Tensor:
        """In place on device frames: the same step with i_filter::new_fir(kernel)."""
        ka = np.ascontiguousarray(kernel, np.float64)
        self._bind_stream()
        rc = self.L.rspt_gpu_prefilter_fir(self.h, frames.data_ptr(), frames.numel() // self.frame_bytes,
                                           ka.ctypes.data_as(C.POINTER(C.c_double)), len(ka))
        check(rc, self.h, "rspt_gpu_prefilter_fir")
        return frames

    def build_index(self, batch: CompressedBatch, status: torch.Tensor | None = None) -> CompressedBatch:
        """Give a batch that came without a decode index (CPU-written frames) one, on the device."""
        n = batch.n_frames
        sc = torch.empty(self.L.rspt_gpu_sidecar_bytes(self.h, n), dtype=torch.uint8, device=self._dev())
        nbp = batch.frame_nb.data_ptr() if batch.frame_nb is not None else None
        self._bind_stream()
        rc = self.L.rspt_gpu_build_index(self.h, batch.stream.data_ptr(), batch.offsets.data_ptr(), n, nbp, sc.data_ptr(),
                                         status.data_ptr() if status is not None else None)
        check(rc, self.h, "rspt_gpu_build_index")
        return CompressedBatch(batch.stream, batch.offsets, batch.frame_nb, sc, n)

    def verify_batch(self, batch: CompressedBatch, status: torch.Tensor | None = None) -> torch.Tensor:
        """hzr_verify (hzr_decode.c:569-624) over every block of every frame: int32 status per frame,
        0 = ok, -4 = malformed framing, -6 = CRC-32C mismatch.  Asynchronous on the current stream."""
        n = batch.n_frames
        if status is None:
            status = torch.zeros(n, dtype=torch.int32, device=self._dev())
        nbp = batch.frame_nb.data_ptr() if batch.frame_nb is not None else None
        self._bind_stream()
        rc = self.L.rspt_gpu_verify_batch(self.h, batch.stream.data_ptr(), batch.offsets.data_ptr(), n, nbp, status.data_ptr())
        check(rc, self.h, "rspt_gpu_verify_batch")
        return status

    def decompress_stream(self, stream: bytes, offsets, frame_nb=None) -> tuple[np.ndarray, np.ndarray]:
        """Decode frames produced elsewhere (e.g. by the CPU reference): no decode index."""
        dev = self._dev()
        off = torch.tensor(list(offsets), dtype=torch.int64, device=dev)
        n = off.numel() - 1
        buf = torch.from_numpy(np.frombuffer(stream, np.uint8).copy()).to(dev)
        pad = torch.zeros(16, dtype=torch.uint8, device=dev)
        buf = torch.cat([buf, pad])
        nb = torch.tensor(list(frame_nb), dtype=torch.uint8, device=dev) if frame_nb is not None else None
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        b = CompressedBatch(buf, off, nb, None, n)
        out = self.decompress_batch(b, status=status, use_sidecar=False)
        torch.cuda.synchronize(dev)
        return out.cpu().numpy().reshape(n, self.frame_bytes), status.cpu().numpy()

    # -- batches, host buffers (end-to-end leg) ------------------------------------------------
    def compress_batch_host(self, frames: np.ndarray, dst: np.ndarray, offsets: np.ndarray) -> int:
        n = frames.size // self.frame_bytes
        rc = self.L.rspt_gpu_compress_batch_host(self.h, frames.ctypes.data, n, dst.ctypes.data, dst.size, offsets.ctypes.data)
        check(rc, self.h, "rspt_gpu_compress_batch_host")
        return int(offsets[n])

    def decompress_batch_host(self, src: np.ndarray, offsets: np.ndarray, out: np.ndarray) -> None:
        n = offsets.size - 1
        rc = self.L.rspt_gpu_decompress_batch_host(self.h, src.ctypes.data, offsets.ctypes.data, n, out.ctypes.data)
        check(rc, self.h, "rspt_gpu_decompress_batch_host")

    # -- stage-level (parity tests) -----------------------------------------------------------
    def debug_planes(self, frames: torch.Tensor):
        n = frames.numel() // self.frame_bytes
        unsigned = C.c_uint(0)
        nb_alloc = (self.max_compressed_size - 1 - self.header_bytes) // (4 + 4 + 7 * ((self.ch * self.ns + 65535) // 65536) + self.ch * self.ns)
        planes = torch.empty(n * nb_alloc * self.ch * self.ns, dtype=torch.uint8, device=self._dev())
        hdr = torch.zeros(max(1, n * self.header_bytes), dtype=torch.uint8, device=self._dev())
        check(self.L.rspt_gpu_debug_planes(self.h, frames.data_ptr(), n, planes.data_ptr(), hdr.data_ptr()), self.h, "debug_planes")
        torch.cuda.synchronize(self._dev())
        return planes.cpu().numpy().reshape(n, nb_alloc, self.ch * self.ns), hdr.cpu().numpy()[: n * self.header_bytes].reshape(n, -1)

    def debug_hzr_tables(self, block: np.ndarray):
        dev = self._dev()
        b = torch.from_numpy(np.ascontiguousarray(block, np.uint8)).to(dev)
        pad = torch.zeros(64, dtype=torch.uint8, device=dev)
        b = torch.cat([b, pad])
        hist = torch.zeros(264, dtype=torch.int32, device=dev)
        codes = torch.zeros(264, dtype=torch.int32, device=dev)
        info = torch.zeros(4, dtype=torch.int32, device=dev)
        check(self.L.rspt_gpu_debug_hzr_tables(self.h, b.data_ptr(), block.size, hist.data_ptr(), codes.data_ptr(), info.data_ptr()),
              self.h, "debug_hzr_tables")
        torch.cuda.synchronize(dev)
        return (hist.cpu().numpy().view(np.uint32)[:261], codes.cpu().numpy().view(np.uint32)[:261],
                info.cpu().numpy().view(np.uint32))


class IngestRing:
    """Pinned-host packet ring in front of a packer (the reference's io_buffer, ring_buffers.h:150-201):
    the producer writes frames into `next_packet()`, the consumer calls `drain()`."""

    def __init__(self, packer: SignalPacker, nr_max_packets: int):
        self.p = packer
        self.L = packer.L
        self.h = C.c_void_p()
        check(self.L.rspt_gpu_ingest_create(packer.h, nr_max_packets, C.byref(self.h)), packer.h, "rspt_gpu_ingest_create")

    def close(self):
        if self.h:
            self.L.rspt_gpu_ingest_destroy(self.h)
            self.h = C.c_void_p()

    def next_packet(self):
        """numpy view of the next packet to fill (frame_bytes bytes), or None when the ring is full."""
        addr = self.L.rspt_gpu_ingest_next_address_to_fill(self.h)
        if not addr:
            return None
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(self.p.frame_bytes,))

    def drain(self, dst: np.ndarray, offsets: np.ndarray, max_frames: int | None = None, flush: bool = False) -> int:
        """Compress the filled packets into dst; offsets[0..n] are filled in.  Returns n."""
        n = C.c_size_t(offsets.size - 1 if max_frames is None else min(max_frames, offsets.size - 1))
        rc = self.L.rspt_gpu_ingest_drain(self.h, int(flush), dst.ctypes.data, dst.size, offsets.ctypes.data, C.byref(n))
        check(rc, self.p.h, "rspt_gpu_ingest_drain")
        return int(n.value)


def crc32c(data: np.ndarray) -> int:
    t = torch.from_numpy(np.array(data, dtype=np.uint8, copy=True)).cuda() if data.size else torch.zeros(1, dtype=torch.uint8, device="cuda")
    out = C.c_uint32(0)
    check(_lib.lib().rspt_gpu_crc32c(t.data_ptr(), data.size, C.byref(out), torch.cuda.current_stream().cuda_stream), None, "crc32c")
    return out.value


def synth_ecg(first_frame: int, n_frames: int, bps: int, ch: int, ns: int, seed: int = 42, amplitude: int = 20000,
              sigma: int = 3, out: torch.Tensor | None = None, device=None) -> torch.Tensor:
    """Synthetic ECG-like frames generated on the device (include/rspt_synth.h)."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    if out is None:
        out = torch.empty(n_frames * bps * ch * ns, dtype=torch.uint8, device=dev)
    check(_lib.lib().rspt_gpu_synth_ecg(out.data_ptr(), first_frame, n_frames, bps, ch, ns, seed, amplitude, sigma,
                                        torch.cuda.current_stream(dev).cuda_stream), None, "synth_ecg")
    return out


def prdn_terms(orig: torch.Tensor, dec: torch.Tensor, n_frames: int, bps: int, ch: int, ns: int) -> tuple[float, float]:
    """Numerator and denominator sums of PRDN (rspt_test.cpp:98-111) over the given frames, so that a
    caller can accumulate them over batches."""
    out = (C.c_double * 2)()
    check(_lib.lib().rspt_gpu_prdn_terms(orig.data_ptr(), dec.data_ptr(), n_frames, bps, ch, ns, out,
                                         torch.cuda.current_stream().cuda_stream), None, "prdn")
    return float(out[0]), float(out[1])


def prdn(orig: torch.Tensor, dec: torch.Tensor, n_frames: int, bps: int, ch: int, ns: int) -> float:
    num, den = prdn_terms(orig, dec, n_frames, bps, ch, ns)
    return float(np.sqrt(num / den) * 100.0) if den > 0 else 0.0
