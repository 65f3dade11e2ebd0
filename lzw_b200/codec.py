"""Codec: a libslzw context plus the batched entry points (host and device memory)."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import Batch, Params


class SlzwError(RuntimeError):
    """A launch-level failure (SLZW_RC_*), not a per-stream codec error."""


class Codec:
    """One slzw_ctx bound to one GPU.  Not shared between threads (one per thread)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        rc = self._lib.slzw_create(device, C.byref(h))
        if rc != _lib.RC_OK:
            raise SlzwError(
                f"slzw_create(device={device}) failed with rc={rc}: "
                + ("no usable sm_100 CUDA device -- lzw_b200 has no CPU fallback"
                   if rc == _lib.RC_NO_DEVICE else "CUDA error"))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.slzw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != _lib.RC_OK:
            raise SlzwError(f"{what} failed (rc={rc}): {self._lib.slzw_last_error(self._h).decode()}")

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.slzw_kernel_launches(self._h))

    def last_deferred(self, cap: int = 1024) -> np.ndarray:
        """Ids of the streams the last decode call handed to the exact-emulation kernel."""
        ids = np.zeros(max(cap, 1), dtype=np.uint32)
        n = int(self._lib.slzw_last_deferred(self._h, ids.ctypes.data, cap))
        return ids[: min(n, cap)] if n <= cap else np.concatenate([ids[:cap], np.full(n - cap, 0xFFFFFFFF, np.uint32)])

    def last_encode_shares(self) -> np.ndarray:
        """Input bytes of the last encode call by kind of warp (tensor-memory warp, shared-memory
        warp, shared-memory lanes, global-memory lanes)."""
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.slzw_last_encode_shares(self._h, out.ctypes.data), "slzw_last_encode_shares")
        return out

    def encode_bound(self, params: Params, n: int) -> int:
        return int(self._lib.slzw_encode_bound(C.byref(params), n))

    # ---- device-resident batches: raw pointers (ints), asynchronous on `stream` ----------
    def _dev(self, fn, params, n, in_ptr, in_off_ptr, out_ptr, out_off_ptr, out_len_ptr,
             status_ptr, detail_ptr, code_size_ptr, stream, what):
        b = Batch(in_ptr, in_off_ptr, out_ptr, out_off_ptr, out_len_ptr, status_ptr, detail_ptr,
                  code_size_ptr or None, n)
        self._check(fn(self._h, C.byref(params), C.byref(b), stream or None), what)

    def encode_batch_device(self, params, n, in_ptr, in_off_ptr, out_ptr, out_off_ptr,
                            out_len_ptr, status_ptr, detail_ptr, code_size_ptr=0, stream=0):
        self._dev(self._lib.slzw_encode_batch_device, params, n, in_ptr, in_off_ptr, out_ptr,
                  out_off_ptr, out_len_ptr, status_ptr, detail_ptr, code_size_ptr, stream,
                  "slzw_encode_batch_device")

    def decode_batch_device(self, params, n, in_ptr, in_off_ptr, out_ptr, out_off_ptr,
                            out_len_ptr, status_ptr, detail_ptr, code_size_ptr=0, stream=0):
        self._dev(self._lib.slzw_decode_batch_device, params, n, in_ptr, in_off_ptr, out_ptr,
                  out_off_ptr, out_len_ptr, status_ptr, detail_ptr, code_size_ptr, stream,
                  "slzw_decode_batch_device")

    def decoded_sizes_batch_device(self, params, n, in_ptr, in_off_ptr, out_len_ptr, status_ptr,
                                   detail_ptr, code_size_ptr=0, stream=0):
        self._dev(self._lib.slzw_decoded_sizes_batch_device, params, n, in_ptr, in_off_ptr, 0, 0,
                  out_len_ptr, status_ptr, detail_ptr, code_size_ptr, stream,
                  "slzw_decoded_sizes_batch_device")

    def compact_device(self, src_ptr, src_off_ptr, len_ptr, n, dst_ptr, dst_off_ptr, align=1,
                       stream=0):
        self._check(self._lib.slzw_compact_device(self._h, src_ptr, src_off_ptr, len_ptr, n, align,
                                                  dst_ptr, dst_off_ptr, stream or None),
                    "slzw_compact_device")

    # ---- TIFF Predictor = 2 (horizontal differencing, 8-bit samples) -------------------------
    DIFFERENCE, ACCUMULATE = 0, 1

    def tiff_predictor_device(self, direction, data_ptr, off_ptr, n, row_bytes, samples_per_pixel,
                              len_ptr=0, stream=0):
        """In place on device memory: stream i is data[off[i] .. off[i] + len[i]) (len_ptr == 0: up
        to off[i + 1]), rows of row_bytes bytes."""
        self._check(self._lib.slzw_tiff_predictor_device(self._h, direction, data_ptr, off_ptr,
                                                         len_ptr or None, n, row_bytes,
                                                         samples_per_pixel, stream or None),
                    "slzw_tiff_predictor_device")

    def set_tiff_predictor(self, row_bytes=0, samples_per_pixel=0):
        """Host batch calls difference before encoding / accumulate after decoding, on the device.
        (0, 0) switches the predictor off."""
        self._check(self._lib.slzw_set_tiff_predictor(self._h, row_bytes, samples_per_pixel),
                    "slzw_set_tiff_predictor")

    # ---- host-resident batches: numpy arrays -----------------------------------------------
    def _host(self, fn, params, in_buf, in_off, out_off, code_size, what, out=None):
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        n = in_off.size - 1
        if out is None:
            out = np.zeros(max(int(out_off[-1]), 1), dtype=np.uint8)
        elif out.dtype != np.uint8 or out.size < int(out_off[-1]) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint8 array of at least out_off[-1] bytes")
        out_len = np.zeros(max(n, 1), dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        detail = np.zeros(max(n, 1), dtype=np.uint32)
        cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
        b = Batch(in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data, out_off.ctypes.data,
                  out_len.ctypes.data, status.ctypes.data, detail.ctypes.data,
                  None if cs is None else cs.ctypes.data, n)
        self._check(fn(self._h, C.byref(params), C.byref(b)), what)
        return out, out_len[:n], status[:n], detail[:n]

    def encode_batch(self, params, in_buf, in_off, out_off=None, code_size=None):
        """Encodes streams in_buf[in_off[i]:in_off[i+1]].  out_off defaults to worst-case slots.
        Returns (out, out_off, out_len, status, detail)."""
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        if out_off is None:
            lens = np.diff(in_off)
            # slzw_encode_bound, vectorised
            slots = ((lens + 3 + lens // 3838 + 1) * 12 + 7) // 8
            slots = (slots + 15) // 16 * 16
            out_off = np.concatenate([[0], np.cumsum(slots)]).astype(np.uint64)
        out, out_len, status, detail = self._host(self._lib.slzw_encode_batch_host, params, in_buf,
                                                  in_off, out_off, code_size,
                                                  "slzw_encode_batch_host")
        return out, np.asarray(out_off, dtype=np.uint64), out_len, status, detail

    def encode_batch_dense(self, params, in_buf, in_off, code_size=None, align=1, out=None):
        """Encodes into a dense buffer (strips back to back).  `out` may be a preallocated
        (pinned) uint8 array; by default it is sized for the worst case.
        Returns (dense, dense_off, status, detail)."""
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        n = in_off.size - 1
        if out is None:
            lens = np.diff(in_off)
            worst = int((((lens + 3 + lens // 3838 + 1) * 12 + 7) // 8 + align).sum())
            out = np.empty(max(worst, 1), dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        detail = np.zeros(max(n, 1), dtype=np.uint32)
        cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
        needed = C.c_uint64(0)
        rc = self._lib.slzw_encode_batch_host_dense(
            self._h, C.byref(params), in_buf.ctypes.data, in_off.ctypes.data, n,
            None if cs is None else cs.ctypes.data, align, out.ctypes.data, out.size,
            out_off.ctypes.data, status.ctypes.data, detail.ctypes.data, C.byref(needed))
        self._check(rc, "slzw_encode_batch_host_dense")
        return out[: int(out_off[-1])], out_off, status[:n], detail[:n]

    def encode_batch_dense_begin(self, params, in_buf, in_off, code_size=None, align=1):
        """First phase of the two-phase dense encode: returns (dense_off, status, detail, total);
        the encoded bytes stay on the device until encode_batch_dense_finish."""
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        n = in_off.size - 1
        out_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        detail = np.zeros(max(n, 1), dtype=np.uint32)
        cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
        total = C.c_uint64(0)
        self._check(self._lib.slzw_encode_batch_host_dense_begin(
            self._h, C.byref(params), in_buf.ctypes.data, in_off.ctypes.data, n,
            None if cs is None else cs.ctypes.data, align, out_off.ctypes.data, status.ctypes.data,
            detail.ctypes.data, C.byref(total)), "slzw_encode_batch_host_dense_begin")
        return out_off, status[:n], detail[:n], int(total.value)

    def encode_batch_dense_finish(self, out):
        self._check(self._lib.slzw_encode_batch_host_dense_finish(self._h, out.ctypes.data, out.size),
                    "slzw_encode_batch_host_dense_finish")
        return out

    def decode_batch(self, params, in_buf, in_off, out_off, code_size=None, out=None):
        """Decodes streams into capacity slots out_off.  `out` may be a preallocated (pinned) uint8
        array.  Returns (out, out_len, status, detail)."""
        return self._host(self._lib.slzw_decode_batch_host, params, in_buf, in_off, out_off,
                          code_size, "slzw_decode_batch_host", out=out)

    # ---- single stream ------------------------------------------------------------------------
    def encode(self, params: Params, data, cap: int | None = None):
        """Returns (status, detail, bytes produced)."""
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) \
            else np.ascontiguousarray(data, dtype=np.uint8)
        cap = self.encode_bound(params, a.size) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        out_len, detail = C.c_uint64(0), C.c_uint32(0)
        st = self._lib.slzw_encode(self._h, C.byref(params), a.ctypes.data if a.size else None,
                                   a.size, out.ctypes.data, cap, C.byref(out_len), C.byref(detail))
        if st < 0:
            self._check(st, "slzw_encode")
        return st, detail.value, out[: out_len.value].tobytes()

    def decoded_size(self, params: Params, data):
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) \
            else np.ascontiguousarray(data, dtype=np.uint8)
        out_len, detail = C.c_uint64(0), C.c_uint32(0)
        st = self._lib.slzw_decode(self._h, C.byref(params), a.ctypes.data if a.size else None,
                                   a.size, None, 0, C.byref(out_len), C.byref(detail))
        if st < 0:
            self._check(st, "slzw_decode(size)")
        return st, detail.value, int(out_len.value)

    def decode(self, params: Params, data, cap: int | None = None):
        """Returns (status, detail, bytes produced).  cap=None sizes the output first, which is
        what decode_to_vec's growing Vec amounts to (decoder.rs:163-172)."""
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) \
            else np.ascontiguousarray(data, dtype=np.uint8)
        if cap is None:
            _, _, cap = self.decoded_size(params, a)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        out_len, detail = C.c_uint64(0), C.c_uint32(0)
        st = self._lib.slzw_decode(self._h, C.byref(params), a.ctypes.data if a.size else None,
                                   a.size, out.ctypes.data, cap, C.byref(out_len), C.byref(detail))
        if st < 0:
            self._check(st, "slzw_decode")
        return st, detail.value, out[: out_len.value].tobytes()

    def status_message(self, is_decoder: bool, status: int, detail: int, code_size: int = 0) -> str:
        return status_message(is_decoder, status, detail, code_size)


class MultiCodec:
    """One host batch over several GPUs of one box (slzw_multi_*): the batch is sharded by stream,
    one context and one worker thread per device, no exchange between devices; results are those
    of the single-device calls."""

    def __init__(self, devices=None, n_devices: int = 0):
        self._lib = _lib.lib()
        h = C.c_void_p()
        if devices is None:
            rc = self._lib.slzw_multi_create(None, n_devices, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._lib.slzw_multi_create(arr, len(devices), C.byref(h))
        if rc != _lib.RC_OK:
            raise SlzwError(f"slzw_multi_create failed with rc={rc}"
                            + (": no usable sm_100 CUDA device -- lzw_b200 has no CPU fallback"
                               if rc == _lib.RC_NO_DEVICE else ""))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.slzw_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return int(self._lib.slzw_multi_device_count(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.slzw_multi_kernel_launches(self._h))

    def _check(self, rc, what):
        if rc != _lib.RC_OK:
            raise SlzwError(f"{what} failed (rc={rc}): {self._lib.slzw_multi_last_error(self._h).decode()}")

    def _host(self, fn, params, in_buf, in_off, out_off, code_size, what, out=None):
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        n = in_off.size - 1
        if out is None:
            out = np.zeros(max(int(out_off[-1]), 1), dtype=np.uint8)
        out_len = np.zeros(max(n, 1), dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        detail = np.zeros(max(n, 1), dtype=np.uint32)
        cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
        b = Batch(in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data, out_off.ctypes.data,
                  out_len.ctypes.data, status.ctypes.data, detail.ctypes.data,
                  None if cs is None else cs.ctypes.data, n)
        self._check(fn(self._h, C.byref(params), C.byref(b)), what)
        return out, out_len[:n], status[:n], detail[:n]

    def encode_batch(self, params, in_buf, in_off, out_off, code_size=None, out=None):
        """Returns (out, out_len, status, detail) for capacity slots out_off."""
        return self._host(self._lib.slzw_multi_encode_batch_host, params, in_buf, in_off, out_off, code_size,
                          "slzw_multi_encode_batch_host", out=out)

    def decode_batch(self, params, in_buf, in_off, out_off, code_size=None, out=None):
        return self._host(self._lib.slzw_multi_decode_batch_host, params, in_buf, in_off, out_off, code_size,
                          "slzw_multi_decode_batch_host", out=out)

    def encode_batch_dense(self, params, in_buf, in_off, code_size=None, align=1, out=None):
        """Returns (dense, dense_off, status, detail); the shards of the devices lie back to back."""
        in_buf = np.ascontiguousarray(in_buf, dtype=np.uint8)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        n = in_off.size - 1
        if out is None:
            lens = np.diff(in_off)
            worst = int((((lens + 3 + lens // 3838 + 1) * 12 + 7) // 8 + align).sum())
            out = np.empty(max(worst, 1), dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint32)
        detail = np.zeros(max(n, 1), dtype=np.uint32)
        cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
        needed = C.c_uint64(0)
        rc = self._lib.slzw_multi_encode_batch_host_dense(
            self._h, C.byref(params), in_buf.ctypes.data, in_off.ctypes.data, n,
            None if cs is None else cs.ctypes.data, align, out.ctypes.data, out.size,
            out_off.ctypes.data, status.ctypes.data, detail.ctypes.data, C.byref(needed))
        self._check(rc, "slzw_multi_encode_batch_host_dense")
        return out[: int(out_off[-1])], out_off, status[:n], detail[:n]


def partition_streams(off, parts: int) -> np.ndarray:
    """slzw_partition_streams: contiguous stream ranges balanced by bytes (parts + 1 bounds)."""
    off = np.ascontiguousarray(off, dtype=np.uint64)
    bounds = np.zeros(parts + 1, dtype=np.uint64)
    _lib.lib().slzw_partition_streams(off.ctypes.data, off.size - 1, parts, bounds.ctypes.data)
    return bounds


class PinnedBuffer:
    """Page-locked host memory from slzw_host_alloc as a uint8 numpy array (`.array`).  The host
    entry points copy to and from pinned memory asynchronously, and the encoder reads pinned input
    in place."""

    def __init__(self, nbytes: int):
        self._lib = _lib.lib()
        self._p = self._lib.slzw_host_alloc(max(int(nbytes), 1))
        if not self._p:
            raise MemoryError(f"slzw_host_alloc({nbytes}) failed")
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(int(nbytes), 1),))[:nbytes]

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            self._lib.slzw_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def status_message(is_decoder: bool, status: int, detail: int, code_size: int = 0) -> str:
    """The reference's Display text for a result (encoder.rs:31-44, decoder.rs:27-42)."""
    buf = C.create_string_buffer(160)
    _lib.lib().slzw_status_message(int(is_decoder), status, detail, code_size & 0xFF, buf, len(buf))
    return buf.value.decode()


_default = threading.local()


def default_codec(device: int = 0) -> Codec:
    """A per-thread Codec on `device`, created on first use (the reference's functions are
    stateless, so the facade needs an implicit context)."""
    key = f"codec{device}"
    c = getattr(_default, key, None)
    if c is None:
        c = Codec(device)
        setattr(_default, key, c)
    return c
