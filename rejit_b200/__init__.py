"""rejit_b200 — Python mirror of the rejit interface over the C ABI.

The product is the shared library `librejit_b200.so` (host front end in C++,
matching on hand-written sm_100a kernels; see include/rejit.h for the C++
interface and include/rejit_b200.h for the C boundary).  This module is the thin
ctypes binding the tests and bench.py use: same operation names and argument
meaning as `rejit::Regej` (/root/reference/include/rejit.h:105-138).

There is no CPU fallback here: if the library is missing, or no CUDA device is
present, the matching calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RJ_LIB") or os.path.join(_HERE, "librejit_b200.so")     # RJ_LIB: tuning builds (scripts/ab_build.sh)

_lib = None


class RejitError(RuntimeError):
    pass


class ParserError(RejitError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("scan_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("launches", ctypes.c_uint32), ("reruns", ctypes.c_uint32),
                ("candidates", ctypes.c_uint64), ("matches", ctypes.c_uint64),
                ("strategy", ctypes.c_int32), ("large_path", ctypes.c_int32)]


class Carry(ctypes.Structure):
    _fields_ = [("cur", ctypes.c_uint64), ("tail", ctypes.c_uint64)]


# every symbol include/rejit_b200.h declares
EXPORTED = [
    "rejit_b200_parse", "rejit_b200_ir_free", "rejit_b200_ir_dump", "rejit_b200_compile",
    "rejit_b200_program_free", "rejit_b200_program_describe", "rejit_b200_program_is_shardable", "rejit_b200_match_all",
    "rejit_b200_match_all_alloc", "rejit_b200_match_first", "rejit_b200_match_full",
    "rejit_b200_match_anywhere", "rejit_b200_match_all_multi_gpu", "rejit_b200_device_count",
    "rejit_b200_device_alloc", "rejit_b200_device_free", "rejit_b200_pinned_alloc",
    "rejit_b200_pinned_free", "rejit_b200_copy_to_device", "rejit_b200_copy_from_device",
    "rejit_b200_flush_l2", "rejit_b200_match_all_device", "rejit_b200_match_all_device_slab", "rejit_b200_free",
    "rejit_b200_text_upload", "rejit_b200_text_from_device", "rejit_b200_text_free", "rejit_b200_match_all_text",
    "rejit_b200_set_create", "rejit_b200_set_free", "rejit_b200_set_describe", "rejit_b200_set_kmer_tables",
    "rejit_b200_match_all_set_text", "rejit_b200_match_all_set_device", "rejit_b200_match_all_set_device_slab",
    "rejit_b200_replace_all", "rejit_b200_replace_all_text", "rejit_b200_replace_all_set_text",
    "rejit_b200_stitch_open", "rejit_b200_stitch_connect", "rejit_b200_stitch_close", "rejit_b200_stitch_exchange",
    "rejit_b200_match_all_set_device_stitched", "rejit_b200_text_length", "rejit_b200_text_device_ptr", "rejit_b200_text_download",
]


def lib():
    """Loads librejit_b200.so (built by `python -m rejit_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RejitError("librejit_b200.so is not built: run `python -m rejit_b200.build` "
                         "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
    u64p = ctypes.POINTER(ctypes.c_uint64)
    L.rejit_b200_parse.argtypes = [cp, sz, ctypes.c_int, ctypes.POINTER(vp), cp, sz]
    L.rejit_b200_parse.restype = ctypes.c_int
    L.rejit_b200_ir_free.argtypes = [vp]
    L.rejit_b200_ir_dump.argtypes = [vp, cp, sz]
    L.rejit_b200_ir_dump.restype = sz
    L.rejit_b200_compile.argtypes = [vp, cp, sz]
    L.rejit_b200_compile.restype = vp
    L.rejit_b200_program_free.argtypes = [vp]
    L.rejit_b200_program_describe.argtypes = [vp]
    L.rejit_b200_program_describe.restype = cp
    L.rejit_b200_program_is_shardable.argtypes = [vp]
    L.rejit_b200_program_is_shardable.restype = ctypes.c_int
    L.rejit_b200_match_all.argtypes = [vp, cp, sz, u64p, sz, cp, sz]
    L.rejit_b200_match_all.restype = ctypes.c_int64
    L.rejit_b200_match_all_alloc.argtypes = [vp, vp, sz, ctypes.POINTER(u64p), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_alloc.restype = ctypes.c_int64
    L.rejit_b200_match_first.argtypes = [vp, cp, sz, u64p, cp, sz]
    L.rejit_b200_match_full.argtypes = [vp, cp, sz, cp, sz]
    L.rejit_b200_match_anywhere.argtypes = [vp, cp, sz, cp, sz]
    L.rejit_b200_match_all_multi_gpu.argtypes = [vp, vp, sz, ctypes.c_int, ctypes.POINTER(u64p),
                                                 ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_multi_gpu.restype = ctypes.c_int64
    L.rejit_b200_device_count.restype = ctypes.c_int
    L.rejit_b200_device_alloc.argtypes = [ctypes.c_int, sz]
    L.rejit_b200_device_alloc.restype = vp
    L.rejit_b200_device_free.argtypes = [ctypes.c_int, vp]
    L.rejit_b200_pinned_alloc.argtypes = [sz]
    L.rejit_b200_pinned_alloc.restype = vp
    L.rejit_b200_pinned_free.argtypes = [vp]
    L.rejit_b200_copy_to_device.argtypes = [ctypes.c_int, vp, vp, sz]
    L.rejit_b200_copy_from_device.argtypes = [ctypes.c_int, vp, vp, sz]
    L.rejit_b200_flush_l2.argtypes = [ctypes.c_int]
    L.rejit_b200_match_all_device.argtypes = [vp, ctypes.c_int, vp, sz, vp, sz, ctypes.POINTER(Carry),
                                              ctypes.POINTER(Carry), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_device.restype = ctypes.c_int64
    L.rejit_b200_match_all_device_slab.argtypes = [vp, ctypes.c_int, vp, sz, ctypes.c_uint64, ctypes.c_uint64,
                                                   ctypes.c_uint64, vp, sz, ctypes.POINTER(Carry),
                                                   ctypes.POINTER(Carry), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_device_slab.restype = ctypes.c_int64
    L.rejit_b200_free.argtypes = [vp]
    L.rejit_b200_text_upload.argtypes = [ctypes.c_int, vp, sz, cp, sz]
    L.rejit_b200_text_upload.restype = vp
    L.rejit_b200_text_from_device.argtypes = [ctypes.c_int, vp, sz, cp, sz]
    L.rejit_b200_text_from_device.restype = vp
    L.rejit_b200_text_free.argtypes = [vp]
    L.rejit_b200_match_all_text.argtypes = [vp, vp, ctypes.POINTER(u64p), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_text.restype = ctypes.c_int64
    L.rejit_b200_set_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.rejit_b200_set_create.restype = vp
    L.rejit_b200_set_free.argtypes = [vp]
    L.rejit_b200_set_describe.argtypes = [vp]
    L.rejit_b200_set_kmer_tables.argtypes = [vp, vp, vp, vp]
    L.rejit_b200_set_kmer_tables.restype = ctypes.c_int
    L.rejit_b200_set_describe.restype = cp
    L.rejit_b200_match_all_set_text.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(u64p),
                                                ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_set_device.argtypes = [vp, ctypes.c_int, vp, sz, ctypes.POINTER(ctypes.c_int64),
                                                  ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_match_all_set_device_slab.argtypes = [vp, ctypes.c_int, vp, sz, ctypes.c_uint64, ctypes.c_uint64,
                                                       ctypes.c_uint64, ctypes.POINTER(Carry), ctypes.POINTER(Carry),
                                                       ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_replace_all.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(vp), ctypes.POINTER(sz),
                                         ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_replace_all.restype = ctypes.c_int64
    L.rejit_b200_replace_all_text.argtypes = [vp, vp, vp, sz, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(Stats),
                                              cp, sz]
    L.rejit_b200_replace_all_text.restype = vp
    L.rejit_b200_replace_all_set_text.argtypes = [ctypes.POINTER(vp), ctypes.c_int, vp, ctypes.POINTER(cp), ctypes.POINTER(sz),
                                                  ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_replace_all_set_text.restype = vp
    L.rejit_b200_stitch_open.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, cp, sz]
    L.rejit_b200_stitch_connect.argtypes = [ctypes.c_int, vp, vp, cp, sz]
    L.rejit_b200_stitch_close.argtypes = [ctypes.c_int]
    L.rejit_b200_stitch_exchange.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Carry), ctypes.c_uint64,
                                             ctypes.POINTER(Carry), ctypes.POINTER(ctypes.c_uint32), cp, sz]
    L.rejit_b200_match_all_set_device_stitched.argtypes = [vp, ctypes.c_int, vp, sz, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                                           ctypes.POINTER(Carry), ctypes.POINTER(Carry), ctypes.POINTER(ctypes.c_uint32),
                                                           ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(Stats), cp, sz]
    L.rejit_b200_text_length.argtypes = [vp]
    L.rejit_b200_text_length.restype = sz
    L.rejit_b200_text_device_ptr.argtypes = [vp]
    L.rejit_b200_text_device_ptr.restype = vp
    L.rejit_b200_text_download.argtypes = [vp, vp, sz, cp, sz]
    _lib = L
    return L


def device_count() -> int:
    return lib().rejit_b200_device_count()


def _as_bytes(x) -> bytes:
    return x.encode("latin-1") if isinstance(x, str) else bytes(x)


class DeviceText:
    """Text resident in HBM (16-byte aligned, padded allocation)."""

    def __init__(self, data=None, device: int = 0, nbytes: Optional[int] = None, borrowed_ptr: Optional[int] = None):
        L = lib()
        self.device = device
        if borrowed_ptr is not None:
            # a device buffer owned by the caller (16-byte aligned, readable 32 bytes past nbytes: rejit_b200.h)
            self.nbytes, self.ptr, self._borrowed = int(nbytes), ctypes.c_void_p(borrowed_ptr), True
            return
        self._borrowed = False
        self.nbytes = len(data) if data is not None else int(nbytes)
        self.ptr = L.rejit_b200_device_alloc(device, max(self.nbytes, 1))
        if not self.ptr:
            raise RejitError("device allocation failed (is a CUDA device present?)")
        if data is not None and self.nbytes:
            self.upload(data)

    def upload(self, data, pinned_ptr: Optional[int] = None):
        L = lib()
        if pinned_ptr is not None:
            src = pinned_ptr
        elif hasattr(data, "ctypes"):              # numpy array
            import numpy as np
            keep = np.ascontiguousarray(data, dtype=np.uint8)
            src = ctypes.c_void_p(keep.ctypes.data)
        else:
            keep = bytes(data)
            src = ctypes.cast(ctypes.c_char_p(keep), ctypes.c_void_p)
        if L.rejit_b200_copy_to_device(self.device, self.ptr, src, self.nbytes) != 0:
            raise RejitError("host to device copy failed")

    def free(self):
        if self.ptr:
            if not getattr(self, "_borrowed", False):
                lib().rejit_b200_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Text:
    """An uploaded text (rejit_b200_text): searched and rewritten on the device;
    `replace_all` returns a new Text, so substitutions chain without host copies."""

    def __init__(self, data=None, device: int = 0, _handle=None, device_ptr=None, nbytes: int = 0):
        L = lib()
        if _handle is not None:
            self._h = ctypes.c_void_p(_handle)
            return
        if device_ptr is not None:                 # bytes already on the device: copied into a text of its own
            err = ctypes.create_string_buffer(512)
            h = L.rejit_b200_text_from_device(device, ctypes.c_void_p(device_ptr), nbytes, err, len(err))
            if not h:
                raise RejitError(err.value.decode("latin-1"))
            self._h = ctypes.c_void_p(h)
            return
        if hasattr(data, "ctypes"):
            import numpy as np
            keep = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, n = ctypes.c_void_p(keep.ctypes.data), keep.size
        else:
            keep = _as_bytes(data)
            ptr, n = ctypes.cast(ctypes.c_char_p(keep), ctypes.c_void_p), len(keep)
        err = ctypes.create_string_buffer(512)
        h = L.rejit_b200_text_upload(device, ptr, n, err, len(err))
        if not h:
            raise RejitError(err.value.decode("latin-1"))
        self._h = ctypes.c_void_p(h)

    def __len__(self) -> int:
        return int(lib().rejit_b200_text_length(self._h))

    def device_ptr(self) -> int:
        return int(lib().rejit_b200_text_device_ptr(self._h) or 0)

    def download(self) -> bytes:
        n = len(self)
        buf = ctypes.create_string_buffer(max(n, 1))
        err = ctypes.create_string_buffer(512)
        if lib().rejit_b200_text_download(self._h, buf, n, err, len(err)) != 0:
            raise RejitError(err.value.decode("latin-1"))
        return buf.raw[:n]

    def free(self):
        if getattr(self, "_h", None):
            lib().rejit_b200_text_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Regej:
    """Mirror of rejit::Regej.  `status` is 0 (RejitSuccess) or -1 (ParserError);
    like the reference, matching on a Regej that failed to parse returns
    False / 0 (/root/reference/src/rejit.cc:229-232)."""

    def __init__(self, pattern, parser_opt: bool = True):
        L = lib()
        self.pattern = _as_bytes(pattern)
        self._ir = ctypes.c_void_p()
        self._prog = None
        err = ctypes.create_string_buffer(512)
        self.status = L.rejit_b200_parse(self.pattern, len(self.pattern), 1 if parser_opt else 0,
                                         ctypes.byref(self._ir), err, len(err))
        self.status_string = err.value.decode("latin-1")

    # -- plumbing ------------------------------------------------------------
    def compile(self) -> bool:
        if self.status != 0:
            return False
        if self._prog:
            return True
        err = ctypes.create_string_buffer(512)
        p = lib().rejit_b200_compile(self._ir, err, len(err))
        if not p:
            raise RejitError(err.value.decode("latin-1"))
        self._prog = ctypes.c_void_p(p)
        return True

    def describe(self) -> str:
        self.compile()
        return lib().rejit_b200_program_describe(self._prog).decode("latin-1")

    def shardable(self) -> bool:
        """False for re-entrant patterns: the slab entry points refuse them (rejit_b200_program_is_shardable)."""
        return bool(self.compile() and lib().rejit_b200_program_is_shardable(self._prog))

    def ir_dump(self) -> str:
        n = lib().rejit_b200_ir_dump(self._ir, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        lib().rejit_b200_ir_dump(self._ir, buf, n + 1)
        return buf.value.decode("latin-1")

    def __del__(self):
        try:
            if self._prog:
                lib().rejit_b200_program_free(self._prog)
            if self._ir:
                lib().rejit_b200_ir_free(self._ir)
        except Exception:
            pass

    # -- matching --------------------------------------------------------------
    def match_all(self, text, stats: Optional[Stats] = None, n_gpus: int = 0) -> List[Tuple[int, int]]:
        if not self.compile():
            return []
        L = lib()
        t = _as_bytes(text)
        pairs = ctypes.POINTER(ctypes.c_uint64)()
        err = ctypes.create_string_buffer(512)
        st = stats if stats is not None else Stats()
        buf = ctypes.cast(ctypes.c_char_p(t), ctypes.c_void_p)
        if n_gpus and n_gpus > 0:
            n = L.rejit_b200_match_all_multi_gpu(self._prog, buf, len(t), n_gpus, ctypes.byref(pairs),
                                                 ctypes.byref(st), err, len(err))
        else:
            n = L.rejit_b200_match_all_alloc(self._prog, buf, len(t), ctypes.byref(pairs),
                                             ctypes.byref(st), err, len(err))
        if n < 0:
            raise RejitError(err.value.decode("latin-1"))
        out = [(pairs[2 * i], pairs[2 * i + 1]) for i in range(n)]
        L.rejit_b200_free(pairs)
        return out

    def match_all_array(self, data, stats: Optional[Stats] = None, n_gpus: int = 0):
        """MatchAll over a numpy uint8 array / bytes; returns an (n, 2) uint64 array."""
        import numpy as np
        if not self.compile():
            return np.zeros((0, 2), dtype=np.uint64)
        L = lib()
        arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
        pairs = ctypes.POINTER(ctypes.c_uint64)()
        err = ctypes.create_string_buffer(512)
        st = stats if stats is not None else Stats()
        ptr = ctypes.c_void_p(arr.ctypes.data)
        if n_gpus and n_gpus > 0:
            n = L.rejit_b200_match_all_multi_gpu(self._prog, ptr, arr.size, n_gpus, ctypes.byref(pairs),
                                                 ctypes.byref(st), err, len(err))
        else:
            n = L.rejit_b200_match_all_alloc(self._prog, ptr, arr.size, ctypes.byref(pairs), ctypes.byref(st),
                                             err, len(err))
        if n < 0:
            raise RejitError(err.value.decode("latin-1"))
        out = np.ctypeslib.as_array(pairs, shape=(max(int(n), 1), 2))[:int(n)].copy() if n else np.zeros((0, 2), dtype=np.uint64)
        L.rejit_b200_free(pairs)
        return out

    def match_all_parallel(self, text, n_gpus: int) -> List[Tuple[int, int]]:
        return self.match_all(text, n_gpus=n_gpus)

    def match_all_count(self, text) -> int:
        return len(self.match_all(text))

    def match_first(self, text) -> Optional[Tuple[int, int]]:
        if not self.compile():
            return None
        t = _as_bytes(text)
        pair = (ctypes.c_uint64 * 2)()
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_match_first(self._prog, t, len(t), pair, err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return (pair[0], pair[1]) if r == 1 else None

    def match_full(self, text) -> bool:
        if not self.compile():
            return False
        t = _as_bytes(text)
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_match_full(self._prog, t, len(t), err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return r == 1

    def match_anywhere(self, text) -> bool:
        if not self.compile():
            return False
        t = _as_bytes(text)
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_match_anywhere(self._prog, t, len(t), err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return r == 1

    # -- ReplaceAll (rebuilt on the device) ---------------------------------------
    def replace_all(self, text, with_, stats: Optional[Stats] = None):
        """Regej::ReplaceAll: returns (rebuilt bytes, number of matches)."""
        if not self.compile():
            return _as_bytes(text), 0
        t, w = _as_bytes(text), _as_bytes(with_)
        out, n_out = ctypes.c_void_p(), ctypes.c_size_t()
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_replace_all(self._prog, ctypes.cast(ctypes.c_char_p(t), ctypes.c_void_p), len(t),
                                         ctypes.cast(ctypes.c_char_p(w), ctypes.c_void_p), len(w),
                                         ctypes.byref(out), ctypes.byref(n_out),
                                         ctypes.byref(stats) if stats is not None else None, err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        try:
            return ctypes.string_at(out.value, n_out.value), int(r)
        finally:
            lib().rejit_b200_free(out)

    def replace_all_text(self, text: "Text", with_, stats: Optional[Stats] = None):
        """The same on an uploaded Text; returns (new Text, number of matches)."""
        if not self.compile():
            raise ParserError(self.status_string)
        w = _as_bytes(with_)
        n = ctypes.c_int64()
        err = ctypes.create_string_buffer(512)
        h = lib().rejit_b200_replace_all_text(self._prog, text._h, ctypes.cast(ctypes.c_char_p(w), ctypes.c_void_p),
                                              len(w), ctypes.byref(n),
                                              ctypes.byref(stats) if stats is not None else None, err, len(err))
        if not h:
            raise RejitError(err.value.decode("latin-1"))
        return Text(_handle=h), int(n.value)

    def match_all_text(self, text: "Text", stats: Optional[Stats] = None) -> List[Tuple[int, int]]:
        if not self.compile():
            return []
        pairs = ctypes.POINTER(ctypes.c_uint64)()
        err = ctypes.create_string_buffer(512)
        n = lib().rejit_b200_match_all_text(self._prog, text._h, ctypes.byref(pairs),
                                            ctypes.byref(stats) if stats is not None else None, err, len(err))
        if n < 0:
            raise RejitError(err.value.decode("latin-1"))
        try:
            return [(int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(n)]
        finally:
            lib().rejit_b200_free(pairs)

    # -- device-resident text --------------------------------------------------
    def match_all_device(self, dtext: DeviceText, length: Optional[int] = None, out_ptr=None,
                         capacity: int = 0, stats: Optional[Stats] = None,
                         carry_in: Optional[Carry] = None, carry_out: Optional[Carry] = None,
                         own: Optional[Tuple[int, int]] = None, base_offset: int = 0) -> int:
        if not self.compile():
            return 0
        err = ctypes.create_string_buffer(512)
        if own is not None:
            n = lib().rejit_b200_match_all_device_slab(
                self._prog, dtext.device, dtext.ptr, dtext.nbytes if length is None else length,
                own[0], own[1], base_offset, out_ptr, capacity,
                ctypes.byref(carry_in) if carry_in is not None else None,
                ctypes.byref(carry_out) if carry_out is not None else None,
                ctypes.byref(stats) if stats is not None else None, err, len(err))
            if n < 0:
                raise RejitError(err.value.decode("latin-1"))
            return int(n)
        n = lib().rejit_b200_match_all_device(
            self._prog, dtext.device, dtext.ptr, dtext.nbytes if length is None else length,
            out_ptr, capacity,
            ctypes.byref(carry_in) if carry_in is not None else None,
            ctypes.byref(carry_out) if carry_out is not None else None,
            ctypes.byref(stats) if stats is not None else None, err, len(err))
        if n < 0:
            raise RejitError(err.value.decode("latin-1"))
        return int(n)


class RegejSet:
    """Several patterns matched against the same text; fused into one scan when
    every member is a fixed-length, anchor-free alternation."""

    def __init__(self, patterns):
        self.members = [p if isinstance(p, Regej) else Regej(p) for p in patterns]
        for m in self.members:
            if not m.compile():
                raise ParserError(m.status_string)
        arr = (ctypes.c_void_p * len(self.members))(*[m._prog for m in self.members])
        self._set = ctypes.c_void_p(lib().rejit_b200_set_create(arr, len(self.members)))

    def describe(self) -> str:
        return lib().rejit_b200_set_describe(self._set).decode("latin-1")

    def kmer_tables(self):
        """(info, bitmap, mask16) of the set's k-mer index as numpy arrays, or None
        (rejit_b200_set_kmer_tables; tests emulate the kernel's arithmetic with it)."""
        import numpy as np
        info = np.zeros(14, dtype=np.uint32)
        bitmap = np.zeros(32768, dtype=np.uint32)
        mask16 = np.zeros(65536, dtype=np.uint32)
        ok = lib().rejit_b200_set_kmer_tables(self._set, info.ctypes.data, bitmap.ctypes.data, mask16.ctypes.data)
        if not ok:
            return None
        return info, bitmap[:1 << (2 * (7 + int(info[13])) - 5)], mask16

    def __del__(self):
        try:
            if self._set:
                lib().rejit_b200_set_free(self._set)
        except Exception:
            pass

    def match_all_device(self, dtext: "DeviceText", stats: Optional[Stats] = None, own=None, base_offset: int = 0,
                         carry_in=None, carry_out=None) -> List[int]:
        counts = (ctypes.c_int64 * len(self.members))()
        err = ctypes.create_string_buffer(512)
        if own is not None:
            r = lib().rejit_b200_match_all_set_device_slab(
                self._set, dtext.device, dtext.ptr, dtext.nbytes, own[0], own[1], base_offset, carry_in, carry_out,
                counts, ctypes.byref(stats) if stats is not None else None, err, len(err))
            if r < 0:
                raise RejitError(err.value.decode("latin-1"))
            return list(counts)
        r = lib().rejit_b200_match_all_set_device(self._set, dtext.device, dtext.ptr, dtext.nbytes, counts,
                                                  ctypes.byref(stats) if stats is not None else None, err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return list(counts)

    def match_all_device_stitched(self, dtext: "DeviceText", own, base_offset: int, stats: Optional[Stats] = None):
        """Slab call + device-side stitch in one step (rejit_b200_match_all_set_device_stitched).
        Returns (counts, carries_out [(cur, tail)] in buffer coordinates, arrived [(cur, tail)] global, redo mask)."""
        k = len(self.members)
        counts = (ctypes.c_int64 * k)()
        cout, arr = (Carry * k)(), (Carry * k)()
        redo = ctypes.c_uint32()
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_match_all_set_device_stitched(self._set, dtext.device, dtext.ptr, dtext.nbytes, own[0], own[1], base_offset,
                                                           cout, arr, ctypes.byref(redo), counts,
                                                           ctypes.byref(stats) if stats is not None else None, err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return list(counts), [(int(c.cur), int(c.tail)) for c in cout], [(int(a.cur), int(a.tail)) for a in arr], int(redo.value)

    def match_all_text(self, text: "Text", stats: Optional[Stats] = None) -> List[int]:
        """Counts per member over an uploaded Text (no match lists copied back)."""
        k = len(self.members)
        counts = (ctypes.c_int64 * k)()
        err = ctypes.create_string_buffer(512)
        r = lib().rejit_b200_match_all_set_text(self._set, text._h, counts, None,
                                                ctypes.byref(stats) if stats is not None else None, err, len(err))
        if r < 0:
            raise RejitError(err.value.decode("latin-1"))
        return list(counts)

    def match_all(self, data) -> List[List[Tuple[int, int]]]:
        """Uploads `data` once and returns one match list per member."""
        L = lib()
        t = _as_bytes(data) if not hasattr(data, "ctypes") else data
        err = ctypes.create_string_buffer(512)
        if hasattr(t, "ctypes"):
            import numpy as np
            keep = np.ascontiguousarray(t, dtype=np.uint8)
            ptr, n = ctypes.c_void_p(keep.ctypes.data), keep.size
        else:
            ptr, n = ctypes.cast(ctypes.c_char_p(t), ctypes.c_void_p), len(t)
        h = L.rejit_b200_text_upload(0, ptr, n, err, len(err))
        if not h:
            raise RejitError(err.value.decode("latin-1"))
        try:
            k = len(self.members)
            counts = (ctypes.c_int64 * k)()
            pairs = (ctypes.POINTER(ctypes.c_uint64) * k)()
            r = L.rejit_b200_match_all_set_text(self._set, h, counts, pairs, None, err, len(err))
            if r < 0:
                raise RejitError(err.value.decode("latin-1"))
            out = []
            for j in range(k):
                out.append([(pairs[j][2 * i], pairs[j][2 * i + 1]) for i in range(counts[j])])
                L.rejit_b200_free(pairs[j])
            return out
        finally:
            L.rejit_b200_text_free(h)


def replace_all_set_text(patterns, text: "Text", withs, stats: Optional[Stats] = None):
    """ReplaceAll(patterns[0], withs[0]), then patterns[1], ... on an uploaded Text, as ONE pass when every pattern
    matches exactly one byte (rejit_b200_replace_all_set_text).  Returns (new Text, [matches per pattern])."""
    regs = [p if isinstance(p, Regej) else Regej(p) for p in patterns]
    for r in regs:
        if not r.compile():
            raise ParserError(r.status_string)
    k = len(regs)
    progs = (ctypes.c_void_p * k)(*[r._prog for r in regs])
    ws = [_as_bytes(w) for w in withs]
    warr = (ctypes.c_char_p * k)(*ws)
    lens = (ctypes.c_size_t * k)(*[len(w) for w in ws])
    counts = (ctypes.c_int64 * k)()
    err = ctypes.create_string_buffer(512)
    h = lib().rejit_b200_replace_all_set_text(progs, k, text._h, warr, lens, counts,
                                              ctypes.byref(stats) if stats is not None else None, err, len(err))
    if not h:
        raise RejitError(err.value.decode("latin-1"))
    return Text(_handle=h), list(counts)


def match_all(pattern, text):
    return Regej(pattern).match_all(text)
