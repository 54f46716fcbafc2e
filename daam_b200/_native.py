"""ctypes binding of ``libdaam_b200.so`` (C ABI: ``include/daam_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``daam_b200/build.py``). There is no CPU or torch
fallback behind these calls: if the shared object is missing, or a call fails, the caller gets an exception.
ctypes releases the GIL around every foreign call.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

LIB_NAME = 'libdaam_b200.so'
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

DAAM_F32, DAAM_F16, DAAM_BF16 = 0, 1, 2
ACC_AUTO, ACC_FORCE_SIMT, ACC_FORCE_MMA = 0, 1, 2
ACC_RMW_AUTO, ACC_RMW_LDST, ACC_RMW_RED = 0x00, 0x10, 0x20
ACC_NO_PDL = 0x100
ACC_EARLY_LOADS = 0x200   # see include/daam_b200.h: only valid when q/k were complete before the previous kernel started
ABI_VERSION = 3
E_INVALID, E_UNSUPPORTED, E_CUDA = -1, -2, -3
TOKENS = 77
EXPAND_SCRATCH_FLOATS = 64   # DAAM_EXPAND_SCRATCH_FLOATS: per word

EXPORTS = ('daam_accumulate', 'daam_attention_probs', 'daam_accumulate_probs', 'daam_finalize', 'daam_finalize_per_key', 'daam_word_heat_map', 'daam_expand_as', 'daam_expand_words', 'daam_side_launcher_create', 'daam_side_launcher_destroy',
           'daam_side_launcher_launch', 'daam_side_launcher_join', 'daam_side_launcher_idle', 'daam_abi_version',
           'daam_last_error', 'daam_device_info', 'daam_launch_count')


class DaamLayer(ctypes.Structure):
    """``struct daam_layer`` (include/daam_b200.h)."""
    _fields_ = [
        ('q', ctypes.c_void_p), ('k', ctypes.c_void_p), ('acc', ctypes.c_void_p),
        ('q_stride_prompt', ctypes.c_int64), ('q_stride_pixel', ctypes.c_int64), ('q_stride_head', ctypes.c_int64),
        ('k_stride_prompt', ctypes.c_int64), ('k_stride_token', ctypes.c_int64), ('k_stride_head', ctypes.c_int64),
        ('n_prompts', ctypes.c_int32), ('heads', ctypes.c_int32), ('hw', ctypes.c_int32), ('tokens', ctypes.c_int32),
        ('head_dim', ctypes.c_int32), ('dtype', ctypes.c_int32), ('scale', ctypes.c_float),
        ('reserved', ctypes.c_int32),
    ]


class DaamKeyGroup(ctypes.Structure):
    """``struct daam_key_group`` (include/daam_b200.h)."""
    _fields_ = [
        ('acc', ctypes.c_void_p), ('heads', ctypes.c_int32), ('h', ctypes.c_int32), ('w', ctypes.c_int32),
        ('tokens', ctypes.c_int32), ('head_sel', ctypes.c_int32), ('reserved', ctypes.c_int32),
    ]


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'libdaam_b200: {message} (status {code})')
        self.code = code


_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library once and declare the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `python -m daam_b200.build`). daam_b200 has no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    lib.daam_abi_version.argtypes = []
    lib.daam_abi_version.restype = ctypes.c_int
    if lib.daam_abi_version() != ABI_VERSION:
        raise RuntimeError(f'{LIB_PATH} has ABI version {lib.daam_abi_version()}, this package needs {ABI_VERSION}: '
                           f'rebuild it (`python -m daam_b200.build --force`)')
    i32, u32, i64, vp, f32 = ctypes.c_int32, ctypes.c_uint32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float
    lib.daam_accumulate.argtypes = [ctypes.POINTER(DaamLayer), i32, u32, vp]
    lib.daam_accumulate.restype = ctypes.c_int
    lib.daam_attention_probs.argtypes = [ctypes.POINTER(DaamLayer), vp, vp]
    lib.daam_attention_probs.restype = ctypes.c_int
    lib.daam_accumulate_probs.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.daam_accumulate_probs.restype = ctypes.c_int
    lib.daam_finalize.argtypes = [ctypes.POINTER(DaamKeyGroup), i32, i32, i32, i32, vp, vp]
    lib.daam_finalize.restype = ctypes.c_int
    lib.daam_finalize_per_key.argtypes = [ctypes.POINTER(DaamKeyGroup), i32, i32, i32, i32, vp, vp]
    lib.daam_finalize_per_key.restype = ctypes.c_int
    lib.daam_word_heat_map.argtypes = [vp, i32, i32, ctypes.POINTER(i32), i32, vp, vp]
    lib.daam_word_heat_map.restype = ctypes.c_int
    lib.daam_expand_as.argtypes = [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp]
    lib.daam_expand_as.restype = ctypes.c_int
    lib.daam_expand_words.argtypes = [vp, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), i32, i32, i32, i32, i32, f32,
                                      vp, vp, vp, vp]
    lib.daam_expand_words.restype = ctypes.c_int
    lib.daam_side_launcher_create.argtypes = [ctypes.POINTER(vp)]
    lib.daam_side_launcher_create.restype = ctypes.c_int
    lib.daam_side_launcher_destroy.argtypes = [vp]
    lib.daam_side_launcher_destroy.restype = None
    lib.daam_side_launcher_launch.argtypes = [vp, ctypes.POINTER(DaamLayer), i32, u32, vp, vp]
    lib.daam_side_launcher_launch.restype = ctypes.c_int
    lib.daam_side_launcher_join.argtypes = [vp, vp]
    lib.daam_side_launcher_join.restype = ctypes.c_int
    lib.daam_side_launcher_idle.argtypes = [vp]
    lib.daam_side_launcher_idle.restype = ctypes.c_int
    lib.daam_last_error.argtypes = []
    lib.daam_last_error.restype = ctypes.c_char_p
    lib.daam_device_info.argtypes = [ctypes.POINTER(i32)] * 3
    lib.daam_device_info.restype = ctypes.c_int
    lib.daam_launch_count.argtypes = []
    lib.daam_launch_count.restype = i64
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise NativeError(rc, load().daam_last_error().decode('utf-8', 'replace'))


class PackedLayers:
    """A ready-made ``daam_layer[]`` (host array) for call sites that replay the same layer calls."""

    def __init__(self, layers: Sequence[DaamLayer]):
        self.n = len(layers)
        self.array = (DaamLayer * max(self.n, 1))(*layers)


class SideLauncher:
    """``daam_side_launcher``: the event pair behind the tracer's per-step launch on its side stream."""

    def __init__(self):
        self._lib = load()
        handle = ctypes.c_void_p()
        _check(self._lib.daam_side_launcher_create(ctypes.byref(handle)))
        self._h = handle

    def launch(self, packed: 'PackedLayers', flags: int, producer_stream: int, side_stream: int):
        rc = self._lib.daam_side_launcher_launch(self._h, packed.array, packed.n, flags, producer_stream, side_stream)
        if rc != 0:
            _check(rc)

    def join(self, stream: int):
        rc = self._lib.daam_side_launcher_join(self._h, stream)
        if rc != 0:
            _check(rc)

    def idle(self) -> bool:
        rc = self._lib.daam_side_launcher_idle(self._h)
        if rc < 0:
            _check(rc)
        return rc == 1

    def close(self):
        if self._h is not None and self._h.value:
            self._lib.daam_side_launcher_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def accumulate(layers, stream: int, flags: int = ACC_AUTO):
    """``layers``: a sequence of :class:`DaamLayer` or a :class:`PackedLayers`."""
    packed = layers if isinstance(layers, PackedLayers) else PackedLayers(layers)
    if packed.n == 0:
        return
    rc = load().daam_accumulate(packed.array, packed.n, flags, stream)
    if rc != 0:
        _check(rc)


def attention_probs(layer: DaamLayer, probs_ptr: int, stream: int):
    _check(load().daam_attention_probs(ctypes.byref(layer), probs_ptr, stream))


def accumulate_probs(probs_ptr: int, dtype: int, first_row: int, n_rows: int, hw: int, tokens: int, acc_ptr: int,
                     stream: int):
    _check(load().daam_accumulate_probs(probs_ptr, dtype, first_row, n_rows, hw, tokens, acc_ptr, stream))


def finalize(groups: Sequence[DaamKeyGroup], x: int, n_rows: int, normalize: bool, out_ptr: int, stream: int):
    n = len(groups)
    arr = (DaamKeyGroup * max(n, 1))(*groups)
    _check(load().daam_finalize(arr, n, x, n_rows, int(bool(normalize)), ctypes.c_void_p(out_ptr),
                                ctypes.c_void_p(stream)))


def finalize_per_key(groups: Sequence[DaamKeyGroup], x: int, n_rows: int, normalize: bool, out_ptr: int, stream: int):
    n = len(groups)
    arr = (DaamKeyGroup * max(n, 1))(*groups)
    _check(load().daam_finalize_per_key(arr, n, x, n_rows, int(bool(normalize)), ctypes.c_void_p(out_ptr),
                                        ctypes.c_void_p(stream)))


def word_heat_map(maps_ptr: int, n_rows: int, x: int, rows: Sequence[int], out_ptr: int, stream: int):
    arr = (ctypes.c_int32 * max(len(rows), 1))(*rows)
    _check(load().daam_word_heat_map(ctypes.c_void_p(maps_ptr), n_rows, x, arr, len(rows), ctypes.c_void_p(out_ptr),
                                     ctypes.c_void_p(stream)))


def expand_as(map_ptr: int, x: int, out_h: int, out_w: int, absolute: bool, threshold: Optional[float], out_ptr: int,
              scratch_ptr: int, stream: int):
    use_thr = bool(threshold)   # the reference's `if threshold:` (daam/heatmap.py:87)
    _check(load().daam_expand_as(ctypes.c_void_p(map_ptr), x, out_h, out_w, int(bool(absolute)), int(use_thr),
                                 float(threshold) if use_thr else 0.0, ctypes.c_void_p(out_ptr),
                                 ctypes.c_void_p(scratch_ptr), ctypes.c_void_p(stream)))


def expand_words(maps_ptr: int, n_rows: int, x: int, rows_per_word: Sequence[Sequence[int]], out_h: int, out_w: int,
                 absolute: bool, threshold: Optional[float], word_maps_ptr: Optional[int], out_ptr: int, scratch_ptr: int,
                 stream: int):
    """``rows_per_word[w]``: the rows of ``maps`` word ``w`` averages (already offset for SOS)."""
    flat = [r for rows in rows_per_word for r in rows]
    begin = [0]
    for rows in rows_per_word:
        begin.append(begin[-1] + len(rows))
    rows_arr = (ctypes.c_int32 * max(len(flat), 1))(*flat)
    begin_arr = (ctypes.c_int32 * len(begin))(*begin)
    use_thr = bool(threshold)   # the reference's `if threshold:` (daam/heatmap.py:87)
    _check(load().daam_expand_words(ctypes.c_void_p(maps_ptr), n_rows, x, rows_arr, begin_arr, len(rows_per_word), out_h,
                                    out_w, int(bool(absolute)), int(use_thr), float(threshold) if use_thr else 0.0,
                                    ctypes.c_void_p(word_maps_ptr) if word_maps_ptr else None,
                                    ctypes.c_void_p(out_ptr), ctypes.c_void_p(scratch_ptr), ctypes.c_void_p(stream)))


def device_info():
    sm, major, minor = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _check(load().daam_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return {'sm_count': sm.value, 'cc': (major.value, minor.value)}


def launch_count() -> int:
    return int(load().daam_launch_count())


def abi_version() -> int:
    return int(load().daam_abi_version())
