"""Host-side mirror of the reference's operator surface for the hot path, over the C ABI.

Names follow the reference:
  flat_knn      <-> VectorIndex::tryBruteForceSearch / VIWithColumnInPart::searchWithoutIndex
                    (src/VectorIndex/Common/BruteForceSearch.h:63, VIWithDataPart.h:342)
  part_scan     <-> MergeTreeVSManager::vectorScanWithoutIndex + searchWrapper
                    (src/VectorIndex/Storages/MergeTreeVSManager.cpp:960-1679)
  Corpus.search <-> Search::VectorIndex(FLAT)::search via VIWithColumnInPart::search
                    (src/VectorIndex/Common/VIWithDataPart.cpp:858-957)
  topk_merge_device <-> MergeTreeBaseSearchManager::getTotalTopSearchResultImpl
                    (src/VectorIndex/Storages/MergeTreeBaseSearchManager.cpp:207-299)
All arithmetic happens in libb200search.so on the GPU; numpy is used for buffers only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib

L2, IP, COSINE, HAMMING, JACCARD = 0, 1, 2, 3, 4
METRIC_NAMES = {"L2": L2, "IP": IP, "COSINE": COSINE, "HAMMING": HAMMING, "JACCARD": JACCARD}
F32, BF16, BIN = 0, 1, 2


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200 error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _bits(a):
    return None if a is None else np.ascontiguousarray(a, np.uint8)


def flat_knn(metric, x, y, k, alive_bits=None):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    nx, d = x.shape
    ny = y.shape[0]
    dis = np.empty((nx, k), np.float32)
    ids = np.empty((nx, k), np.int64)
    ab = _bits(alive_bits)
    _check(lib().b200_flat_knn(C.c_int(metric), _p(x, C.c_float), C.c_int64(nx), _p(y, C.c_float), C.c_int64(ny),
                               C.c_int(d), C.c_int(k), _p(ab, C.c_uint8), _p(dis, C.c_float), _p(ids, C.c_int64)))
    return dis, ids


def binary_knn(metric, x, y, k, alive_bits=None):
    x = np.ascontiguousarray(x, np.uint8)
    y = np.ascontiguousarray(y, np.uint8)
    nx, nb = x.shape
    ny = y.shape[0]
    dis = np.empty((nx, k), np.float32)
    ids = np.empty((nx, k), np.int64)
    ab = _bits(alive_bits)
    _check(lib().b200_binary_knn(C.c_int(metric), _p(x, C.c_uint8), C.c_int64(nx), _p(y, C.c_uint8), C.c_int64(ny),
                                 C.c_int(nb), C.c_int(k), _p(ab, C.c_uint8), _p(dis, C.c_float), _p(ids, C.c_int64)))
    return dis, ids


def part_scan(metric, x, y, k, block_rows=8192, row_exists=None, filter_bits=None):
    binary = metric in (HAMMING, JACCARD)
    if binary:
        x = np.ascontiguousarray(x, np.uint8)
        y = np.ascontiguousarray(y, np.uint8)
        d = x.shape[1] * 8
    else:
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        d = x.shape[1]
    nx, ny = x.shape[0], y.shape[0]
    dis = np.empty((nx, k), np.float32)
    ids = np.empty((nx, k), np.int64)
    re_ = None if row_exists is None else np.ascontiguousarray(row_exists, np.uint8)
    fb = _bits(filter_bits)
    _check(lib().b200_part_scan(C.c_int(metric), x.ctypes.data_as(C.c_void_p), C.c_int64(nx),
                                y.ctypes.data_as(C.c_void_p), C.c_int64(ny), C.c_int(d), C.c_int(k),
                                C.c_int64(block_rows), _p(re_, C.c_uint8), _p(fb, C.c_uint8), _p(dis, C.c_float),
                                _p(ids, C.c_int64)))
    return dis, ids


class Corpus:
    """Device-resident FLAT index / cached part column."""

    def __init__(self, metric, d, dtype=F32, capacity=0):
        self._h = C.c_void_p()
        self.metric, self.d, self.dtype = metric, d, dtype
        _check(lib().b200_corpus_create(C.c_int(metric), C.c_int(dtype), C.c_int(d), C.c_int64(capacity), C.byref(self._h)))

    def append(self, rows):
        rows = np.ascontiguousarray(rows, np.uint8 if self.dtype == BIN else np.float32)
        _check(lib().b200_corpus_append(self._h, rows.ctypes.data_as(C.c_void_p), C.c_int64(rows.shape[0])))
        return self

    def adopt_device(self, data_ptr: int, n: int):
        _check(lib().b200_corpus_adopt_device(self._h, C.c_void_p(data_ptr), C.c_int64(n)))
        return self

    def set_path(self, path: int):
        _check(lib().b200_corpus_set_path(self._h, C.c_int(path)))
        return self

    @property
    def size(self):
        n = C.c_int64()
        _check(lib().b200_corpus_size(self._h, C.byref(n)))
        return n.value

    def search(self, queries, k, alive_bits=None):
        q = np.ascontiguousarray(queries, np.uint8 if self.dtype == BIN else np.float32)
        nq = q.shape[0]
        dis = np.empty((nq, k), np.float32)
        ids = np.empty((nq, k), np.int64)
        ab = _bits(alive_bits)
        _check(lib().b200_corpus_search(self._h, q.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(nq), C.c_int(k),
                                        _p(ab, C.c_uint8), _p(dis, C.c_float), _p(ids, C.c_int64)))
        return dis, ids

    def search_device(self, q_ptr: int, nq: int, k: int, out_dis_ptr: int, out_ids_ptr: int, id_offset=0,
                      alive_ptr: int = 0, stream: int = 0):
        _check(lib().b200_corpus_search_device(self._h, C.c_void_p(q_ptr), C.c_int64(nq), C.c_int(k),
                                               C.c_void_p(alive_ptr or None), C.c_int64(id_offset),
                                               C.c_void_p(out_dis_ptr), C.c_void_p(out_ids_ptr),
                                               C.c_void_p(stream or None)))

    def enable_timing(self, on=True):
        _check(lib().b200_corpus_enable_timing(self._h, C.c_int(1 if on else 0)))

    def kernel_time(self, reset=False):
        """(total ms, launches) of the dominant kernel since the last reset (CUDA events)."""
        ms, n = C.c_double(), C.c_int64()
        _check(lib().b200_corpus_kernel_time(self._h, C.c_int(1 if reset else 0), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self):
        if self._h:
            lib().b200_corpus_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topk_merge_device(dis_ptr, ids_ptr, n_lists, nq, k, descending, out_dis_ptr, out_ids_ptr, stream=0):
    _check(lib().b200_topk_merge_device(C.c_void_p(dis_ptr), C.c_void_p(ids_ptr), C.c_int(n_lists), C.c_int64(nq),
                                        C.c_int(k), C.c_int(1 if descending else 0), C.c_void_p(out_dis_ptr),
                                        C.c_void_p(out_ids_ptr), C.c_void_p(stream or None)))


def topk_merge_device_strided(dis_ptr, ids_ptr, n_lists, dis_stride, ids_stride, nq, k, descending, out_dis_ptr,
                              out_ids_ptr, stream=0):
    _check(lib().b200_topk_merge_device_strided(C.c_void_p(dis_ptr), C.c_void_p(ids_ptr), C.c_int(n_lists),
                                                C.c_int64(dis_stride), C.c_int64(ids_stride), C.c_int64(nq), C.c_int(k),
                                                C.c_int(1 if descending else 0), C.c_void_p(out_dis_ptr),
                                                C.c_void_p(out_ids_ptr), C.c_void_p(stream or None)))


def launch_count(reset=False) -> int:
    return int(lib().b200_launch_count(C.c_int(1 if reset else 0)))
