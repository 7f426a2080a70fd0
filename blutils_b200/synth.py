"""Seeded synthetic workloads -- thin wrapper over libblu_synth.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

from . import _ffi

BASE_SEED = 20261018


class SynthWorkload:
    def __init__(self, n_taxa: int, seed: int = BASE_SEED):
        self._l = _ffi.synth_lib()
        self._h = C.c_void_p(self._l.blu_synth_create(n_taxa, seed))
        if not self._h:
            raise ValueError("blu_synth_create failed")
        self.n_taxa = n_taxa

    def lineages(self, numeric: bool = False):
        """(taxids int64 ndarray, offsets uint64 ndarray[n+1], blob bytes-like ndarray)"""
        import numpy as np

        n = self.n_taxa
        tot = self._l.blu_synth_lineages(self._h, int(numeric), None, None, None)
        ids = np.empty(n, dtype=np.int64)
        off = np.empty(n + 1, dtype=np.uint64)
        blob = np.empty(max(tot, 1), dtype=np.uint8)
        self._l.blu_synth_lineages(self._h, int(numeric), ids.ctypes.data, off.ctypes.data, blob.ctypes.data)
        return ids, off, blob

    def write_json(self, path: str) -> str:
        if self._l.blu_synth_write_json(self._h, os.fspath(path).encode()):
            raise IOError(path)
        return path

    def hits_into(self, dst_ptr: int, cap: int, q_begin: int, n_queries: int, hits: int, zipf: bool = False,
                  threads: Optional[int] = None) -> Tuple[int, int]:
        """Generates queries [q_begin, q_begin+n_queries) into a caller buffer; returns (bytes, rows)."""
        ln, nr = C.c_uint64(), C.c_uint64()
        rc = self._l.blu_synth_hits(self._h, q_begin, n_queries, 1 if zipf else 0, hits, threads or (os.cpu_count() or 1), dst_ptr, cap,
                                    C.byref(ln), C.byref(nr))
        if rc == 2:
            e = MemoryError(f"buffer too small: need {ln.value} bytes")
            e.needed = ln.value
            raise e
        if rc:
            raise RuntimeError("blu_synth_hits failed")
        return ln.value, nr.value

    def hits(self, q_begin: int, n_queries: int, hits: int, zipf: bool = False, threads: Optional[int] = None) -> bytes:
        cap = n_queries * (hits if not zipf else min(hits, 600)) * 96 + 4096
        buf = C.create_string_buffer(cap)
        try:
            n, _ = self.hits_into(C.addressof(buf), cap, q_begin, n_queries, hits, zipf, threads)
        except MemoryError as e:  # (Zipf tables: the size is only known once generated)
            cap = int(getattr(e, "needed", 0)) + 4096
            buf = C.create_string_buffer(cap)
            n, _ = self.hits_into(C.addressof(buf), cap, q_begin, n_queries, hits, zipf, threads)
        return buf.raw[:n]

    def close(self):
        if self._h:
            self._l.blu_synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
