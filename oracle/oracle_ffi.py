"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_build/libblu_oracle.so (the C++ CPU restatement).
May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Dict, List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libblu_oracle.so")
TAXON = {"fungi": 0, "bacteria": 1, "eukaryotes": 2, "custom": 3}
STRATEGY = {"cautious": 0, "relaxed": 1}
_CUSTOM_KEYS = ["domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"]


class OracleDataError(Exception):
    pass


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "blu_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.blu_oracle_create.restype = C.c_void_p
        l.blu_oracle_create.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_char_p, C.c_int]
        l.blu_oracle_destroy.argtypes = [C.c_void_p]
        l.blu_oracle_run.restype = C.c_int
        l.blu_oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        l.blu_oracle_free.argtypes = [C.c_void_p]
        l.blu_oracle_checksum_jsonl.restype = C.c_uint64
        l.blu_oracle_checksum_jsonl.argtypes = [C.c_char_p, C.c_uint64]
        l.blu_oracle_interpolate.restype = C.c_int
        l.blu_oracle_interpolate.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int]
        _lib = l
    return _lib


def _custom8(custom: Optional[Dict[str, Optional[int]]]):
    if custom is None:
        return None
    arr = (C.c_int * 8)(*[(-1 if custom.get(k) is None else int(custom[k])) for k in _CUSTOM_KEYS])
    return arr


def read_taxonomy_json(path: str, use_taxid: bool):
    """mod.rs:246-327 via Python's json module: returns (taxids int64 list, lineage strings)."""
    d = json.load(open(path))
    for k in ("blutilsVersion", "sourceDatabase", "taxonomies"):
        if k not in d:
            raise IOError("taxonomies json: missing " + k)
    ids, lin = [], []
    for u in d["taxonomies"]:
        f = float(u["taxid"])
        if f >= 9223372036854775808.0:
            continue
        ids.append(int(f))
        lin.append(u["numericLineage"] if use_taxid else u["textLineage"])
    return ids, lin


class Oracle:
    def __init__(self, taxids: Sequence[int], lineages: Sequence[str], taxon: str, strategy: str,
                 custom: Optional[dict] = None, threads: int = 0):
        import numpy as np

        self.threads = threads or (os.cpu_count() or 1)
        enc = [s.encode("utf-8") for s in lineages]
        off = np.zeros(len(enc) + 1, dtype=np.uint64)
        if enc:
            off[1:] = np.cumsum([len(b) for b in enc], dtype=np.uint64)
        blob = b"".join(enc)
        ids = np.asarray(list(taxids), dtype=np.int64)
        err = C.create_string_buffer(512)
        c8 = _custom8(custom)
        self._h = lib().blu_oracle_create(ids.ctypes.data, off.ctypes.data, blob, len(enc), TAXON[taxon],
                                          C.cast(c8, C.c_void_p) if c8 is not None else None, STRATEGY[strategy],
                                          self.threads, err, 512)
        if not self._h:
            raise OracleDataError(err.value.decode())

    @classmethod
    def from_json(cls, path: str, taxon: str, strategy: str, use_taxid: bool = False, custom=None, threads: int = 0):
        ids, lin = read_taxonomy_json(path, use_taxid)
        return cls(ids, lin, taxon, strategy, custom, threads)

    def run_raw(self, text, nbytes: Optional[int] = None, headers: Optional[List[str]] = None, threads: Optional[int] = None):
        """text: bytes or an integer address.  Returns (jsonl bytes sorted by query, n_queries, n_rows)."""
        if isinstance(text, (bytes, bytearray)):
            buf = (C.c_char * len(text)).from_buffer_copy(text) if len(text) else C.create_string_buffer(1)
            addr, n = C.addressof(buf), len(text)
        else:
            addr, n = int(text), int(nbytes)
        out = C.c_void_p()
        ol, nq, nr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        err = C.create_string_buffer(512)
        hb = "\n".join(headers).encode() if headers else None
        rc = lib().blu_oracle_run(self._h, addr, n, threads or self.threads, hb, len(hb) if hb else 0, C.byref(out), C.byref(ol),
                                  C.byref(nq), C.byref(nr), err, 512)
        if rc == 2:
            raise OracleDataError(err.value.decode())
        if rc != 0:
            raise RuntimeError(err.value.decode())
        try:
            return C.string_at(out, ol.value), nq.value, nr.value
        finally:
            lib().blu_oracle_free(out)

    def run(self, text: bytes, headers: Optional[List[str]] = None) -> List[dict]:
        js, _, _ = self.run_raw(text, headers=headers)
        return [json.loads(l) for l in js.decode("utf-8").splitlines()]

    def close(self):
        if self._h:
            lib().blu_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interpolate(ranks: Sequence[str], taxon: str, custom: Optional[dict] = None) -> List[float]:
    out = (C.c_double * 128)()
    c8 = _custom8(custom)
    n = lib().blu_oracle_interpolate("\n".join(ranks).encode(), TAXON[taxon], C.cast(c8, C.c_void_p) if c8 is not None else None,
                                     out, 128)
    if n < 0:
        raise OracleDataError("interpolate failed")
    return [out[i] for i in range(n)]


def checksum_jsonl(js: bytes) -> int:
    return lib().blu_oracle_checksum_jsonl(js, len(js))
