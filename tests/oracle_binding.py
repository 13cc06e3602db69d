"""ctypes binding of oracle/_build/libqk_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's count path (oracle/qk_oracle.c).  It is
the checker; nothing under quick-mer2_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_build" / "libqk_oracle.so"
EXE = ROOT / "oracle" / "_build" / "qk_oracle"


class Dict(C.Structure):
    _fields_ = [("k", C.c_uint8), ("n_slots", C.c_uint64), ("first", C.c_uint64), ("keys", C.POINTER(C.c_uint64)),
                ("next", C.POINTER(C.c_uint32))]


class Stats(C.Structure):
    _fields_ = [("total_kmers", C.c_uint64), ("hits", C.c_uint64), ("lines", C.c_uint64), ("bases", C.c_uint64),
                ("undefined_lines", C.c_uint64), ("fastq", C.c_int)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Pass1(C.Structure):
    """qko_pass1 (oracle/qk_oracle.c)."""
    _fields_ = [("hash_size", C.c_uint64), ("distinct", C.c_uint64), ("unique", C.c_uint64), ("resizes", C.c_uint64),
                ("keys", C.POINTER(C.c_uint64)), ("occ", C.POINTER(C.c_uint8))]


class Oracle:
    def __init__(self):
        if not LIB.exists():
            subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "port"], check=True)
        L = C.CDLL(str(LIB))
        L.qko_djb.restype = C.c_uint64
        L.qko_djb.argtypes = [C.c_uint64]
        L.qko_dict_load.argtypes = [C.c_char_p, C.POINTER(Dict)]
        L.qko_dict_free.argtypes = [C.POINTER(Dict)]
        L.qko_chain_length.restype = C.c_uint64
        L.qko_chain_length.argtypes = [C.POINTER(Dict)]
        L.qko_chunk_keys.restype = C.c_uint64
        L.qko_chunk_keys.argtypes = [C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.qko_frame_file.restype = C.c_size_t
        L.qko_frame_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(Stats)]
        L.qko_count_file.argtypes = [C.POINTER(Dict), C.c_char_p, C.c_void_p, C.POINTER(Stats)]
        L.qko_chain_gather.restype = C.c_uint64
        L.qko_chain_gather.argtypes = [C.POINTER(Dict), C.c_void_p, C.c_void_p, C.c_uint64]
        L.qko_count_t.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint, C.POINTER(Stats)]
        L.qko_fifo_padding.restype = C.c_uint64
        L.qko_fifo_padding.argtypes = [C.POINTER(Dict), C.c_uint64, C.c_void_p]
        L.qko_search_pass1.argtypes = [C.c_char_p, C.c_uint8, C.c_uint64, C.POINTER(Pass1)]
        L.qko_pass1_free.argtypes = [C.POINTER(Pass1)]
        self.L = L

    def search_pass1(self, fasta, k: int, hash_size: int) -> dict:
        """`search` pass 1 (Q.c:824-923): the k-mer occurrence table of a reference FASTA.  Returns the figures the
        reference prints (distinct = its `total`, unique) and the table as {key: occurrences}."""
        p = Pass1()
        rc = self.L.qko_search_pass1(os.fsencode(str(fasta)), k, hash_size, C.byref(p))
        assert rc == 0, f"oracle search pass 1 failed: {rc}"
        try:
            keys = np.ctypeslib.as_array(p.keys, shape=(p.hash_size,)).copy()
            occ = np.ctypeslib.as_array(p.occ, shape=(p.hash_size,)).copy()
            return {"hash_size": int(p.hash_size), "distinct": int(p.distinct), "unique": int(p.unique), "resizes": int(p.resizes),
                    "keys": keys, "occ": occ}
        finally:
            self.L.qko_pass1_free(C.byref(p))

    def djb(self, key: int) -> int:
        return self.L.qko_djb(key)

    def chunk_keys(self, k: int, chunk: bytes) -> np.ndarray:
        """Canonical keys the reference codec (Q.c:399-420) emits for a framed chunk."""
        buf = np.frombuffer(chunk, dtype=np.uint8) if chunk else np.zeros(0, np.uint8)
        cap = max(1, len(chunk))
        out = np.zeros(cap, dtype=np.uint64)
        n = self.L.qko_chunk_keys(k, buf.ctypes.data, buf.size, out.ctypes.data, cap)
        assert n <= cap
        return out[:n]

    def frame_file(self, path) -> tuple[bytes, dict]:
        st = Stats()
        size = os.path.getsize(path) + 16
        out = np.zeros(size, dtype=np.uint8)
        n = self.L.qko_frame_file(os.fsencode(str(path)), out.ctypes.data, size, C.byref(st))
        assert n != C.c_size_t(-1).value and n <= size
        return out[:n].tobytes(), st.as_dict()

    def count(self, ref_prefix, reads, out_prefix, threads: int = 0) -> dict:
        """The whole command: writes <out_prefix>.bin (and .txt when <ref_prefix>.qgc exists).  threads = the
        reference's -t: only whether it is 0 matters (the batch padding of Q.c:458-466)."""
        st = Stats()
        rc = self.L.qko_count_t(os.fsencode(str(ref_prefix)), os.fsencode(str(reads)), os.fsencode(str(out_prefix)),
                                threads, C.byref(st))
        assert rc == 0, f"oracle count failed: {rc}"
        return st.as_dict()

    def count_bin(self, qm_path, reads, threads: int = 0) -> tuple[np.ndarray, dict]:
        """Depths in chain order (= .bin contents) for a dictionary file and a reads file (threads: see count)."""
        d = Dict()
        rc = self.L.qko_dict_load(os.fsencode(str(qm_path)), C.byref(d))
        assert rc == 0, f"oracle dict load failed: {rc}"
        try:
            depth = np.zeros(d.n_slots, dtype=np.uint16)
            st = Stats()
            assert self.L.qko_count_file(C.byref(d), os.fsencode(str(reads)), depth.ctypes.data, C.byref(st)) == 0
            if threads:
                self.L.qko_fifo_padding(C.byref(d), st.total_kmers, depth.ctypes.data)
            n = self.L.qko_chain_length(C.byref(d))
            out = np.zeros(n, dtype=np.uint16)
            self.L.qko_chain_gather(C.byref(d), depth.ctypes.data, out.ctypes.data, n)
        finally:
            self.L.qko_dict_free(C.byref(d))
        return out, st.as_dict()


def write_qm(path, k: int, keys: np.ndarray, nxt: np.ndarray, first: int):
    """Write a QM11 file (layout of Q.c:1284-1299) from arrays."""
    hdr = bytearray(24)
    hdr[0:4] = b"QM11"
    hdr[4] = k
    hdr[5], hdr[6], hdr[7] = 0, 100, 100
    hdr[8:16] = int(keys.size).to_bytes(8, "little")
    hdr[16:24] = int(first).to_bytes(8, "little")
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(np.ascontiguousarray(keys, dtype="<u8").tobytes())
        f.write(np.ascontiguousarray(nxt, dtype="<u4").tobytes())


def build_qm_arrays(oracle: Oracle, ordered_keys, n_slots: int):
    """Place keys (in chain order; duplicates allowed) with the reference's probe rule
    (Q.c:90-99 as used by the inserters Q.c:866-887, 209-216) and link the chain."""
    keys = np.zeros(n_slots, dtype=np.uint64)
    nxt = np.zeros(n_slots, dtype=np.uint32)
    slots = []
    for key in ordered_keys:
        key = int(key)
        s = oracle.djb(key) & (n_slots - 1)
        step = -1 if s & (n_slots >> 1) else 1
        while keys[s] != 0:      # duplicates take the next free slot, as `index` does
            s += step
        keys[s] = key
        slots.append(s)
    for a, b in zip(slots, slots[1:] + slots[:1]):
        nxt[a] = b
    return keys, nxt, slots[0]
