"""ctypes face of quick-mer2_b200/bin/libqk_synth_gpu.so (tools/qk_synth_gpu.cu): seeded synthetic
genomes, QM11 dictionaries and reads generated on the GPU, at any size up to the human-scale
configurations of BASELINE.json.  Test / bench data only -- the product never loads it."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = ROOT / "quick-mer2_b200" / "bin" / "libqk_synth_gpu.so"

FRAMED, FASTA, FASTQ = 0, 1, 2


class DictInfo(C.Structure):
    _fields_ = [("slots", C.c_uint64), ("n_unique", C.c_uint64), ("first", C.c_uint64), ("t1_slots", C.c_uint64),
                ("k", C.c_uint32), ("pad", C.c_uint32), ("occ_s", C.c_double), ("select_s", C.c_double), ("place_s", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "pad"}


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(LIB_PATH))
        P, U64, U32 = C.c_void_p, C.c_uint64, C.c_uint32
        sig = {
            "qs_last_error": (C.c_char_p, [P]),
            "qs_genome_create": (C.c_int, [C.POINTER(P), C.c_int, U64, U32, U64, U64, U64, U32, U64]),
            "qs_genome_from_host": (C.c_int, [C.POINTER(P), C.c_int, P, U64, P, U32]),
            "qs_genome_download": (C.c_int, [P, U64, P, U64]),
            "qs_genome_contigs": (C.c_int, [P, P, U32, C.POINTER(U32)]),
            "qs_genome_device_ptr": (P, [P]),
            "qs_genome_destroy": (None, [P]),
            "qs_dict_build": (C.c_int, [P, U32, U64, U64, C.c_int, C.POINTER(DictInfo)]),
            "qs_dict_download": (C.c_int, [P, P, P, P]),
            "qs_dict_write": (C.c_int, [P, C.c_char_p, C.c_int]),
            "qs_dict_free": (C.c_int, [P]),
            "qs_reads_record_bytes": (U64, [U32, C.c_int]),
            "qs_reads_generate": (C.c_int, [P, U64, U64, U64, U32, U32, C.c_int, P, P, P]),
            "qs_device_alloc": (P, [C.c_int, C.c_size_t]),
            "qs_device_free": (None, [P]),
            "qs_pinned_alloc": (P, [C.c_size_t]),
            "qs_pinned_free": (None, [P]),
            "qs_copy_to_host": (C.c_int, [P, P, C.c_size_t]),
            "qs_device_mem_info": (C.c_int, [C.c_int, C.POINTER(U64), C.POINTER(U64)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def record_bytes(length: int, fmt: int) -> int:
    return int(lib().qs_reads_record_bytes(length, fmt))


def hifi_lengths(n: int, seed: int, median: int = 15000, sigma: float = 0.5, lo: int = 1000, hi: int = 99998) -> np.ndarray:
    """Log-normal read lengths clipped to [lo, hi] (config 4: HiFi-like, up to the 100k line buffer)."""
    rng = np.random.default_rng(seed)
    return np.clip(median * np.exp(sigma * rng.standard_normal(n)), lo, hi).astype(np.uint32)


class DeviceBuffer:
    def __init__(self, device: int, nbytes: int):
        self.ptr = lib().qs_device_alloc(device, nbytes)
        if not self.ptr:
            raise MemoryError(f"cudaMalloc of {nbytes} bytes failed")
        self.nbytes = nbytes

    def free(self):
        if self.ptr:
            lib().qs_device_free(self.ptr)
            self.ptr = None

    def to_host(self, nbytes=None, offset=0) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        if lib().qs_copy_to_host(out.ctypes.data, self.ptr + offset, n):
            raise RuntimeError("D2H failed")
        return out


class Genome:
    def __init__(self, handle, device):
        self._h, self.device, self._lib = handle, device, lib()

    @classmethod
    def create(cls, n_bases: int, contigs: int = 1, seed: int = 1, dup_period: int = 0, dup_len: int = 0, div_ppm: int = 5000,
               nblock: int = 0, device: int = 0):
        h = C.c_void_p()
        rc = lib().qs_genome_create(C.byref(h), device, n_bases, contigs, seed, dup_period, dup_len, div_ppm, nblock)
        g = cls(h, device)
        g._check(rc)
        return g

    @classmethod
    def from_fasta(cls, path, device: int = 0):
        """Upload a (small) FASTA: contigs concatenated, newlines removed -- as qk_synth reads it."""
        seq, cur = [], None
        for line in Path(path).read_bytes().split(b"\n"):
            if line.startswith(b">"):
                if cur is not None:
                    seq.append(b"".join(cur))
                cur = []
            elif line and cur is not None:
                cur.append(line)
        if cur is not None:
            seq.append(b"".join(cur))
        starts = [0]
        for s_ in seq:
            starts.append(starts[-1] + len(s_))
        data = np.frombuffer(b"".join(seq), dtype=np.uint8)
        st = np.asarray(starts, dtype=np.uint64)
        h = C.c_void_p()
        rc = lib().qs_genome_from_host(C.byref(h), device, data.ctypes.data, data.size, st.ctypes.data, len(starts) - 1)
        g = cls(h, device)
        g._check(rc)
        return g

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"qk_synth_gpu error {rc}: {self._lib.qs_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            self._lib.qs_genome_destroy(self._h)
            self._h = C.c_void_p()

    def download(self, offset: int, count: int) -> bytes:
        out = np.empty(count, dtype=np.uint8)
        self._check(self._lib.qs_genome_download(self._h, offset, out.ctypes.data, count))
        return out.tobytes()

    def contigs(self) -> np.ndarray:
        n = C.c_uint32()
        st = np.zeros(257, dtype=np.uint64)
        self._check(self._lib.qs_genome_contigs(self._h, st.ctypes.data, st.size, C.byref(n)))
        return st[: n.value + 1]

    # -- dictionary ------------------------------------------------------------------------
    def build_dict(self, k: int = 30, slots: int = 0, ctrl_block: int = 0, with_qgc: bool = True) -> dict:
        info = DictInfo()
        self._check(self._lib.qs_dict_build(self._h, k, slots, ctrl_block, int(with_qgc), C.byref(info)))
        self.info = info.as_dict()
        return self.info

    def write_dict(self, prefix, threads: int = 8):
        self._check(self._lib.qs_dict_write(self._h, os.fsencode(str(prefix)), threads))

    def download_dict(self):
        slots, n = self.info["slots"], self.info["n_unique"]
        keys, nxt, qgc = np.empty(slots, np.uint64), np.empty(slots, np.uint32), np.empty(n, np.uint16)
        self._check(self._lib.qs_dict_download(self._h, keys.ctypes.data, nxt.ctypes.data, qgc.ctypes.data))
        return keys, nxt, qgc

    def free_dict(self):
        self._lib.qs_dict_free(self._h)

    # -- reads -----------------------------------------------------------------------------
    def reads_into(self, dev_ptr: int, seed: int, first: int, n: int, length: int = 150, err_ppm: int = 2000, fmt: int = FRAMED,
                   lens: np.ndarray | None = None, offsets: np.ndarray | None = None):
        lp = op = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            lp, op = lens.ctypes.data, offsets.ctypes.data
        self._check(self._lib.qs_reads_generate(self._h, seed, first, n, length, err_ppm, fmt, lp, op, dev_ptr))

    def reads_bytes(self, seed: int, first: int, n: int, length: int = 150, err_ppm: int = 2000, fmt: int = FASTQ,
                    lens: np.ndarray | None = None) -> bytes:
        """Small cases: the records as host bytes."""
        total, offsets = layout(n, length, fmt, lens)
        buf = DeviceBuffer(self.device, total + 64)
        try:
            self.reads_into(buf.ptr, seed, first, n, length, err_ppm, fmt, lens, offsets)
            return buf.to_host(total).tobytes()
        finally:
            buf.free()

    def reads_to_file(self, path, seed: int, first: int, n: int, length: int = 150, err_ppm: int = 2000, fmt: int = FASTQ,
                      lens: np.ndarray | None = None, piece_reads: int = 4 << 20):
        """Write records [first, first + n) to a file, a piece at a time through the device."""
        with open(path, "wb") as f:
            at = 0
            while at < n:
                m = min(piece_reads, n - at)
                sub = None if lens is None else lens[at: at + m]
                f.write(self.reads_bytes(seed, first + at, m, length, err_ppm, fmt, sub))
                at += m


def layout(n: int, length: int, fmt: int, lens: np.ndarray | None):
    """(total bytes, per-record offsets or None) of n records."""
    if lens is None:
        return n * record_bytes(length, fmt), None
    lens = np.asarray(lens, dtype=np.uint64)
    rec = lens + 1 if fmt == FRAMED else 13 + lens + 1 if fmt == FASTA else 13 + 2 * (lens + 1) + 2
    offsets = np.zeros(n, dtype=np.uint64)
    np.cumsum(rec[:-1], out=offsets[1:])
    return int(rec.sum()), offsets
