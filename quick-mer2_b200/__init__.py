"""quickmer2_b200 -- Python face of the B200 `quicKmer2 count` path.

The product is the C-ABI shared library ``libquickmer2_b200.so`` (CUDA kernels for sm_100a
plus the host C code) declared in ``include/quickmer2_b200.h`` and ``include/qk_host.h``;
this module is a thin ctypes binding used by the tests, ``bench.py`` and
``__graft_entry__.py``.  It mirrors the reference's only public interface, the command
``quicKmer2 count [-t N] ref.fa sample.fast[a/q] Out_prefix`` (QuicKmer.c:298-302,
main_count QuicKmer.c:304-545), as :func:`count`.

There is no CPU fallback: if the library is missing, or no CUDA device is present,
everything here raises.

The directory name ``quick-mer2_b200`` is not an importable identifier; load the package
with ``tests/conftest.py::load_package`` / ``bench.py`` (importlib, module name
``quickmer2_b200``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libquickmer2_b200.so"
CLI_PATH = PKG_DIR / "bin" / "quicKmer2_b200"
SYNTH_PATH = PKG_DIR / "bin" / "qk_synth"

GC_BINS = 401
MAX_LINE_BYTES = 99999

ERRORS = {1: "QK_ERR_CUDA", 2: "QK_ERR_ARG", 3: "QK_ERR_NOMEM", 4: "QK_ERR_STATE", 5: "QK_ERR_FORMAT", 6: "QK_ERR_IO"}


class QkError(RuntimeError):
    def __init__(self, code: int, message: str = ""):
        self.code = code
        super().__init__(f"{ERRORS.get(code, code)}: {message}" if message else str(ERRORS.get(code, code)))


class TableDesc(C.Structure):
    """struct qk_table_desc (include/quickmer2_b200.h)."""

    _fields_ = [
        ("n_kmers", C.c_uint64),
        ("n_buckets", C.c_uint64),
        ("stash_slots", C.c_uint64),
        ("stash_used", C.c_uint64),
        ("table_bytes", C.c_uint64),
        ("stash_bytes", C.c_uint64),
        ("k", C.c_uint32),
        ("bucket_bits", C.c_uint32),
        ("ord_bits", C.c_uint32),
        ("rem_bits", C.c_uint32),
        ("skipped_keys", C.c_uint64),
        ("ext_bytes", C.c_uint64),
        ("cont_bytes", C.c_uint64),
        ("has_ext", C.c_uint32),
        ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


class QmHeader(C.Structure):
    _fields_ = [("k", C.c_uint8), ("hash_size", C.c_uint64), ("first_idx", C.c_uint64)]


class FramerStats(C.Structure):
    _fields_ = [
        ("lines", C.c_uint64),
        ("bases", C.c_uint64),
        ("raw_bytes", C.c_uint64),
        ("long_lines", C.c_uint64),
        ("unterminated", C.c_uint64),
        ("fastq", C.c_int),
        ("sink_bytes", C.c_uint64),
    ]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/*.h declares
_P = C.c_void_p
_U64P = C.POINTER(C.c_uint64)
SIGNATURES = {
    # quickmer2_b200.h
    "qk_version": (C.c_char_p, []),
    "qk_device_count": (C.c_int, []),
    "qk_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_uint32, C.c_size_t]),
    "qk_ctx_destroy": (None, [_P]),
    "qk_last_error": (C.c_char_p, [_P]),
    "qk_ctx_info": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]),
    "qk_dict_begin": (C.c_int, [_P, C.c_uint8, C.c_uint64, C.c_uint64]),
    "qk_dict_upload_keys": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "qk_dict_upload_chain": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "qk_dict_upload_from_slot": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64]),
    "qk_dict_build": (C.c_int, [_P, _U64P]),
    "qk_table_geometry": (C.c_int, [C.c_uint64, C.c_uint32, C.POINTER(TableDesc)]),
    "qk_dict_describe": (C.c_int, [_P, C.POINTER(TableDesc)]),
    "qk_dict_adopt": (C.c_int, [_P, C.POINTER(TableDesc)]),
    "qk_dict_device_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "qk_dict_ext_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "qk_slot_host_buffer": (_P, [_P, C.c_uint32]),
    "qk_submit": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t, _P, C.c_uint32]),
    "qk_submit_device": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t]),
    "qk_raw_begin": (C.c_int, [_P, C.c_int, C.c_int]),
    "qk_raw_begin_state": (C.c_int, [_P, C.c_int, C.c_uint32]),
    "qk_raw_state": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "qk_submit_raw": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t]),
    "qk_raw_stats": (C.c_int, [_P, _U64P, _U64P, _U64P]),
    "qk_host_is_pinned": (C.c_int, [_P]),
    "qk_wait_slot": (C.c_int, [_P, C.c_uint32]),
    "qk_slot_ready": (C.c_int, [_P, C.c_uint32]),
    "qk_sync": (C.c_int, [_P]),
    "qk_stats": (C.c_int, [_P, _U64P, _U64P, _U64P]),
    "qk_stats_ext": (C.c_int, [_P, _U64P]),
    "qk_stats_probes": (C.c_int, [_P, _U64P, _U64P]),
    "qk_counters_device_ptr": (C.c_int, [_P, C.POINTER(_P), _U64P]),
    "qk_reset_counters": (C.c_int, [_P]),
    "qk_reset_counters_async": (C.c_int, [_P]),
    "qk_add_depth": (C.c_int, [_P, C.c_uint64, C.c_uint32]),
    "qk_submit_packed": (C.c_int, [_P, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32]),
    "qk_counters_select": (C.c_int, [_P, C.c_uint32]),
    "qk_slot_stream": (_P, [_P, C.c_uint32]),
    "qk_counters_download": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "qk_finish": (C.c_int, [_P, _P, C.c_uint64]),
    "qk_finish_pieces": (C.c_int, [_P, _P, _P]),
    "qk_finish_async": (C.c_int, [_P, _P, C.c_uint64]),
    "qk_finish_wait": (C.c_int, [_P]),
    "qk_gc_curve": (C.c_int, [_P, _P, C.c_uint64, _P, _P, _P]),
    "qk_gc_begin": (C.c_int, [_P]),
    "qk_gc_from_slot": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_uint64]),
    "qk_gc_end": (C.c_int, [_P, _P, _P, _P, _U64P]),
    "qk_est_begin": (C.c_int, [_P, C.c_uint64]),
    "qk_est_upload_from_slot": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64]),
    "qk_est_windows": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P]),
    "qk_est_end": (C.c_int, [_P]),
    "qk_gc_curve_file": (C.c_int, [_P, C.c_char_p, C.c_uint64, _P, _P, _P, _U64P, _U64P]),
    "qk_timing": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), _U64P]),
    "qk_span_begin": (C.c_int, [_P]),
    "qk_span_end": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "qk_bench_gather": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_double)]),
    "qk_bench_h2d": (C.c_int, [_P, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "qk_multi_create": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.c_size_t]),
    "qk_multi_destroy": (None, [_P]),
    "qk_multi_last_error": (C.c_char_p, [_P]),
    "qk_multi_size": (C.c_uint32, [_P]),
    "qk_multi_ctx": (_P, [_P, C.c_uint32]),
    "qk_multi_replicate": (C.c_int, [_P]),
    "qk_multi_reduce": (C.c_int, [_P]),
    # qk_host.h
    "qk_qm_read_header": (C.c_int, [C.c_char_p, C.POINTER(QmHeader)]),
    "qk_qm_load": (C.c_int, [_P, C.c_char_p, C.POINTER(QmHeader), _U64P]),
    "qk_framer_open": (_P, [C.c_char_p]),
    "qk_framer_open_fd": (_P, [C.c_int, C.c_int]),
    "qk_framer_open_mem": (_P, [_P, C.c_size_t, C.c_int]),
    "qk_framer_next": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t), _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "qk_framer_get_stats": (None, [_P, C.POINTER(FramerStats)]),
    "qk_framer_close": (None, [_P]),
    "qk_stream_open": (_P, [C.c_char_p]),
    "qk_stream_open_fd": (_P, [C.c_int, C.c_int]),
    "qk_stream_read": (C.c_ssize_t, [_P, _P, C.c_size_t]),
    "qk_stream_is_gzip": (C.c_int, [_P]),
    "qk_stream_is_bam": (C.c_int, [_P]),
    "qk_stream_seekable": (C.c_int, [_P]),
    "qk_stream_close": (None, [_P]),
    "qk_count_raw_stream": (C.c_int, [_P, _P, C.POINTER(FramerStats)]),
    "qk_write_bin": (C.c_int, [C.c_char_p, _P, C.c_uint64]),
    "qk_write_bin_from_device": (C.c_int, [_P, C.c_char_p]),
    "qk_write_gc_txt": (C.c_int, [C.c_char_p, _P, _P, _P, C.POINTER(C.c_double)]),
    "qk_count_file": (C.c_int, [_P, C.c_char_p, C.POINTER(FramerStats)]),
    "qk_count_framer": (C.c_int, [_P, _P, C.POINTER(FramerStats)]),
    "qk_count_raw_mem": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.POINTER(FramerStats)]),
    "qk_count_raw_fd": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(FramerStats)]),
    "qk_count_raw_file": (C.c_int, [_P, C.c_char_p, C.POINTER(FramerStats)]),
    "qk_frame_mem_mt": (C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FramerStats)]),
    "qk_bench_framer": (C.c_int, [_P, C.c_size_t, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qk_bench_host_memory": (C.c_int, [C.c_size_t, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qk_count_mem_mt": (C.c_int, [C.POINTER(_P), C.c_uint32, _P, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(FramerStats)]),
    "qk_count_file_mt": (C.c_int, [C.POINTER(_P), C.c_uint32, C.c_char_p, C.c_uint32, C.POINTER(FramerStats)]),
    "qk_shard_bounds": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _U64P, _U64P]),
    "qk_fastq_state_guess": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "qk_count_raw_range": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(FramerStats),
                                     C.POINTER(C.c_uint32)]),
    "qk_count_raw_file_mt": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.POINTER(FramerStats)]),
    "qk_count_raw_range_mt": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32,
                                        C.POINTER(FramerStats), C.POINTER(C.c_uint32)]),
    "qk_count_file_multi": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.POINTER(FramerStats)]),
    "qk_count_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "qk_est_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "qk_est_reduce": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, C.c_double, _P, _P, C.c_uint64, C.POINTER(_P), C.POINTER(_P), _U64P]),
}

_lib = None


def build(verbose: bool = False) -> None:
    """Compile the library and the host binaries in-tree (nvcc, sm_100a)."""
    res = subprocess.run(["make", "-C", str(PKG_DIR), "all"], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode:
        raise RuntimeError("building libquickmer2_b200.so failed")


def lib() -> C.CDLL:
    """The C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `make -C {PKG_DIR}` (or __graft_entry__.build()). "
                "There is no Python/CPU fallback for the count path."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU: struct qk_ctx.  Mirrors the life cycle of main_count (Q.c:304-545)."""

    def __init__(self, device: int = 0, n_slots: int = 4, chunk_capacity: int = 32 << 20):
        self._h = C.c_void_p()
        self._lib = lib()
        rc = self._lib.qk_ctx_create(C.byref(self._h), device, n_slots, chunk_capacity)
        if rc:
            msg = self._lib.qk_last_error(self._h).decode() if self._h else "no CUDA device (no CPU fallback)"
            if self._h:
                self._lib.qk_ctx_destroy(self._h)
                self._h = C.c_void_p()
            raise QkError(rc, msg)
        self.device = device
        self.n_slots = n_slots
        cap = C.c_size_t()
        self._lib.qk_ctx_info(self._h, None, C.byref(cap))
        self.chunk_capacity = cap.value
        self.n_kmers = 0

    # -- plumbing -------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc:
            raise QkError(rc, self._lib.qk_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.qk_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- dictionary -----------------------------------------------------------------------
    def load_dictionary(self, qm_path) -> int:
        """Stream a QM11 file to the device and build the table (Q.c:345-359, 483)."""
        n = C.c_uint64()
        hdr = QmHeader()
        self._check(self._lib.qk_qm_load(self._h, os.fsencode(str(qm_path)), C.byref(hdr), C.byref(n)))
        self.n_kmers = n.value
        self.header = {"k": hdr.k, "hash_size": hdr.hash_size, "first_idx": hdr.first_idx}
        return self.n_kmers

    def load_dictionary_arrays(self, k: int, keys: np.ndarray, nxt: np.ndarray, first_idx: int) -> int:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        nxt = np.ascontiguousarray(nxt, dtype=np.uint32)
        assert keys.shape == nxt.shape
        self._check(self._lib.qk_dict_begin(self._h, k, keys.size, first_idx))
        self._check(self._lib.qk_dict_upload_keys(self._h, 0, _np_ptr(keys), keys.size))
        self._check(self._lib.qk_dict_upload_chain(self._h, 0, _np_ptr(nxt), nxt.size))
        n = C.c_uint64()
        self._check(self._lib.qk_dict_build(self._h, C.byref(n)))
        self.n_kmers = n.value
        return self.n_kmers

    def table_desc(self) -> TableDesc:
        d = TableDesc()
        self._check(self._lib.qk_dict_describe(self._h, C.byref(d)))
        return d

    def adopt(self, desc: TableDesc):
        self._check(self._lib.qk_dict_adopt(self._h, C.byref(desc)))
        self.n_kmers = int(desc.n_kmers)

    def table_device_ptrs(self):
        t, s = C.c_void_p(), C.c_void_p()
        self._check(self._lib.qk_dict_device_ptrs(self._h, C.byref(t), C.byref(s)))
        return t.value, s.value

    def table_images(self):
        """(device pointer, bytes) of everything a replica needs: table, stash, extension arrays."""
        d = self.table_desc()
        t, s = self.table_device_ptrs()
        out = [(t, int(d.table_bytes)), (s, int(d.stash_bytes))]
        if d.has_ext:
            a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
            self._check(self._lib.qk_dict_ext_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
            out += [(a.value, int(d.ext_bytes))]      # one array: last / first base and continuation bits per 16 ordinals
        return out

    def counters_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.qk_counters_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def counters(self) -> np.ndarray:
        """Raw uint32 counters by ordinal (before the 16-bit wrap of :meth:`finish`)."""
        out = np.empty(self.n_kmers, dtype=np.uint32)
        self._check(self._lib.qk_counters_download(self._h, 0, _np_ptr(out), out.size))
        return out

    # -- counting -------------------------------------------------------------------------
    def count_file(self, reads_path, host_framer: bool = False, threads: int = 0) -> dict:
        """Count a FASTA/FASTQ file (Q.c:393-479).  Default: raw pieces cut at line ends go to
        the device, which frames them (`threads` readers fill the pinned buffers);
        host_framer=True frames on the host instead."""
        st = FramerStats()
        if host_framer:
            rc = self._lib.qk_count_file(self._h, os.fsencode(str(reads_path)), C.byref(st))
        else:
            rc = self._lib.qk_count_raw_file_mt(self._h, os.fsencode(str(reads_path)), threads, C.byref(st))
        if rc == 6:
            raise QkError(rc, f"cannot read {reads_path}")
        self._check(rc)
        return st.as_dict()

    def count_raw(self, data: bytes, seekable: bool = True, host_framer: bool = False) -> dict:
        """Count an in-memory FASTA/FASTQ byte string (device framing unless host_framer)."""
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)
        return self.count_mem(buf.ctypes.data, buf.size, seekable, host_framer)

    def submit_chunk(self, chunk: bytes, slot: int = 0, n_lines: int = 0):
        """Count an already framed chunk (sequence lines only, each ending in a newline)."""
        assert len(chunk) <= self.chunk_capacity
        self._check(self._lib.qk_wait_slot(self._h, slot))
        host = self._lib.qk_slot_host_buffer(self._h, slot)
        C.memmove(host, chunk, len(chunk))
        self._check(self._lib.qk_submit(self._h, slot, host, len(chunk), None, n_lines))

    def submit_device(self, dev_ptr: int, n_bytes: int, slot: int = 0):
        self._check(self._lib.qk_submit_device(self._h, slot, dev_ptr, n_bytes))

    def sync(self):
        self._check(self._lib.qk_sync(self._h))

    def reset(self):
        self._check(self._lib.qk_reset_counters(self._h))

    def add_depth(self, key: int, n: int):
        """n more occurrences of a canonical key (the reference's -t N batch padding: key 0, Q.c:458-466)."""
        self._check(self._lib.qk_add_depth(self._h, key, n))

    def select_counters(self, which: int):
        """Use counter buffer 0 or 1 for what is issued from now on."""
        self._check(self._lib.qk_counters_select(self._h, which))

    def reset_async(self):
        """Stream-ordered reset (no host sync) for jobs run back to back."""
        self._check(self._lib.qk_reset_counters_async(self._h))

    def slot_stream(self, slot: int = 0) -> int:
        """cudaStream_t of a slot as an integer (e.g. for torch.cuda.ExternalStream)."""
        return int(self._lib.qk_slot_stream(self._h, slot) or 0)

    def stats(self) -> dict:
        t, h, l = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.qk_stats(self._h, C.byref(t), C.byref(h), C.byref(l)))
        e = C.c_uint64()
        self._check(self._lib.qk_stats_ext(self._h, C.byref(e)))
        p, w = C.c_uint64(), C.c_uint64()
        self._check(self._lib.qk_stats_probes(self._h, C.byref(p), C.byref(w)))
        return {"total_kmers": t.value, "hits": h.value, "lines": l.value, "ext_verified": e.value,
                "bucket_probes": p.value, "walks": w.value}

    def timing(self) -> dict:
        k, h, n = C.c_double(), C.c_double(), C.c_uint64()
        self._check(self._lib.qk_timing(self._h, C.byref(k), C.byref(h), C.byref(n)))
        return {"kernel_ms": k.value, "h2d_ms": h.value, "launches": n.value}

    # -- results --------------------------------------------------------------------------
    def finish(self, out: np.ndarray | None = None) -> np.ndarray:
        """Depths in .bin order, uint16 with the reference's wrap (Q.c:498-518).  `out` may be
        a caller-owned (ideally pinned) uint16 array of n_kmers entries."""
        if out is None:
            out = np.empty(self.n_kmers, dtype=np.uint16)
        assert out.dtype == np.uint16 and out.size == self.n_kmers and out.flags.c_contiguous
        self._check(self._lib.qk_finish(self._h, _np_ptr(out), self.n_kmers))
        return out

    def finish_async(self, out: np.ndarray):
        """Enqueue the download of the selected counter buffer into page-locked `out`; finish_wait() completes it."""
        assert out.dtype == np.uint16 and out.size == self.n_kmers and out.flags.c_contiguous
        self._check(self._lib.qk_finish_async(self._h, _np_ptr(out), self.n_kmers))

    def finish_wait(self):
        self._check(self._lib.qk_finish_wait(self._h))

    def write_bin(self, path):
        """Depths straight from the device to a .bin file, piece by piece."""
        rc = self._lib.qk_write_bin_from_device(self._h, os.fsencode(str(path)))
        if rc == 6:
            raise QkError(rc, f"cannot write {path}")
        self._check(rc)

    def gc_curve(self, qgc: np.ndarray):
        qgc = np.ascontiguousarray(qgc, dtype=np.uint16)
        assert qgc.size == self.n_kmers
        s = np.zeros(GC_BINS, dtype=np.uint64)
        q = np.zeros(GC_BINS, dtype=np.int64)
        c = np.zeros(GC_BINS, dtype=np.uint64)
        self._check(self._lib.qk_gc_curve(self._h, _np_ptr(qgc), qgc.size, _np_ptr(s), _np_ptr(q), _np_ptr(c)))
        return s, q, c

    def span_begin(self):
        self._check(self._lib.qk_span_begin(self._h))

    def span_end(self) -> float:
        ms = C.c_double()
        self._check(self._lib.qk_span_end(self._h, C.byref(ms)))
        return ms.value

    def submit_host(self, host_ptr: int, n_bytes: int, slot: int = 0, n_lines: int = 0):
        """Count a framed chunk straight from (ideally pinned) host memory: async H2D + kernel."""
        self._check(self._lib.qk_submit(self._h, slot, host_ptr, n_bytes, None, n_lines))

    def submit_packed(self, packed: np.ndarray, n_positions: int, slot: int = 0, n_lines: int = 0):
        """Count a packed chunk (24 bytes per 64 positions: qk_submit_packed) held in a uint8 array; waits for the copy."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if packed.size < n_positions // 64 * 24:
            raise ValueError("packed chunk shorter than n_positions / 64 * 24 bytes")
        self._check(self._lib.qk_submit_packed(self._h, slot, _np_ptr(packed), n_positions, n_lines))
        self._check(self._lib.qk_wait_slot(self._h, slot))

    def count_mem(self, host_ptr: int, n_bytes: int, seekable: bool = True, host_framer: bool = False) -> dict:
        """Count raw FASTA/FASTQ bytes at a host address (the whole reads stream).  Pinned
        memory is DMA'd from directly; the device does the record framing unless host_framer."""
        st = FramerStats()
        if not host_framer:
            self._check(self._lib.qk_count_raw_mem(self._h, host_ptr, n_bytes, int(seekable), C.byref(st)))
            return st.as_dict()
        fr = self._lib.qk_framer_open_mem(host_ptr, n_bytes, int(seekable))
        try:
            self._check(self._lib.qk_count_framer(self._h, fr, C.byref(st)))
        finally:
            self._lib.qk_framer_close(fr)
        return st.as_dict()

    def count_mem_mt(self, host_ptr: int, n_bytes: int, seekable: bool = True, threads: int = 0, peers=()) -> dict:
        """Count raw FASTA/FASTQ bytes at a host address with the framing done by `threads` host workers
        (only sequence lines cross the host link); `peers` = further contexts (other GPUs, same
        dictionary and slot geometry) that take chunks from the same queue."""
        st = FramerStats()
        hs = (C.c_void_p * (1 + len(peers)))(self._h, *[p._h for p in peers])
        self._check(self._lib.qk_count_mem_mt(hs, 1 + len(peers), host_ptr, n_bytes, int(seekable), threads, C.byref(st)))
        return st.as_dict()

    def count_file_mt(self, reads_path, threads: int = 0, peers=()) -> dict:
        """The same for a reads file (mapped; pipes and gzip fall back to the sequential stream path)."""
        st = FramerStats()
        hs = (C.c_void_p * (1 + len(peers)))(self._h, *[p._h for p in peers])
        rc = self._lib.qk_count_file_mt(hs, 1 + len(peers), os.fsencode(str(reads_path)), threads, C.byref(st))
        if rc == 6:
            raise QkError(rc, f"cannot read {reads_path}")
        self._check(rc)
        return st.as_dict()

    def count_range(self, reads_path, begin: int, end: int, fastq: bool, line_state: int):
        """Count bytes [begin, end) of a reads file starting in `line_state`; returns (stats, final state)."""
        st, fin = FramerStats(), C.c_uint32()
        rc = self._lib.qk_count_raw_range(self._h, os.fsencode(str(reads_path)), begin, end, int(fastq), line_state,
                                          C.byref(st), C.byref(fin))
        if rc == 6:
            raise QkError(rc, f"cannot read {reads_path}")
        self._check(rc)
        return st.as_dict(), fin.value

    def submit_raw(self, host_ptr: int, n_bytes: int, slot: int = 0):
        self._check(self._lib.qk_submit_raw(self._h, slot, host_ptr, n_bytes))

    def raw_begin(self, fastq: bool, skip_first_line: bool):
        self._check(self._lib.qk_raw_begin(self._h, int(fastq), int(skip_first_line)))

    # -- measurement ----------------------------------------------------------------------
    def bench_gather(self, table_bytes: int, gran: int = 32, loads_in_flight: int = 4, n_gathers: int = 1 << 30) -> float:
        g = C.c_double()
        self._check(self._lib.qk_bench_gather(self._h, table_bytes, gran, loads_in_flight, n_gathers, C.byref(g)))
        return g.value

    def bench_h2d(self, n_bytes: int, repeats: int = 8) -> float:
        g = C.c_double()
        self._check(self._lib.qk_bench_h2d(self._h, n_bytes, repeats, C.byref(g)))
        return g.value


def write_gc_txt(path, s: np.ndarray, q: np.ndarray, c: np.ndarray) -> float:
    mean = C.c_double()
    rc = lib().qk_write_gc_txt(os.fsencode(str(path)), _np_ptr(s), _np_ptr(q), _np_ptr(c), C.byref(mean))
    if rc:
        raise QkError(rc, f"cannot write {path}")
    return mean.value


def read_stream(path=None, fd=None, seekable: bool = True, piece: int = 1 << 16):
    """All bytes of a plain or gzip stream through qk_stream_* (host only); returns (bytes, is_gzip)."""
    L = lib()
    s = L.qk_stream_open(os.fsencode(str(path))) if path is not None else L.qk_stream_open_fd(fd, int(seekable))
    if not s:
        raise QkError(6, f"cannot open {path if path is not None else fd}")
    out, buf = [], np.empty(piece, dtype=np.uint8)
    try:
        gz = bool(L.qk_stream_is_gzip(s))
        while True:
            n = L.qk_stream_read(s, _np_ptr(buf), buf.size)
            if n < 0:
                raise QkError(6, "stream read failed (I/O error or corrupt gzip data)")
            if n == 0:
                break
            out.append(buf[:n].tobytes())
    finally:
        L.qk_stream_close(s)
    return b"".join(out), gz


def frame(data: bytes, seekable: bool = True, chunk_capacity: int = 1 << 20, with_offsets: bool = False):
    """Host framer only (no GPU): FASTA/FASTQ bytes -> list of framed chunks, stats."""
    L = lib()
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
    fr = L.qk_framer_open_mem(_np_ptr(buf), buf.size, int(seekable))
    chunks, offsets = [], []
    dst = np.empty(chunk_capacity, dtype=np.uint8)
    off = np.empty(chunk_capacity // 2 + 2, dtype=np.uint32)
    n, nl = C.c_size_t(), C.c_uint32()
    try:
        while True:
            r = L.qk_framer_next(fr, _np_ptr(dst), dst.size, C.byref(n), _np_ptr(off) if with_offsets else None,
                                 off.size, C.byref(nl))
            if r < 0:
                raise QkError(-r, "framer")
            if r == 0:
                break
            chunks.append(dst[: n.value].tobytes())
            if with_offsets:
                offsets.append(off[: nl.value + 1].copy())
        st = FramerStats()
        L.qk_framer_get_stats(fr, C.byref(st))
    finally:
        L.qk_framer_close(fr)
    return (chunks, offsets, st.as_dict()) if with_offsets else (chunks, st.as_dict())


class ChunkSink(C.Structure):
    """struct qk_chunk_sink (include/qk_host.h)."""
    BUFFER = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32)
    READY = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32)
    SUBMIT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_size_t, C.c_uint32)
    _fields_ = [("user", C.c_void_p), ("n_ctx", C.c_uint32), ("n_slots", C.c_uint32), ("cap", C.c_size_t),
                ("buffer", BUFFER), ("ready", READY), ("wait", READY), ("submit", SUBMIT), ("packed", C.c_int)]


def unpack_chunk(packed: bytes) -> tuple[np.ndarray, np.ndarray]:
    """A packed chunk (struct qk_chunk_sink: 24 bytes per 64 positions) -> (2-bit code, reset flag) per position."""
    g = np.frombuffer(packed, dtype=np.uint8).reshape(-1, 24)
    words = g[:, :16].copy().view("<u4")                                    # (groups, 4): 16 positions each, the first in the top pair
    shifts = (2 * (15 - np.arange(16))).astype(np.uint32)
    codes = ((words[:, :, None] >> shifts[None, None, :]) & 3).astype(np.uint8).reshape(-1)
    flags = np.unpackbits(g[:, 16:24].copy(), axis=1, bitorder="little").reshape(-1)
    return codes, flags


def frame_mt(data: bytes, seekable: bool = True, threads: int = 4, n_ctx: int = 1, n_slots: int = 3, cap: int = 1 << 20,
             busy_every: int = 0, packed: bool = False):
    """Host-only: the multi-threaded framer (qk_frame_mem_mt) into Python-owned buffers.  Returns
    (chunks in stream order, per-chunk (consumer, lines), stats).  busy_every > 0 makes ready()
    answer "busy" now and then, as a GPU whose copy is still in flight would.  packed: the chunks are
    packed ones (cap and their sizes in positions; see unpack_chunk)."""
    import threading
    L = lib()
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
    bufs = [[np.zeros(cap, dtype=np.uint8) for _ in range(n_slots)] for _ in range(n_ctx)]
    got, lock, calls = {}, threading.Lock(), [0]

    def ready(_u, c, s):
        with lock:
            calls[0] += 1
            return 0 if busy_every and calls[0] % busy_every == 0 else 1

    def submit(_u, c, s, seq, n_bytes, n_lines):
        with lock:
            got[seq] = (bufs[c][s][:n_bytes // 64 * 24 if packed else n_bytes].tobytes(), c, n_lines)
        return 0

    sink = ChunkSink(None, n_ctx, n_slots, cap, ChunkSink.BUFFER(lambda _u, c, s: bufs[c][s].ctypes.data), ChunkSink.READY(ready),
                     ChunkSink.READY(lambda _u, c, s: 0), ChunkSink.SUBMIT(submit), int(packed))
    st = FramerStats()
    rc = L.qk_frame_mem_mt(C.byref(sink), _np_ptr(buf), buf.size, int(seekable), threads, C.byref(st))
    if rc:
        raise QkError(rc, "qk_frame_mem_mt")
    order = sorted(got)
    assert order == list(range(len(order))), order
    return [got[i][0] for i in order], [(got[i][1], got[i][2]) for i in order], st.as_dict()


def bench_framer(host_ptr: int, n_bytes: int, threads: int = 0, repeats: int = 3):
    """GB/s (raw in, framed out) of the multi-threaded host framer alone."""
    a, b = C.c_double(), C.c_double()
    rc = lib().qk_bench_framer(host_ptr, n_bytes, threads, repeats, C.byref(a), C.byref(b))
    if rc:
        raise QkError(rc, "qk_bench_framer")
    return a.value, b.value


def bench_host_memory(bytes_per_thread: int = 256 << 20, threads: int = 8):
    a, b = C.c_double(), C.c_double()
    rc = lib().qk_bench_host_memory(bytes_per_thread, threads, C.byref(a), C.byref(b))
    if rc:
        raise QkError(rc, "qk_bench_host_memory")
    return a.value, b.value


def count(ref_prefix, reads_path, out_prefix, threads: int = 0, device: int = 0, host_framer: bool = False,
          n_slots: int = 4, chunk_capacity: int = 32 << 20) -> dict:
    """``quicKmer2 count [-t N] ref.fa reads Out_prefix`` on the GPU, in-process.

    Same files as the reference: reads ``<ref_prefix>.qm`` (and ``.qgc`` if present), writes
    ``<out_prefix>.bin`` (and ``.txt``).  Returns the counting statistics.
    """
    with Context(device=device, n_slots=n_slots, chunk_capacity=chunk_capacity) as ctx:
        n = ctx.load_dictionary(f"{ref_prefix}.qm")
        st = ctx.count_file(reads_path, host_framer=host_framer)
        st.update(ctx.stats())
        if threads & 0xFF:      # the reference's batch padding with -t N (Q.c:458-466; uint8_t thread_count, Q.c:306)
            ctx.add_depth(0, 4096 - st["total_kmers"] % 4096)
        counts = ctx.finish()
        counts.tofile(f"{out_prefix}.bin")
        qgc_path = Path(f"{ref_prefix}.qgc")
        if qgc_path.exists():
            qgc = np.zeros(n, dtype=np.uint16)
            raw = np.fromfile(qgc_path, dtype=np.uint16, count=n)
            qgc[: raw.size] = raw
            s, q, c = ctx.gc_curve(qgc)
            st["mean_depth"] = write_gc_txt(f"{out_prefix}.txt", s, q, c)
        st["n_kmers"] = n
        st.update(ctx.timing())
    return st


def run_cli(args, **kw) -> subprocess.CompletedProcess:
    """Run the C command ``quicKmer2_b200`` (the drop-in for ``quicKmer2``)."""
    return subprocess.run([str(CLI_PATH), *map(str, args)], capture_output=True, text=True, **kw)
