"""Shared test plumbing.

* ``gpu`` marker: tests that need a B200 (run with ``-m gpu`` under gpurun).
* ``load_package()``: import ``quick-mer2_b200/`` (not an identifier) as ``quickmer2_b200``.
* builds the C-ABI library, the host binaries and the oracle once per session.

The oracle (``oracle/``) is imported here and nowhere in the product.
"""
import importlib.util
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "quick-mer2_b200"
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu under gpurun")


def load_package():
    if "quickmer2_b200" in sys.modules:
        return sys.modules["quickmer2_b200"]
    spec = importlib.util.spec_from_file_location("quickmer2_b200", PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["quickmer2_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_dist():
    """quick-mer2_b200/dist.py as quickmer2_b200.dist."""
    import importlib
    load_package()
    return importlib.import_module("quickmer2_b200.dist")


@pytest.fixture(scope="session")
def built():
    """Build everything that can be built on this machine (nvcc cross-compiles without a GPU)."""
    for d in (PKG_DIR, ROOT / "oracle"):
        res = subprocess.run(["make", "-s", "-C", str(d), "all"], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
    return True


@pytest.fixture(scope="session")
def qk(built):
    return load_package()


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_binding
    return oracle_binding.Oracle()


@pytest.fixture(scope="session")
def ref_binary():
    """The compiled, unmodified reference (present where `make -C oracle ref` could run)."""
    p = ROOT / "oracle" / "_ref" / "quicKmer2"
    return p if p.exists() else None


def golden_cases():
    return sorted(p.parent.name for p in GOLDEN.glob("*/meta.json"))


def golden_meta(case):
    return json.loads((GOLDEN / case / "meta.json").read_text())


@pytest.fixture(scope="session")
def synth(built):
    exe = PKG_DIR / "bin" / "qk_synth"

    def run(*args, cwd=None):
        res = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, cwd=cwd)
        assert res.returncode == 0, res.stderr
        return res.stdout
    return run


@pytest.fixture(scope="session")
def gpu_ctx(qk):
    """One context for the whole GPU session (creating one costs pinned allocations)."""
    ctx = qk.Context(device=0, n_slots=3, chunk_capacity=8 << 20)
    yield ctx
    ctx.close()


def weird_stream(rng, seq, fastq_like, n_lines):
    """Lines in arbitrary order -- not a valid FASTA/FASTQ -- so that the reference's line state
    machine (Q.c:397-398, 451-455) is driven through every phase: '>' where a read is expected,
    quality lines that start with '>' or '@', empty lines, N runs, CR."""
    out = []
    for _ in range(n_lines):
        kind = rng.integers(0, 10)
        a = int(rng.integers(0, len(seq) - 400))
        body = seq[a:a + int(rng.integers(0, 300))]
        if kind == 0:
            out.append(">" + body[:20])
        elif kind == 1:
            out.append("@" + body[:30])
        elif kind == 2:
            out.append("+")
        elif kind == 3:
            out.append("")
        elif kind == 4:
            out.append(">" * int(rng.integers(1, 4)) + "III@@>>")
        elif kind == 5:
            out.append(body[:50] + "N" * int(rng.integers(1, 5)) + body[50:] + "\r")
        else:
            out.append(body)
    first = "@first" if fastq_like else rng.choice([">first", seq[5:160], ""])
    return "\n".join([first] + out) + "\n"
