import sys, time, os
sys.path.insert(0, "/root/repo/tests")
from conftest import load_package
qk = load_package()
qm = sys.argv[1]
for thr in (1, 4, 8, 12):
    os.environ["QK_READER_THREADS"] = str(thr)
    t0 = time.perf_counter()
    ctx = qk.Context(n_slots=12, chunk_capacity=64 << 20)
    t1 = time.perf_counter()
    ctx.load_dictionary(qm)
    t2 = time.perf_counter()
    ctx.close()
    print(f"threads {thr}: ctx {t1-t0:.3f} s, load+build {t2-t1:.3f} s", flush=True)
