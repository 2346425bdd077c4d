"""A/B of the asynchronous position-map prefetch (PB2_MAP_ASYNC) on the extra workloads, same box, alternating"""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyoomph_b200.assembly import load_library
lib = load_library()
lib.pb2_set_device(0) if hasattr(lib, "pb2_set_device") else None
for wl, n in (("ale_freesurface", 512),):
    for v in ("0", "1", "0", "1"):
        os.environ["PB2_MAP_ASYNC"] = v
        out = bench.run_extra_workload(lib, 0, wl, n, 10, 6556.2, 36.9)
        print(wl, "async=" + v, "ms", round(out["ms_per_step"], 3), flush=True)
