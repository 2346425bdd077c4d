"""A/B of the register-pipelined batch table (PB2_BT_PIPE_S scatter role, PB2_BT_PIPE_G gather role), same box"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyoomph_b200.assembly import load_library
lib = load_library()
for wl, n in (("ns_cavity", 1024), ("poisson", 2048), ("ns_swirl_hvp", 512), ("heat3d", 126)):
    for s_, g_ in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1"), ("0", "0"), ("1", "1")):
        os.environ["PB2_BT_PIPE_S"], os.environ["PB2_BT_PIPE_G"] = s_, g_
        out = bench.run_extra_workload(lib, 0, wl, n, 10, 6556.2, 36.9)
        print(wl, "scatter=" + s_, "gather=" + g_, "ms", round(out["ms_per_step"], 3), flush=True)
