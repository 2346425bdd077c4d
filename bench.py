#!/usr/bin/env python
"""bench.py -- element residual+Jacobian assemblies/s of the B200 assembly engine (BASELINE.json metric).

One "step" = one full residual+Jacobian assembly (flag 1) of the workload mesh: ONE persistent, warp-specialised launch of the
generated routine (gather -> points -> contraction -> coloured scatter into the fixed CSR pattern; DESIGN.md sections 2-3), plus one
pack and one add kernel per neighbour and a grouped NCCL send/recv when the mesh is split over several GPUs.  Pattern, position maps
and schedule are setup and are not timed (the reference's Jacobian_setup_time includes its per-assembly pattern
build).  `value` is device-resident (CUDA events, max over ranks); `e2e` is the reference-facing call with pinned host buffers, host
dof vector in and host residual / CSR values out, every step; `cpu_baseline` / `--impl reference` time the CPU restatement of the
reference path (oracle/, all host threads) on a bounded sample of the same element class.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ns_cavity|heat3d|poisson] [--n SIZE]

Default workload: BASELINE configs[1], 2D Navier-Stokes lid-driven cavity, Taylor-Hood Q9/Q4, 1024x1024 elements.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "element Jacobian+residual assemblies/sec"
UNIT = "elements/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class gpu_local_cpus:
    """Run the enclosed block on the CPUs of the GPU's NUMA node (pinned host buffers are placed on the node of the allocating
    thread; a buffer on the far socket halves the host-link rate of the end-to-end leg).  No-op where the topology is not exposed."""

    def __init__(self, device):
        self.device, self.saved = device, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(self.device)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
            if node < 0:
                return self
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus |= set(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, cpus)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def build_workload(workload, n, seed=0):
    from problems import smooth_field
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import NavierStokesEquations, PoissonEquation, TransientHeatEquation
    from pyoomph_b200.meshes import CuboidBrickMesh, RectangularQuadMesh, assign_equation_numbers
    from problems import poisson_source
    if workload == "ns_cavity":
        mesh = RectangularQuadMesh((n, int(os.environ["PB2_BENCH_NY"])) if os.environ.get("PB2_BENCH_NY") else n)   # strip meshes: schedule experiments
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0), name="ns")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
        pinned = {"velocity_x": wall, "velocity_y": wall, "pressure": np.array([0])}
        unsteady = False
        label = "2D Navier-Stokes lid-driven cavity, Taylor-Hood QUAD2, %dx%d elements, steady Newton step" % (n, n)
    elif workload == "heat3d":
        mesh = CuboidBrickMesh(n)
        code = FiniteElementCode("Brick3dC2", TransientHeatEquation(), name="heat3d")
        pinned = {"u": mesh.boundaries["left"]}
        unsteady = True
        label = "3D transient heat, C2 bricks %d^3, BDF2" % n
    elif workload == "poisson":
        mesh = RectangularQuadMesh(n)
        code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
        pinned = {"u": np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])}
        unsteady = False
        label = "2D Poisson QUAD2 %dx%d" % (n, n)
    elif workload == "ale_freesurface":
        # BASELINE config 4's element classes: axisymmetric NS Taylor-Hood on a pseudo-elastic moving mesh (49 dofs per element) + the free-surface
        # interface class on the top boundary, assembled into the same matrix (child problem)
        from pyoomph_b200.equations import DeclareFields, NavierStokesFreeSurface, PseudoElasticMesh
        from pyoomph_b200.meshes import boundary_line_mesh
        mesh = RectangularQuadMesh(n)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh() + DeclareFields(_kin_bc="C2"),
                                 name="aleaxiif", coordinate_system="axisymmetric")
        imesh = boundary_line_mesh(mesh, ["top"])
        icode = FiniteElementCode("Line1dC2", NavierStokesFreeSurface(surface_tension=0.7, static_interface=False), name="freesurfmovaxi",
                                  coordinate_system="axisymmetric")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("bottom", "left", "right")]))
        pinned = {"velocity_x": wall, "velocity_y": wall, "_kin_bc": np.setdiff1d(np.arange(mesh.n_node), np.unique(imesh.elem_nodes))}
        pinned_pos = {"coordinate_x": np.unique(np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])), "coordinate_y": mesh.boundaries["bottom"]}
        unsteady = True
        label = "axisymmetric NS Taylor-Hood on a pseudo-elastic moving mesh %dx%d + free-surface interface class (%d line elements) in one matrix, BDF2" % (n, n, imesh.n_elem)
        extra = dict(interface_mesh=imesh, interface_code=icode)
    elif workload in ("ns_swirl_hvp", "ns_azimuthal"):
        # BASELINE config 5's element class: axisymmetric NS Taylor-Hood with swirl (31 dofs per element); timed: one Hessian-vector
        # product d(J.Y)/dU, resp. the real contribution of the azimuthal m = 1 eigenproblem (Jacobian + mass matrix)
        from pyoomph_b200.expressions import AxisymmetryBreakingCoordinateSystem
        mesh = RectangularQuadMesh(n)
        eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0, with_azimuthal_velocity=True)
        if workload == "ns_swirl_hvp":
            code = FiniteElementCode("Quad2dC2", eqs, name="nsswirl", coordinate_system="axisymmetric")
            label = "axisymmetric NS Taylor-Hood with swirl %dx%d: one Hessian-vector product d(J.Y)/dU, BDF2" % (n, n)
        else:
            code = FiniteElementCode("Quad2dC2", eqs, name="nsazi", coordinate_system=AxisymmetryBreakingCoordinateSystem("azimuthal_m"))
            label = "axisymmetric NS Taylor-Hood with swirl %dx%d: real contribution of the azimuthal m=1 eigenproblem (J and M), BDF2" % (n, n)
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("right", "bottom", "top")]))
        pinned = {"velocity_x": np.unique(np.concatenate([wall, mesh.boundaries["left"]])), "velocity_y": wall,
                  "velocity_phi": np.unique(np.concatenate([mesh.boundaries["right"], mesh.boundaries["left"]]))}
        unsteady = True
    else:
        raise SystemExit("unknown workload " + workload)
    dofmap = assign_equation_numbers(mesh, code, pinned, locals().get("pinned_pos"))
    T, nval = code.history_levels(), code.n_nodal_values
    vals = np.zeros((T, mesh.n_node, nval))
    for t in range(T):
        for f in range(nval):
            vals[t, :, f] = smooth_field(mesh.node_pos, f + 3 * t, seed) * (1.0 - 0.05 * t)
    pos_hist = None
    if code.coordinates_as_dofs:
        pos_hist = np.stack([mesh.node_pos + 1e-3 / n * (1 + 0.3 * t) * np.stack(
            [smooth_field(mesh.node_pos, 10 + d + 2 * t, seed) for d in range(mesh.dim)], axis=1) for t in range(T)])
    out = dict(kind=workload, code=code, mesh=mesh, dofmap=dofmap, vals=vals, pos_hist=pos_hist, unsteady=unsteady, params={}, label=label)
    out.update(locals().get("extra") or {})
    return out


def cpu_run(workload, n_sample, steps, warmup, threads):
    """Reference path on the host cores: oracle plugin (gcc -O3 -march=native, SystemCCompiler flags) driven by the
    restated serial element loop + vectors_of_pairs scatter, static element-range split over `threads`."""
    from problems import make_oracle
    pb = build_workload(workload, n_sample)
    op = make_oracle(pb)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        op.assemble(flag=1, nthreads=threads, fetch=False)      # residual + CSR arrays produced in memory; no copy into numpy
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    ne = pb["mesh"].n_elem
    op.close()
    return ne, float(np.mean(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="ns_cavity")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle window check of the distributed matrix (N > 1)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads (BASELINE configs 1 and 3) of the default run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    default_n = {"ns_cavity": 1024, "heat3d": 126, "poisson": 2048}[args.workload]
    n = args.n or default_n
    cores = os.cpu_count() or 1
    sample_n = {"ns_cavity": 160, "heat3d": 14, "poisson": 256}[args.workload]

    if args.impl == "reference":
        # the reference's CPU path (oracle port: generated-format C plugin + restated element loop + vectors_of_pairs CSR build) on all host
        # threads; EXACTLY --steps timed assemblies after --warmup untimed ones, each step one assembly of a bounded sample mesh of the
        # same element class (the full mesh needs minutes per run on the host cores)
        if rank != 0:
            return
        ref_n = int(os.environ.get("PB2_REF_SAMPLE_N", {"ns_cavity": 320, "heat3d": 20, "poisson": 512}[args.workload]))
        ne, sec = cpu_run(args.workload, ref_n, max(1, args.steps), max(0, args.warmup), cores)
        val = ne / sec
        wl = build_label(args.workload, n)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(0, args.warmup),
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "one assembly of %s at %d^%d = %d elements per step (same element class and data as the full workload), oracle C plugin gcc -O3 -march=native, %d threads" % (
                                     args.workload, ref_n, 3 if args.workload == "heat3d" else 2, ne, cores)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    from pyoomph_b200.assembly import load_library
    lib = load_library()
    lib.pb2_host_alloc.restype = ctypes.c_void_p
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        import datetime
        torch.cuda.set_device(local_rank)
        # a desynchronised collective fails within two minutes instead of sitting in the NCCL watchdog's default ten
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=120))
        dist = dist_mod
    device = local_rank

    t_setup = t_start = time.time()
    verbose = os.environ.get("PB2_BENCH_VERBOSE") and rank == 0

    def note(msg):
        if verbose:
            print("[bench %.1fs] %s" % (time.time() - t_start, msg), file=sys.stderr, flush=True)
    pb = build_workload(args.workload, n)
    note("workload built")
    mesh = pb["mesh"]
    from pyoomph_b200.assembly import B200Assembly
    dasm = None
    if world > 1:
        # element blocks in mesh order (x-strips), contiguous CSR row block per GPU, interface rows exchanged over NCCL
        from pyoomph_b200.distributed import DistributedAssembly, GPULocalAssembler

        def make_local(el, dm, extra):
            a_ = B200Assembly(pb["code"], mesh, dm, name=pb["code"].name, device=device, elements=el, extra_pattern=extra)
            return GPULocalAssembler(a_, device)
        dasm = DistributedAssembly.create(pb["code"], mesh, pb["dofmap"], rank, world, make_local, dist=dist, device="cuda:%d" % device)
        asm = dasm.local.asm
    else:
        asm = B200Assembly(pb["code"], mesh, pb["dofmap"], name=pb["code"].name, device=device)
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    if pb["unsteady"]:
        from problems import TIME
        asm.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    note("assembler ready")
    t_setup = time.time() - t_setup
    n_elem_rank = asm.n_elem
    info = asm.info

    def barrier():
        lib.pb2_device_synchronize()
        if dist is not None:
            dist.barrier()
        lib.pb2_device_synchronize()

    # ---- device-resident timing (value)
    def step():
        if dasm is not None:
            dasm.assemble(flag=1)
        else:
            asm.assemble(flag=1)

    sampler = ClockSampler(device)
    sampler.start()                      # nvidia-smi needs a few 100 ms to deliver its first sample: start before the warm-up
    # Warm-up.  Every step() of a multi-GPU run contains matched NCCL sends/receives, so ALL ranks must run the SAME number of steps:
    # the count is fixed per group, and whether another group follows (>= 0.6 s of load, so that clocks and the nvidia-smi sampler have
    # settled) is decided by an all-reduce, never by a rank-local clock.
    barrier()
    t_w = time.time()
    n_w = 0
    group = max(3, args.warmup)
    while True:
        for _ in range(group):
            step()
        n_w += group
        lib.pb2_device_synchronize()
        more = 1.0 if (time.time() - t_w < 0.6 and n_w < 4096) else 0.0
        if dist is not None:
            import torch
            flag_t = torch.tensor([more], device="cuda", dtype=torch.float64)
            dist.all_reduce(flag_t, op=dist.ReduceOp.MAX)
            more = float(flag_t.item())
        if more == 0.0:
            break
        group = 16
    barrier()
    note("warm-up done (%d steps)" % n_w)
    launches = 0
    lib.pb2_event_record(0, None)
    for _ in range(args.steps):
        step()
        launches += asm.launch_count() + ((len(dasm.send) + len(dasm.recv)) if dasm is not None else 0)   # + one pack / add kernel per neighbour
    lib.pb2_event_record(1, None)
    ms = ctypes.c_float()
    lib.pb2_event_elapsed_ms(0, 1, ctypes.byref(ms))
    barrier()
    clocks = sampler.stop()
    ms_step = ms.value / args.steps
    if dist is not None:
        import torch
        tt = torch.tensor([ms_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())
    total_elems = mesh.n_elem
    value = total_elems / (ms_step * 1e-3)
    exch_max = exch_sum = int(dasm.exchange_bytes) if dasm is not None else 0
    if dist is not None:
        import torch
        et = torch.tensor([float(exch_max)], device="cuda", dtype=torch.float64)
        es = et.clone()
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(es, op=dist.ReduceOp.SUM)
        exch_max, exch_sum = int(et.item()), int(es.item())

    # ---- end-to-end through the reference-facing call (host buffers, copies inside the timed region)
    e2e = e2e_dev = None
    if not args.no_e2e and world == 1:
        eq = pb["dofmap"].node_eqn
        ndof, nnz = asm.n_dof, asm.nnz

        def pinned(nelem):
            ptr = lib.pb2_host_alloc(ctypes.c_size_t(nelem * 8))
            if not ptr:
                raise RuntimeError("pinned allocation failed")
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(nelem,))
        with gpu_local_cpus(device):
            h_dofs, h_res, h_jac = pinned(ndof), pinned(ndof), pinned(nnz)
            h_res[:] = 0.0
            h_jac[:] = 0.0          # touch the pages here, on the GPU's NUMA node
        h_dofs[:] = 0.0
        h_dofs[eq[eq >= 0]] = pb["vals"][0][eq >= 0]
        ksteps = max(1, min(args.steps, 5))
        asm.assemble_host(h_dofs, 1, out=(h_res, h_jac, None))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            asm.assemble_host(h_dofs, 1, out=(h_res, h_jac, None))
        barrier()
        sec = (time.perf_counter() - t0) / ksteps
        e2e = {"value": total_elems / sec, "unit": UNIT, "h2d_bytes_per_step": int(ndof * 8), "d2h_bytes_per_step": int((ndof + nnz) * 8),
               "ms_per_step": sec * 1e3, "steps": ksteps}
        # the same step when the linear solver is device-resident (SURVEY N-d, pyoomph_b200/solvers.py): host dofs in, full R+J assembly,
        # only the residual comes back -- the matrix stays in HBM for the solver.  Reported next to e2e, never instead of it.
        t0 = time.perf_counter()
        for _ in range(ksteps):
            asm.set_dofs(h_dofs)
            asm.assemble(flag=1)
            lib.pb2_problem_fetch(asm.prob, h_res.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None, None)
        barrier()
        sec_d = (time.perf_counter() - t0) / ksteps
        e2e_dev = {"value": total_elems / sec_d, "unit": UNIT, "h2d_bytes_per_step": int(ndof * 8), "d2h_bytes_per_step": int(ndof * 8),
                   "ms_per_step": sec_d * 1e3, "steps": ksteps,
                   "note": "Jacobian left on the device for a device-resident solver plugin (DeviceLinearSystemSolver); not the headline e2e"}
    elif not args.no_e2e:
        # N GPUs: every rank feeds the host dof values of its local rows and reads its owned CSR row block back (N host links);
        # local kernels + the NCCL interface exchange sit between the copies.  A failure on one rank must not desynchronise the
        # collectives: the timing all-reduce below is unconditional.
        # The calls below contain collectives (the interface exchange), so nothing here is wrapped in try/except: a failure on one rank
        # ends that process, torchrun ends the others, the run fails loudly instead of leaving peers in an unmatched send/recv.
        import torch
        ksteps = max(1, min(args.steps, 5))
        part = dasm.part
        eq = part.local_dofmap.node_eqn
        nnz_owned = int(dasm.indptr[dasm.n_owned])
        with gpu_local_cpus(device):
            h_dofs = torch.zeros(asm.n_dof, dtype=torch.float64).pin_memory()
            outb = (torch.zeros(dasm.n_owned, dtype=torch.float64).pin_memory(), torch.zeros(nnz_owned, dtype=torch.float64).pin_memory(), None)
        m_ = eq >= 0
        h_dofs.numpy()[eq[m_]] = pb["vals"][0][m_]
        h2d, d2h = int(asm.n_dof * 8), int((dasm.n_owned + nnz_owned) * 8)
        barrier()
        dasm.assemble_host(h_dofs, 1, out=outb)          # warm-up (ends with a device synchronisation, like every call)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            dasm.assemble_host(h_dofs, 1, out=outb)
        barrier()
        sec = (time.perf_counter() - t0) / ksteps
        tt = torch.tensor([sec, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        sec = float(mx[0].item())
        e2e = {"value": total_elems / sec, "unit": UNIT, "h2d_bytes_per_step": int(tt[1].item()), "d2h_bytes_per_step": int(tt[2].item()),
               "ms_per_step": sec * 1e3, "steps": ksteps, "note": "bytes summed over ranks; every rank copies its own row block"}

    # ---- N GPUs: correctness of THIS run's distributed matrix, on the hardware that was timed.  Every rank compares rows of its owned
    # block inside 4x4-element windows that straddle its partition interfaces (the rows the NCCL exchange completes) and windows in its
    # interior with the CPU oracle assembled on the window alone (tests/windows.py); worst error and row count are reduced over ranks.
    mgp = None
    if dasm is not None and args.workload == "ns_cavity" and not args.no_parity:
        import torch
        from problems import make_oracle
        from windows import check_windows
        dasm.assemble(flag=1)
        rb, re_, ip_, gc_, jv_, res_ = dasm.owned_block()
        wins = []
        for q in range(1, world):                       # element blocks are x-strips in mesh order: interface q at element column n*q/world
            ix0 = (n * n * q // world) // n
            for jy in (0, n // 2 - 2, n - 4, (37 * q) % (n - 4)):
                wins += [(max(0, min(n - 4, ix0 - 2)), jy), (max(0, min(n - 4, ix0 - 3)), jy), (max(0, min(n - 4, ix0 - 1)), jy)]
        lo, hi = (n * n * rank // world) // n, (n * n * (rank + 1) // world) // n
        wins += [(max(0, min(n - 4, (lo + hi) // 2)), n // 3), (max(0, min(n - 4, lo + 1)), 5)]
        st_ = {}
        try:          # no collective inside: a failed comparison on one rank is carried into the reduction below, not raised past it
            worst = check_windows(pb, make_oracle, ip_, gc_, jv_, res_, sorted(set(wins)), w=4, tol=1e-12, new_of_old=dasm.part.new_of_old,
                                  row_begin=rb, row_end=re_, stats=st_)
        except AssertionError as exc:
            print("[bench] rank %d: distributed window parity FAILED: %r" % (rank, exc), file=sys.stderr, flush=True)
            worst = float("inf")
        tw = torch.tensor([worst, -float(st_.get("rows", 0))], device="cuda", dtype=torch.float64)
        tr = torch.tensor([float(st_.get("rows", 0))], device="cuda", dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        dist.all_reduce(tr, op=dist.ReduceOp.SUM)
        mgp = {"checked": "rows of 4x4-element windows at every partition interface and inside every rank's block against the CPU oracle",
               "rows_compared": int(tr.item()), "min_rows_on_a_rank": int(-tw[1].item()), "worst_row_scaled_error": float(tw[0].item()),
               "tolerance": 1e-12, "ok": bool(tw[0].item() <= 1e-12 and -tw[1].item() > 0)}
        del jv_

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    b_el = float(info.alg_bytes_per_elem[1])
    if not pb["unsteady"]:
        b_el -= float(info.alg_bytes_per_hist_level) * (info.n_hist_val - 1)   # steady: history levels are not read
    # dominant (only) kernel: the generated ResidualAndJacobian routine, one launch per colour
    achieved = b_el * n_elem_rank / (ms_step * 1e-3) / 1e9
    traffic, flops = profile_numbers(args.workload, n) if world == 1 else (None, None)
    fp64_peak = ctypes.c_double(0.0)
    if lib.pb2_measure_fp64_peak(device, ctypes.byref(fp64_peak)) != 0:
        fp64_peak.value = 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "alg_bytes_per_element": b_el, "kernel": "pb2_%s_r0_f1" % pb["code"].name,
                "launches_per_step": launches // max(1, args.steps), "tiles_per_step": asm.num_launches(), "avg_launch_ms": ms_step / max(1, launches // max(1, args.steps)),
                "fp64_peak_tflops": fp64_peak.value or None, "fp64_peak_source": "measured in this run (pb2_measure_fp64_peak: dependent-chain-free DFMA kernel)"}
    if flops is not None and fp64_peak.value > 0:
        tf = flops / (ms_step * 1e-3) / 1e12
        roofline.update({"fp64_tflops": tf, "frac_fp64": tf / fp64_peak.value, "fp64_flops_per_element": flops / n_elem_rank})
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": pb["label"], "elements": int(total_elems), "dofs": int(asm.n_dof), "nnz": int(asm.nnz), "ndof_el": int(info.ndof_el),
                       "colours": asm.num_colours(), "tiles": asm.num_launches(), "cache": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2" % (b_el * n_elem_rank / 1e9),
                       "setup_s": round(t_setup, 1), "pattern_setup_s": round(asm.setup_seconds, 2),
                       "ms_per_step_incl_pattern_setup": ms_step + asm.setup_seconds * 1e3,
                       "parallelism": "element blocks x%d" % world,
                       "exchange_bytes_per_step_max_rank": exch_max, "exchange_bytes_per_step_all_ranks": exch_sum},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches)}
    if e2e is not None:
        line["e2e"] = e2e
    if e2e_dev is not None:
        line["e2e_matrix_stays_on_device"] = e2e_dev
    if mgp is not None:
        line["multi_gpu_parity"] = mgp
    if not args.no_cpu_baseline and world == 1:
        ne, sec = cpu_run(args.workload, sample_n, 2, 1, cores)
        line["cpu_baseline"] = {"value": ne / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%s at sample size %d (%d elements), oracle C plugin gcc -O3 -march=native, %d threads, 2 assemblies" % (args.workload, sample_n, ne, cores)}
        ne1, sec1 = cpu_run(args.workload, max(8, sample_n // 3), 1, 0, 1)
        line["cpu_baseline"]["value_1core"] = ne1 / sec1
        # the port's element loop scales with threads, its vectors-of-pairs merge does not: perfect scaling of the 1-thread rate is the
        # most any CPU run of this port could reach on this box
        line["cpu_baseline"]["value_1core_times_cores"] = ne1 / sec1 * cores
    if world == 1 and not args.no_extra and args.workload == "ns_cavity" and not args.n:
        # BASELINE configs 1, 3 and the element classes of 4 and 5 on the same GPU, so that the driver's record carries them too (config 2 above is the headline)
        asm.close()
        line["extra_workloads"] = [run_extra_workload(lib, device, wl_, n_, max(3, min(args.steps, 10)), peak, fp64_peak.value)
                                   for (wl_, n_) in (("heat3d", 126), ("poisson", 2048), ("ale_freesurface", 512), ("ns_swirl_hvp", 512), ("ns_azimuthal", 512))]
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def profile_numbers(workload, n):
    """DRAM bytes and executed fp64 flops per launch from the committed `ncu --set full` capture of this very workload (profiles/r0X_traffic.json,
    newest round first); (None, None) when no capture of this workload/size is committed."""
    for rnd in ("r02", "r01"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "%s_traffic.json" % rnd)))
            e = tj["%s:%d" % (workload, n)]
            return float(e["traffic_bytes"]), (float(e["fp64_flops"]) if "fp64_flops" in e else None)
        except Exception:
            continue
    return None, None


def run_extra_workload(lib, device, workload, n, steps, peak_hbm, peak_fp64):
    """One more BASELINE config on the same GPU, device-resident timing only: ms per assembly, elements/s and both roofline fractions."""
    from pyoomph_b200.assembly import B200Assembly
    t0 = time.time()
    pb = build_workload(workload, n)
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, device=device)
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    if pb["pos_hist"] is not None:
        for t in range(pb["pos_hist"].shape[0]):
            asm.set_nodal_positions(t, pb["pos_hist"][t])
    child = None
    if "interface_code" in pb:
        child = B200Assembly(pb["interface_code"], pb["interface_mesh"], pb["dofmap"], name=pb["interface_code"].name, parent=asm)
    if pb["unsteady"]:
        from problems import TIME
        asm.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    if workload == "ns_azimuthal":
        asm.set_parameters(azimuthal_m=1.0)
    t_setup = time.time() - t0
    if workload == "ns_swirl_hvp":
        Y = np.random.default_rng(3).standard_normal(asm.n_dof)[None, :]
        Yp = Y.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

        def step():          # the vector is copied to the device inside the call (asm.n_dof doubles): part of what a tracker pays per product
            if lib.pb2_problem_assemble_hessian(asm.prob, 0, 1, 1, Yp, None) != 0:
                raise RuntimeError(lib.pb2_last_error().decode())
    elif workload == "ns_azimuthal":
        rn = pb["code"].coordinate_system.real_contribution_name

        def step():
            asm.assemble(flag=2, residual=rn)
    else:
        def step():
            asm.assemble(flag=1)     # a parent assembles its child element classes behind its own launch
    t_w, n_w = time.time(), 0
    while n_w < 3 or (time.time() - t_w < 0.4 and n_w < 2000):
        step()
        n_w += 1
        if n_w % 4 == 0:
            lib.pb2_device_synchronize()
    lib.pb2_device_synchronize()
    lib.pb2_event_record(2, None)
    for _ in range(steps):
        step()
    lib.pb2_event_record(3, None)
    ms = ctypes.c_float()
    lib.pb2_event_elapsed_ms(2, 3, ctypes.byref(ms))
    ms_step = ms.value / steps
    info = asm.info
    b_el = float(info.alg_bytes_per_elem[2 if workload == "ns_azimuthal" else 1])
    if not pb["unsteady"]:
        b_el -= float(info.alg_bytes_per_hist_level) * (info.n_hist_val - 1)
    ne = pb["mesh"].n_elem
    hbm = b_el * ne / (ms_step * 1e-3) / 1e9
    traffic, flops = profile_numbers(workload, n)
    out = {"workload": pb["label"], "elements": int(ne), "ndof_el": int(info.ndof_el), "nnz": int(asm.nnz), "steps": steps, "ms_per_step": ms_step,
           "value": ne / (ms_step * 1e-3), "unit": UNIT, "setup_s": round(t_setup, 1), "pattern_setup_s": round(asm.setup_seconds, 2),
           "roofline": {"bound": "hbm", "achieved": hbm, "peak": peak_hbm, "unit": "GB/s", "frac": hbm / peak_hbm, "alg_bytes_per_element": b_el,
                        "traffic": traffic, "kernel": {"ns_swirl_hvp": "pb2_%s_h0_f1", "ns_azimuthal": "pb2_%s_r1_f2"}.get(workload, "pb2_%s_r0_f1") % pb["code"].name}}
    if child is not None:
        out["launches_per_step"] = 2
        out["interface_elements"] = int(pb["interface_mesh"].n_elem)
    if flops is not None and peak_fp64:
        tf = flops / (ms_step * 1e-3) / 1e12
        out["roofline"].update({"fp64_tflops": tf, "fp64_peak_tflops": peak_fp64, "frac_fp64": tf / peak_fp64, "fp64_flops_per_element": flops / ne})
        # the kernel is bounded by whichever floor is higher: HBM time of the algorithmic bytes or fp64-pipe time of the executed flops
        if tf / peak_fp64 > hbm / peak_hbm:
            out["roofline"]["bound"] = "fp64"
    asm.close()
    return out


def build_label(workload, n):
    return {"ns_cavity": "2D Navier-Stokes lid-driven cavity, Taylor-Hood QUAD2, %dx%d elements, steady Newton step" % (n, n),
            "heat3d": "3D transient heat, C2 bricks %d^3, BDF2" % n, "poisson": "2D Poisson QUAD2 %dx%d" % (n, n)}[workload]


if __name__ == "__main__":
    main()
