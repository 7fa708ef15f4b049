#!/usr/bin/env python
"""bench.py -- fluid Newton-iterations/s (assembly + GMRES) on the synthetic 10M-tet pipe.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores
    (N > 1: launched by torch.distributed.run, one rank per GPU; strong scaling, the 10M-tet mesh
     is split into N axial slabs)
    python bench.py --scaling weak --solver ns --gpus 8      # BASELINE configs[4]: 5M tets per GPU, NSSOLVER
    python bench.py --physics heat --gpus 8                  # BASELINE configs[3]: heatS + CGRADS

One "step" = one Newton iteration of the reference's hot path as SURVEY.md 8d defines it,
1 / (t_zero + t_asm + t_bc + t_commuR + t_solve + t_update), S/MAIN.f:111-206: the generalised-alpha
initiator (PICI; PICP first, so that every step starts from the same old state and does the same
work), LSALLOC zeroing + CONSTRUCT_FLUID element loop + block-CSR assembly, the outlet flux (IntegV)
and the resistance Neumann face (BASSEMNEUBC / BFLUID), COMMU(R), FSILS_SOLVE (Jacobi scaling +
restarted GMRES with the face's rank-one resistance term, ADDBCMUL) and the corrector PICC.  `value`
times it with the state resident in HBM; `e2e` times assembly + face + COMMU + solve through the
C-ABI with HOST buffers (Ag, Yg uploaded from pinned memory, the increment downloaded: the two call
sites of SURVEY.md 8b with the time integrator left on the host) inside the timed region.  Prints ONE
JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload: BASELINE.json configs[2] / SURVEY.md 8d "C3 10M": 64 x 64 x 408 Kuhn lattice
DIMS = (64, 64, 408)
R_PIPE, L_PIPE = 2.0, 30.0
RHO, MU, DT = 1.06, 0.04, 5e-3
RHO_INF = 0.2
F_BODY = (0.0, 0.0, 0.0)
# FSILS GMRES + diagonal (FSILS) preconditioner; svFSI-Tests-like pipe settings
LS = dict(relTol=1e-3, absTol=1e-12, maxItr=10, dimKry=50)
R_OUT = 100.0  # outlet resistance r [dyn s / cm^5] (bc%r, S/SETBC.f:282-283): h = r * Q; res = gamma*dt*r
BF_STAB = 0.2  # backflow stabilisation coefficient (svFSI default, S/READFILES.f)


def gen_alpha(rho_inf):
    am = 0.5 * (3.0 - rho_inf) / (1.0 + rho_inf)
    af = 1.0 / (1.0 + rho_inf)
    return dict(am=am, af=af, gam=0.5 + am - af, beta=0.25 * (1.0 + am - af) ** 2)


GA = gen_alpha(RHO_INF)


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md).
    NVML in a thread (one sample every ~5 ms; the main thread sits in ctypes calls that release the
    GIL), so even a 200 ms timed region gets tens of samples; `nvidia-smi -lms` as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, uuid=None):
        self.device, self.uuid = device, uuid
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.nv = None
        self.handle = None
        self.stop_flag = threading.Event()
        self.t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    u = str(uuid)
                    h = pynvml.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = device
                if vis:
                    ent = vis.split(",")[device].strip()
                    if ent.isdigit():
                        idx = int(ent)
                    else:
                        h = pynvml.nvmlDeviceGetHandleByUUID(ent)
                if h is None:
                    h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.handle = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(1.0)        # nvidia-smi needs ~0.5 s before its first sample
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nv is not None:
            self.stop_flag.set()
            if self.t:
                self.t.join(timeout=1.0)
            nv = self.nv
            bits = dict(hw_slowdown=nv.nvmlClocksEventReasonHwSlowdown,
                        hw_thermal_slowdown=nv.nvmlClocksEventReasonHwThermalSlowdown,
                        sw_thermal_slowdown=nv.nvmlClocksEventReasonSwThermalSlowdown,
                        sw_power_cap=nv.nvmlClocksEventReasonSwPowerCap,
                        hw_power_brake=nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)
            sm = [x[0] for x in self.samples]
            reasons = sorted(k for k, b in bits.items() if any(x[1] & b for x in self.samples))
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_min_mhz=min(sm) if sm else None,
                        sm_max_mhz=self.max_mhz, reasons=reasons, samples=len(sm), source="nvml")
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx,
                    reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


PHYSICS = "fluid"      # --physics heat switches to BASELINE.json configs[3] (heatS, CGRADS)
SOLVER = "gmres"       # --solver ns switches to FSILS_NSSOLVER with the FSILS defaults (configs[4])
HEAT = dict(nu=1.0, s=0.0, rho=1.0)
HEAT_LS = dict(relTol=1e-6, absTol=1e-12, maxItr=1000)
# FP64 operations per element of the default (gather) fluid assembly, 2 x DFMA + DMUL + DADD, from ncu's
# smsp__sass_thread_inst_executed_op_dfma/dmul/dadd_pred_on on the kernels of this commit at 10.03M tets
# (profiles/r02_ncu_asm.md): records kernel v4 883 / 297 / 122 per element = 2186 flop, tangent gather 768 / 320
# / 0 = 1857, residual gather 16 adds; the reference's loop needs ~15 kflop
ASM_FLOP_PER_ELEM = 4059
FP64_PEAK_TFLOPS = 37.2


def setup_rank(api, mesh, dims, rank, nparts):
    gnNo, p = mesh.build_rank_problem(*dims, rank=rank, nparts=nparts, R=R_PIPE,
                                      L=L_PIPE * dims[2] / DIMS[2])
    if PHYSICS == "heat":
        api.FSILS_LHS_CREATE(gnNo, p.rm.nNo, p.colPtr.size, p.rm.ltg, p.rowPtr, p.colPtr, 2)
        for fi, name in enumerate(("inlet", "outlet"), start=1):
            fa = p.faces[name]
            api.FSILS_BC_CREATE(fi, fa["gN"].size, 1, api.BC_TYPE_Dir, fa["gN"], None)
        p.Ag = np.ascontiguousarray(p.Yg[:, 0] * 0.1)        # dT/dt
        p.Yg = np.ascontiguousarray(p.Yg[:, 2])              # T
    else:
        api.FSILS_LHS_CREATE(gnNo, p.rm.nNo, p.colPtr.size, p.rm.ltg, p.rowPtr, p.colPtr, 3)
        for fi, name in enumerate(("inlet", "wall", "outlet"), start=1):
            fa = p.faces[name]
            api.FSILS_BC_CREATE(fi, fa["gN"].size, 3,
                                api.BC_TYPE_Neu if fa["bc"] == "Neu" else api.BC_TYPE_Dir, fa["gN"],
                                fa["val"])
    api.mesh_create(p.rm.IEN, p.rm.x)
    if PHYSICS == "fluid":
        fo = p.faces["outlet"]           # every rank creates it (the flux is summed over ranks)
        api.face_create(3, fo["gN"], fo["IEN"], fo["gE"])
    return gnNo, p


def make_ls(api):
    if PHYSICS == "heat":
        return api.FSILS_LS_CREATE(api.LS_TYPE_CG, **HEAT_LS)
    if SOLVER == "ns":
        return api.FSILS_LS_CREATE(api.LS_TYPE_NS)          # L/LS.f:70-78 defaults
    return api.FSILS_LS_CREATE(api.LS_TYPE_GMRES, **LS)


def res_out():
    return GA["gam"] * DT * R_OUT


def newton_step_dev(api, variant, p):
    """device-resident Newton iteration (S/MAIN.f:111-206): PICP + PICI, R=0, Val=0, element loop, outlet
    flux + resistance Neumann face, COMMU(R), FSILS_SOLVE, PICC -- no nodal vector crosses PCIe"""
    api.PICP(GA["gam"])
    api.pici(GA["am"], GA["af"])
    if PHYSICS == "heat":
        api.construct_heats_dev(HEAT["nu"], HEAT["s"], HEAT["rho"], DT, GA["af"], GA["am"], GA["gam"],
                                variant)
        api.commu_dev(1)
        ls = make_ls(api)
        api.solve_dev(ls, 1, incL=[1, 1])
        api.picc(GA["gam"], GA["beta"], DT)
        return ls
    api.construct_fluid_dev(RHO, MU, F_BODY, DT, GA["af"], GA["am"], GA["gam"], variant)
    q = api.IntegV(3, which=1, s=1)                       # Q = int Yn . n dGamma over all ranks
    api.BASSEMNEUBC_FLUID(3, np.full(p.faces["outlet"]["gN"].size, -(R_OUT * q)), RHO, BF_STAB, GA["af"],
                          GA["gam"], DT)                  # h = -bc%r * Q (S/SETBC.f:282-283, :292-306)
    api.commu_dev(4)
    ls = make_ls(api)
    api.solve_dev(ls, 4, incL=[1, 1, 1], res=[0.0, 0.0, res_out()])
    api.picc(GA["gam"], GA["beta"], DT)
    return ls


def newton_step_e2e(api, variant, p, Ag, Yg, Rout):
    """assembly + face + COMMU + solve through the reference-facing calls with HOST buffers (the time
    integrator stays on the host, as at the two call sites of SURVEY.md 8b)"""
    if PHYSICS == "heat":
        api.CONSTRUCT_HEATS(Ag, Yg, HEAT["nu"], HEAT["s"], HEAT["rho"], DT, GA["af"], GA["am"],
                            GA["gam"], variant)
        api.commu_dev(1)
        ls = make_ls(api)
        api.solve_dev(ls, 1, incL=[1, 1])
        api._check(api.lib().gpu_get_r_(api._ci(1), api._d(Rout)))
        return ls
    api.CONSTRUCT_FLUID(Ag, Yg, None, RHO, MU, F_BODY, DT, GA["af"], GA["am"], GA["gam"], variant)
    q = api.IntegV(3, which=0, s=1)
    api.BASSEMNEUBC_FLUID(3, np.full(p.faces["outlet"]["gN"].size, -(R_OUT * q)), RHO, BF_STAB, GA["af"],
                          GA["gam"], DT)
    api.commu_dev(4)
    ls = make_ls(api)
    api.solve_dev(ls, 4, incL=[1, 1, 1], res=[0.0, 0.0, res_out()])
    api._check(api.lib().gpu_get_r_(api._ci(4), api._d(Rout)))
    return ls


def cpu_newton_sample(nz_sample, threads_note="1 (scalar port, single simulated rank)", rebuild=True):
    """Oracle port (-O3 -march=native, the reference's own flags) on a slice of the same pipe:
    one Newton iteration (assembly incl. the reference's per-element overheads + COMMU + GMRES
    with the bench's solver settings), scaled to the full mesh by element count."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as ora
    from svfsi_b200 import mesh
    if rebuild:
        ora.build(native=True, force=True)  # -march=native must be compiled on THIS host
    dims = (DIMS[0], DIMS[1], nz_sample)
    Ls = L_PIPE * nz_sample / DIMS[2]
    gnNo, p = mesh.build_rank_problem(*dims, rank=0, nparts=1, R=R_PIPE, L=Ls)
    par = ora.fluid_par(RHO, MU, F_BODY, DT, GA["af"], GA["am"], GA["gam"])
    w = ora.World(gnNo, [p.rm.ltg], [p.rowPtr], [p.colPtr], 3, native=True)
    for fi, name in enumerate(("inlet", "wall", "outlet"), start=1):
        fa = p.faces[name]
        w.bc_create(fi, [fa["gN"]], 3, ora.BC_TYPE_Neu if fa["bc"] == "Neu" else ora.BC_TYPE_Dir,
                    None if fa["val"] is None else [fa["val"]])
    fo = p.faces["outlet"]
    p.Yg[:, 3] += R_OUT * ora.integ_v(p.rm.x, p.rm.IEN, fo["IEN"], fo["gE"], p.Yg[:, :3])   # as the GPU arm
    t0 = time.perf_counter()
    An, Yn = ora.picp(p.Ag, p.Yg, GA["gam"])                      # old state = (Ag, Yg) of the GPU arm
    Ag, Yg = ora.pici(p.Ag, An, p.Yg, Yn, GA["am"], GA["af"])
    Rr, Vr = ora.construct_fluid(par, p.rm.IEN, p.rm.x, Ag, Yg, np.zeros((p.rm.nNo, 3)),
                                 p.rowPtr, p.colPtr, faithful=True, native=True)
    q = ora.integ_v(p.rm.x, p.rm.IEN, fo["IEN"], fo["gE"], Yn[:, :3])
    hg = np.zeros(p.rm.nNo); hg[fo["gN"] - 1] = -(R_OUT * q)
    ora.bassem_neu_fluid(p.rm.x, p.rm.IEN, fo["IEN"], fo["gE"], hg, Yg, p.rowPtr, p.colPtr, Rr, Vr, RHO,
                         BF_STAB, GA["af"], GA["gam"], DT)
    t_asm = time.perf_counter() - t0
    ls = ora.ls_create(ora.LS_TYPE_GMRES, **LS)
    t1 = time.perf_counter()
    w.solve(ls, 4, [Rr], [Vr], incL=[1, 1, 1], res=[0.0, 0.0, res_out()])
    ora.picc(An, Yn, Rr, GA["gam"], GA["beta"], DT)
    t_sol = time.perf_counter() - t1
    nEl_full = 6 * DIMS[0] * DIMS[1] * DIMS[2]
    scale = nEl_full / p.rm.nEl
    return dict(t_asm=t_asm, t_sol=t_sol, nEl=p.rm.nEl, itr=ls.RI.itr, scale=scale,
                t_full=(t_asm + t_sol) * scale, threads=threads_note)


def _cpu_worker(args):
    nz_sample, barrier = args
    if barrier is not None:
        barrier.wait()
    t0 = time.perf_counter()
    r = cpu_newton_sample(nz_sample, rebuild=False)
    r["t_wall"] = time.perf_counter() - t0
    return r


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_newton_parallel(nz_sample, procs=None):
    """The CPU arm on ALL host cores: `procs` processes (one per core, like `mpiexec -np procs
    svFSI`), each running one Newton iteration of the oracle port on its own slab of the pipe at
    the same time (they contend for memory bandwidth as MPI ranks would; the halo exchange between
    slabs is NOT included, which flatters the CPU).  Whole-mesh time = slowest process x
    (slabs of the full mesh / procs)."""
    import multiprocessing as mp
    from oracle import oracle as ora
    ora.build(native=True, force=True)      # -march=native must be compiled on THIS host
    procs = procs or max(1, min(host_cores(), 64))
    if procs == 1:
        rs = [_cpu_worker((nz_sample, None))]
    else:
        ctx = mp.get_context("fork")
        with ctx.Manager() as mgr:
            bar = mgr.Barrier(procs)
            with ctx.Pool(procs) as pool:
                rs = pool.map(_cpu_worker, [(nz_sample, bar)] * procs, chunksize=1)
    t_max = max(r["t_asm"] + r["t_sol"] for r in rs)
    nEl_full = 6 * DIMS[0] * DIMS[1] * DIMS[2]
    slabs = nEl_full / rs[0]["nEl"]
    return dict(procs=procs, t_max=t_max, t_full=t_max * slabs / procs, nEl=rs[0]["nEl"],
                itr=rs[0]["itr"], t_asm=max(r["t_asm"] for r in rs), t_sol=max(r["t_sol"] for r in rs),
                slabs=slabs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--nz", type=int, default=DIMS[2], help="axial cells (default = 10M-tet workload)")
    ap.add_argument("--variant", default="gather", choices=["atomic", "colored", "gather"])
    ap.add_argument("--cpu-nz", type=int, default=24, help="axial cells of the CPU-baseline slice")
    ap.add_argument("--cpu-procs", type=int, default=None,
                    help="processes of the CPU arm (default: all host cores, at most 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--solver", default="gmres", choices=["gmres", "ns"],
                    help="non-default: FSILS_NSSOLVER with FSILS defaults (BASELINE configs[4])")
    ap.add_argument("--physics", default="fluid", choices=["fluid", "heat"],
                    help="non-default: heatS + CGRADS (BASELINE configs[3])")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="weak: --nz-per-gpu axial cells PER GPU (default 80 on a 102 x 102 cross-section = "
                         "5M tets per GPU, 40M on 8: BASELINE configs[4])")
    ap.add_argument("--nx", type=int, default=None, help="cross-section cells (default 64; weak: 102)")
    ap.add_argument("--nz-per-gpu", type=int, default=80)
    args = ap.parse_args()
    global PHYSICS, SOLVER
    PHYSICS, SOLVER = args.physics, args.solver
    if PHYSICS != "fluid" or SOLVER != "gmres":
        args.no_cpu = True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling == "weak":
        nx = args.nx or 102
        dims = (nx, nx, args.nz_per_gpu * max(world, 1))
    else:
        nx = args.nx or DIMS[0]
        dims = (nx, nx, args.nz)
    nEl_total = 6 * dims[0] * dims[1] * dims[2]
    if PHYSICS == "heat":
        what = f"unsteady heat diffusion (dof=1), FSILS CGRADS(relTol={HEAT_LS['relTol']}) + diagonal preconditioner"
    elif SOLVER == "ns":
        what = "unsteady VMS Navier-Stokes, FSILS NSSOLVER (BIPN, FSILS defaults) + diagonal preconditioner"
    else:
        what = (f"unsteady VMS Navier-Stokes, FSILS GMRES(sD={LS['dimKry']}, relTol={LS['relTol']}, "
                f"mItr={LS['maxItr']}) + diagonal preconditioner")
    workload = (f"synthetic {nEl_total / 1e6:.2f}M-tet cylinder {dims[0]}x{dims[1]}x{dims[2]} Kuhn lattice, "
                + what)
    config = dict(workload=workload, partition=f"{max(world, 1)} axial slabs",
                  l2="inputs larger than L2 (Val = 128 B x nnz >> 126 MB)", assembly=args.variant,
                  dt=DT, rho=RHO, mu=MU,
                  step=("PICP+PICI, element loop, outlet IntegV + resistance Neumann face (r=%g, ADDBCMUL in the "
                        "solve), COMMU(R), FSILS_SOLVE, PICC (SURVEY.md 8d)" % R_OUT) if PHYSICS == "fluid"
                  else "PICP+PICI, element loop, COMMU(R), FSILS_SOLVE, PICC")
    metric, unit = "fluid_newton_iters_per_sec", "Newton-iter/s"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        times = []
        for _ in range(args.warmup + args.steps):
            times.append(cpu_newton_parallel(args.cpu_nz, args.cpu_procs))
        times = times[args.warmup:] if len(times) > args.warmup else times
        t_full = float(np.mean([t["t_full"] for t in times]))
        s = times[-1]
        sample = (f"{s['procs']} processes at once (one per host core, as mpiexec -np {s['procs']} would; no "
                  f"halo exchange between them), each one Newton iteration (faithful CONSTRUCT_FLUID + "
                  f"FSILS GMRES, same settings) of the oracle port (-O3 -march=native) on a {dims[0]}x{dims[1]}x"
                  f"{args.cpu_nz} slab ({s['nEl']} tets, {s['itr']} SpMVs; slowest: assembly {s['t_asm']:.2f}s, "
                  f"GMRES {s['t_sol']:.2f}s); whole mesh = {s['slabs']:.0f} slabs / {s['procs']} at a time")
        val = 1.0 / t_full
        out = dict(metric=metric, value=val, unit=unit, n_gpus=args.gpus, steps=args.steps,
                   warmup=args.warmup, ms_per_step=t_full * 1e3, higher_is_better=True,
                   scaling=args.scaling, vs_baseline=None, dtype="f64", data="synthetic", config=config,
                   impl="reference",
                   cpu_baseline=dict(value=val, unit=unit, cores=s["procs"], kind="port-extrapolated",
                                     sample=sample),
                   e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ native arm (CUDA)
    import torch
    from svfsi_b200 import api, mesh
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
    if world > 1:
        # stdout carries ONE JSON line: NCCL prints its "NCCL version ..." banner to stdout at every debug
        # level from VERSION up (WARN included).  Drop the level unless more than warnings was asked for,
        # and send whatever NCCL logs to stderr.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        api.init_distributed(device=local)
    else:
        dist = None
        api.init(device=0, rank=0, nranks=1)
    variant = dict(atomic=api.ASM_ATOMIC, colored=api.ASM_COLORED, gather=api.ASM_GATHER)[args.variant]
    t_setup = time.perf_counter()
    gnNo, p = setup_rank(api, mesh, dims, rank, world)
    dof = 1 if PHYSICS == "heat" else 4
    api.state_upload(dof, p.Ag, p.Yg, None)
    api.pic_init(dof, p.Ag, p.Yg)         # old state (Ao, Yo) of the time integrator = the step's linearisation point
    if PHYSICS == "fluid":
        # make the synthetic state consistent with the resistance outlet: p(outlet) = r * Q, as in a developed
        # flow (otherwise the step starts from a 6000 dyn/cm^2 traction jump at the outlet)
        q0 = api.IntegV(3, which=1, s=1)
        p.Yg[:, 3] += R_OUT * q0
        api.state_upload(dof, p.Ag, p.Yg, None)
        api.pic_init(dof, p.Ag, p.Yg)
    api.sync()
    t_setup = time.perf_counter() - t_setup
    nNo, nnz, nEl = p.rm.nNo, p.colPtr.size, p.rm.nEl

    stream = torch.cuda.ExternalStream(api.stream_ptr(), device=torch.device("cuda", local))

    def barrier():
        api.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last = None
        for _ in range(steps):
            last = fn()
        e1.record(stream)
        api.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, last

    # device-resident leg
    for _ in range(args.warmup):
        ls = newton_step_dev(api, variant, p)
    l0 = api.launch_count()
    try:
        dev_uuid = torch.cuda.get_device_properties(local).uuid
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local, dev_uuid)
    if rank == 0:
        sampler.start()
    ms_dev, ls = timed(lambda: newton_step_dev(api, variant, p), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = api.launch_count() - l0
    # second timed pass over the same K steps with CUDA-event pairs around every kernel group
    # (roofline + phase split); kept separate so that the events do not perturb `value`
    api.prof_reset(); api.prof_enable(True)
    ms_prof, _ = timed(lambda: newton_step_dev(api, variant, p), args.steps)
    prof = api.prof_get()
    spmv_bytes_total, spmv_ops_lib = api.prof_spmv()
    api.prof_enable(False)

    # e2e leg: host buffers (pinned), H2D + D2H inside the timed region
    Ag_h = torch.from_numpy(p.Ag).pin_memory().numpy()
    Yg_h = torch.from_numpy(p.Yg).pin_memory().numpy()
    R_h = torch.empty((nNo, dof) if dof > 1 else (nNo,), dtype=torch.float64).pin_memory().numpy()
    for _ in range(max(1, args.warmup - 1)):
        newton_step_e2e(api, variant, p, Ag_h, Yg_h, R_h)
    ms_e2e, ls2 = timed(lambda: newton_step_e2e(api, variant, p, Ag_h, Yg_h, R_h), args.steps)

    # SpMV roofline: event pairs around every SPARMULVV kernel of the timed region
    spmv_ms, spmv_n = prof["spmv"]
    # SURVEY.md 8d, per launch, this rank (NS solver: mixed shapes -> bytes of the K (3x3) SpMV only approx.)
    alg_bytes = nnz * (8 * dof * dof + 4) + nNo * (8 + 16 * dof)
    peak, peak_src = measured_peak()
    # on the NCCL / unfused paths every FSILS_SPARMUL is TWO launches of the same kernel (boundary
    # rows, then interior rows while the halo is in flight); the roofline unit is the whole SpMV
    comm = api.comm_mode()
    per_op = 2 if comm in (1, 2) else 1
    spmv_ops = spmv_n / per_op
    if spmv_ops_lib:    # the library's own byte count (exact for the mixed shapes of NSSOLVER / CG)
        spmv_ops = spmv_ops_lib
        alg_bytes = spmv_bytes_total / spmv_ops_lib
    achieved = alg_bytes / (spmv_ms / max(spmv_ops, 1) * 1e-3) / 1e9 if spmv_n else None
    # DRAM traffic of the same kernel from the committed ncu --set full capture (same mesh only)
    traffic = None
    tp = next((q for q in (os.path.join(ROOT, "profiles", "r02_spmv_traffic.json"),
                           os.path.join(ROOT, "profiles", "r01_spmv_traffic.json")) if os.path.exists(q)), None)
    if tp and dof == 4 and SOLVER == "gmres":
        with open(tp) as fh:
            tj = json.load(fh)
        kname = "spmv_vv4_quad_kernel" if api.spmv_variant() & 1 else "spmv_vv4_kernel"
        if tj.get("nnz") == int(nnz) and tj.get("nNo") == int(nNo) and tj.get("kernel") == kname:
            traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
    asm_ms, asm_n = prof["asm"]
    # PROF_ASM holds the element loop AND the face kernels: one assembly per step
    asm_per_step_ms = asm_ms / args.steps if asm_n else None
    melem = (nEl * world) / (asm_per_step_ms * 1e-3) / 1e6 if asm_n else None
    # compulsory HBM bytes of one assembly (this rank) and the kernel's own FP64 count
    # (ncu smsp__sass_thread_inst_executed_op_dfma/dmul/dadd)
    asm_bytes = (nEl * 16 + nNo * 8 * (3 + 4 + 4 + 3) + nnz * 128 + nNo * 32) if dof == 4 else \
                (nEl * 16 + nNo * 8 * (3 + 1 + 1) + nnz * 8 + nNo * 8)
    asm_flop = nEl * ASM_FLOP_PER_ELEM if dof == 4 else None

    if rank == 0:
        per = ms_dev / args.steps
        out = dict(metric=metric, value=1e3 / per, unit=unit, n_gpus=world, steps=args.steps,
                   warmup=args.warmup, ms_per_step=per, higher_is_better=True, scaling=args.scaling,
                   vs_baseline=None, dtype="f64", data="synthetic", config=config, clocks=clocks,
                   e2e=dict(value=1e3 / (ms_e2e / args.steps), unit=unit,
                            h2d_bytes_per_step=int(Ag_h.nbytes + Yg_h.nbytes),
                            d2h_bytes_per_step=int(R_h.nbytes), ms_per_step=ms_e2e / args.steps),
                   gpu_launches=int(launches), comm=api.COMM_MODES[comm],
                   roofline=dict(bound="hbm", kernel=((("spmv_vv4_quad_fused_kernel" if api.spmv_variant() & 2 else
                                                        "spmv_vv4_fused_kernel") if comm == 3 else
                                                       ("spmv_vv4_quad_kernel" if api.spmv_variant() & 1 else "spmv_vv4_kernel")) +
                                         " (FSILS_SPARMULVV dof=4)" if dof == 4 and SOLVER == "gmres"
                                                    else "spmv kernels (flat-row, mixed shapes: average over the step's SpMVs)"),
                                 achieved=achieved, peak=peak, unit="GB/s",
                                 frac=(achieved / peak) if achieved else None, peak_source=peak_src,
                                 traffic=traffic, algorithmic_bytes_per_launch=int(alg_bytes),
                                 avg_launch_ms=spmv_ms / max(spmv_ops, 1), launches=int(spmv_n),
                                 launches_per_spmv=per_op),
                   detail=dict(nEl_rank0=int(nEl), nNo_rank0=int(nNo), nnz_rank0=int(nnz),
                               gmres_spmv_count=int(ls.RI.itr), gmres_suc=int(ls.RI.suc), gm_itr=int(ls.GM.itr), cg_itr=int(ls.CG.itr),
                               iNorm=ls.RI.iNorm, fNorm=ls.RI.fNorm,
                               assembly_Melem_per_s=melem,
                               phase_ms_per_step={k: v[0] / args.steps for k, v in prof.items()},
                               profiled_pass_ms_per_step=ms_prof / args.steps,
                               setup_s=t_setup))
        if asm_per_step_ms:
            out["roofline_assembly"] = dict(
                ms=asm_per_step_ms, Melem_per_s=melem,
                hbm=dict(compulsory_bytes=int(asm_bytes), achieved=asm_bytes / (asm_per_step_ms * 1e-3) / 1e9,
                         peak=peak, unit="GB/s", frac=asm_bytes / (asm_per_step_ms * 1e-3) / 1e9 / peak),
                fp64=(dict(flop=int(asm_flop), achieved=asm_flop / (asm_per_step_ms * 1e-3) / 1e12,
                           peak=FP64_PEAK_TFLOPS, unit="TFLOP/s",
                           frac=asm_flop / (asm_per_step_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                           peak_source="64 DFMA per clock per SM x 148 SMs x 1.965 GHz (nominal)")
                      if asm_flop else None))
        if world == 1 and not args.no_cpu:
            c = cpu_newton_parallel(args.cpu_nz, args.cpu_procs)
            out["cpu_baseline"] = dict(
                value=1.0 / c["t_full"], unit=unit, cores=c["procs"], kind="port-extrapolated",
                sample=(f"oracle port (-O3 -march=native), {c['procs']} processes at once (one per host "
                        f"core, no halo exchange), each one Newton iteration on a {dims[0]}x{dims[1]}x"
                        f"{args.cpu_nz} slab ({c['nEl']} tets; slowest: assembly {c['t_asm']:.2f}s, GMRES "
                        f"{c['t_sol']:.2f}s / {c['itr']} SpMVs); whole mesh = {c['slabs']:.0f} slabs / "
                        f"{c['procs']} at a time"))
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    api.finalize()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
