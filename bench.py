#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ice-dynamics hot path (contract: see the task statement).

Workload (BASELINE.json configs[2]): SSA ice stream/shelf on a flat synthetic bed, ~1M-vertex mesh (4M AaAc
vertices), MISMIP_mod physics switches with the analytical grounding-line flux, through the region time loop
(thickness update -> general data -> SIA -> SSA viscosity/SOR -> CFL).  A "step" is one pass of that loop.

  value  model-years per wall-hour, state resident in HBM                      (ufm_run_model)
  e2e    the same steps in drop-in mode: host buffers, H2D/D2H every step      (ufm_run_model_host)
  roofline  the SOR sweep kernel: algorithmic bytes sum_i(80+20 n_i) per iteration / CUDA-event time
  cpu_baseline  the CPU oracle (restatement of the Fortran, OpenMP threads as MPI ranks) on a bounded sample

`--impl reference` times that CPU restatement alone (the Fortran reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "model_yr_per_wall_hr"
UNIT = "model-yr/wall-hr"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_workload(nv, seed=20211103, order="random"):
    """BASELINE configs[2].  `order` is the HOST vertex numbering the reference-layout arrays are built in: "random" mimics the poor index
    locality of refinement-ordered meshes (the default, as in round 1), "morton" is a locality-preserving numbering (the CPU arm's best case).
    The device path renumbers at upload and does not care."""
    from ufemism_b200 import mesh as M
    from ufemism_b200 import scenarios as S

    c = S.CONFIG3
    t = time.time()
    m = M.square_mesh_with_nv(c["half_width"], nv, seed=seed, order=order)
    st = S.state_ssa_icestream(m, scale=1.0, Hb=c["Hb"], H_shelf=c["H_shelf"])
    log(f"[bench] mesh nV={m.nV} nAc={m.nAc} nVAaAc={m.nVAaAc} (host vertex order: {order}) built in {time.time() - t:.1f}s")
    return m, st


def workload_name(m):
    return f"config3_ssa_icestream_flatbed_nV{m.nV}_AaAc{m.nVAaAc}_MISMIP_mod_GLflux"


def bench_config(m, args, world):
    """The `config` object of the JSON line -- the SAME function in both arms, so the driver's same_config check compares like with like."""
    part = world > 1 and args.multi == "partition"
    return {"workload": workload_name(m),
            "parallelism": (f"one region, vertex-partitioned into {world} x-strips (one per GPU), NVLink P2P halo pushes" if part
                            else f"{world} independent region(s), one per GPU (no data-path collective)"),
            "flush": "inputs larger than L2 (SOR streams ~1 GB of coefficients per iteration)", "exact_xy": int(args.exact_xy),
            "host_vertex_order": args.order}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a background thread every
    10 ms (the library calls release the GIL), nvidia-smi -lms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        import threading

        self.p = None
        self.thread = None
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid_order = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(uuid_order.split(",")[index]) if uuid_order and all(t.strip().isdigit() for t in uuid_order.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self._sample()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in self.BITS.items():
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self.stop_flag.wait(0.01):
            try:
                self._sample()
            except Exception:
                break

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = self.sm[1:] if len(self.sm) > 1 else self.sm      # the first sample predates the timed region
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 10 ms period"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def ssa_solve_time(step_ms, rows):
    """"SSA solve time / step" (BASELINE metric) from per-step device times: mean time of the steps that ran solve_SSA minus that of the
    steps that did not.  Reporting only: any problem yields an empty dict, never an exception."""
    try:
        w_ = [t for t, x in zip(step_ms, rows) if x["ssa"]]
        wo = [t for t, x in zip(step_ms, rows) if not x["ssa"]]
        d = {"ms_per_step_with_ssa_solve": sum(w_) / len(w_) if w_ else None,
             "ms_per_step_without_ssa_solve": sum(wo) / len(wo) if wo else None,
             "step_ms": [round(float(t), 3) for t in step_ms]}
        if w_ and wo:
            n_solves = max(sum(int(x["ssa"]) for x in rows), 1)
            d["ms_per_ssa_solve"] = sum(w_) / len(w_) - sum(wo) / len(wo)
            d["n_outer_per_solve"] = sum(x["n_outer"] for x in rows) / n_solves
            d["n_sor_per_solve"] = sum(x["n_sor"] for x in rows) / n_solves
        return d
    except Exception:  # noqa: BLE001
        return {}


# ------------------------------------------------------------------------------------------------
# CPU restatement run FOR REAL: the same region loop, step by step, on all host cores -- the CPU arm and the full-size parity check
# ------------------------------------------------------------------------------------------------
PARITY_FIELDS = ("Hi", "U_SSA", "V_SSA")
PARITY_GATES = {"rel_l2_U": 1e-10, "rel_l2_V": 1e-10, "rel_l2_Hi": 1e-8}   # north_star; stop tests at ice_dynamics_module.f90:601-691


def cpu_trajectory(m, st, nthreads, n_warm, n_steps):
    """n_warm + n_steps passes of ora_run_model (max_steps = 1) from the workload's start state.  Returns per-step wall seconds and
    iteration counts, the model time, and the final Hi / U_SSA / V_SSA (reference vertex order)."""
    from oracle.oracle import Oracle

    o = Oracle(m, benchmark=st["benchmark"], nthreads=nthreads, use_analytical_GL_flux=1)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    r = o.region(0.0)
    rows, secs = [], []
    t_warm_end = 0.0
    for k in range(n_warm + n_steps):
        a = (r.n_sia, r.n_ssa, r.n_outer_total, r.n_sor_total)
        t = time.perf_counter()
        o.run_model(r, 1e12, max_steps=1)
        secs.append(time.perf_counter() - t)
        rows.append(dict(dt=r.dt, sia=int(r.n_sia - a[0]), ssa=int(r.n_ssa - a[1]), n_outer=int(r.n_outer_total - a[2]), n_sor=int(r.n_sor_total - a[3])))
        if k == n_warm - 1:
            t_warm_end = r.time
    return {"rows": rows, "step_s": secs, "time": r.time, "time_after_warmup": t_warm_end,
            "fields": {f: np.array(o[f], dtype=np.float64, copy=True) for f in PARITY_FIELDS}}


def field_digest(fields):
    import hashlib

    return {f: hashlib.sha256(np.ascontiguousarray(fields[f]).tobytes()).hexdigest()[:16] for f in PARITY_FIELDS}


def parity_report(gpu_fields, gpu_rows, gpu_time, cpu):
    """GPU trajectory vs the CPU restatement's on the same mesh and inputs after the same steps (north_star gates)."""
    def rel(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    cr = cpu["rows"]
    n = min(len(cr), len(gpu_rows))
    out = {"rel_l2_U": rel(gpu_fields["U_SSA"], cpu["fields"]["U_SSA"]), "rel_l2_V": rel(gpu_fields["V_SSA"], cpu["fields"]["V_SSA"]),
           "rel_l2_Hi": rel(gpu_fields["Hi"], cpu["fields"]["Hi"]),
           "max_abs_dHi_m": float(np.max(np.abs(gpu_fields["Hi"] - cpu["fields"]["Hi"]))),
           "n_sor_equal": all(cr[k]["n_sor"] == gpu_rows[k]["n_sor"] for k in range(n)),
           "n_outer_equal": all(cr[k]["n_outer"] == gpu_rows[k]["n_outer"] for k in range(n)),
           "dt_equal": all(cr[k]["dt"] == gpu_rows[k]["dt"] for k in range(n)),
           "steps_compared": n, "model_time_equal": bool(cpu["time"] == gpu_time),
           "bit_identical": {f: bool(np.array_equal(gpu_fields[f], cpu["fields"][f])) for f in PARITY_FIELDS},
           "gates": PARITY_GATES, "umax_m_per_yr": float(np.max(np.abs(cpu["fields"]["U_SSA"])))}
    out["passed"] = bool(all(out[k] <= g for k, g in PARITY_GATES.items()) and out["n_sor_equal"] and out["n_outer_equal"] and out["dt_equal"])
    out["rel_Hi"] = out["rel_l2_Hi"]   # the name the round-1 review used for the same number
    return out


def cpu_line(traj, n_warm, nthreads, m, what):
    timed_s = float(sum(traj["step_s"][n_warm:]))
    yrs = traj["time"] - traj["time_after_warmup"]
    n = max(len(traj["step_s"]) - n_warm, 1)
    return {"value": yrs / timed_s * 3600.0, "unit": UNIT, "cores": nthreads, "kind": "port", "ms_per_step": timed_s / n * 1e3,
            "sample": (f"{what}: the oracle's region loop run for real on the {m.nV}-vertex workload, {n_warm} warm-up + {n} timed steps "
                       f"({sum(traj['step_s']):.1f} s of CPU work on {nthreads} threads), same steps as the GPU arm"),
            "model_years": yrs, "n_sor": int(sum(x["n_sor"] for x in traj["rows"][n_warm:])), "n_outer": int(sum(x["n_outer"] for x in traj["rows"][n_warm:]))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nthreads = os.cpu_count() or 1
    m, st = build_workload(args.nv, order=args.order)
    t = time.time()
    traj = cpu_trajectory(m, st, nthreads, args.warmup, args.steps)
    log(f"[bench] CPU trajectory: {args.warmup}+{args.steps} steps in {time.time() - t:.1f}s")
    base = cpu_line(traj, args.warmup, nthreads, m, "CPU restatement (oracle/)")
    # for the record / the GPU arm's cross-check: final fields of this trajectory
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", "reference_arm_trajectory.npz"), workload=workload_name(m), order=args.order, warmup=args.warmup, steps=args.steps,
                 rows=json.dumps(traj["rows"]), step_s=np.array(traj["step_s"]), time=traj["time"], **traj["fields"])
    except Exception as ex:  # noqa: BLE001
        log(f"[bench] could not write the trajectory file: {ex}")
    base["final_field_sha256_16"] = field_digest(traj["fields"])
    # the same job on the other host vertex numbering (locality-preserving if the main one is random and vice versa): the restatement keeps the
    # reference's strided column-major ELL rows, so its speed depends on the numbering the mesh generator happened to produce
    other = "morton" if args.order == "random" else "random"
    if not args.no_other_order:
        try:
            n2 = min(args.steps, 6)        # a shorter window keeps this arm within a few minutes; compared with the main run over the same steps
            m2, st2 = build_workload(args.nv, order=other)
            t = time.time()
            traj2 = cpu_trajectory(m2, st2, nthreads, args.warmup, n2)
            log(f"[bench] CPU trajectory ({other} order): {time.time() - t:.1f}s")
            same_window = {"rows": traj["rows"][: args.warmup + n2], "step_s": traj["step_s"][: args.warmup + n2], "time_after_warmup": traj["time_after_warmup"],
                           "time": traj["time_after_warmup"] + sum(x["dt"] for x in traj["rows"][args.warmup: args.warmup + n2])}
            base["other_host_vertex_order"] = dict(cpu_line(traj2, args.warmup, nthreads, m2, f"same job, host vertex order '{other}'"), order=other,
                                                   main_order_over_the_same_steps=cpu_line(same_window, args.warmup, nthreads, m, f"host vertex order '{args.order}'")["value"])
            del m2, st2, traj2
        except Exception as ex:  # noqa: BLE001
            base["other_host_vertex_order"] = {"error": str(ex)}
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong" if (world > 1 and args.multi == "partition") else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(m, args, world),
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
           "note": "CPU restatement of the Fortran hot path (oracle/), bit-identical to the reference's own source run through oracle/f90py.py (tests/test_reference_source.py); the Fortran reference itself cannot be built here (no gfortran/MPI/NetCDF)"}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# Per-step streaming kernels: algorithmic bytes of SURVEY 8(d) (from the actual degree sums) / CUDA-event time
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(m):
    N, E, M = m.nV, m.nAc, m.nVAaAc
    dAa, dM = float(np.sum(m.nC)), float(np.sum(m.nCAaAc))
    return {"geom": 200.0 * N + 20.0 * dAa + 290.0 * E,          # K-GEOM: one update_general_ice_model_data (benchmark flow factor)
            "sia": 84.0 * E + 28.0 * N + 4.0 * dAa,              # K-SIA: D_SIA_3D kept in registers, scalar A
            "thk": 60.0 * N + 44.0 * dAa,                        # K-THK
            "cfl": 24.0 * E + 40.0 * N,                          # K-CFL (no thermodynamics)
            "visc": 80.0 * M + 20.0 * dM + 130.0 * M}            # K-VISC (+RN) and K-SLID+LIN: one fused launch here


def kernel_rooflines(g, m, torch, stream, peak, reps=12):
    """Every per-step routine launched `reps` times round robin (so each one finds the L2 full of the others' data: ~2 GB go through
    between two launches of the same routine), each call bracketed by CUDA events on the library's stream, nothing else in between."""
    calls = {"geom": lambda: g.update_general_ice_model_data(0.0), "sia": g.solve_SIA, "thk": lambda: g.calculate_ice_thickness_change(1e-3),
             "cfl": g.determine_timesteps, "visc": g.ssa_viscosity}
    g.update_general_ice_model_data(0.0); g.ssa_prepare()
    ev = {k: [] for k in calls}
    for rep in range(reps + 2):
        for k, fn in calls.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            if rep >= 2:
                ev[k].append((e0, e1))
    torch.cuda.synchronize()
    B = algorithmic_bytes(m)
    out = {}
    for k, pairs in ev.items():
        ms = float(np.median([a.elapsed_time(b) for a, b in pairs]))
        gbs = B[k] / (ms * 1e-3) / 1e9
        out[k] = {"ms": ms, "algorithmic_bytes": B[k], "achieved_GBps": gbs, "frac": gbs / peak}
    out["how"] = (f"median of {reps} launches per routine, round robin, CUDA events around each call on the library's stream; thk with dt = 1e-3 yr (a real update: in-fluxes gather their source's out-flux factor, which a dt of 0 would skip); "
                  "cfl includes its 3-scalar device-to-host read; visc = viscosity + RN partials + sliding term + linear-system setup in one launch (+ the 64-CTA sum)")
    return out


# ------------------------------------------------------------------------------------------------
# Legs after the headline: the other BASELINE configs that fit one GPU, each with its own parity / accuracy figure (reporting only)
# ------------------------------------------------------------------------------------------------
def _timed(torch, stream, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); t = time.perf_counter(); fn(); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1), (time.perf_counter() - t) * 1e3


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def leg_config1_eismint(torch, stream, local, nthreads, nv=10000, years=10000.0):
    """BASELINE configs[0]: EISMINT-1 moving margin, SIA only, ~10 k vertices, ice-free start, 10 kyr (the run behind the reference's only
    published timing: "about 20 seconds" on 2 cores, documentation/UFEMISM_documentation.tex:70).  Thermodynamics does not feed back on the
    dynamics in this experiment (constant flow factor) and is left out on both sides."""
    from oracle.oracle import Oracle
    from ufemism_b200 import mesh as M, scenarios as S
    from ufemism_b200.capi import IceModelGPU

    m = M.square_mesh_with_nv(750e3, nv)
    st = S.state_eismint1(m)
    g = IceModelGPU(m, benchmark=st["benchmark"], device=local)
    g.set_stream(stream.cuda_stream)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    r = g.region(0.0)
    g.reset_counters()
    ms_dev, ms_wall = _timed(torch, stream, lambda: g.run_model(r, years))
    launches = int(g.counters().kernel_launches)
    Hi = g.download("Hi")
    o = Oracle(m, benchmark=st["benchmark"], nthreads=nthreads)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    ro = o.region(0.0)
    t = time.perf_counter(); o.run_model(ro, years); cpu_s = time.perf_counter() - t
    g.close()
    return {"workload": f"config1_EISMINT1_moving_margin_SIA_nV{m.nV}_{years:g}yr", "model_years": years, "steps": int(r.n_steps), "wall_s": ms_wall * 1e-3,
            "device_s": ms_dev * 1e-3, "ms_per_step": ms_wall / max(r.n_steps, 1), "model_yr_per_wall_hr": years / (ms_wall * 1e-3) * 3600.0, "gpu_launches": launches,
            "cpu_restatement": {"wall_s": cpu_s, "cores": nthreads, "steps": int(ro.n_steps), "model_yr_per_wall_hr": years / cpu_s * 3600.0},
            "reference_published": "about 20 seconds for this run on 2 cores (documentation/UFEMISM_documentation.tex:70; other hardware, incl. thermodynamics and output)",
            "parity": {"rel_l2_Hi": _rel(Hi, o["Hi"]), "steps_equal": bool(r.n_steps == ro.n_steps), "time_equal": bool(r.time == ro.time), "gate_rel_l2_Hi": 1e-8,
                       "dome_height_m": float(np.max(Hi))}}


def leg_config2_halfar(torch, stream, local, nthreads, nv=250000, warm=40, timed=200, span=100.0):
    """BASELINE configs[1]: Halfar dome on a ~250 k-vertex mesh.  Started from Halfar_solution(1000 yr) (at t = 0 the as-coded diffusivity clip
    is active and the model deliberately departs from the similarity solution).  Reports ms / step over `timed` steps,
    the error against Halfar_solution after `span` years (src/reference_fields_module.f90:707-745) and parity with the oracle after `warm` steps."""
    from oracle.oracle import Oracle
    from ufemism_b200 import mesh as M, scenarios as S
    from ufemism_b200.capi import IceModelGPU

    m = M.square_mesh_with_nv(750e3, nv)
    t0 = 1000.0
    st = S.state_halfar(m, t=t0)
    g = IceModelGPU(m, benchmark="Halfar", device=local)
    g.set_stream(stream.cuda_stream)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, st[k])
    r = g.region(0.0)
    g.run_model(r, 1e12, max_steps=warm)
    Hi_w = g.download("Hi")
    o = Oracle(m, benchmark="Halfar", nthreads=nthreads)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        o[k][:] = st[k]
    ro = o.region(0.0)
    t = time.perf_counter(); o.run_model(ro, 1e12, max_steps=warm); cpu_s = time.perf_counter() - t
    par = {"rel_l2_Hi": _rel(Hi_w, o["Hi"]), "steps": warm, "time_equal": bool(r.time == ro.time), "gate_rel_l2_Hi": 1e-8}
    t_a = r.time
    g.reset_counters()
    ms_dev, ms_wall = _timed(torch, stream, lambda: g.run_model(r, 1e12, max_steps=timed))
    launches = int(g.counters().kernel_launches)
    yrs = r.time - t_a
    g.run_model(r, span)                       # on to exactly `span` model years
    Hi = g.download("Hi")
    x, y = m.V[:, 0], m.V[:, 1]
    exact, start = S.halfar_H(5000.0, 300000.0, x, y, t0 + span), st["Hi"]
    g.close()
    return {"workload": f"config2_Halfar_dome_SIA_nV{m.nV}", "timed_steps": timed, "ms_per_step": ms_wall / timed, "device_ms_per_step": ms_dev / timed, "model_years_timed": yrs,
            "model_yr_per_wall_hr": yrs / (ms_wall * 1e-3) * 3600.0, "gpu_launches_per_step": launches / timed,
            "cpu_restatement": {"ms_per_step": cpu_s / warm * 1e3, "cores": nthreads, "steps": warm},
            "vs_Halfar_solution": {"years": span, "rel_l2_error": _rel(Hi, exact), "rel_l2_change_of_the_solution": _rel(exact, start),
                                   "dome_height_model_m": float(np.max(Hi)), "dome_height_exact_m": float(np.max(exact)), "steps_total": int(r.n_steps)},
            "parity": par}


def mismip_dome_state(V, edge_index):
    """MISMIP_mod bed (src/reference_fields_module.f90:549-564) under a grounded dome with a thin shelf ring: a marine ice sheet with a
    grounding line, as a 100-yr window of a spun-up MISMIP-style run would see it (the benchmark's own start, 100 m of ice, is inert)."""
    r = np.hypot(V[:, 0], V[:, 1])
    Hi = np.where(r < 600e3, 800.0 - 600.0 * r / 600e3, 0.0) + np.where((r >= 600e3) & (r < 680e3), 150.0, 0.0)
    Hi[edge_index > 0] = 0.0
    return dict(Hi=Hi, Hb=720.0 - 778.5 * r / 750e3, SL=np.zeros(len(r)), SMB_year=np.full(len(r), 0.3), BMB=np.zeros(len(r)))


def leg_config4_mismip(torch, stream, local, nv=1000000, years=100.0, t_update=50.0):
    """BASELINE configs[3]: hybrid SIA/SSA marine ice sheet on the MISMIP_mod bed with a moving grounding line (analytical GL flux), a
    `years`-long window with ONE mesh update at t_update: the host hands over a NEW mesh as primary data (as mesh_update_module does,
    src/UFEMISM_main_model.f90:240-315), the thickness is remapped with ufm_remap_stash / _apply and the library re-derives every secondary
    mesh array (ufm_mesh_upload_primary).  The re-upload is timed separately; building the new mesh and the remapping weights is host work
    of the reference's mesh generator (not part of the path) and is reported but not counted."""
    from scipy.spatial import cKDTree
    from ufemism_b200 import mesh as M
    from ufemism_b200.capi import IceModelGPU

    t = time.perf_counter()
    mA = M.primary_mesh_with_nv(750e3, nv, seed=20211103)
    mB = M.primary_mesh_with_nv(750e3, nv, seed=20211104)
    _, nn = cKDTree(mA.V).query(mB.V)          # stand-in for create_remapping_arrays_conservative: nearest old vertex, weight 1
    host_mesh_s = time.perf_counter() - t
    nB = mB.nV
    vli = np.arange(1, nB + 1, dtype=np.int32)
    sa, sb = mismip_dome_state(mA.V, mA.edge_index), mismip_dome_state(mB.V, mB.edge_index)
    g = IceModelGPU(mA, benchmark="MISMIP_mod", device=local, primary_only=True, use_analytical_GL_flux=1)
    g.set_stream(stream.cuda_stream)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        g.upload(k, sa[k])
    r = g.region(0.0)
    g.reset_counters()
    ms1, wall1 = _timed(torch, stream, lambda: g.run_model(r, t_update))
    steps1, sor1 = int(r.n_steps), int(r.n_sor_total)

    def mesh_update():
        g.remap_stash("Hi")
        g.upload_mesh(mB)
        g.set_stream(stream.cuda_stream)
        g.remap_apply("Hi", vli, vli, (nn + 1).astype(np.int32), np.ones(nB))
        for k in ("Hb", "SL", "SMB_year", "BMB"):
            g.upload(k, sb[k])
    ms_u, wall_u = _timed(torch, stream, mesh_update)
    r.do_[0] = r.do_[1] = 1                    # region%do_SIA / do_SSA after a mesh update (src/UFEMISM_main_model.f90:304-311)
    ms2, wall2 = _timed(torch, stream, lambda: g.run_model(r, years))
    Hi = g.download("Hi")
    launches = int(g.counters().kernel_launches)
    g.close()
    total_ms = wall1 + wall_u + wall2
    return {"workload": f"config4_MISMIP_mod_hybrid_SIA_SSA_GLflux_nV{mA.nV}_{years:g}yr_one_mesh_update", "model_years": years, "steps": int(r.n_steps),
            "n_ssa_solves": int(r.n_ssa), "n_sor": int(r.n_sor_total), "model_yr_per_wall_hr": years / (total_ms * 1e-3) * 3600.0,
            "model_yr_per_wall_hr_without_the_mesh_update": years / ((wall1 + wall2) * 1e-3) * 3600.0,
            "ms_per_step": (wall1 + wall2) / max(r.n_steps, 1), "wall_s": {"before_update": wall1 * 1e-3, "mesh_update_device_side": wall_u * 1e-3, "after_update": wall2 * 1e-3},
            "mesh_update": {"at_year": t_update, "new_mesh_nV": nB, "remap": "1st-order, nearest old vertex (weights built on the host)",
                            "device_side_s": wall_u * 1e-3, "share_of_wall": wall_u / total_ms, "host_mesh_generation_and_weights_s_not_counted": host_mesh_s},
            "steps_before_update": steps1, "n_sor_before_update": sor1, "gpu_launches": launches,
            "sanity": {"Hi_max_m": float(np.max(Hi)), "Hi_finite": bool(np.isfinite(Hi).all())}}


def leg_warm_ssa_solve(torch, stream, g, r, max_extra_steps=200):
    """"SSA solve time per step" from a warm state: the headline window starts cold and every solve_SSA in it hits the cap of
    C%SSA_max_outer_loops viscosity iterations.  The same run is continued until a solve converges by the RN test (fewer outer iterations
    than the cap); that solve's device time and counts are reported."""
    cap = g.P.SSA_max_outer_loops
    last = None
    for k in range(max_extra_steps):
        a = (r.n_ssa, r.n_outer_total, r.n_sor_total)
        ms, _ = _timed(torch, stream, lambda: g.run_model(r, 1e12, max_steps=1))
        if r.n_ssa > a[0]:
            last = {"ms_step_with_this_solve": ms, "n_outer": int(r.n_outer_total - a[1]), "n_sor": int(r.n_sor_total - a[2]), "extra_step": k + 1, "model_time": float(r.time)}
            if last["n_outer"] < cap:
                return dict(last, converged_by_RN=True, cap=int(cap))
    return dict(last or {}, converged_by_RN=False, cap=int(cap), note=f"no solve converged below the cap within {max_extra_steps} further steps")


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ufemism_b200 import scenarios as S
    from ufemism_b200.capi import Counters, IceModelGPU, Region

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1, --multi partition (default): ONE region, SSA solve vertex-partitioned into x-strips over the N GPUs with NVLink
    #   peer pushes after every colour sweep (north_star / SURVEY 8e); total work fixed -> "strong" scaling.
    # N > 1, --multi regions: one independent region per GPU (the reference runs NAM/EAS/GRL/ANT independently,
    #   src/UFEMISM_program.f90:194-229); per-GPU work fixed -> "weak" scaling, no data-path communication.
    part = world > 1 and args.multi == "partition"
    dev = torch.device("cuda", local)
    m, st = build_workload(args.nv, order=args.order)
    t = time.time()
    g = IceModelGPU(m, benchmark=st["benchmark"], device=local, rank=rank if part else 0, nranks=world if part else 1,
                    use_analytical_GL_flux=S.CONFIG3["use_analytical_GL_flux"], exact_xy=args.exact_xy)
    if part:
        g.connect(dist, device=dev)
    log(f"[bench] rank {rank}: mesh upload {time.time() - t:.1f}s")
    stream = torch.cuda.Stream()  # a non-default stream: the library launches on it, torch events time it
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)

    def fresh_state():
        g.upload_mesh(m)  # re-upload: all state zero
        if part:
            g.connect(dist, device=dev)
        g.set_stream(stream.cuda_stream)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
            g.upload(k, st[k])
        return g.region(0.0)

    def one_by_one(r, n, host=None, marks=None):
        """`marks`: a list that receives one CUDA event before the first step and one after every step (per-step device times for the
        "SSA solve time per step" report; the events cost nothing measurable and do not change what e0 / e1 bracket)."""
        rows = []

        def mark():
            if marks is not None:
                try:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(stream)
                    marks.append(ev)
                except Exception:  # noqa: BLE001  (a missing mark only drops the per-step report)
                    pass

        mark()
        for _ in range(n):
            a = (r.n_sia, r.n_ssa, r.n_outer_total, r.n_sor_total)
            if host is None:
                g.run_model(r, 1e12, max_steps=1)
            else:
                g.run_model_host(r, 1e12, 1, host)
            rows.append(dict(dt=r.dt, sia=int(r.n_sia - a[0]), ssa=int(r.n_ssa - a[1]), n_outer=int(r.n_outer_total - a[2]), n_sor=int(r.n_sor_total - a[3])))
            mark()
        return rows

    # ---------------- device-resident run: `value` ----------------
    r = fresh_state()
    warm_rows = one_by_one(r, args.warmup)
    g.reset_counters()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_model0 = r.time
    marks = []
    e0.record(stream)
    rows = one_by_one(r, args.steps, marks=marks)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    step_ms = None
    try:   # per-step device times (reporting only; never allowed to take the bench line down)
        step_ms = [marks[k].elapsed_time(marks[k + 1]) for k in range(len(marks) - 1)]
        if len(step_ms) != len(rows):
            step_ms = None
    except Exception as ex:  # noqa: BLE001
        log(f"[bench] per-step times unavailable: {ex}")
        step_ms = None
    clocks = sampler.stop() if sampler else None
    cnt = g.counters()
    yrs = r.time - t_model0
    gpu_time_final = r.time
    # after the timed region: what the parity check compares (partitioned run: a collective gather by owner)
    gpu_fields = {f: (g.download_global(dist, f, device=dev) if part else g.download(f)) for f in PARITY_FIELDS} if (rank == 0 or part) else None
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    nreg = 1 if part or world == 1 else world   # independent regions add up; a partitioned region is one job
    value = nreg * yrs / (ms * 1e-3) * 3600.0

    # ---------------- drop-in mode with host buffers: `e2e` ----------------
    from ufemism_b200.capi import HostIce

    nV = m.nV
    hb = {n: np.zeros(nV) for n in ("Hi", "Hb", "SL", "dHb_dt", "SMB_year", "BMB", "Hi_prev", "dHi_dt", "Hs", "U_SSA", "V_SSA", "U_SIA", "V_SIA", "D_SIA")}
    hb["mask_noice"] = np.zeros(nV, np.int32); hb["mask"] = np.zeros(nV, np.int32)
    for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
        hb[k][:] = st[k]
    host = HostIce(**{n: hb[n].ctypes.data for n in ("Hb", "SL", "dHb_dt", "SMB_year", "BMB", "mask_noice", "Hi_prev", "dHi_dt", "Hs", "U_SSA", "V_SSA", "U_SIA", "V_SIA", "D_SIA", "mask")},
                   Hi=hb["Hi"].ctypes.data, Hi_out=hb["Hi"].ctypes.data)  # the host's Hi window is read and written in place
    for a_ in hb.values():
        g.host_register(a_)   # what the Fortran shim does once for its shared-memory windows
    r2 = fresh_state()
    one_by_one(r2, args.warmup, host)
    g.reset_counters()
    barrier()
    t2_0 = r2.time
    e0.record(stream)
    one_by_one(r2, args.steps, host)
    e1.record(stream)
    barrier()
    ms2 = e0.elapsed_time(e1)
    cnt2 = g.counters()
    if world > 1:
        tt = torch.tensor([ms2], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms2 = float(tt.item())
    e2e_value = nreg * (r2.time - t2_0) / (ms2 * 1e-3) * 3600.0
    same = abs(r2.time - r.time) <= 1e-9 * max(1.0, abs(r.time))

    # steady-state SOR kernel alone (forced iteration count, no stop test), both cross-term modes, for the roofline discussion
    sor_forced = {}
    if not part:
        for mode in (1, 0):
            g.set_params(exact_xy=mode)
            g.ssa_sor(max_inner=5, force_iters=True)
            g.reset_counters()
            g.ssa_sor(max_inner=100, force_iters=True)
            c_ = g.counters()
            us = c_.sor_ms * 1e3 / max(c_.sor_iterations, 1)
            sor_forced[f"exact_xy={mode}"] = {"us_per_iteration": us, "achieved_GBps": c_.sor_bytes_per_iteration / (us * 1e-6) / 1e9}
        g.set_params(exact_xy=int(args.exact_xy))

    # N > 1: also time the communication-free alternative (one independent region per GPU) for the record
    regions = None
    if part and not args.no_regions:
        g.close()
        g = IceModelGPU(m, benchmark=st["benchmark"], device=local, use_analytical_GL_flux=S.CONFIG3["use_analytical_GL_flux"], exact_xy=args.exact_xy)
        g.set_stream(stream.cuda_stream)
        for k in ("Hi", "Hb", "SL", "SMB_year", "BMB"):
            g.upload(k, st[k])
        r3 = g.region(0.0)
        one_by_one(r3, args.warmup)
        barrier()
        t3 = r3.time
        e0.record(stream)
        one_by_one(r3, args.steps)
        e1.record(stream)
        barrier()
        ms3 = e0.elapsed_time(e1)
        tt = torch.tensor([ms3], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms3 = float(tt.item())
        regions = {"value": world * (r3.time - t3) / (ms3 * 1e-3) * 3600.0, "unit": UNIT, "ms_per_step": ms3 / args.steps, "scaling": "weak",
                   "what": f"{world} independent regions of the same workload, one per GPU, no communication (aggregate model-years)"}
        # the un-partitioned pass doubles as the in-run check of the partitioned one: same steps, same bits
        if rank == 0:
            same_bits = {f: bool(np.array_equal(g.download(f), gpu_fields[f])) for f in PARITY_FIELDS}
            regions["partitioned_bit_identical"] = bool(all(same_bits.values()) and r3.time == gpu_time_final and r3.n_sor_total == r.n_sor_total and r3.n_outer_total == r.n_outer_total)
            regions["partitioned_vs_single_gpu"] = {"fields_bit_identical": same_bits, "model_time_equal": bool(r3.time == gpu_time_final),
                                                    "n_sor": [int(r.n_sor_total), int(r3.n_sor_total)], "n_outer": [int(r.n_outer_total), int(r3.n_outer_total)]}

    # N > 1: BASELINE configs[4], the size the partition is for (~4 M vertices), on the same ranks
    config5 = None
    if part and not args.no_config5:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import config5_probe

            g.close()
            config5 = config5_probe.run(args.nv5, dist=dist, torch=torch, rank=rank, world=world, local=local, log=log)
            ref1 = os.path.join(ROOT, "profiles", "config5_4M_1gpu_r02.json")
            if os.path.exists(ref1):
                one = json.load(open(ref1))
                if one.get("workload") == config5["workload"]:
                    config5["one_gpu"] = {k: one[k] for k in ("ms_per_step", "sor_us_per_iteration_forced", "sor_us_per_iteration_in_solve") if k in one}
                    config5["one_gpu"]["source"] = "profiles/config5_4M_1gpu_r02.json (same script on one B200, this round)"
                    config5["speedup_sor_iteration"] = one["sor_us_per_iteration_forced"] / config5["sor_us_per_iteration_forced"]
                    config5["efficiency_sor_iteration"] = config5["speedup_sor_iteration"] / world
                    config5["speedup_step"] = one["ms_per_step"] / config5["ms_per_step"]
                    config5["efficiency_step"] = config5["speedup_step"] / world
        except Exception as ex:  # noqa: BLE001  (reporting only)
            config5 = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        peak, peak_src = hbm_peak()
        if part:
            peak, peak_src = peak * world, peak_src + f" x {world} GPUs"
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "sor_traffic_r02.json")
        if os.path.exists(tf):
            tj = json.load(open(tf))
            if abs(tj["nVAaAc"] - m.nVAaAc) <= 0.02 * m.nVAaAc and int(args.exact_xy) == 1:
                # ncu dram__bytes_read+write per SOR iteration x mean iterations per launch of this run
                traffic = tj["dram_bytes_per_iteration"] * cnt.sor_iterations / max(cnt.sor_launches, 1)
                traffic_src = "ncu --set full capture of 10 forced iterations (profiles/sor_traffic_r02.json) scaled to this run's mean iterations per launch"
        t_iter = cnt.sor_ms * 1e-3 / max(cnt.sor_iterations, 1)
        achieved = cnt.sor_bytes_per_iteration / t_iter / 1e9 if cnt.sor_iterations else 0.0
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if part else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": bench_config(m, args, world),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": cnt2.h2d_bytes / args.steps, "d2h_bytes_per_step": cnt2.d2h_bytes / args.steps,
                       "ms_per_step": ms2 / args.steps, "same_trajectory_as_value": bool(same)},
               "gpu_launches": int(cnt.kernel_launches),
               "roofline": {"bound": "hbm", "kernel": "k_ssa_sor (five-colour SOR sweep, persistent cooperative)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": cnt.sor_bytes_per_iteration * cnt.sor_iterations / max(cnt.sor_launches, 1), "peak_source": peak_src, "algorithmic_bytes_per_iteration": cnt.sor_bytes_per_iteration,
                            "us_per_iteration": t_iter * 1e6, "iterations": int(cnt.sor_iterations), "launches": int(cnt.sor_launches),
                            "sor_share_of_step": cnt.sor_ms / ms},
               "ssa": {"model_years": yrs, "n_ssa_solves": int(sum(x["ssa"] for x in rows)), "n_outer": int(sum(x["n_outer"] for x in rows)),
                       "n_sor": int(sum(x["n_sor"] for x in rows))}}
        if step_ms:
            out["ssa"].update(ssa_solve_time(step_ms, rows))
        if config5:
            out["config5_4M"] = config5
        if regions:
            out["partitioned_bit_identical"] = regions.pop("partitioned_bit_identical", None)
            out["partitioned_vs_single_gpu"] = regions.pop("partitioned_vs_single_gpu", None)
            out["independent_regions_mode"] = regions
        if sor_forced:
            pk, _ = hbm_peak()
            for v in sor_forced.values():
                v["frac"] = v["achieved_GBps"] / pk
            out["roofline"]["steady_state_100_forced_iterations"] = sor_forced
        if world == 1:
            try:
                out["roofline_kernels"] = kernel_rooflines(g, m, torch, stream, peak)
            except Exception as ex:  # noqa: BLE001  (reporting only)
                out["roofline_kernels"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu:
            # the CPU restatement runs the SAME warm-up + timed steps for real: the CPU baseline and the full-size parity check in one
            nthreads = os.cpu_count() or 1
            t = time.time()
            traj = cpu_trajectory(m, st, nthreads, args.warmup, args.steps)
            log(f"[bench] CPU trajectory: {args.warmup}+{args.steps} steps in {time.time() - t:.1f}s")
            out["cpu_baseline"] = cpu_line(traj, args.warmup, nthreads, m, "CPU restatement (oracle/)")
            out["parity"] = parity_report(gpu_fields, warm_rows + rows, gpu_time_final, traj)
            out["parity"]["what"] = (f"device-resident GPU run vs the CPU restatement on the same {m.nV}-vertex mesh and inputs after the same {args.warmup + args.steps} steps "
                                     "(Hi, U_SSA, V_SSA in reference vertex order; iteration counts and dt per step)")
            ref_file = os.path.join(ROOT, "gpurun_out", "reference_arm_trajectory.npz")
            if os.path.exists(ref_file):   # the --impl reference arm ran here before: its final fields must be the ones this arm's CPU run produced
                try:
                    z = np.load(ref_file)
                    if str(z["workload"]) == workload_name(m) and int(z["warmup"]) == args.warmup and int(z["steps"]) == args.steps and str(z["order"]) == args.order:
                        out["parity"]["reference_arm_file_bit_identical"] = bool(all(np.array_equal(z[f], traj["fields"][f]) for f in PARITY_FIELDS))
                except Exception:  # noqa: BLE001
                    pass
        if world == 1 and not args.no_extras:
            nthreads = os.cpu_count() or 1
            extras = {}
            for name, fn in (("ssa_warm_state_solve", lambda: leg_warm_ssa_solve(torch, stream, g, r)),
                             ("config1", lambda: leg_config1_eismint(torch, stream, local, nthreads)),
                             ("config2", lambda: leg_config2_halfar(torch, stream, local, nthreads)),
                             ("config4", lambda: leg_config4_mismip(torch, stream, local, nv=args.nv))):
                t = time.time()
                try:
                    extras[name] = fn()
                except Exception as ex:  # noqa: BLE001  (reporting only: the headline line must not depend on these)
                    extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
                log(f"[bench] leg {name}: {time.time() - t:.1f}s")
                if name == "ssa_warm_state_solve":
                    g.close()     # the headline handle is no longer needed: free its 3 GB before the other legs
            out["ssa"]["warm_state_solve"] = extras.pop("ssa_warm_state_solve")
            out["other_configs"] = extras
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nv", type=int, default=1000000)
    ap.add_argument("--exact-xy", dest="exact_xy", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU restatement run (no cpu_baseline, no parity object)")
    ap.add_argument("--order", default="random", choices=["random", "morton", "lattice"], help="host vertex numbering of the workload mesh (see build_workload)")
    ap.add_argument("--no-other-order", dest="no_other_order", action="store_true", help="--impl reference: skip the second trajectory on the other host vertex numbering")
    ap.add_argument("--no-extras", dest="no_extras", action="store_true", help="skip the legs after the headline (configs 1, 2, 4, warm-state solve)")
    ap.add_argument("--no-regions", action="store_true", help="N > 1: skip the extra independent-regions measurement")
    ap.add_argument("--no-config5", dest="no_config5", action="store_true", help="N > 1: skip the ~4 M-vertex leg (BASELINE configs[4])")
    ap.add_argument("--nv5", type=int, default=4000000, help="vertices of the config-5 leg")
    ap.add_argument("--multi", default="partition", choices=["partition", "regions"], help="what N > 1 GPUs do (see run_ours)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
