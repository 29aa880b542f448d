"""CPU-side logic of bench.py (no GPU): the per-step report, the CPU cost model of the reference arm, the roofline denominator."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ssa_solve_time_from_per_step_times():
    rows = [dict(ssa=0, n_outer=0, n_sor=0), dict(ssa=1, n_outer=50, n_sor=150), dict(ssa=0, n_outer=0, n_sor=0), dict(ssa=1, n_outer=50, n_sor=160)]
    d = bench.ssa_solve_time([2.0, 42.0, 2.0, 44.0], rows)
    assert d["ms_per_step_with_ssa_solve"] == 43.0 and d["ms_per_step_without_ssa_solve"] == 2.0 and d["ms_per_ssa_solve"] == 41.0
    assert d["n_outer_per_solve"] == 50.0 and d["n_sor_per_solve"] == 155.0 and d["step_ms"] == [2.0, 42.0, 2.0, 44.0]
    # every step solves: no difference can be formed, nothing raises
    d = bench.ssa_solve_time([40.0, 41.0], rows[1::2])
    assert "ms_per_ssa_solve" not in d and d["ms_per_step_without_ssa_solve"] is None
    assert bench.ssa_solve_time(None, rows) == {} and bench.ssa_solve_time([1.0], [{}]) == {}


def test_cpu_trajectory_and_parity_report():
    """The CPU arm runs the oracle's region loop for real; the parity report compares a (here: second CPU) trajectory with it field by
    field and step by step, and fails on a perturbed field, a different iteration count or a different time step."""
    import copy

    import numpy as np

    m, st = bench.build_workload(4000)
    tr = bench.cpu_trajectory(m, st, 2, 2, 3)
    assert len(tr["rows"]) == 5 == len(tr["step_s"]) and tr["time"] > tr["time_after_warmup"] > 0.0
    assert tr["rows"][0]["sia"] == 1 and tr["rows"][0]["ssa"] == 1 and tr["rows"][0]["n_sor"] > 0     # first step: both solvers due at once
    line = bench.cpu_line(tr, 2, 2, m, "test")
    assert line["kind"] == "port" and line["cores"] == 2 and line["value"] > 0 and abs(line["model_years"] - (tr["time"] - tr["time_after_warmup"])) < 1e-12
    same = bench.parity_report(tr["fields"], tr["rows"], tr["time"], tr)
    assert same["passed"] and all(same["bit_identical"].values()) and same["rel_l2_U"] == 0.0 and same["steps_compared"] == 5
    bad = {k: v.copy() for k, v in tr["fields"].items()}
    bad["U_SSA"][10] += 1e-6 * max(1.0, abs(bad["U_SSA"]).max())
    rep = bench.parity_report(bad, tr["rows"], tr["time"], tr)
    assert not rep["passed"] and not rep["bit_identical"]["U_SSA"] and rep["bit_identical"]["Hi"] and rep["rel_l2_U"] > 1e-10
    rows = copy.deepcopy(tr["rows"]); rows[1]["n_sor"] += 1
    assert not bench.parity_report(tr["fields"], rows, tr["time"], tr)["passed"]
    rows = copy.deepcopy(tr["rows"]); rows[2]["dt"] *= 1.0 + 1e-15
    assert not bench.parity_report(tr["fields"], rows, tr["time"], tr)["dt_equal"]
    tr2 = bench.cpu_trajectory(m, st, 1, 2, 3)                                       # other thread count: same bits (rank-independent sums)
    assert bench.parity_report(tr2["fields"], tr2["rows"], tr2["time"], tr)["passed"]


def test_both_arms_print_the_same_config_object():
    import argparse

    m, _ = bench.build_workload(3000)
    a = argparse.Namespace(exact_xy=1, order="random", multi="partition")
    assert bench.bench_config(m, a, 1) == bench.bench_config(m, a, 1) and "model" not in bench.bench_config(m, a, 1)
    assert "x-strips" in bench.bench_config(m, a, 4)["parallelism"] and bench.bench_config(m, a, 1)["workload"] == bench.workload_name(m)
    b = bench.algorithmic_bytes(m)
    assert b["geom"] > b["thk"] > b["cfl"] > 0 and b["visc"] > b["geom"] * 0.9


def test_roofline_denominator_is_the_measured_peak_when_present():
    peak, src = bench.hbm_peak()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and src.startswith("measured")
    else:
        assert src.startswith("fallback") and 6000.0 < peak < 8000.0


def test_ours_arm_bookkeeping_with_a_fake_device(monkeypatch, capsys):
    """bench.py's own arm end to end on a FAKE device: torch.cuda and the ctypes wrapper are replaced by stand-ins that only count, so
    this checks the harness (warm-up / timed split, per-step marks, the JSON contract keys, the CPU leg at a small size) and nothing about
    the product.  The real arm needs a B200 (`python bench.py`)."""
    import argparse
    import time

    import torch

    from ufemism_b200 import capi

    class FakeEvent:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    class FakeStream:
        cuda_stream = 0

    class FakeModel:
        def __init__(self, mesh, **kw):
            self.cnt = capi.Counters()
            self.cnt.sor_bytes_per_iteration = 200.0 * mesh.nVAaAc
            self.k = 0
            self.nV = mesh.nV

        def upload_mesh(self, m): self.k = 0
        def download(self, name): import numpy as np; return np.zeros(self.nV)
        def pow_mode(self): return 3
        def set_stream(self, s): pass
        def upload(self, k, v): pass
        def host_register(self, a): pass
        def set_params(self, **kw): pass
        def close(self): pass
        def region(self, t): return capi.Region()
        def reset_counters(self): self.cnt.kernel_launches = self.cnt.sor_iterations = self.cnt.sor_launches = 0; self.cnt.sor_ms = self.cnt.h2d_bytes = self.cnt.d2h_bytes = 0.0
        def counters(self): return capi.Counters.from_buffer_copy(self.cnt)      # a snapshot, like ufm_counters_get

        def _step(self, r, solve):
            r.dt = 0.5; r.time += 0.5; r.n_steps += 1; r.n_sia += 1
            self.cnt.kernel_launches += 7
            if solve:
                r.n_ssa += 1; r.n_outer_total += 50; r.n_sor_total += 150
                self.cnt.kernel_launches += 150; self.cnt.sor_iterations += 150; self.cnt.sor_launches += 50; self.cnt.sor_ms += 0.05
                time.sleep(0.002)

        def run_model(self, r, t_end, max_steps=1):
            self._step(r, self.k % 2 == 0); self.k += 1

        def run_model_host(self, r, t_end, n, host):
            self.cnt.h2d_bytes += 52e6; self.cnt.d2h_bytes += 76e6
            self.run_model(r, t_end)

        def ssa_sor(self, max_inner=0, force_iters=False):
            self.cnt.sor_iterations += max_inner; self.cnt.sor_ms += 0.17 * max_inner; self.cnt.sor_launches += 1

    class FakeSampler:
        def __init__(self, index=0): pass
        def stop(self): return {"sm_mhz": 1900.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3, "source": "fake"}

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "set_stream", lambda s: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(capi, "IceModelGPU", FakeModel)
    monkeypatch.setattr(bench, "ClockSampler", FakeSampler)
    monkeypatch.delenv("WORLD_SIZE", raising=False); monkeypatch.delenv("RANK", raising=False)
    bench.run_ours(argparse.Namespace(gpus=1, steps=4, warmup=3, impl="ours", nv=6000, exact_xy=1, no_cpu=False, no_regions=True, multi="partition",
                                      order="random", no_extras=True, no_other_order=True))
    line = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
              "e2e", "gpu_launches", "roofline", "cpu_baseline", "ssa"):
        assert k in out, k
    assert out["steps"] == 4 and out["warmup"] == 3 and out["n_gpus"] == 1 and out["dtype"] == "f64" and out["vs_baseline"] is None and "model" not in out["config"]
    assert out["ssa"]["model_years"] == 2.0 and out["ssa"]["n_ssa_solves"] == 2 and out["gpu_launches"] == 4 * 7 + 2 * 150
    assert len(out["ssa"]["step_ms"]) == 4 and out["ssa"]["ms_per_ssa_solve"] > 1.0 and out["ssa"]["n_sor_per_solve"] == 150.0
    assert out["e2e"]["h2d_bytes_per_step"] == 52e6 and out["e2e"]["d2h_bytes_per_step"] == 76e6 and out["e2e"]["same_trajectory_as_value"]
    assert set(out["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and out["roofline"]["bound"] == "hbm"
    assert abs(out["roofline"]["frac"] - out["roofline"]["achieved"] / out["roofline"]["peak"]) < 1e-12
    assert set(out["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and out["cpu_baseline"]["kind"] == "port"
    assert set(out["parity"]) >= {"rel_l2_U", "rel_l2_V", "rel_l2_Hi", "n_sor_equal", "n_outer_equal", "passed"} and out["parity"]["steps_compared"] == 7
    assert not out["parity"]["passed"]                        # the fake device returns zeros: the report must say so
    assert out["config"] == bench.bench_config(bench.build_workload(6000)[0], argparse.Namespace(exact_xy=1, order="random", multi="partition"), 1)
