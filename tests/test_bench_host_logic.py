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


def test_cpu_cost_model_adds_what_a_step_runs():
    T = dict(geom=1.0, sia=2.0, thk=4.0, cfl=8.0, ssa_prepare=16.0, visc=32.0, slid=64.0, sor_iter=128.0)
    assert bench.cpu_step_seconds(T, 0, 0, 0, 0) == 13.0                       # thickness + geometry + CFL every step
    assert bench.cpu_step_seconds(T, 1, 0, 0, 0) == 15.0
    assert bench.cpu_step_seconds(T, 1, 1, 2, 3) == 15.0 + 16.0 + 2 * 96.0 + 3 * 128.0


def test_recorded_step_counts_match_the_bench_workload():
    """profiles/config3_step_counts.json (written by the GPU arm) is what --impl reference scales its unit costs by."""
    steps, how = bench.load_counts(1000785, 11)
    assert steps is not None and len(steps) == 11 and "config3_step_counts" in how
    assert all(set(s) >= {"dt", "sia", "ssa", "n_outer", "n_sor"} for s in steps)
    assert steps[0]["sia"] == 1 and steps[0]["ssa"] == 1 and steps[0]["dt"] > 0.0    # first step: both solvers due at once (UFEMISM_main_model.f90:352-390)
    assert bench.load_counts(250000, 11) == (None, None)                              # another mesh: counts do not apply


def test_roofline_denominator_is_the_measured_peak_when_present():
    peak, src = bench.hbm_peak()
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        assert peak == float(json.load(open(p))["hbm_gbs"]) and src.startswith("measured")
    else:
        assert src.startswith("fallback") and 6000.0 < peak < 8000.0
