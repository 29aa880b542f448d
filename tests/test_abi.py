"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
ctypes/Fortran mirrors agree with the header, and the product path fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ufemism_b200.h")


@pytest.fixture(scope="module")
def lib():
    from ufemism_b200 import build as B
    from ufemism_b200 import capi

    B.build_all()
    return capi.load_library()


def header_functions():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ufm_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from ufemism_b200 import capi

    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ufemism_b200.h but not exported"
    assert set(capi.EXPORTED) <= set(names)
    assert lib.ufm_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """sizeof/offsetof as the C compiler sees them vs the ctypes mirrors used by tests and bench."""
    from ufemism_b200 import capi

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ufemism_b200.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ufm_params), sizeof(ufm_mesh_desc), sizeof(ufm_ssa_stats), sizeof(ufm_counters),'
                   ' sizeof(ufm_region), sizeof(ufm_host_ice), offsetof(ufm_params, benchmark), offsetof(ufm_region, n_steps), offsetof(ufm_mesh_desc, colour_nV));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(capi.Params), ctypes.sizeof(capi.MeshDesc), ctypes.sizeof(capi.SsaStats), ctypes.sizeof(capi.Counters),
            ctypes.sizeof(capi.Region), ctypes.sizeof(capi.HostIce), capi.Params.benchmark.offset, capi.Region.n_steps.offset, capi.MeshDesc.colour_nV.offset]
    assert got == want
    # restart / help_fields structs (row N4)
    from ufemism_b200 import restart as R

    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ufemism_b200.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ufm_nc_mesh), sizeof(ufm_restart_frame), sizeof(ufm_restart_frame_out),'
                   ' offsetof(ufm_nc_mesh, V), offsetof(ufm_nc_mesh, w_transect), sizeof(ufm_mesh_primary), offsetof(ufm_mesh_primary, xmin),'
                   ' offsetof(ufm_mesh_primary, thermo));return 0;}')
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(R.NcMesh), ctypes.sizeof(R.RestartFrame), ctypes.sizeof(R.RestartFrameOut), R.NcMesh.V.offset, R.NcMesh.w_transect.offset,
                   ctypes.sizeof(capi.MeshPrimary), capi.MeshPrimary.xmin.offset, capi.MeshPrimary.thermo.offset]


def test_fortran_shim_field_ids_match_header():
    from ufemism_b200 import capi

    f90 = open(os.path.join(ROOT, "ufemism_b200", "fortran", "ufemism_b200_shim.f90")).read()
    ids = dict((k, int(v)) for k, v in re.findall(r"UFM_F_(\w+)\s*=\s*(\d+)", f90))
    assert len(ids) >= 25
    for k, v in ids.items():
        assert capi.FIELD_IDS[k] == v, (k, v, capi.FIELD_IDS[k])
    bms = dict((k, int(v)) for k, v in re.findall(r"UFM_BM_(\w+)\s*=\s*(\d+)", f90))
    for k, v in bms.items():
        assert capi.BENCHMARKS[{"NONE": "none", "HALFAR": "Halfar", "BUELER": "Bueler", "MISMIP_MOD": "MISMIP_mod",
                                "MESH_GENERATION_TEST": "mesh_generation_test", "SSA_ICESTREAM": "SSA_icestream"}.get(k, k)] == v
    # every C function the shim binds exists in the header
    for name in re.findall(r"NAME='(\w+)'", f90):
        assert name in header_functions()


def test_oracle_and_product_share_benchmark_ids():
    from oracle import oracle as O
    from ufemism_b200 import capi

    assert O.BENCHMARKS == capi.BENCHMARKS


def test_no_gpu_means_loud_failure(lib):
    """No CPU fallback: without a CUDA device ufm_create must fail with a message (skipped on a GPU box)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from ufemism_b200 import capi

    h = ctypes.c_void_p()
    P = capi.default_params()
    rc = lib.ufm_create(0, ctypes.byref(P), ctypes.byref(h))
    assert rc < 0 and not h.value
    assert b"CPU fallback" in lib.ufm_last_error() or b"CUDA" in lib.ufm_last_error()
    # argument validation happens before any device work
    P.nZ = 99
    assert lib.ufm_create(0, ctypes.byref(P), ctypes.byref(h)) == -2


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "ufemism_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".f90")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "ufm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
