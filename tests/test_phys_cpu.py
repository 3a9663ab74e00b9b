"""The production column-physics sources (isca_b200/csrc/physics*.cu: kernels, launch functions, staging, the C ABI entry points)
compiled UNCHANGED for the host by tests/host/build_phys_cpu.py (kernel launches rewritten to OpenMP loops, CUDA runtime calls
served by tests/host/cuda_on_cpu/cuda_runtime.h) -- and the GPU parity tests of those entry points run against that build in a
subprocess.  This puts every line of the column-physics kernels under `pytest -m "not gpu"`; what only a B200 run can still falsify
is device arithmetic (fused multiply-adds, device `pow`/`exp`) and launch geometry.  TEST INFRASTRUCTURE: the host build is never
loaded by the isca_b200 package (its library fails at create time without a CUDA device)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "host"))


@pytest.fixture(scope="module")
def cpu_lib():
    import build_phys_cpu
    return build_phys_cpu.build()


def run_on_cpu_build(cpu_lib, args):
    env = dict(os.environ, ISCA_B200_TESTS_ON_CPU_BUILD=cpu_lib)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + args, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=3000)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_column_physics_entry_points_on_the_host_build(cpu_lib):
    """tests/test_gpu_physics.py (saturation tables, lscale_cond, grey radiation, Rayleigh damping, vertical diffusion + mixed layer incl.
    prescribed SST, Monin-Obukhov known answers, surface_flux, diffusivity incl. free_atm_diff, simplified Betts-Miller, FATAL paths)
    except the GPU bandwidth assertion"""
    tail = run_on_cpu_build(cpu_lib, ["tests/test_gpu_physics.py", "-k", "not full_size_timing"])
    assert " passed" in tail and "50 passed" in tail, tail


def test_option_kernels_on_the_host_build(cpu_lib):
    """the column-level tests of tests/test_gpu_rows_f.py: dry convection, full Betts-Miller, compute_es_k tables, do_seasonal
    insolation and do_read_co2 of the grey schemes"""
    tail = run_on_cpu_build(cpu_lib, ["tests/test_gpu_rows_f.py", "-k",
                                      "dry_convection_parity or betts_miller_parity or full_saturation or do_read_co2 or seasonal_insolation"])
    assert "13 passed" in tail, tail


def test_host_build_is_not_reachable_from_the_package():
    """no module of the isca_b200 package mentions the host build or its switch"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "isca_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libisca_phys_cpu" not in text and "ISCA_B200_TESTS_ON_CPU_BUILD" not in text and "cuda_on_cpu" not in text, f
