"""The Fortran side of the boundary (fortran/isca_b200_c.F90, fortran/atmosphere.F90).  No Fortran compiler exists in this image, so:
* the bind(C) derived types of the interface module are checked field by field against the C structs of the headers;
* every C function the two files bind exists in the built library with that name;
* the call sequence of the shim (cold start, steps, restart dump, restart load, steps) is made by a C driver with Fortran-ordered
  arrays (tests/host/shim_driver.c, gcc): built and linked here, run on the GPU box."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER_SRC = os.path.join(ROOT, "tests", "host", "shim_driver.c")
DRIVER = os.path.join(ROOT, "tests", "host", "_build", "shim_driver")


def test_bind_c_types_match_the_headers():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_interface.py"), "--check"])
    assert r.returncode == 0, "fortran/isca_b200_c.F90 is out of step with include/*.h: run tools/gen_fortran_interface.py"


def test_bound_names_exist_in_the_library(lib_built):
    lib = ctypes.CDLL(lib_built)
    names = set()
    for f in ("isca_b200_c.F90", "atmosphere.F90"):
        names |= set(re.findall(r'bind\(C,\s*name="(\w+)"\)', open(os.path.join(ROOT, "fortran", f)).read()))
    assert len(names) > 30
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_every_namelist_variable_of_the_config_is_forwarded():
    """atmosphere_init assigns every field of IscaConfig (a field added to the header must be forwarded by the shim)"""
    src = open(os.path.join(ROOT, "fortran", "atmosphere.F90")).read()
    mod = open(os.path.join(ROOT, "fortran", "isca_b200_c.F90")).read()
    body = mod[mod.index("type, bind(C) :: isca_config"):mod.index("end type isca_config")]
    fields = re.findall(r"::\s*(\w+)", body)[1:]          # [0] is the type name itself
    assert len(fields) == 61
    missing = [f for f in fields if f != "abi_version" and not re.search(r"cfg%" + f + r"\b", src)]
    assert not missing, missing


@pytest.fixture(scope="module")
def driver(lib_built):
    os.makedirs(os.path.dirname(DRIVER), exist_ok=True)
    libdir = os.path.dirname(lib_built)
    subprocess.check_call(["gcc", "-O1", "-std=c11", "-Wall", "-o", DRIVER, DRIVER_SRC, "-L" + libdir, "-lisca_b200",
                           "-Wl,-rpath," + libdir, "-lm"])
    return DRIVER


def test_shim_driver_builds_and_fails_loudly_without_a_gpu(driver):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test below")
    r = subprocess.run([driver, "2", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "create:" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_shim_call_sequence_cold_start_and_restart(driver):
    r = subprocess.run([driver, "30", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    vals = dict(re.findall(r"CHECKSUM (\w+) (\S+)", r.stdout))
    assert abs(float(vals["T_restarted"]) - float(vals["T_uninterrupted"])) <= 1e-12 * abs(float(vals["T_uninterrupted"]))
    # the same sequence through the ctypes binding gives the same state
    import numpy as np
    from isca_b200 import api
    import bench
    nml = bench.hs_namelist("T21", 10, True)
    nml["dt_atmos"] = 1200.0
    atm = api.Atmosphere(api.make_config(**nml))
    atm.cold_start()
    atm.atmosphere(38)
    t = atm.get_field(api.F_T)
    w = 1.0 + (np.arange(t.size) % 7)
    assert abs(float((t.ravel() * w).sum()) - float(vals["T_uninterrupted"])) <= 1e-10 * abs(float(vals["T_uninterrupted"]))
    atm.atmosphere_end()


def test_shim_namelist_groups_cover_the_shipped_test_cases():
    """every variable that a script under exp/test_cases sets in a namelist group the shim reads is declared in the shim's group of
    that name (a Fortran namelist read fails on an unknown variable; the shim turns that into FATAL)"""
    import json
    union = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_test_case_namelists.json")))["__union_of_all_test_cases__"]
    src = open(os.path.join(ROOT, "fortran", "atmosphere.F90")).read()
    src = re.sub(r"&\s*\n\s*", " ", src)                                          # join continuation lines
    src = re.sub(r"!.*", "", src)
    groups = {}
    for m in re.finditer(r"namelist\s*/(\w+)/\s*([^\n]*)", src, flags=re.I):
        groups.setdefault(m.group(1).lower(), set()).update(v.strip().lower() for v in m.group(2).split(",") if v.strip())
    for g in ("atmosphere_nml", "spectral_dynamics_nml", "idealized_moist_phys_nml", "mixed_layer_nml", "vert_turb_driver_nml", "diffusivity_nml",
              "surface_flux_nml", "lscale_cond_nml", "qe_moist_convection_nml", "two_stream_gray_rad_nml", "damping_driver_nml",
              "sat_vapor_pres_nml", "hs_forcing_nml"):
        assert g in groups, g
        missing = sorted(v for v in union.get(g, []) if v.lower() not in groups[g])
        if g == "atmosphere_nml":                    # print_interval: atmosphere_nml of the barotropic / shallow-water atmosphere_mod, another module
            missing = [v for v in missing if v != "print_interval"]
        assert not missing, (g, missing)
