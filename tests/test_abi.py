"""C-ABI checks that need no GPU: the library builds, loads, exports every symbol declared in
include/isca_b200.h, the ctypes struct matches the C struct, and the product path fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "isca_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(isca_b200_\w+)\s*\(", txt)))


def test_header_symbols_exported(lib_built):
    from isca_b200 import api
    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/isca_b200.h but not exported"
    assert set(api.EXPORTS) == set(syms)


def test_config_struct_layout_matches_c(lib_built, tmp_path):
    """sizeof/offsets of IscaConfig as seen by a C compiler == the ctypes mirror."""
    from isca_b200 import api
    src = tmp_path / "sz.c"
    fields = [f[0] for f in api.IscaConfigStruct._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(IscaConfig, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200.h"\nint main(){printf("%zu\\n", sizeof(IscaConfig));' + body + "return 0;}")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(api.IscaConfigStruct)
    for f, off in zip(fields, out[1:]):
        assert getattr(api.IscaConfigStruct, f).offset == int(off), f


def test_default_config_is_reference_namelist(lib_built):
    from isca_b200 import api
    c = api.make_config()
    # spectral_dynamics.F90:152-206 defaults
    assert (c.lon_max, c.lat_max, c.num_fourier, c.num_spherical, c.num_levels) == (128, 64, 42, 43, 18)
    assert c.damping_order == 2 and abs(c.damping_coeff - 1.15740741e-4) < 1e-20
    assert c.robert_coeff == 0.04 and c.alpha_implicit == 0.5 and c.raw_filter_coeff == 1.0
    assert c.reference_sea_level_press == 101325.0
    # hs_forcing.F90:78-83
    assert (c.t_zero, c.t_strat, c.delh, c.delv, c.sigma_b, c.ka, c.ks, c.kf) == (315., 200., 60., 10., 0.7, -40., -4., -1.)
    with pytest.raises(api.IscaError):
        api.make_config(not_a_namelist_variable=1)
    with pytest.raises(api.IscaError):
        api.make_config(vert_coord_option="nonsense")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lib_built):
    from isca_b200 import api
    with pytest.raises(api.IscaError) as e:
        api.Atmosphere(api.make_config())
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """The shipped package must never route through the oracle."""
    pkg = os.path.join(ROOT, "isca_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt, f"{f} mentions the oracle"


PHYS_HEADER = os.path.join(ROOT, "include", "isca_b200_physics.h")


def test_physics_header_symbols_exported_and_struct_layout(lib_built, tmp_path):
    from isca_b200 import api, physics
    lib = api.load_library()
    txt = re.sub(r"/\*.*?\*/", "", open(PHYS_HEADER).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(isca_b200_\w+)\s*\(", txt)))
    from isca_b200 import moist
    assert set(syms) == set(physics.PHYSICS_EXPORTS) | set(moist.MOIST_EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/isca_b200_physics.h but not exported"
    fields = [f[0] for f in physics.IscaPhysicsConfigStruct._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(IscaPhysicsConfig, {f}));' for f in fields)
    src = tmp_path / "psz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200_physics.h"\nint main(){printf("%zu\\n", sizeof(IscaPhysicsConfig));' + body + "return 0;}")
    exe = tmp_path / "psz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(physics.IscaPhysicsConfigStruct)
    for f, off in zip(fields, out[1:]):
        assert getattr(physics.IscaPhysicsConfigStruct, f).offset == int(off), f
    # namelist defaults: two_stream_gray_rad.F90:72-113, lscale_cond.F90:48-52, damping_driver.f90:60-78
    c = physics.IscaPhysicsConfigStruct()
    lib.isca_b200_physics_default_config(C.byref(c))
    assert (c.solar_constant, c.del_sol, c.ir_tau_eq, c.ir_tau_pole, c.linear_tau, c.wv_exponent) == (1360.0, 1.4, 6.0, 1.5, 0.1, 4.0)
    assert (c.hc, c.do_evap, c.trayfric, c.sponge_pbottom) == (1.0, 0, 0.0, 50.0)
    # the other rad_scheme values (two_stream_gray_rad.F90:96-118)
    assert (c.abi_version, c.rad_scheme, c.window, c.carbon_conc, c.bog_a, c.bog_b, c.lw_tau_0_gp, c.single_albedo) == \
        (4, 0, 0.3732, 360.0, 0.8678, 1997.9, 80.0, 0.8)
    # diffusivity_nml free_atm_diff parameters (diffusivity.F90:132-143)
    assert (c.free_atm_diff, c.free_atm_skyhi_diff, c.ampns, c.rich_crit_diff, c.mix_len, c.rich_prandtl, c.ampns_max) == (0, 0, 0, 0.25, 30.0, 1.0, 1.0e20)
    assert c.sat_vapor_pres_do_simple == 1


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_physics_no_cpu_fallback(lib_built):
    from isca_b200 import api, physics
    with pytest.raises(api.IscaError) as e:
        physics.ColumnPhysics(8, 4, 5)
    assert "CUDA" in str(e.value)


def test_moist_config_struct_layout(lib_built, tmp_path):
    from isca_b200 import moist
    fields = [f[0] for f in moist.IscaMoistConfigStruct._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(IscaMoistConfig, {f}));' for f in fields)
    src = tmp_path / "msz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200_physics.h"\nint main(){printf("%zu\\n", sizeof(IscaMoistConfig));' + body + "return 0;}")
    exe = tmp_path / "msz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(moist.IscaMoistConfigStruct)
    for f, off in zip(fields, out[1:]):
        assert getattr(moist.IscaMoistConfigStruct, f).offset == int(off), f
    c = moist.IscaMoistConfigStruct()
    moist._lib().isca_b200_moist_default_config(C.byref(c))
    # idealized_moist_phys.F90:136-138, mixed_layer.F90:92-95
    assert (c.roughness_mom, c.roughness_heat, c.roughness_moist, c.mixed_layer_depth, c.albedo_value) == (0.05, 0.05, 0.05, 40.0, 0.06)
    assert (c.abi_version, c.use_tau, c.constant_gust) == (2, 1, 1.0)            # vert_turb_driver.F90:109,116


RRTM_HEADER = os.path.join(ROOT, "include", "isca_b200_rrtm.h")


def test_rrtm_header_symbols_exported_and_struct_layout(lib_built, tmp_path):
    from isca_b200 import api, rrtm
    lib = api.load_library()
    txt = re.sub(r"/\*.*?\*/", "", open(RRTM_HEADER).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(isca_b200_\w+)\s*\(", txt)))
    assert set(syms) == set(rrtm.RRTM_EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/isca_b200_rrtm.h but not exported"
    fields = [f[0] for f in rrtm.IscaRrtmConfigStruct._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(IscaRrtmConfig, {f}));' for f in fields)
    src = tmp_path / "rsz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200_rrtm.h"\nint main(){printf("%zu\\n", sizeof(IscaRrtmConfig));' + body + "return 0;}")
    exe = tmp_path / "rsz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(rrtm.IscaRrtmConfigStruct)
    for f, off in zip(fields, out[1:]):
        assert getattr(rrtm.IscaRrtmConfigStruct, f).offset == int(off), f
    fields = [f[0] for f in rrtm.IscaRrtmDriverConfigStruct._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(IscaRrtmDriverConfig, {f}));' for f in fields)
    src = tmp_path / "dsz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "isca_b200_rrtm.h"\nint main(){printf("%zu\\n", sizeof(IscaRrtmDriverConfig));' + body + "return 0;}")
    exe = tmp_path / "dsz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(rrtm.IscaRrtmDriverConfigStruct)
    for f, off in zip(fields, out[1:]):
        assert getattr(rrtm.IscaRrtmDriverConfigStruct, f).offset == int(off), f
    d = rrtm.driver_config()
    # rrtm_radiation.F90:163-189 and astronomy.f90:141-164 defaults
    assert (d.dt_rad, d.dt_rad_avg, d.do_rad_time_avg, d.store_intermediate_rad, d.solday, d.equinox_day) == (0, -1, 1, 1, 0, 0.75)
    assert (d.ecc, d.obliq, d.per, d.num_angles) == (0.0, 23.439, 102.932, 3600)
    # rrtm_radiation_nml defaults (rrtm_radiation.F90:117-226)
    c = rrtm.default_config()
    assert (c.co2ppmv, c.h2o_lower_limit, c.temp_lower_limit, c.temp_upper_limit, c.solrad, c.solr_cnst) == (300.0, 2.0e-7, 100.0, 370.0, 1.0, 1368.22)
    assert (c.include_secondary_gases, c.convert_sphum_to_vmr, c.input_o3_file_is_mmr, c.lonstep) == (0, 1, 1, 1)
    assert os.path.exists(rrtm.TABLE_FILE)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_rrtm_no_cpu_fallback(lib_built):
    from isca_b200 import api, rrtm
    with pytest.raises(api.IscaError) as e:
        rrtm.Rrtm()
    assert "CUDA" in str(e.value)
    with pytest.raises(api.IscaError) as e:
        rrtm.Rrtm(num_lon=128, lonstep=3)
    assert "lonstep" in str(e.value)


def _prototypes(header):
    txt = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*|IscaHandle)\s+(isca_b200_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if params == ["void"] or params == [""] else params
    return protos


def _expected_ctype(param):
    """C parameter declaration -> the ctypes class family the binding must use"""
    p = re.sub(r"\s+", " ", param)
    if "[" in p:                                   # array parameter = pointer
        return "ptr"
    if "double*" in p.replace(" *", "*"):
        return "double*"
    if "int*" in p.replace(" *", "*"):
        return "int*"
    if "char*" in p.replace(" *", "*"):
        return "char*"
    if "*" in p or re.match(r"(const )?Isca(Rrtm|Moist|Physics|Handle)\b", p) and not re.search(r"Config", p):
        return "ptr"
    if p.startswith("long long"):
        return "longlong"
    if p.startswith("double"):
        return "double"
    if p.startswith("int"):
        return "int"
    return "ptr"


def test_rrtm_ctypes_bindings_match_the_header(lib_built):
    """every binding of isca_b200/rrtm.py has the parameter count and the scalar / pointer classes of its C prototype (a wrong
    count or an int passed where a double is expected would only show up on the GPU box otherwise)"""
    from isca_b200 import rrtm
    lib = rrtm._lib()
    protos = _prototypes(RRTM_HEADER)
    assert set(protos) == set(rrtm.RRTM_EXPORTS)
    for name, params in protos.items():
        fn = getattr(lib, name)
        at = fn.argtypes
        assert at is not None, f"{name} has no argtypes"
        assert len(at) == len(params), f"{name}: {len(at)} argtypes for {len(params)} parameters"
        for t, p in zip(at, params):
            want = _expected_ctype(p)
            if want == "double":
                assert t is C.c_double, (name, p)
            elif want == "int":
                assert t is C.c_int, (name, p)
            elif want == "longlong":
                assert t is C.c_longlong, (name, p)
            elif want == "char*":
                assert t is C.c_char_p, (name, p)
            elif want == "double*":
                assert t in (C.POINTER(C.c_double), C.c_void_p), (name, p)
            else:
                assert t is C.c_void_p or issubclass(t, C._Pointer), (name, p)


def _check_bindings(lib, protos):
    for name, params in protos.items():
        at = getattr(lib, name).argtypes
        assert at is not None, f"{name} has no argtypes"
        assert len(at) == len(params), f"{name}: {len(at)} argtypes for {len(params)} parameters"
        for t, p in zip(at, params):
            want = _expected_ctype(p)
            if want == "double":
                assert t is C.c_double, (name, p)
            elif want == "int":
                assert t is C.c_int, (name, p)
            elif want == "longlong":
                assert t is C.c_longlong, (name, p)
            elif want == "char*":
                assert t is C.c_char_p, (name, p)
            else:
                assert t is C.c_void_p or issubclass(t, C._Pointer), (name, p)


def test_physics_and_moist_ctypes_bindings_match_the_header(lib_built):
    from isca_b200 import physics, moist
    physics._lib()
    lib = moist._lib()
    _check_bindings(lib, _prototypes(PHYS_HEADER))


def test_core_ctypes_bindings_match_the_header(lib_built):
    from isca_b200 import api
    lib = api.load_library()
    protos = _prototypes(HEADER)
    bound = {n: p for n, p in protos.items() if getattr(getattr(lib, n), "argtypes", None) is not None}
    assert len(bound) >= 30
    _check_bindings(lib, bound)
