#!/usr/bin/env python
"""The namelists of the reference's shipped test cases (exp/test_cases/<case>/<case>_test_case.py: the dict literal handed to
`Namelist(...)`, parsed with ast.literal_eval -- nothing of the reference is executed) as tests/golden/reference_test_case_namelists.json.
The reference's own call sites of the hot path: tests/test_reference_python_pins.py checks the library's test-case helpers against them.
Note: a dict literal with a repeated key keeps the LAST value, exactly as Python does when the reference runs the script (the
axisymmetric case repeats 'diffusivity_nml').  Run: python tests/golden/make_reference_namelists.py   (needs /root/reference)"""
import ast
import json
import os

CASES = {"frierson": "frierson/frierson_test_case.py", "MiMA": "MiMA/MiMA_test_case.py", "axisymmetric": "axisymmetric/axisymmetric_test_case.py",
         "held_suarez": "held_suarez/held_suarez_test_case.py", "giant_planet": "giant_planet/giant_planet_test_case.py",
         "variable_co2_grey": "variable_co2_concentration/variable_co2_grey.py", "top_down_test": "top_down_test/top_down_test.py",
         "realistic_continents_fixed_sst": "realistic_continents/realistic_continents_fixed_sst_test_case.py"}
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_test_case_namelists.json")

out = {}
for name, rel in CASES.items():
    path = os.path.join("/root/reference/exp/test_cases", rel)
    if not os.path.exists(path):
        continue
    src = open(path).read()
    if "Namelist({" not in src:
        continue
    i = src.index("Namelist({") + len("Namelist(")
    depth = 0
    for j in range(i, len(src)):
        depth += src[j] == "{"
        depth -= src[j] == "}"
        if depth == 0:
            break
    try:
        out[name] = ast.literal_eval(src[i:j + 1])
    except (ValueError, SyntaxError):
        continue                                        # namelists built from expressions: not a literal
# the union over ALL scripts under exp/test_cases (every dict literal handed to Namelist(...) or to an .update(...)) of the variables set
# per namelist group: what an input.nml produced by the shipped test cases can contain
import glob
union = {}
for path in sorted(glob.glob("/root/reference/exp/test_cases/*/*.py")):
    src = open(path).read()
    for marker in ("Namelist({", ".update({"):
        start = 0
        while True:
            i = src.find(marker, start)
            if i < 0:
                break
            k = src.index("{", i)
            depth = 0
            for j in range(k, len(src)):
                depth += src[j] == "{"
                depth -= src[j] == "}"
                if depth == 0:
                    break
            start = j
            try:
                d = ast.literal_eval(src[k:j + 1])
            except (ValueError, SyntaxError):
                continue
            for g, vals in d.items():
                if isinstance(vals, dict):
                    union.setdefault(g, set()).update(vals)
out["__union_of_all_test_cases__"] = {g: sorted(v) for g, v in union.items()}
json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
print("wrote", OUT, sorted(out))
