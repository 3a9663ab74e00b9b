#!/usr/bin/env python
"""Regenerate the three bind(C) derived types of fortran/isca_b200_c.F90 (isca_config, isca_physics_config, isca_moist_config), the ABI
version and the ISCA_F_* field ids from include/isca_b200.h and include/isca_b200_physics.h, in place.  `--check` only reports whether the file is in step
(exit code 1 if not); tests/test_fortran_shim.py runs the same comparison."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FT = {'int32_t': 'integer(c_int32_t)', 'int': 'integer(c_int)', 'double': 'real(c_double)', 'const double*': 'type(c_ptr)'}
STRUCTS = (('isca_b200.h', 'IscaConfig', 'isca_config'), ('isca_b200_physics.h', 'IscaPhysicsConfig', 'isca_physics_config'),
           ('isca_b200_physics.h', 'IscaMoistConfig', 'isca_moist_config'))


def struct_fields(text, name):
    body = text[text.index('typedef struct ' + name + ' {'):]
    body = body[body.index('{') + 1: body.index('} ' + name)]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out = []
    for stmt in body.split(';'):
        stmt = ' '.join(stmt.split())
        if not stmt:
            continue
        m = re.match(r'(const double\*|int32_t|int|double)\s+(.*)', stmt)
        ctype, names = m.group(1), m.group(2)
        for nm in names.split(','):
            nm = nm.strip()
            arr = None
            ma = re.match(r'(\w+)\[(\d+)\]', nm)
            if ma:
                nm, arr = ma.group(1), int(ma.group(2))
            out.append((ctype, nm, arr))
    return out


def type_body(fields, tname):
    lines = [f'  type, bind(C) :: {tname}']
    for ctype, nm, arr in fields:
        lines.append(f'    {FT[ctype]:20s} :: {nm}' + (f'({arr})' if arr else ''))
    lines.append(f'  end type {tname}')
    return '\n'.join(lines)


def constants(src):
    """ISCA_B200_ABI_VERSION and the ISCA_F_* field ids from include/isca_b200.h"""
    text = open(os.path.join(ROOT, 'include', 'isca_b200.h')).read()
    ver = re.search(r'#define ISCA_B200_ABI_VERSION (\d+)', text).group(1)
    src = re.sub(r'(ISCA_B200_ABI_VERSION = )\d+', r'\g<1>' + ver, src)
    nocom = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    ids = re.findall(r'\b(ISCA_F_\w+)\s*=\s*(\d+)', nocom)
    items = [f'{n} = {v}' for n, v in ids]
    lines, cur = [], '  integer(c_int), parameter :: '
    for k, it in enumerate(items):
        piece = it + (', ' if k + 1 < len(items) else '')
        if len(cur) + len(piece) > 126:
            lines.append(cur.rstrip() + ' &')
            cur = ' ' * 31
        cur += piece
    lines.append(cur.rstrip())
    a = src.index('  integer(c_int), parameter :: ISCA_F_PS')
    b = src.index('  integer(c_int), parameter :: ISCA_S_VOR')
    return src[:a] + '\n'.join(lines) + '\n' + src[b:]


def regenerate(src):
    src = constants(src)
    for header, cname, tname in STRUCTS:
        text = open(os.path.join(ROOT, 'include', header)).read()
        a = src.index('  type, bind(C) :: ' + tname)
        b = src.index('  end type ' + tname) + len('  end type ' + tname)
        src = src[:a] + type_body(struct_fields(text, cname), tname) + src[b:]
    return src


if __name__ == '__main__':
    path = os.path.join(ROOT, 'fortran', 'isca_b200_c.F90')
    old = open(path).read()
    new = regenerate(old)
    if '--check' in sys.argv:
        sys.exit(0 if new == old else 1)
    open(path, 'w').write(new)
    print('fortran/isca_b200_c.F90', 'unchanged' if new == old else 'updated')
