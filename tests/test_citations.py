"""Every `file:line` citation of the reference in the header, the docs, the oracle and the kernels must point into the file it
names (checked where /root/reference is present), and the routines the scope table names must start and end where cited."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not here")

FILES = ["include/gritas_b200.h", "DESIGN.md", "INTEGRATION.md", "README.md", "oracle/gritas_oracle.c", "oracle/oracle.py",
         "oracle/oracle_np.py", "oracle/ref.py", "oracle/ref_harness.c", "oracle/f2c_subset.py", "gritas_b200/csrc/kernels.cuh",
         "gritas_b200/csrc/gritas_b200.cu", "gritas_b200/csrc/xcov.cu", "gritas_b200/m_gritas_grids.py", "gritas_b200/cabi.py",
         "fortran/m_gritas_b200.F90", "tests/test_oracle_vs_ref.py"]
NAMES = {"m_gritas_grids.f", "gritas.f", "grisas.f", "gritas_kxkt.f", "m_gritas_masks.F90", "gritas.h", "etc.f", "m_gritas_binning.F"}
CITE = re.compile(r"\b(" + "|".join(re.escape(n) for n in sorted(NAMES, key=len, reverse=True)) + r"):(\d+)(?:-(\d+))?")


def ref_path(name):
    for dirpath, _, files in os.walk(os.path.join(REF, "src")):
        if name in files:
            return os.path.join(dirpath, name)
    return None


def test_cited_lines_exist():
    lengths = {n: sum(1 for _ in open(ref_path(n), errors="replace")) for n in NAMES if ref_path(n)}
    bad, seen = [], 0
    for f in FILES:
        p = os.path.join(ROOT, f)
        if not os.path.exists(p):
            continue
        for m in CITE.finditer(open(p, errors="replace").read()):
            name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if name not in lengths:
                continue
            seen += 1
            if not (1 <= a <= b <= lengths[name]):
                bad.append(f"{f}: {m.group(0)} (file has {lengths[name]} lines)")
    assert seen > 200, seen
    assert not bad, bad


ROUTINES = [("GrITAS_Pint_", "m_gritas_grids.f", 217, 296), ("GrITAS_Accum_", "m_gritas_grids.f", 307, 478),
            ("GrITAS_Norm_", "m_gritas_grids.f", 490, 656), ("GrITAS_XCov_Accum_", "m_gritas_grids.f", 667, 848),
            ("GrITAS_XCov_Norm_", "m_gritas_grids.f", 859, 1008), ("GrITAS_MinMax_", "m_gritas_grids.f", 1019, 1167),
            ("GrITAS_3CH_Norm_", "m_gritas_grids.f", 1178, 1341), ("scores2_", "m_gritas_grids.f", 1671, 1795),
            ("GrITAS_XCov_Prs_Accum_", "m_gritas_grids.f", 1862, 2059), ("GrITAS_KxKt", "gritas_kxkt.f", 10, 119),
            ("maskout_", "m_gritas_masks.F90", 137, 186)]


@pytest.mark.parametrize("name,fname,a,b", ROUTINES)
def test_routines_start_and_end_where_cited(name, fname, a, b):
    lines = open(ref_path(fname), errors="replace").read().split("\n")
    assert re.search(rf"subroutine\s+{re.escape(name)}\b", lines[a - 1], re.I), lines[a - 1]
    assert re.match(r"^\s*end\b", lines[b - 1], re.I), lines[b - 1]
