"""The Fortran drop-in cannot be compiled here (no Fortran compiler in the image); tools/check_fortran_shim.py checks
what a compiler's front end would: bind(C) interfaces vs include/gritas_b200.h, no module cycle, call arities,
and that the replacement routines keep the reference's argument lists.  The mutation cases prove the checker bites."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    spec = importlib.util.spec_from_file_location("check_fortran_shim", os.path.join(ROOT, "tools", "check_fortran_shim.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shim_is_consistent_with_the_header_and_acyclic():
    m = load()
    errs, ifaces, graph = m.run()
    assert not errs, errs
    for nm in ("gritas_b200_create", "gritas_b200_accum", "gritas_b200_accum_ods", "gritas_b200_norm", "gritas_b200_norm_sharded",
               "gritas_b200_norm_reduce", "gritas_b200_comm_init", "gritas_b200_peer_attach", "gritas_b200_wait"):
        assert nm in ifaces, nm
    assert "m_gritas_grids" not in graph["m_gritas_b200"]
    assert "m_gritas_b200" in graph["m_gritas_grids"]


def _mutated(tmp_path, shim_edit=None, inc_edit=None):
    m = load()
    shim = open(m.SHIM).read()
    inc = open(m.INC).read()
    if shim_edit:
        new = shim_edit(shim)
        assert new != shim
        shim = new
    if inc_edit:
        new = inc_edit(inc)
        assert new != inc
        inc = new
    a, b = tmp_path / "shim.F90", tmp_path / "inc.inc"
    a.write_text(shim)
    b.write_text(inc)
    m.SHIM, m.INC = str(a), str(b)
    return m.run()[0]


def test_checker_catches_a_circular_use(tmp_path):
    errs = _mutated(tmp_path, lambda s: s.replace("use m_odsmeta, only : KTMAX, KXMAX", "use m_gritas_grids, only : im\n      use m_odsmeta, only : KTMAX, KXMAX", 1))
    assert any("circular" in e or "cycle" in e for e in errs), errs


def test_checker_catches_a_wrong_prototype(tmp_path):
    # nobs by reference instead of by value
    errs = _mutated(tmp_path, lambda s: s.replace("""            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*), del(*)""", """            type(c_ptr),    value             :: h
            integer(c_int), intent(in)        :: nobs
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*), del(*)""", 1))
    assert any("gritas_b200_accum: argument 2" in e for e in errs), errs
    # a missing argument
    errs = _mutated(tmp_path, lambda s: s.replace("gritas_b200_norm_reduce ( h, root, nrmlz, ntresh,", "gritas_b200_norm_reduce ( h, nrmlz, ntresh,", 1))
    assert any("gritas_b200_norm_reduce" in e for e in errs), errs


def test_checker_catches_a_wrong_call_arity(tmp_path):
    errs = _mutated(tmp_path, inc_edit=lambda s: s.replace("call B200_Accum ( im, jm, km, nlevs, lat,", "call B200_Accum ( im, jm, km, lat,", 1))
    assert any("b200_accum: called with" in e for e in errs), errs


def test_checker_catches_what_implicit_none_would(tmp_path):
    """an undeclared variable, a misspelt module variable, a lost `end if`, a lost parenthesis"""
    errs = _mutated(tmp_path, lambda s: s.replace("      integer(c_int) :: rc\n", "      integer(c_int) :: rcx\n", 1))
    assert any("`rc` is not declared" in e for e in errs), errs
    errs = _mutated(tmp_path, lambda s: s.replace("      normed = .true.", "      normd = .true.", 1))
    assert any("`normd` is not declared" in e for e in errs), errs
    errs = _mutated(tmp_path, lambda s: s.replace("      end if\n", "\n", 1))
    assert any("unterminated" in e or "unbalanced" in e for e in errs), errs
    errs = _mutated(tmp_path, lambda s: s.replace("      sent3 = 0\n", "      sent3 = max(0, 1\n", 1))
    assert any("unbalanced parentheses" in e for e in errs), errs
