"""The C-ABI shared library loads and exports every symbol include/covidcurve_b200.h declares (no compute calls)."""
import os
import re

from tests.conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "covidcurve_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported(built_lib):
    from covidcurve_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SYMBOLS) == declared, "covidcurve_b200/_lib.py must bind exactly the header's functions"
    for name in declared:
        assert hasattr(built_lib, name), "missing export: %s" % name
    assert built_lib.cc_abi_version() == _lib.CC_ABI_VERSION


def test_struct_layouts_match_header():
    """ctypes mirrors must list the header's struct fields in the same order."""
    from covidcurve_b200 import _lib
    src = open(os.path.join(ROOT, "include", "covidcurve_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, cls in (("cc_model_desc", _lib.ModelDesc), ("cc_mcmc_opts", _lib.McmcOpts), ("cc_call_args", _lib.CallArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
            names = re.sub(r"^(char\s*\*\s*)?const\s*\*", "", names.strip())          # `const char* const* names`
            for n in names.split(","):
                fields.append(re.sub(r"[\*\s]|\[.*?\]", "", n))
        assert fields == [f[0] for f in cls._fields_], cname


def test_no_device_is_a_loud_error(built_lib, modfit):
    """Without a GPU every compute entry point fails with CC_E_CUDA - there is no CPU fallback."""
    import pytest
    from covidcurve_b200 import _lib
    if built_lib.cc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    spec, _ = modfit
    with pytest.raises(_lib.CCError, match="no CUDA device"):
        _lib.Model(spec)
    with pytest.raises(_lib.CCError):
        _lib.fp64_peak_probe()


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may touch oracle/."""
    pkg = os.path.join(ROOT, "covidcurve_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".R")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libcc_oracle" not in txt and "cc_oracle.h" not in txt and "nmath_port" not in txt, f


def test_r_shim_compiles():
    """The R .Call layer (covidcurve_b200/r/ccb200_shim.c) type-checks against the C ABI header and the R API declarations in
    tests/r_stubs/ (R itself is absent from the build image), and registers the reference's own entry points with their arity
    (src/RcppExports.cpp:37-41)."""
    import subprocess
    shim = os.path.join(ROOT, "covidcurve_b200", "r", "ccb200_shim.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-I", os.path.join(ROOT, "tests", "r_stubs"),
                        "-I", os.path.join(ROOT, "include"), shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(shim).read()
    for sym in ("_COVIDCurve_natcubspline_loglike_binomial", "_COVIDCurve_natcubspline_loglike_logit", "CC_loglike"):
        assert re.search(r'\{"%s", \(DL_FUNC\)&%s, 4\}' % (sym, sym), src), sym
    assert "cold_rung_last" in src and 'opt_int(opts, "thin", 1)' in src
