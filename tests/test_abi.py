"""The C-ABI library loads without a GPU and exports exactly what include/mpcb.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "mpcb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcb_[a-z0-9_]+)\s*\(", txt)))


def test_binding_table_matches_header():
    from mpc_arpo_project_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _header_functions()


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from mpc_arpo_project_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), name
    assert _lib.load().mpcb_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_field_order():
    """Field names of the ctypes mirrors follow the header's declaration order."""
    from mpc_arpo_project_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "mpcb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for cname, cls in (("mpcb_problem", _lib.MpcbProblem), ("mpcb_sim_out", _lib.MpcbSimOut), ("mpcb_counters", _lib.MpcbCounters)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?\s*$", part.strip())[0])
        assert names == [f for f, _ in cls._fields_], cname


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpc_arpo_project_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mpc_arpo_project_b200 as M
    from oracle.gen_golden import make_params
    sc, mp, fp, _ = make_params(M, dict(Nx=10, sigma=0.1))
    with pytest.raises(M._lib.MpcbError):
        M.Engine(M.build_problem(sc, mp, fp, None))
