"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol that
include/b200mlip.h declares; host mirrors keep the reference's names; the product path fails loudly
without a GPU instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from distmlip_b200 import build, _lib

    build.build()
    return _lib.load_library()


def test_exports_match_header(lib):
    from distmlip_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "b200mlip.h")).read()
    declared = set(re.findall(r"\b(b2m_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback(lib):
    import torch

    from distmlip_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B2MError) as ei:
        _lib.Engine(n_elem=89, dim=64, max_n=9, max_f=4, n_blocks=4, cutoff=5.0, three_body_cutoff=3.0,
                    cutoff_exponent=5)
    assert "no CUDA device" in str(ei.value)


def test_create_rejects_bad_args(lib):
    import ctypes as C

    from distmlip_b200 import _lib

    h = C.c_void_p()
    dev = (C.c_int * 1)(0)
    bad = _lib.ModelDesc(89, 32, 9, 4, 4, 5, 5.0, 3.0, 0.0, 1.0)  # dim != 64
    assert lib.b2m_create(C.byref(bad), dev, 1, C.byref(h)) == -1
    assert b"dim=64" in lib.b2m_last_error(None)
    ok = _lib.ModelDesc(89, 64, 9, 4, 4, 5, 5.0, 3.0, 0.0, 1.0)
    assert lib.b2m_create(C.byref(ok), dev, 0, C.byref(h)) == -2  # ndev in [1,16] (one ordinal per partition)
    assert lib.b2m_create(C.byref(ok), dev, 17, C.byref(h)) == -2
    assert lib.b2m_create(None, dev, 1, C.byref(h)) == -1
    bondgt = _lib.ModelDesc(89, 64, 9, 4, 4, 5, 3.0, 5.0, 0.0, 1.0)  # bond_r > r (fpis.c:436)
    assert lib.b2m_create(C.byref(bondgt), dev, 1, C.byref(h)) == -1


def test_reference_surface_names():
    from distmlip_b200.implementations.matgl import (CHGNet_Dist, MolecularDynamics, PESCalculator_Dist,
                                                      Potential_Dist, Relaxer)
    from distmlip_b200.distributed.dist import Distributed

    for name in ("from_existing", "enable_distributed_mode", "potential_forward_dist", "dist_forward",
                 "predict_structure_dist"):
        assert hasattr(CHGNet_Dist, name)
    for name in ("create_distributed", "num_atoms", "num_bonds", "num_bond_edges", "num_atom_border_nodes",
                 "num_bond_border_nodes", "cartesian_to_wrapped_fractional"):
        assert hasattr(Distributed, name)
    assert PESCalculator_Dist.implemented_properties == ("energy", "free_energy", "forces", "stress", "hessian",
                                                         "magmoms")
    assert Potential_Dist.__version__ == 2 and CHGNet_Dist.__version__ == 1
    assert Relaxer and MolecularDynamics


def test_from_existing_and_cpu_partitions_rejected():
    from distmlip_b200.implementations.matgl import CHGNet_Dist
    from tests._util import make_model

    m = CHGNet_Dist.from_existing(make_model())
    assert m.dist_enabled is False and float(m.cutoff) == 5.0 and m.n_blocks == 4
    with pytest.raises(RuntimeError):
        m.enable_distributed_mode(["cpu", "cpu"])


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under distmlip_b200/ may import it, and bench.py only inside its
    CPU-baseline / reference-arm functions."""
    import ast

    pkg = os.path.join(ROOT, "distmlip_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                for node in ast.walk(ast.parse(src)):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n) for n in ast.walk(fn))
        if uses:
            assert fn.name in ("cpu_reference_step",), fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)


def test_random_init_container_is_accepted():
    from distmlip_b200.implementations.matgl import CHGNet_Dist
    from distmlip_b200.random_init import RandomCHGNet
    from tests._util import make_model

    m = CHGNet_Dist.from_existing(RandomCHGNet(seed=1))
    ref = make_model().state_dict()
    assert set(m._state_dict) == set(ref) and all(m._state_dict[k].shape == ref[k].shape for k in ref)
