"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares;
host-side logic that needs no device (index helpers, matdims, version, error convention)."""
import ctypes
import os
import re

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "eigenexa_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b(eigen\w*|eigenexa_b200_\w+)\s*\(", txt)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_library_exports_every_declared_symbol():
    import eigenexa_b200 as E
    L = E.lib()
    syms = _declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_index_helpers_match_oracle():
    import eigenexa_b200 as E
    for nnod in (1, 2, 3, 4, 8):
        for inod in range(1, nnod + 1):
            for i in (1, 2, 5, 17, 1000):
                assert E.eigen_loop_start(i, nnod, inod) == O.loop_start(i, nnod, inod)
                assert E.eigen_loop_end(i, nnod, inod) == O.loop_end(i, nnod, inod)
                assert E.eigen_translate_l2g(i, nnod, inod) == O.translate_l2g(i, nnod, inod)
                assert E.eigen_translate_g2l(i, nnod, inod) == O.translate_g2l(i, nnod, inod)
                assert E.eigen_owner_node(i, nnod, inod) == O.owner_node(i, nnod, inod)
                assert E.eigen_owner_index(i, nnod, inod) == O.owner_index(i, nnod, inod)


def test_matdims_single_rank_matches_reference_values():
    import eigenexa_b200 as E
    # not initialised: grid defaults to 1x1, values are those of the reference on one process
    for n in (1, 2, 100, 1000, 4000, 10000):
        assert E.eigen_get_matdims(n) == O.get_matdims(n, 1, 1)
        assert E.eigen_get_matdims(n, mode="M") == (n, n)
    assert E.eigen_get_matdims(-3) == (-1, -1)
    # deliberate deviation: 64-bit indexing accepts the headline size the reference rejects
    nx, ny = E.eigen_get_matdims(50000)
    assert nx >= 50000 and ny >= 50000 and O.get_matdims(50000, 1, 1) == (-1, -1)


def test_version_and_uninitialised_calls_are_silent():
    import eigenexa_b200 as E
    v, date, code = E.eigen_get_version()
    assert v == 21300 and "2024" in date and code.startswith("tamakazura")
    # eigen_s before eigen_init returns without touching w (src/eigen_s.F:81-84)
    n = 4
    a = np.eye(n, order="F"); w = np.full(n, 7.0); z = np.zeros((n, n), order="F")
    E.eigen_s(n, a, w, z)
    assert np.all(w == 7.0)
    E.eigen_free()  # no-op when not initialised


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    """On a box without a CUDA device eigen_init must report the missing device (there is no CPU path), and the
    stage-level entry points must refuse to run instead of computing anything on the host."""
    import torch
    import pytest
    import eigenexa_b200 as E
    if torch.cuda.is_available():
        pytest.skip("this check is for CPU-only boxes")
    E.eigen_init(None, "C")
    assert "no CUDA device" in E.last_error() and "no CPU fallback" in E.last_error()
    assert E.eigen_get_procs()[0] in (0, 1)
    n = 8
    a = np.asfortranarray(np.eye(n))
    for call in (lambda: E.eigen_trd(n, a.copy(order="F")), lambda: E.eigen_prd(n, a.copy(order="F")),
                 lambda: E.eigen_dc(n, np.ones(n), np.zeros(n), np.zeros((n, n), order="F")),
                 lambda: E.eigen_bisect(n, np.ones(n), np.zeros(n)),
                 lambda: E.eigen_bisect2(n, np.ones(n), np.zeros(n), np.zeros(n))):
        with pytest.raises(RuntimeError):
            call()
    # the drivers return silently like the reference does without eigen_init (src/eigen_s.F:81-84)
    w = np.full(n, 3.0); z = np.zeros((n, n), order="F")
    E.eigen_sx(n, a.copy(order="F"), w, z)
    assert np.all(w == 3.0)


def test_fatal_errors_unwind_to_the_entry_point_instead_of_aborting():
    """A device / allocation / NCCL failure must not abort() the host application (the reference calls MPI_Abort,
    src/eigen_devel.F:148-164): the entry point returns 99, errinfo = -1, the message is kept."""
    import eigenexa_b200 as E
    L = E.lib()
    assert L.eigenexa_b200_debug_raise() == 99
    assert "debug_raise" in E.last_error()
    assert E.eigen_get_errinfo() == -1
