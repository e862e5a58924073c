"""CPU: the C-ABI library builds, loads, and exports every symbol include/lrds_b200.h declares; argument
validation answers without a GPU."""
import ctypes as C
import os
import re

from sde_sampler_lrds_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    N.build()
    lib = N.lib()
    header = open(os.path.join(ROOT, "include", "lrds_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(lrds_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lrds_abi_version() == N.ABI_VERSION


def test_struct_sizes_match_header():
    # field-by-field mirror of include/lrds_b200.h (natural alignment, no packing pragmas)
    assert C.sizeof(N.Mlp) == 16 + 6 * 8
    assert C.sizeof(N.Gmm) == 8 + 4 * 8 + 24
    assert C.sizeof(N.Phi4) == 16
    assert C.sizeof(N.LogReg) == 16 + 24 + 24
    assert C.sizeof(N.Distr) == 8 + C.sizeof(N.Gmm) + C.sizeof(N.Phi4) + C.sizeof(N.LogReg)
    assert C.sizeof(N.Spec) == 40 + 24 + 8 + C.sizeof(N.Mlp) + C.sizeof(N.Distr) + 2 * C.sizeof(N.Gmm)


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = N.lib()
    assert lib.lrds_rollout(None, None, None, 0, 0, None, None, None, None) == -1
    assert b"NULL" in lib.lrds_last_error()
    spec = N.Spec()
    spec.abi_version = 99
    assert lib.lrds_rollout(C.byref(spec), None, None, 0, 0, None, None, None, None) == -1
    assert lib.lrds_estimator_blocks(1) == 1 and lib.lrds_estimator_blocks(8193) == 2
    assert lib.lrds_estimator_partials(None, 0, None, None, None) == -1
