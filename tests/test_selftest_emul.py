"""CPU: the self-test cases of tests/selftest_cases.py against the HOST build of the kernel headers (tests/host_emul), at
reduced sizes.  Pins the Python limb model (tests/limb_model.py) and the case logic without a GPU; the device forms of
the same primitives are pinned by tests/test_gpu_selftest.py."""
import ctypes

import numpy as np
import pytest

from tests import selftest_cases as C


class EmulEngine:
    def __init__(self, emul):
        self.emul = emul
        emul.emul_gtable_build()

    def selftest(self, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 8)
        b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, 8)
        out = np.zeros((a.shape[0], 16), dtype=np.uint32)
        self.emul.emul_selftest(int(op), a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_size_t(a.shape[0]), out.ctypes.data_as(ctypes.c_void_p))
        return out


@pytest.fixture(scope="module")
def eng(emul):
    return EmulEngine(emul)


def test_u256_and_field_edges(eng):
    C.case_u256_products_and_add_sub_exact(eng)
    C.case_field_edge_all_pairs_raw_limbs(eng)
    C.case_field_unary_ops_edge(eng)


def test_rare_branches_and_random(eng):
    C.case_field_rare_branches_constructed(eng)
    C.case_field_mul_ten_million_random_pairs(eng, total=100_000)
    C.case_field_other_ops_random_pairs(eng, m=50_000)


def test_scalar_and_recoding(eng):
    C.case_scalar_ops_edge_and_rare_folds(eng)
    C.case_scalar_mul_two_million_random_pairs(eng, m=50_000)
    C.case_glv_split_and_recoding_on_device(eng)


def test_ecmult_kat(eng):
    C.case_ecmult_kat_through_device_comb_table(eng)
