"""GPU: the inline-PTX arithmetic of the engine, primitive by primitive, against Python integers.

The cases live in tests/selftest_cases.py; here they run on the device through the C ABI (sv_selftest_host -> k_selftest),
at full size: edge tables, operands constructed for every rare-carry branch, 10^7 random fe_mul pairs, 2x10^6 for the
other binary operations.  Branch hit counts go to gpurun_out/selftest_coverage.json (copied to profiles/)."""
import pytest

from tests import selftest_cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _coverage():
    C.COVERAGE.clear()
    yield
    C.write_coverage("selftest_coverage.json")


def test_u256_products_and_add_sub_exact(engine):
    C.case_u256_products_and_add_sub_exact(engine)


def test_field_edge_all_pairs_raw_limbs(engine):
    C.case_field_edge_all_pairs_raw_limbs(engine)


def test_field_unary_ops_edge(engine):
    C.case_field_unary_ops_edge(engine)


def test_field_rare_branches_constructed(engine):
    C.case_field_rare_branches_constructed(engine)


def test_field_mul_ten_million_random_pairs(engine):
    C.case_field_mul_ten_million_random_pairs(engine)


def test_field_other_ops_random_pairs(engine):
    C.case_field_other_ops_random_pairs(engine)


def test_scalar_ops_edge_and_rare_folds(engine):
    C.case_scalar_ops_edge_and_rare_folds(engine)


def test_scalar_mul_two_million_random_pairs(engine):
    C.case_scalar_mul_two_million_random_pairs(engine)


def test_glv_split_and_recoding_on_device(engine):
    C.case_glv_split_and_recoding_on_device(engine)


def test_ecmult_kat_through_device_comb_table(engine):
    C.case_ecmult_kat_through_device_comb_table(engine)
