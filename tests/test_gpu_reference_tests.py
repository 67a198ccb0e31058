"""The reference's own batch tests (voxtree.rs:1284-2185) through the C ABI on the GPU."""
import pytest

import reference_suite as rs

pytestmark = pytest.mark.gpu

NEEDS_NONEMPTY_APPLY = {"batch_double_apply", "batch_solid_fill_half_one_by_one"}


@pytest.mark.parametrize("case", rs.ALL, ids=lambda f: f.__name__)
def test_reference_batch_tests_gpu(gpu_api, case):
    case(gpu_api)
