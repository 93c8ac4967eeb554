"""The reference's known-answer tests on the CUDA path, through the C ABI.  Bit-exact."""
import pytest

import golden_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases.ALL, ids=lambda f: f.__name__)
def test_reference_golden_on_cuda(case, cir):
    case(cir)
