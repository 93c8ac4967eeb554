"""Pins the CPU oracle against every known-answer test the reference holds for the path
(SURVEY.md §8c): libs/vkjit-core/src/test.rs (13), libs/vkjit-rust/src/types.rs (2), src/main.rs."""
import pytest

import golden_cases


@pytest.mark.parametrize("case", golden_cases.ALL, ids=lambda f: f.__name__)
def test_reference_golden_on_oracle(case, oir):
    case(oir)
