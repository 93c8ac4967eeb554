"""tests/golden/reference_known_answers.json must agree with what tests/golden_cases.py replays, and the oracle
must reproduce every vector in it (bit-exact)."""
import json
import os

import numpy as np

import golden_cases
from vkjit_b200.ir import VarType as T

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fixture_covers_every_replayed_case():
    cases = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))["cases"]
    assert len(cases) == len(golden_cases.ALL) == 16
    for c in cases:
        assert c["source"].split(":")[0] in ("libs/vkjit-core/src/test.rs", "libs/vkjit-rust/src/types.rs", "src/main.rs")


def test_oracle_reproduces_fixture_vectors(oir):
    cases = {c["name"]: c for c in json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))["cases"]}
    ir = oir
    x = ir.linspace(T.F32, ir.const_f32(2.), ir.const_f32(4.), 4)
    ir.eval([x])
    assert ir.as_slice(x, T.F32).tolist() == cases["test_linspace_f32"]["expected"]
    y = ir.linspace(T.F32, ir.const_f32(10.), ir.const_f32(20.), 10)
    ir.eval([y])
    assert ir.as_slice(y, T.F32).tolist() == cases["test_linspace_eval2"]["expected"]
    z = ir.add(ir.array_u32([1, 2]), ir.const_i32(-1))
    ir.eval([z])
    assert ir.as_slice(z, T.I32).tolist() == cases["autocast"]["expected"]
    c = ir.cast(ir.arange(T.U32, 3), T.F32)
    ir.eval([c])
    assert ir.as_slice(c, T.F32).tolist() == cases["cast_u32_to_f32"]["expected"]
    s = ir.sub(ir.array_i32([0, 1, 2]), ir.array_i32([1, 2, 3]))
    ir.eval([s])
    assert np.array_equal(ir.as_slice(s, T.I32), np.array(cases["test_sub_i32"]["expected"], np.int32))
