"""Writes tests/golden/reference_known_answers.json: the known-answer vectors that the reference's own tests
hold for the hot path, transcribed from the reference sources (file:line given per case).  The reference is
Rust + Vulkan and cannot be executed in this image, so these are the ASSERTED values of its test-suite, not
outputs generated here.  Run from the repository root:  python tests/golden/make_golden.py"""
import json
import os

CASES = [
    {"name": "test_linspace_f32", "source": "libs/vkjit-core/src/test.rs:10-20", "dtype": "f32",
     "trace": "linspace(F32, const 2.0, const 4.0, 4)", "expected": [2.0, 2.5, 3.0, 3.5]},
    {"name": "test_linspace_eval2", "source": "libs/vkjit-core/src/test.rs:22-44", "dtype": "f32",
     "trace": "second eval on the same Ir: linspace(F32, 10.0, 20.0, 10)",
     "expected": [10.0, 11.0, 12.0, 13.0, 14.0, 15.0, 16.0, 17.0, 18.0, 19.0]},
    {"name": "test_add_f32", "source": "libs/vkjit-core/src/test.rs:47-58", "dtype": "f32", "trace": "arange(F32,3)+arange(F32,3)", "expected": [0.0, 2.0, 4.0]},
    {"name": "test_add_u32", "source": "libs/vkjit-core/src/test.rs:60-71", "dtype": "u32", "trace": "arange(U32,3)+arange(U32,3)", "expected": [0, 2, 4]},
    {"name": "test_add_i32", "source": "libs/vkjit-core/src/test.rs:73-84", "dtype": "i32", "trace": "arange(I32,3)+arange(I32,3)", "expected": [0, 2, 4]},
    {"name": "test_sub_f32", "source": "libs/vkjit-core/src/test.rs:87-98", "dtype": "f32", "trace": "[1,2,3]-[0,1,2]", "expected": [1.0, 1.0, 1.0]},
    {"name": "test_sub_u32", "source": "libs/vkjit-core/src/test.rs:100-111", "dtype": "u32", "trace": "[1,2,3]-[0,1,2]", "expected": [1, 1, 1]},
    {"name": "test_sub_i32", "source": "libs/vkjit-core/src/test.rs:113-124", "dtype": "i32", "trace": "[0,1,2]-[1,2,3]", "expected": [-1, -1, -1]},
    {"name": "test_scatter_f32", "source": "libs/vkjit-core/src/test.rs:127-142", "dtype": "f32",
     "trace": "scatter([0,1,2]+1 -> y=[0,0,0], idx=arange(U32,3)); read y", "expected": [1.0, 2.0, 3.0]},
    {"name": "test_scatter_conditional", "source": "libs/vkjit-core/src/test.rs:144-161", "dtype": "f32",
     "trace": "scatter(x=[0..4] -> y=zeros(5), idx=arange(5), active=idx<3); read y", "expected": [0.0, 1.0, 2.0, 0.0, 0.0]},
    {"name": "cast_u32_to_f32", "source": "libs/vkjit-core/src/test.rs:164-173", "dtype": "f32", "trace": "cast(arange(U32,3), F32)", "expected": [0.0, 1.0, 2.0]},
    {"name": "autocast", "source": "libs/vkjit-core/src/test.rs:176-187", "dtype": "i32", "trace": "array_u32([1,2]) + const_i32(-1), read as i32", "expected": [0, 1]},
    {"name": "dec_ref_count", "source": "libs/vkjit-core/src/test.rs:190-207", "dtype": "whitebox",
     "trace": "y = x + c; dec(c); dec(x); eval(y)", "expected": {"vars_len": 3, "vars0_ref_count": 0, "arrays_len": 1}},
    {"name": "setattr", "source": "libs/vkjit-rust/src/types.rs:214-228", "dtype": "f32",
     "trace": "st = zeros(Struct[F32,F32]); st.setattr([1,2,3], 0); eval(st.0, st.1)", "expected": [[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]]},
    {"name": "test_scatter", "source": "libs/vkjit-rust/src/types.rs:230-242", "dtype": "f32",
     "trace": "y = 7.0; y.scatter(x=[1,2,3], arange(U32,3)); eval(y); read x", "expected": [7.0, 7.0, 7.0]},
    {"name": "main_rs", "source": "src/main.rs:4-12", "dtype": "u32", "trace": "arange(U32, 10); eval; dbg", "expected": list(range(10))},
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
    with open(out, "w") as f:
        json.dump({"provenance": "asserted values of the reference's own tests (see `source` per case); transcribed by "
                                 "tests/golden/make_golden.py, replayed by tests/golden_cases.py", "cases": CASES}, f, indent=1)
    print("wrote", out, len(CASES), "cases")
