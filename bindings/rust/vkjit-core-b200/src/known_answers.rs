//! The reference's asserted vectors (libs/vkjit-core/src/test.rs:9-207, tests/golden/reference_known_answers.json in
//! this repository) replayed through this crate, table-driven.  Needs a B200 (`Ir::new` binds the device).
use crate::internal::{Ir, Red};
use crate::vartype::VarType;

#[test]
fn linspace_twice_on_one_ir() {
    // test.rs:10-44; the second vector hits an exact round-to-even tie at 3/10*10
    let mut ir = Ir::new();
    for (lo, hi, n, want) in [
        (2.0f32, 4.0f32, 4usize, vec![2.0f32, 2.5, 3.0, 3.5]),
        (10.0, 20.0, 10, (10..20).map(|v| v as f32).collect()),
    ] {
        let (a, b) = (ir.const_f32(lo), ir.const_f32(hi));
        let x = ir.linspace(VarType::F32, a, b, n);
        ir.eval(&[x]);
        assert_eq!(ir.as_slice::<f32>(x), &want[..]);
    }
}

#[test]
fn add_and_sub_in_every_numeric_type() {
    // test.rs:47-124
    let mut ir = Ir::new();
    let (x, y) = (ir.arange(VarType::F32, 3), ir.arange(VarType::F32, 3));
    let z = ir.add(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<f32>(z), &[0.0, 2.0, 4.0]);
    let (x, y) = (ir.arange(VarType::U32, 3), ir.arange(VarType::U32, 3));
    let z = ir.add(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<u32>(z), &[0, 2, 4]);
    let (x, y) = (ir.arange(VarType::I32, 3), ir.arange(VarType::I32, 3));
    let z = ir.add(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<i32>(z), &[0, 2, 4]);
    let (x, y) = (ir.array_f32(&[1.0, 2.0, 3.0]), ir.array_f32(&[0.0, 1.0, 2.0]));
    let z = ir.sub(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<f32>(z), &[1.0, 1.0, 1.0]);
    let (x, y) = (ir.array_u32(&[1, 2, 3]), ir.array_u32(&[0, 1, 2]));
    let z = ir.sub(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<u32>(z), &[1, 1, 1]);
    let (x, y) = (ir.array_i32(&[0, 1, 2]), ir.array_i32(&[1, 2, 3]));
    let z = ir.sub(x, y);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<i32>(z), &[-1, -1, -1]);
}

#[test]
fn scatter_plain_and_masked() {
    // test.rs:127-161: the TARGET is read, not the scatter var
    let mut ir = Ir::new();
    let idx = ir.arange(VarType::U32, 3);
    let src = ir.array_f32(&[0.0, 1.0, 2.0]);
    let one = ir.const_f32(1.0);
    let src = ir.add(src, one);
    let dst = ir.array_f32(&[0.0; 3]);
    let s = ir.scatter(src, dst, idx, None);
    ir.eval(&[s]);
    assert_eq!(ir.as_slice::<f32>(dst), &[1.0, 2.0, 3.0]);

    let idx = ir.arange(VarType::U32, 5);
    let src = ir.array_f32(&[0.0, 1.0, 2.0, 3.0, 4.0]);
    let dst = ir.array_f32(&[0.0; 5]);
    let three = ir.const_u32(3);
    let active = ir.lt(idx, three);
    let s = ir.scatter(src, dst, idx, Some(active));
    ir.eval(&[s]);
    assert_eq!(ir.as_slice::<f32>(dst), &[0.0, 1.0, 2.0, 0.0, 0.0]);
}

#[test]
fn casts() {
    // test.rs:164-187: explicit U32 -> F32, and U32 + I32 promotes to I32 (bit reinterpretation)
    let mut ir = Ir::new();
    let x = ir.arange(VarType::U32, 3);
    let y = ir.cast(x, &VarType::F32);
    ir.eval(&[y]);
    assert_eq!(ir.as_slice::<f32>(y), &[0.0, 1.0, 2.0]);
    let x = ir.array_u32(&[1, 2]);
    let m = ir.const_i32(-1);
    let z = ir.add(x, m);
    ir.eval(&[z]);
    assert_eq!(ir.as_slice::<i32>(z), &[0, 1]);
    assert_eq!(ir.cast(z, &VarType::I32), z); // same type: the var itself comes back (internal.rs:283-290)
}

#[test]
fn released_operands_die_with_the_eval() {
    // test.rs:190-207 (white box: vars.len(), vars[0].ref_count, arrays.len())
    let mut ir = Ir::new();
    let x = ir.array_f32(&[1.0, 2.0, 3.0]);
    let c = ir.const_f32(1.0);
    let y = ir.add(x, c);
    ir.dec_ref_count(c);
    ir.dec_ref_count(x);
    ir.eval(&[y]);
    assert_eq!(ir.var_count(), 3);
    assert_eq!(ir.var(x).ref_count(), 0);
    assert_eq!(ir.array_count(), 1);
}

#[test]
fn walks_and_extensions() {
    let mut ir = Ir::new();
    let i = ir.arange(VarType::U32, 1000);
    let k = ir.const_u32(3);
    let v = ir.mul(i, k);
    let order: Vec<_> = ir.iter_dep(&[v]).collect();
    assert_eq!(order, vec![v, i, k]); // pre-order, first dependency first
    let total = ir.reduce(Red::Sum, v);
    assert_eq!(ir.as_slice::<u32>(total), &[3 * (999 * 1000 / 2)]);
    assert!(!ir.is_buffer(&v)); // a reduction does not materialise its operand
    let seven = ir.const_u32(7);
    let rem = ir.and(i, seven);
    let zero = ir.const_u32(0);
    let mask = ir.eq(rem, zero);
    let (idx, count) = ir.compress(mask);
    assert_eq!(count, 125);
    assert_eq!(ir.as_slice::<u32>(idx)[..3], [0, 8, 16]);
}
