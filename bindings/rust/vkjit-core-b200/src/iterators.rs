//! Walks over the trace as the reference exposes them through `Ir::iter_dep` / `Ir::iter_se`
//! (libs/vkjit-core/src/iterators.rs:6-61): depth-first, pre-order, every var once.  `DepIterator` follows
//! dependencies only (what kernel-size inference looks at, internal.rs:712), `SeIterator` additionally follows
//! side effects (scatter targets).  The edges are read from the native trace (`vkjit_var_deps`); the reference's
//! `MutSeVisitor` (the ref-count cascade, :63-88) has no counterpart here because releasing happens inside the library.
use std::collections::HashSet;

use crate::internal::{Ir, VarId};

fn step(ir: &Ir, stack: &mut Vec<VarId>, seen: &mut HashSet<VarId>, with_side_effects: bool) -> Option<VarId> {
    while let Some(id) = stack.pop() {
        if !seen.insert(id) {
            continue;
        }
        let var = ir.var(id);
        // children are pushed in reverse so that the first dependency is visited first; side effects go below the
        // dependencies on the stack, i.e. they are visited after them
        if with_side_effects {
            for s in var.side_effects().into_iter().rev() {
                if !seen.contains(&s) {
                    stack.push(s);
                }
            }
        }
        for d in var.deps().into_iter().rev() {
            if !seen.contains(&d) {
                stack.push(d);
            }
        }
        return Some(id);
    }
    None
}

pub struct DepIterator<'a> {
    pub ir: &'a Ir,
    pub stack: Vec<VarId>,
    pub discovered: HashSet<VarId>,
}

impl<'a> Iterator for DepIterator<'a> {
    type Item = VarId;

    fn next(&mut self) -> Option<VarId> {
        step(self.ir, &mut self.stack, &mut self.discovered, false)
    }
}

pub struct SeIterator<'a> {
    pub ir: &'a Ir,
    pub stack: Vec<VarId>,
    pub discovered: HashSet<VarId>,
}

impl<'a> Iterator for SeIterator<'a> {
    type Item = VarId;

    fn next(&mut self) -> Option<VarId> {
        step(self.ir, &mut self.stack, &mut self.discovered, true)
    }
}
