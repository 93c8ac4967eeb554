//! `vkjit_core` on a B200: the reference crate's public surface (libs/vkjit-core/src/lib.rs:1-12) over the C ABI of
//! `libvkjit_b200.so` (include/vkjit_b200.h).  Module names and re-exports are the reference's, so
//! `use vkjit_core::{Ir, VarId}`, `use vkjit_core::vartype::VarType` and `vkjit_core::internal::...` keep resolving.
#[allow(dead_code)]
pub mod backend;
#[allow(dead_code)]
pub mod internal;
mod iterators;
pub mod vartype;

#[cfg(test)]
mod known_answers;

pub use internal::{Ir, Red, Var, VarId};
pub use vartype::{AsVarType, VarType};
