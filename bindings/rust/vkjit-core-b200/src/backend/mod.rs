//! The reference's backend seam (libs/vkjit-core/src/backend/mod.rs:8-25) with the CUDA implementation behind it.
//! `Backend::execute(Kernel, ..)` is SPIR-V specific in the reference (it takes the rspirv-built `Kernel` by value,
//! mod.rs:24) and is therefore not part of this trait: compile + launch happen inside `vkjit_eval`.
pub mod cuda;

pub trait Array {
    fn device_address(&self) -> u64;
    fn map(&self) -> &[u8];
    fn size(&self) -> usize;
}

pub trait Backend {
    type Array: Array;

    fn create() -> Self;
    fn create_array_from_slice(&self, data: &[u8]) -> Self::Array;
    fn create_array(&self, size: usize) -> Self::Array;
}
