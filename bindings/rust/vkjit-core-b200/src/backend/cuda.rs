//! `CudaBackend` / `CudaArray`: what `VulkanBackend` / its `Array` are in the reference
//! (backend/vulkan/mod.rs:14-86).  The reference's arrays are host-visible mapped Vulkan buffers, so `map()` is a
//! zero-copy view there (vulkan/mod.rs:29-34); device memory on a B200 is not host-visible, so `map()` reads the array
//! back once (D2H through the library's pinned staging ring) and keeps the bytes for the lifetime of the `CudaArray`.
use std::cell::OnceCell;
use std::os::raw::c_void;

use vkjit_sys as sys;

use super::{Array, Backend};
use crate::internal::check;

/// A device array owned by a Binding var of an `Ir` (`ir`, `id`), or a free-standing allocation (`id == None`).
pub struct CudaArray {
    pub(crate) ir: *mut sys::vkjit_ir,
    pub(crate) id: Option<sys::vkjit_var>,
    pub(crate) ptr: u64,
    pub(crate) bytes: usize,
    pub(crate) host: OnceCell<Box<[u8]>>,
    /// true: created by `Backend::create_array*` (holds the var's only reference); false: a view of a var of an `Ir`
    pub(crate) owned: bool,
}

impl Drop for CudaArray {
    fn drop(&mut self) {
        if let (true, Some(id)) = (self.owned, self.id) {
            unsafe {
                sys::vkjit_dec_ref(self.ir, id);
            }
        }
    }
}

impl Array for CudaArray {
    fn device_address(&self) -> u64 {
        self.ptr
    }
    fn map(&self) -> &[u8] {
        self.host.get_or_init(|| {
            let mut buf = vec![0u8; self.bytes].into_boxed_slice();
            if let Some(id) = self.id {
                let mut ty: sys::vkjit_type = 0;
                unsafe {
                    check(sys::vkjit_var_type(self.ir, id, &mut ty));
                    check(sys::vkjit_read(self.ir, id, ty, buf.as_mut_ptr() as *mut c_void, self.bytes));
                }
            }
            buf
        })
    }
    fn size(&self) -> usize {
        self.bytes
    }
}

/// Owns nothing: the device, stream, memory pool and kernel cache live inside `libvkjit_b200.so` and are shared by
/// every `Ir` of the process (`vkjit_init`).  `create()` binds device `$LOCAL_RANK` (or 0).
pub struct CudaBackend {
    pub(crate) scratch_ir: *mut sys::vkjit_ir,
}

impl Backend for CudaBackend {
    type Array = CudaArray;

    fn create() -> Self {
        unsafe {
            check(sys::vkjit_init(-1));
            let mut ir: *mut sys::vkjit_ir = std::ptr::null_mut();
            check(sys::vkjit_ir_create(&mut ir));
            CudaBackend { scratch_ir: ir }
        }
    }
    fn create_array_from_slice(&self, data: &[u8]) -> CudaArray {
        assert!(data.len() % 4 == 0, "arrays are made of 4-byte words");
        let mut id: sys::vkjit_var = 0;
        let mut ptr: u64 = 0;
        unsafe {
            check(sys::vkjit_array_u32(self.scratch_ir, data.as_ptr() as *const u32, data.len() / 4, &mut id));
            check(sys::vkjit_var_device_ptr(self.scratch_ir, id, &mut ptr));
        }
        CudaArray { ir: self.scratch_ir, id: Some(id), ptr, bytes: data.len(), host: OnceCell::new(), owned: true }
    }
    fn create_array(&self, size: usize) -> CudaArray {
        assert!(size % 4 == 0, "arrays are made of 4-byte words");
        let mut id: sys::vkjit_var = 0;
        let mut ptr: u64 = 0;
        unsafe {
            check(sys::vkjit_array_empty(self.scratch_ir, sys::VKJIT_TY_U32, size / 4, &mut id));
            check(sys::vkjit_var_device_ptr(self.scratch_ir, id, &mut ptr));
        }
        CudaArray { ir: self.scratch_ir, id: Some(id), ptr, bytes: size, host: OnceCell::new(), owned: true }
    }
}

impl Drop for CudaBackend {
    fn drop(&mut self) {
        unsafe {
            sys::vkjit_ir_destroy(self.scratch_ir);
        }
    }
}
