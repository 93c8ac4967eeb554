//! Raw FFI declarations for `libvkjit_b200.so`, generated from `include/vkjit_b200.h`
//! (one `extern "C"` item per entry point of the C ABI; see INTEGRATION.md for the safe `Ir` wrapper that
//! replaces `vkjit_core::Ir`, reference libs/vkjit-core/src/internal.rs:126-542).
//! NOT compiled in the build image (no Rust toolchain there).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type vkjit_status = i32;
pub type vkjit_type = u32;
pub type vkjit_var = u32;
#[repr(C)]
pub struct vkjit_ir {
    _private: [u8; 0],
}

pub const VKJIT_OK: vkjit_status = 0;
pub const VKJIT_ERR_INVALID: vkjit_status = 1;
pub const VKJIT_ERR_TYPE: vkjit_status = 2;
pub const VKJIT_ERR_SIZE: vkjit_status = 3;
pub const VKJIT_ERR_UNSUPPORTED: vkjit_status = 4;
pub const VKJIT_ERR_NO_DEVICE: vkjit_status = 5;
pub const VKJIT_ERR_CUDA: vkjit_status = 6;
pub const VKJIT_ERR_COMPILE: vkjit_status = 7;
pub const VKJIT_ERR_DIST: vkjit_status = 8;

pub const VKJIT_TY_VOID: vkjit_type = 1;
pub const VKJIT_TY_BOOL: vkjit_type = 2;
pub const VKJIT_TY_U32: vkjit_type = 3;
pub const VKJIT_TY_I32: vkjit_type = 4;
pub const VKJIT_TY_F32: vkjit_type = 5;
pub const VKJIT_TY_STRUCT_BASE: vkjit_type = 16;

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct vkjit_stats_t {
    pub cache_hits: u64,
    pub cache_misses: u64,
    pub trace_launches: u64,
    pub prim_launches: u64,
    pub last_compile_ns: u64,
    pub last_eval_ns: u64,
    pub bytes_h2d: u64,
    pub bytes_d2h: u64,
    pub pool_bytes_live: u64,
    pub collectives: u64,
    pub disk_hits: u64,
}

extern "C" {
    pub fn vkjit_init(device: i32) -> vkjit_status;
    pub fn vkjit_shutdown() -> vkjit_status;
    pub fn vkjit_is_initialized() -> i32;
    pub fn vkjit_last_error() -> *const c_char;
    pub fn vkjit_abi_version() -> u32;
    pub fn vkjit_stream(out_stream: *mut *mut c_void) -> vkjit_status;
    pub fn vkjit_device(out_device: *mut i32) -> vkjit_status;
    pub fn vkjit_sync() -> vkjit_status;
    pub fn vkjit_host_alloc(bytes: usize, out_ptr: *mut *mut c_void) -> vkjit_status;
    pub fn vkjit_host_free(ptr: *mut c_void) -> vkjit_status;
    pub fn vkjit_ir_create(out_ir: *mut *mut vkjit_ir) -> vkjit_status;
    pub fn vkjit_ir_destroy(ir: *mut vkjit_ir) -> vkjit_status;
    pub fn vkjit_type_struct(ir: *mut vkjit_ir, elems: *const vkjit_type, n: usize, out_ty: *mut vkjit_type) -> vkjit_status;
    pub fn vkjit_type_struct_len(ir: *mut vkjit_ir, ty: vkjit_type, out_n: *mut usize) -> vkjit_status;
    pub fn vkjit_type_struct_elem(ir: *mut vkjit_ir, ty: vkjit_type, i: usize, out_elem: *mut vkjit_type) -> vkjit_status;
    pub fn vkjit_const_f32(ir: *mut vkjit_ir, v: f32, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_const_i32(ir: *mut vkjit_ir, v: i32, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_const_u32(ir: *mut vkjit_ir, v: u32, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_const_bool(ir: *mut vkjit_ir, v: i32, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_f32(ir: *mut vkjit_ir, data: *const f32, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_i32(ir: *mut vkjit_ir, data: *const i32, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_u32(ir: *mut vkjit_ir, data: *const u32, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_bool(ir: *mut vkjit_ir, data: *const u32, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_empty(ir: *mut vkjit_ir, ty: vkjit_type, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_wrap_device(ir: *mut vkjit_ir, ty: vkjit_type, device_ptr: u64, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_wrap_device_owned(ir: *mut vkjit_ir, ty: vkjit_type, device_ptr: u64, n: usize, release: Option<unsafe extern "C" fn(*mut c_void)>, ctx: *mut c_void, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_var_to_dlpack(ir: *mut vkjit_ir, id: vkjit_var, out_managed_tensor: *mut *mut c_void) -> vkjit_status;
    pub fn vkjit_var_from_dlpack(ir: *mut vkjit_ir, managed_tensor: *mut c_void, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_dlpack_delete(managed_tensor: *mut c_void);
    pub fn vkjit_arange(ir: *mut vkjit_ir, ty: vkjit_type, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_linspace(ir: *mut vkjit_ir, ty: vkjit_type, start: vkjit_var, stop: vkjit_var, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_zeros(ir: *mut vkjit_ir, ty: vkjit_type, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_ones(ir: *mut vkjit_ir, ty: vkjit_type, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_cast(ir: *mut vkjit_ir, src: vkjit_var, ty: vkjit_type, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_bop(ir: *mut vkjit_ir, kind: i32, lhs: vkjit_var, rhs: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_uop(ir: *mut vkjit_ir, kind: i32, src: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_bitcast(ir: *mut vkjit_ir, src: vkjit_var, ty: vkjit_type, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_select(ir: *mut vkjit_ir, cond: vkjit_var, lhs: vkjit_var, rhs: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_struct_init(ir: *mut vkjit_ir, elems: *const vkjit_var, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_getattr(ir: *mut vkjit_ir, src: vkjit_var, idx: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_setattr(ir: *mut vkjit_ir, dst: vkjit_var, src: vkjit_var, idx: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_gather(ir: *mut vkjit_ir, src: vkjit_var, idx: vkjit_var, has_active: i32, active: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_scatter(ir: *mut vkjit_ir, src: vkjit_var, dst: vkjit_var, idx: vkjit_var, has_active: i32, active: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_scatter_add(ir: *mut vkjit_ir, src: vkjit_var, dst: vkjit_var, idx: vkjit_var, has_active: i32, active: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_var_type(ir: *mut vkjit_ir, id: vkjit_var, out_ty: *mut vkjit_type) -> vkjit_status;
    pub fn vkjit_var_ref_count(ir: *mut vkjit_ir, id: vkjit_var, out_: *mut u32) -> vkjit_status;
    pub fn vkjit_var_count(ir: *mut vkjit_ir, out_: *mut usize) -> vkjit_status;
    pub fn vkjit_var_deps(ir: *mut vkjit_ir, id: vkjit_var, deps: *mut vkjit_var, cap: usize, out_ndeps: *mut usize, out_has_side_effect: *mut i32, out_side_effect: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_count(ir: *mut vkjit_ir, out_: *mut usize) -> vkjit_status;
    pub fn vkjit_is_buffer(ir: *mut vkjit_ir, id: vkjit_var, out_: *mut i32) -> vkjit_status;
    pub fn vkjit_var_size(ir: *mut vkjit_ir, id: vkjit_var, out_elems: *mut usize) -> vkjit_status;
    pub fn vkjit_var_device_ptr(ir: *mut vkjit_ir, id: vkjit_var, out_ptr: *mut u64) -> vkjit_status;
    pub fn vkjit_inc_ref(ir: *mut vkjit_ir, id: vkjit_var) -> vkjit_status;
    pub fn vkjit_dec_ref(ir: *mut vkjit_ir, id: vkjit_var) -> vkjit_status;
    pub fn vkjit_ir_repr(ir: *mut vkjit_ir, buf: *mut c_char, cap: usize, out_len: *mut usize) -> vkjit_status;
    pub fn vkjit_var_repr(ir: *mut vkjit_ir, id: vkjit_var, buf: *mut c_char, cap: usize, out_len: *mut usize) -> vkjit_status;
    pub fn vkjit_schedule(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize) -> vkjit_status;
    pub fn vkjit_eval(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize) -> vkjit_status;
    pub fn vkjit_read(ir: *mut vkjit_ir, id: vkjit_var, ty: vkjit_type, dst: *mut c_void, bytes: usize) -> vkjit_status;
    pub fn vkjit_reduce(ir: *mut vkjit_ir, red: i32, id: vkjit_var, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_prefix_sum(ir: *mut vkjit_ir, id: vkjit_var, exclusive: i32, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_compress(ir: *mut vkjit_ir, mask: vkjit_var, out_indices: *mut vkjit_var, out_count: *mut usize) -> vkjit_status;
    pub fn vkjit_compress_values(ir: *mut vkjit_ir, values: vkjit_var, mask: vkjit_var, out_values: *mut vkjit_var, out_count: *mut usize) -> vkjit_status;
    pub fn vkjit_dist_unique_id(out_id128: *mut c_void) -> vkjit_status;
    pub fn vkjit_dist_init(rank: i32, world: i32, id128: *const c_void) -> vkjit_status;
    pub fn vkjit_dist_mailbox_handle(out_handle64: *mut c_void) -> vkjit_status;
    pub fn vkjit_dist_mailbox_open(handles: *const c_void, world: i32) -> vkjit_status;
    pub fn vkjit_dist_init_env() -> vkjit_status;
    pub fn vkjit_debug_rendezvous(rank: i32, world: i32, addr: *const c_char, port: i32, blob64: *const c_void, root128: *mut c_void, out_all: *mut c_void, timeout_s: f64) -> vkjit_status;
    pub fn vkjit_dist_set_p2p(on: i32) -> vkjit_status;
    pub fn vkjit_dist_shutdown() -> vkjit_status;
    pub fn vkjit_dist_info(out_rank: *mut i32, out_world: *mut i32) -> vkjit_status;
    pub fn vkjit_shard_range(n: usize, rank: i32, world: i32, out_lo: *mut usize, out_hi: *mut usize) -> vkjit_status;
    pub fn vkjit_arange_sharded(ir: *mut vkjit_ir, ty: vkjit_type, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_sharded(ir: *mut vkjit_ir, ty: vkjit_type, data: *const c_void, n: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_array_shard_local(ir: *mut vkjit_ir, ty: vkjit_type, data: *const c_void, n_local: usize, out_: *mut vkjit_var) -> vkjit_status;
    pub fn vkjit_var_is_sharded(ir: *mut vkjit_ir, id: vkjit_var, out_: *mut i32) -> vkjit_status;
    pub fn vkjit_var_shard_base(ir: *mut vkjit_ir, id: vkjit_var, out_: *mut u64) -> vkjit_status;
    pub fn vkjit_stats(out_: *mut vkjit_stats_t) -> vkjit_status;
    pub fn vkjit_stats_reset() -> vkjit_status;
    pub fn vkjit_cache_clear() -> vkjit_status;
    pub fn vkjit_debug_codegen(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize, compile: i32, buf: *mut c_char, cap: usize, out_len: *mut usize, out_cubin_bytes: *mut usize) -> vkjit_status;
    pub fn vkjit_debug_walk_ns(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize, reps: u32, out_ns: *mut u64, out_nodes: *mut u32) -> vkjit_status;
    pub fn vkjit_debug_eval_bookkeeping(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize) -> vkjit_status;
    pub fn vkjit_debug_codegen_reduce(ir: *mut vkjit_ir, id: vkjit_var, red: i32, compile: i32, buf: *mut c_char, cap: usize, out_len: *mut usize, out_cubin_bytes: *mut usize) -> vkjit_status;
    pub fn vkjit_debug_codegen_scan(ir: *mut vkjit_ir, ids: *const vkjit_var, n: usize, mode: i32, compile: i32, buf: *mut c_char, cap: usize, out_len: *mut usize, out_cubin_bytes: *mut usize) -> vkjit_status;
    pub fn vkjit_debug_reduce_trace(out_: *mut u64, cap_words: usize, out_launches: *mut usize) -> vkjit_status;
}
