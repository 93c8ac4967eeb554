//! `Ir`, `Var`, `VarId` — the reference's libs/vkjit-core/src/internal.rs:79-542 with the same public methods, every one
//! of them forwarding to the C ABI of `libvkjit_b200.so` (include/vkjit_b200.h).  The trace itself (vars, ref-counts,
//! schedule, arrays) lives inside the native library; this side keeps only what Rust's borrow rules force it to own:
//!
//! * `Ir::var(&self, id) -> &Var` has to return a reference, so a small `Var` (id + type) is cached per id;
//! * `Ir::as_slice::<T>(&self, id) -> &[T]` and `Array::map(&self) -> &[u8]` return borrows of host memory.  The
//!   reference's arrays are host-visible mapped buffers (backend/vulkan/mod.rs:29-34); device memory here is not, so
//!   the first call reads the array back (one D2H copy through the library's pinned staging ring) and keeps the words
//!   until the next `&mut self` call that can change them (`eval`, `dec_ref_count`).  A borrow handed out earlier
//!   cannot outlive such a call — the borrow checker already forbids it — which is what makes the cache sound.
//!
//! The reference panics on every misuse (`assert!`, `unimplemented!()`, index errors); the C ABI returns a status, and
//! `check` turns it back into a panic carrying the library's message, so front-end behaviour is unchanged.
//!
//! Written blind (no Rust toolchain in the build image).
use std::any::TypeId;
use std::cell::{OnceCell, RefCell};
use std::collections::{HashMap, HashSet};
use std::ffi::CStr;
use std::fmt::Debug;
use std::ops::Deref;
use std::os::raw::{c_char, c_void};

use vkjit_sys as sys;

use crate::backend::cuda::{CudaArray, CudaBackend};
use crate::backend::Backend;
use crate::iterators::{DepIterator, SeIterator};
use crate::vartype::*;

/// status -> panic with the library's thread-local message (the reference panics at the same places)
pub(crate) fn check(status: sys::vkjit_status) {
    if status != sys::VKJIT_OK {
        let msg = unsafe {
            let p = sys::vkjit_last_error();
            if p.is_null() {
                String::from("unknown error")
            } else {
                CStr::from_ptr(p).to_string_lossy().into_owned()
            }
        };
        panic!("vkjit_b200 [status {}]: {}", status, msg);
    }
}

/// Two-call protocol of the string-returning entry points (size query, then fill).
fn read_string(f: impl Fn(*mut c_char, usize, *mut usize) -> sys::vkjit_status) -> String {
    let mut len: usize = 0;
    check(f(std::ptr::null_mut(), 0, &mut len));
    let mut buf = vec![0u8; len + 1];
    check(f(buf.as_mut_ptr() as *mut c_char, buf.len(), &mut len));
    buf.truncate(len);
    String::from_utf8_lossy(&buf).into_owned()
}

// ---- VarId (internal.rs:79-103) -------------------------------------------------------------------------------
#[derive(Clone, Copy, Hash, PartialEq, Eq)]
pub struct VarId(usize);
impl Debug for VarId {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{}", self.0)
    }
}
impl From<usize> for VarId {
    fn from(val: usize) -> Self {
        Self(val)
    }
}
impl VarId {
    pub fn get_id(&self) -> usize {
        self.0
    }
    fn raw(&self) -> sys::vkjit_var {
        self.0 as sys::vkjit_var
    }
}
impl Deref for VarId {
    type Target = usize;

    fn deref(&self) -> &Self::Target {
        &self.0
    }
}

// ---- Var (internal.rs:105-119) --------------------------------------------------------------------------------
/// What a front-end can see of a var: its type (`ty()`) and its `{:?}` text.  Op, deps, side effects and the ref
/// count are read from the native trace when asked for.
#[derive(Clone)]
pub struct Var {
    ir: *mut sys::vkjit_ir,
    id: VarId,
    ty: VarType,
}
impl Var {
    pub fn ty(&self) -> &VarType {
        &self.ty
    }
    /// `var.ref_count` (a `pub(crate)` field in the reference; its own test reads it, test.rs:205)
    pub fn ref_count(&self) -> usize {
        let mut rc: u32 = 0;
        unsafe { check(sys::vkjit_var_ref_count(self.ir, self.id.raw(), &mut rc)) };
        rc as usize
    }
    /// `var.deps` / `var.side_effects` (internal.rs:108-109)
    pub fn deps(&self) -> Vec<VarId> {
        self.edges().0
    }
    pub fn side_effects(&self) -> Vec<VarId> {
        self.edges().1
    }
    fn edges(&self) -> (Vec<VarId>, Vec<VarId>) {
        let mut n: usize = 0;
        let mut has_se: i32 = 0;
        let mut se: sys::vkjit_var = 0;
        unsafe {
            check(sys::vkjit_var_deps(self.ir, self.id.raw(), std::ptr::null_mut(), 0, &mut n, &mut has_se, &mut se));
            let mut deps = vec![0 as sys::vkjit_var; n];
            check(sys::vkjit_var_deps(self.ir, self.id.raw(), deps.as_mut_ptr(), n, &mut n, &mut has_se, &mut se));
            (
                deps.into_iter().map(|d| VarId(d as usize)).collect(),
                if has_se != 0 { vec![VarId(se as usize)] } else { vec![] },
            )
        }
    }
}
impl Debug for Var {
    /// `Var { op: .., deps: [..], side_effects: [..], ty: .., ref_count: .. }` — the derive(Debug) text of the reference
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let (ir, id) = (self.ir, self.id.raw());
        f.write_str(&read_string(|b, c, l| unsafe { sys::vkjit_var_repr(ir, id, b, c, l) }))
    }
}

/// Horizontal reductions (extension: the reference has none)
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Red {
    Sum = 0,
    Min = 1,
    Max = 2,
}

// ---- Ir (internal.rs:126-542) ---------------------------------------------------------------------------------
pub struct Ir {
    pub(crate) backend: CudaBackend,
    h: *mut sys::vkjit_ir,
    vars: RefCell<HashMap<usize, Box<Var>>>,
    arrays: RefCell<HashMap<usize, Box<CudaArray>>>,
    slices: RefCell<HashMap<usize, Box<[u32]>>>,
}

// The native Ir is guarded by its own mutex; the front-ends keep the Rust side in a `Mutex<Ir>` (vkjit-rust/src/lib.rs:9-11).
unsafe impl Send for Ir {}

impl Debug for Ir {
    /// `{:#?}`: `Ir { [0]: Var { .. }, [1]: .. }` (internal.rs:133-144) — produced by the native library in the same format
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let h = self.h;
        f.write_str(&read_string(|b, c, l| unsafe { sys::vkjit_ir_repr(h, b, c, l) }))
    }
}

impl Drop for Ir {
    fn drop(&mut self) {
        self.invalidate_host_views();
        unsafe {
            sys::vkjit_ir_destroy(self.h);
        }
    }
}

macro_rules! bop {
    ($name:ident, $code:expr) => {
        pub fn $name(&mut self, lhs: VarId, rhs: VarId) -> VarId {
            self.bop($code, lhs, rhs)
        }
    };
}
macro_rules! uop {
    ($name:ident, $code:expr) => {
        pub fn $name(&mut self, src: VarId) -> VarId {
            let mut out: sys::vkjit_var = 0;
            unsafe { check(sys::vkjit_uop(self.h, $code, src.raw(), &mut out)) };
            self.register(out)
        }
    };
}

impl Ir {
    /// `Ir::new` (internal.rs:167-182): creates the backend (binds the B200) and an empty trace.
    pub fn new() -> Self {
        let backend = CudaBackend::create();
        let mut h: *mut sys::vkjit_ir = std::ptr::null_mut();
        unsafe { check(sys::vkjit_ir_create(&mut h)) };
        Self { backend, h, vars: RefCell::default(), arrays: RefCell::default(), slices: RefCell::default() }
    }

    // -- plumbing ----------------------------------------------------------------------------------------------
    /// records the type of a freshly returned var (a recycled id may have carried another type before)
    fn register(&mut self, raw: sys::vkjit_var) -> VarId {
        let id = VarId(raw as usize);
        let ty = self.query_type(raw);
        self.vars.borrow_mut().insert(id.0, Box::new(Var { ir: self.h, id, ty }));
        self.arrays.borrow_mut().remove(&id.0);
        self.slices.borrow_mut().remove(&id.0);
        id
    }
    fn query_type(&self, raw: sys::vkjit_var) -> VarType {
        let mut code: sys::vkjit_type = 0;
        unsafe { check(sys::vkjit_var_type(self.h, raw, &mut code)) };
        self.type_from_code(code)
    }
    /// `VarType` -> C ABI type code; struct types are interned per Ir (`vkjit_type_struct`)
    fn type_code(&self, ty: &VarType) -> sys::vkjit_type {
        match ty {
            VarType::Struct(elems) => {
                let codes: Vec<sys::vkjit_type> = elems.iter().map(|e| self.type_code(e)).collect();
                let mut out: sys::vkjit_type = 0;
                unsafe { check(sys::vkjit_type_struct(self.h, codes.as_ptr(), codes.len(), &mut out)) };
                out
            }
            scalar => scalar.scalar_code().unwrap(),
        }
    }
    fn type_from_code(&self, code: sys::vkjit_type) -> VarType {
        if let Some(t) = VarType::from_scalar_code(code) {
            return t;
        }
        let mut n: usize = 0;
        unsafe { check(sys::vkjit_type_struct_len(self.h, code, &mut n)) };
        let mut elems = Vec::with_capacity(n);
        for i in 0..n {
            let mut e: sys::vkjit_type = 0;
            unsafe { check(sys::vkjit_type_struct_elem(self.h, code, i, &mut e)) };
            elems.push(self.type_from_code(e));
        }
        VarType::Struct(elems)
    }
    /// device contents may change: drop every host copy handed out through `as_slice` / `array().map()`
    fn invalidate_host_views(&mut self) {
        self.slices.get_mut().clear();
        self.arrays.get_mut().clear();
    }
    fn bop(&mut self, kind: i32, lhs: VarId, rhs: VarId) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_bop(self.h, kind, lhs.raw(), rhs.raw(), &mut out)) };
        self.register(out)
    }
    /// the native handle (for the interop entry points of `vkjit_sys`: DLPack, device pointers, multi-GPU)
    pub fn handle(&self) -> *mut sys::vkjit_ir {
        self.h
    }

    // -- reference surface -------------------------------------------------------------------------------------
    /// `Ir::array` (internal.rs:183-185): the array behind an evaluated var
    pub fn array(&self, id: VarId) -> &<CudaBackend as Backend>::Array {
        assert!(self.is_buffer(&id), "no array for var {:?}", id);
        let mut map = self.arrays.borrow_mut();
        let entry = map.entry(id.0).or_insert_with(|| {
            let (mut ptr, mut n): (u64, usize) = (0, 0);
            unsafe {
                check(sys::vkjit_var_device_ptr(self.h, id.raw(), &mut ptr));
                check(sys::vkjit_var_size(self.h, id.raw(), &mut n));
            }
            Box::new(CudaArray { ir: self.h, id: Some(id.raw()), ptr, bytes: n * 4, host: OnceCell::new(), owned: false })
        });
        let p: *const CudaArray = &**entry;
        // Sound: the box is only dropped by `&mut self` methods (register / invalidate_host_views), which cannot run
        // while the returned borrow of `self` is alive.
        unsafe { &*p }
    }
    pub fn var(&self, id: VarId) -> &Var {
        let mut map = self.vars.borrow_mut();
        let entry = map.entry(id.0).or_insert_with(|| Box::new(Var { ir: self.h, id, ty: self.query_type(id.raw()) }));
        let p: *const Var = &**entry;
        unsafe { &*p } // see `array`
    }
    pub fn var_mut(&mut self, id: VarId) -> &mut Var {
        self.var(id);
        self.vars.get_mut().get_mut(&id.0).unwrap()
    }

    // Binary operations (internal.rs:146-166, :218-227): both operands are promoted to max(lhs, rhs) inside the
    // library (`bop!`), comparisons return Bool.
    bop!(add, 0);
    bop!(sub, 1);
    bop!(mul, 2);
    bop!(div, 3);
    bop!(lt, 4);
    bop!(gt, 5);
    bop!(eq, 6);
    bop!(leq, 7);
    bop!(geq, 8);
    bop!(neq, 9);

    pub fn select(&mut self, cond_id: VarId, lhs_id: VarId, rhs_id: VarId) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_select(self.h, cond_id.raw(), lhs_id.raw(), rhs_id.raw(), &mut out)) };
        self.register(out)
    }
    pub fn arange(&mut self, ty: VarType, num: usize) -> VarId {
        let code = self.type_code(&ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_arange(self.h, code, num, &mut out)) };
        self.register(out)
    }
    pub fn linspace(&mut self, ty: VarType, start_id: VarId, stop_id: VarId, num: usize) -> VarId {
        let code = self.type_code(&ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_linspace(self.h, code, start_id.raw(), stop_id.raw(), num, &mut out)) };
        self.register(out)
    }
    pub fn zeros(&mut self, ty: VarType) -> VarId {
        let code = self.type_code(&ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_zeros(self.h, code, &mut out)) };
        self.register(out)
    }
    pub fn ones(&mut self, ty: VarType) -> VarId {
        let code = self.type_code(&ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_ones(self.h, code, &mut out)) };
        self.register(out)
    }
    /// identity-eliding cast (internal.rs:283-290): returns `src` itself when the type already matches
    pub fn cast(&mut self, src: VarId, ty: &VarType) -> VarId {
        let code = self.type_code(ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_cast(self.h, src.raw(), code, &mut out)) };
        if out == src.raw() {
            return src;
        }
        self.register(out)
    }
    pub fn struct_init(&mut self, vars: &[VarId]) -> VarId {
        let raw: Vec<sys::vkjit_var> = vars.iter().map(|v| v.raw()).collect();
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_struct_init(self.h, raw.as_ptr(), raw.len(), &mut out)) };
        self.register(out)
    }
    pub fn const_f32(&mut self, val: f32) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_const_f32(self.h, val, &mut out)) };
        self.register(out)
    }
    pub fn const_i32(&mut self, val: i32) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_const_i32(self.h, val, &mut out)) };
        self.register(out)
    }
    pub fn const_u32(&mut self, val: u32) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_const_u32(self.h, val, &mut out)) };
        self.register(out)
    }
    pub fn const_bool(&mut self, val: bool) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_const_bool(self.h, val as i32, &mut out)) };
        self.register(out)
    }
    /// upload (internal.rs:313-348): the slice may be reused as soon as the call returns
    pub fn array_f32(&mut self, data: &[f32]) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_array_f32(self.h, data.as_ptr(), data.len(), &mut out)) };
        self.register(out)
    }
    pub fn array_i32(&mut self, data: &[i32]) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_array_i32(self.h, data.as_ptr(), data.len(), &mut out)) };
        self.register(out)
    }
    pub fn array_u32(&mut self, data: &[u32]) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_array_u32(self.h, data.as_ptr(), data.len(), &mut out)) };
        self.register(out)
    }
    pub fn getattr(&mut self, src_id: VarId, idx: usize) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_getattr(self.h, src_id.raw(), idx, &mut out)) };
        self.register(out)
    }
    pub fn setattr(&mut self, dst_id: VarId, src_id: VarId, idx: usize) -> VarId {
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_setattr(self.h, dst_id.raw(), src_id.raw(), idx, &mut out)) };
        self.register(out)
    }
    /// `out[i] = active[i] ? src[idx[i]] : 0` (the reference's lowering of Gather is broken, internal.rs:1035-1055;
    /// semantics per SURVEY.md Appendix A.1)
    pub fn gather(&mut self, src_id: VarId, idx_id: VarId, active_id: Option<VarId>) -> VarId {
        let mut out: sys::vkjit_var = 0;
        let (has, act) = match active_id {
            Some(a) => (1, a.raw()),
            None => (0, 0),
        };
        unsafe { check(sys::vkjit_gather(self.h, src_id.raw(), idx_id.raw(), has, act, &mut out)) };
        self.register(out)
    }
    pub fn scatter(&mut self, src_id: VarId, dst_id: VarId, idx_id: VarId, active_id: Option<VarId>) -> VarId {
        let mut out: sys::vkjit_var = 0;
        let (has, act) = match active_id {
            Some(a) => (1, a.raw()),
            None => (0, 0),
        };
        unsafe { check(sys::vkjit_scatter(self.h, src_id.raw(), dst_id.raw(), idx_id.raw(), has, act, &mut out)) };
        self.register(out)
    }
    pub fn is_buffer(&self, id: &VarId) -> bool {
        let mut out: i32 = 0;
        unsafe { check(sys::vkjit_is_buffer(self.h, id.raw(), &mut out)) };
        out != 0
    }
    /// `{:?}` of the buffer contents (internal.rs:404-422); Bool buffers print as raw u8 like the reference
    pub fn str(&self, id: VarId) -> String {
        assert!(self.is_buffer(&id), "str() on a var that is not a buffer");
        let h = self.h;
        read_string(|b, c, l| unsafe { sys::vkjit_var_repr(h, id.raw(), b, c, l) })
    }
    pub fn print_buffer(&self, id: VarId) {
        println!("{}", self.str(id));
    }
    /// `as_slice::<T>` (internal.rs:443-449): asserts that `T` is the var's element type, then views the array's words.
    pub fn as_slice<T: bytemuck::Pod>(&self, id: VarId) -> &[T] {
        assert!(self.var(id).ty.type_id() == TypeId::of::<T>());
        let mut map = self.slices.borrow_mut();
        let entry = map.entry(id.0).or_insert_with(|| {
            let mut n: usize = 0;
            let code = self.type_code(&self.var(id).ty);
            unsafe {
                check(sys::vkjit_var_size(self.h, id.raw(), &mut n));
                let mut words = vec![0u32; n].into_boxed_slice();
                check(sys::vkjit_read(self.h, id.raw(), code, words.as_mut_ptr() as *mut c_void, n * 4));
                words
            }
        });
        let p: *const [u32] = &**entry;
        bytemuck::cast_slice(unsafe { &*p }) // see `array` for why the borrow may outlive the RefCell guard
    }
    pub fn dec_ref_count(&mut self, id: VarId) {
        self.invalidate_host_views();
        unsafe { check(sys::vkjit_dec_ref(self.h, id.raw())) };
    }
    pub fn inc_ref_count(&mut self, id: VarId) {
        unsafe { check(sys::vkjit_inc_ref(self.h, id.raw())) };
    }
    pub fn schedule(&mut self, schedule: &[VarId]) {
        let raw: Vec<sys::vkjit_var> = schedule.iter().map(|v| v.raw()).collect();
        unsafe { check(sys::vkjit_schedule(self.h, raw.as_ptr(), raw.len())) };
    }
    /// `Ir::eval` (internal.rs:482-525): compile-or-lookup + launch, asynchronous; the scheduled vars become
    /// Bindings owning fresh arrays.  Reads (`as_slice`, `str`) synchronise.
    pub fn eval(&mut self, schedule: &[VarId]) {
        self.invalidate_host_views();
        let raw: Vec<sys::vkjit_var> = schedule.iter().map(|v| v.raw()).collect();
        unsafe { check(sys::vkjit_eval(self.h, raw.as_ptr(), raw.len())) };
    }
    pub fn iter_dep(&self, root: &[VarId]) -> DepIterator {
        DepIterator { ir: self, stack: Vec::from(root), discovered: HashSet::default() }
    }
    pub fn iter_se(&self, root: &[VarId]) -> SeIterator {
        SeIterator { ir: self, stack: Vec::from(root), discovered: HashSet::default() }
    }

    // -- white-box counters the reference's own tests read (test.rs:204-206) -------------------------------------
    /// `ir.vars.len()`
    pub fn var_count(&self) -> usize {
        let mut n: usize = 0;
        unsafe { check(sys::vkjit_var_count(self.h, &mut n)) };
        n
    }
    /// `ir.arrays.len()`
    pub fn array_count(&self) -> usize {
        let mut n: usize = 0;
        unsafe { check(sys::vkjit_array_count(self.h, &mut n)) };
        n
    }

    // -- extensions (no reference counterpart) --------------------------------------------
    bop!(and, 16);
    bop!(or, 17);
    bop!(xor, 18);
    bop!(shl, 19);
    bop!(shr, 20);
    bop!(min, 21);
    bop!(max, 22);
    uop!(neg, 0);
    uop!(abs, 1);
    uop!(not, 2);
    uop!(sqrt, 3);
    uop!(exp, 4);
    uop!(log, 5);
    uop!(sin, 6);
    uop!(cos, 7);
    pub fn bitcast(&mut self, src: VarId, ty: &VarType) -> VarId {
        let code = self.type_code(ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_bitcast(self.h, src.raw(), code, &mut out)) };
        self.register(out)
    }
    pub fn array_bool(&mut self, data: &[bool]) -> VarId {
        let words: Vec<u32> = data.iter().map(|b| *b as u32).collect();
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_array_bool(self.h, words.as_ptr(), words.len(), &mut out)) };
        self.register(out)
    }
    /// `dst[idx[i]] += src[i]` atomically (u32/i32 bit-exact, f32 order dependent)
    pub fn scatter_add(&mut self, src_id: VarId, dst_id: VarId, idx_id: VarId, active_id: Option<VarId>) -> VarId {
        let mut out: sys::vkjit_var = 0;
        let (has, act) = match active_id {
            Some(a) => (1, a.raw()),
            None => (0, 0),
        };
        unsafe { check(sys::vkjit_scatter_add(self.h, src_id.raw(), dst_id.raw(), idx_id.raw(), has, act, &mut out)) };
        self.register(out)
    }
    /// horizontal reduction to a 1-element array; on a sharded operand the per-GPU partials are combined over NVLink
    pub fn reduce(&mut self, red: Red, id: VarId) -> VarId {
        self.invalidate_host_views();
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_reduce(self.h, red as i32, id.raw(), &mut out)) };
        self.register(out)
    }
    pub fn prefix_sum(&mut self, id: VarId, exclusive: bool) -> VarId {
        self.invalidate_host_views();
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_prefix_sum(self.h, id.raw(), exclusive as i32, &mut out)) };
        self.register(out)
    }
    /// indices of the lanes whose mask is set (stable order) and their number
    pub fn compress(&mut self, mask: VarId) -> (VarId, usize) {
        self.invalidate_host_views();
        let (mut out, mut count): (sys::vkjit_var, usize) = (0, 0);
        unsafe { check(sys::vkjit_compress(self.h, mask.raw(), &mut out, &mut count)) };
        (self.register(out), count)
    }
    pub fn compress_values(&mut self, values: VarId, mask: VarId) -> (VarId, usize) {
        self.invalidate_host_views();
        let (mut out, mut count): (sys::vkjit_var, usize) = (0, 0);
        unsafe { check(sys::vkjit_compress_values(self.h, values.raw(), mask.raw(), &mut out, &mut count)) };
        (self.register(out), count)
    }
    /// this rank's contiguous shard of `arange(ty, n)` (one process per GPU; see `vkjit_sys::vkjit_dist_init_env`)
    pub fn arange_sharded(&mut self, ty: VarType, num: usize) -> VarId {
        let code = self.type_code(&ty);
        let mut out: sys::vkjit_var = 0;
        unsafe { check(sys::vkjit_arange_sharded(self.h, code, num, &mut out)) };
        self.register(out)
    }
    /// owned copy of an evaluated var (what `Var::to_vec` of vkjit-rust builds from `as_slice`)
    pub fn to_vec<T: bytemuck::Pod>(&self, id: VarId) -> Vec<T> {
        Vec::from(self.as_slice::<T>(id))
    }
    /// wait for everything enqueued so far (eval is asynchronous; reads synchronise on their own)
    pub fn sync(&self) {
        unsafe { check(sys::vkjit_sync()) };
    }
}
