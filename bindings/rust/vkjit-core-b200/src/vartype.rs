//! `VarType` and `AsVarType` — libs/vkjit-core/src/vartype.rs:1-83, minus the SPIR-V conversions (`to_spirv`,
//! `From<VarType> for rspirv::sr::Type`, :84-119), which have no meaning for this backend.
use std::any::TypeId;

use vkjit_sys as sys;

pub trait AsVarType {
    fn as_var_type() -> VarType;
}

impl AsVarType for u32 {
    fn as_var_type() -> VarType {
        VarType::U32
    }
}
impl AsVarType for i32 {
    fn as_var_type() -> VarType {
        VarType::I32
    }
}
impl AsVarType for f32 {
    fn as_var_type() -> VarType {
        VarType::F32
    }
}

/// Variant order matters: `derive(Ord)` is the promotion order `Struct < Void < Bool < U32 < I32 < F32`
/// (vartype.rs:24-33); the C ABI's type codes (VKJIT_TY_*) are numbered the same way.
#[derive(Debug, Clone, PartialEq, Eq, Hash, PartialOrd, Ord)]
pub enum VarType {
    Struct(Vec<VarType>),
    Void,
    Bool,
    U32,
    I32,
    F32,
}

impl VarType {
    pub fn name(&self) -> String {
        match self {
            VarType::Void => "Void".into(),
            VarType::Bool => "Bool".into(),
            VarType::U32 => "UInt32".into(),
            VarType::I32 => "Int32".into(),
            VarType::F32 => "Float32".into(),
            _ => unimplemented!(),
        }
    }
    /// Every scalar occupies one 4-byte word in device memory, Bool included (crevice std140, vartype.rs:45-53).
    pub fn stride(&self) -> usize {
        match self {
            VarType::Void => 0,
            VarType::Bool | VarType::U32 | VarType::I32 | VarType::F32 => 4,
            _ => unimplemented!(),
        }
    }
    pub fn size(&self) -> usize {
        self.stride()
    }
    pub fn alignment(&self) -> usize {
        match self {
            VarType::Void => 0,
            VarType::Bool | VarType::U32 | VarType::I32 | VarType::F32 => 4,
            _ => unimplemented!(),
        }
    }
    pub fn type_id(&self) -> TypeId {
        match self {
            VarType::Bool => TypeId::of::<bool>(),
            VarType::U32 => TypeId::of::<u32>(),
            VarType::I32 => TypeId::of::<i32>(),
            VarType::F32 => TypeId::of::<f32>(),
            _ => panic!("Error: {:?} type has no defined type id!", self),
        }
    }

    /// Scalar type -> C ABI code; struct types are interned per `Ir` (see `Ir::type_code`).
    pub(crate) fn scalar_code(&self) -> Option<sys::vkjit_type> {
        match self {
            VarType::Void => Some(sys::VKJIT_TY_VOID),
            VarType::Bool => Some(sys::VKJIT_TY_BOOL),
            VarType::U32 => Some(sys::VKJIT_TY_U32),
            VarType::I32 => Some(sys::VKJIT_TY_I32),
            VarType::F32 => Some(sys::VKJIT_TY_F32),
            VarType::Struct(_) => None,
        }
    }
    pub(crate) fn from_scalar_code(code: sys::vkjit_type) -> Option<VarType> {
        match code {
            sys::VKJIT_TY_VOID => Some(VarType::Void),
            sys::VKJIT_TY_BOOL => Some(VarType::Bool),
            sys::VKJIT_TY_U32 => Some(VarType::U32),
            sys::VKJIT_TY_I32 => Some(VarType::I32),
            sys::VKJIT_TY_F32 => Some(VarType::F32),
            _ => None,
        }
    }
}
