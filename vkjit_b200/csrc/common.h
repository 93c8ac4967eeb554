// common.h — status codes, error type and small helpers shared by the native library.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/vkjit_b200.h"

namespace vkjit {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Error(code, msg); }

using VarId = uint32_t;
using TypeId = uint32_t;  // scalar: VKJIT_TY_*; struct: VKJIT_TY_STRUCT_BASE + index into Ir::struct_types

inline bool ty_is_scalar(TypeId t) { return t >= VKJIT_TY_BOOL && t <= VKJIT_TY_F32; }
inline bool ty_is_struct(TypeId t) { return t >= VKJIT_TY_STRUCT_BASE; }
inline bool ty_is_num(TypeId t) { return t == VKJIT_TY_U32 || t == VKJIT_TY_I32 || t == VKJIT_TY_F32; }

inline uint32_t f32_bits(float f) { uint32_t w; memcpy(&w, &f, 4); return w; }
inline float bits_f32(uint32_t w) { float f; memcpy(&f, &w, 4); return f; }

uint64_t now_ns();

}  // namespace vkjit
