// C++ client of include/vkjit.hpp, the header-only mirror of the reference's `vkjit-rust` crate.
//   cpp_client            trace construction, typing, ownership, error mapping — needs no GPU
//   cpp_client --device   additionally the reference's programs on the device: src/main.rs, the two front-end tests
//                         of libs/vkjit-rust/src/types.rs:213-242 and the config-1 trace x*y + c
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>

#include "vkjit.hpp"

using namespace vkjit;

#define CHECK(x)                                                          \
  do {                                                                    \
    if (!(x)) {                                                           \
      std::fprintf(stderr, "FAILED: %s (line %d)\n", #x, __LINE__);      \
      return 1;                                                           \
    }                                                                     \
  } while (0)

static uint32_t ref_count(const Var& v) {
  uint32_t rc = 0;
  vkjit_var_ref_count(detail::ir(), v.id(), &rc);
  return rc;
}

static int host_part() {
  Var x = arange(F32, 8);
  Var i = arange(U32, 8);
  // From<f32|i32|u32|bool> and the autocast rules (internal.rs:146-166): U32 < I32 < F32
  CHECK(Var(2.0f).ty() == F32 && Var(7.).ty() == F32 && Var(-3).ty() == I32 && Var(3u).ty() == U32 && Var(true).ty() == Bool);
  CHECK((i + 1u).ty() == U32 && (i + (-1)).ty() == I32 && (i + 1.0f).ty() == F32 && (x * x).ty() == F32);
  CHECK(x.lt(1.0f).ty() == Bool && x.geq(x).ty() == Bool && i.neq(3u).ty() == Bool);
  // operand order of the operators: lhs first
  Var d = 2.0f - x;
  CHECK(d.repr().rfind("Var(Var { op: Bop(Sub), deps: [", 0) == 0);
  // Clone / Drop own one count each (types.rs:128-140); a move does not touch the count
  CHECK(ref_count(x) == 2);  // the handle `x` and d's dependency edge; every temporary above is gone
  {
    Var c = x;  // Clone
    CHECK(c.id() == x.id() && ref_count(x) == 3);
    Var m = std::move(c);
    CHECK(ref_count(x) == 3);
    const Var& alias = x;
    x = alias;  // self-assignment through the by-value operator= is a clone + drop
    CHECK(ref_count(x) == 3);
  }
  CHECK(ref_count(x) == 2);
  // *Assign forms replace the variable (types.rs:55-70)
  Var acc = x;
  const vkjit_var before = acc.id();
  acc += 1.0f;
  acc *= x;
  CHECK(acc.id() != before && acc.ty() == F32 && ref_count(x) == 4);  // handle, d, and both uses inside acc
  // struct vars: From<&[Var]>, getattr / setattr (types.rs:99-105, :149-159)
  Var st = Var::structure({x, Var(2.5f)});
  CHECK(st.getattr(1).ty() == F32);
  Var z = zeros(Struct({F32, U32}));
  z.setattr(i, 1);
  CHECK(z.getattr(1).ty() == U32 && z.getattr(0).ty() == F32);
  // select requires equal operand types (internal.rs:232): the reference panics, the header throws
  bool threw = false;
  try {
    select(x.lt(1.0f), x, 1u);
  } catch (const Error& e) {
    threw = e.status == VKJIT_ERR_TYPE;
  }
  CHECK(threw);
  CHECK(repr_ir().rfind("Ir {", 0) == 0);
  std::ostringstream os;
  os << Var(1.5f);
  CHECK(os.str() == "Var(Var { op: Const(Float32(1.5)), deps: [], side_effects: [], ty: F32, ref_count: 1 })");
  if (!vkjit_is_initialized()) {  // no B200: evaluating and uploading fail loudly, there is no CPU path
    threw = false;
    try {
      eval(acc);
    } catch (const Error& e) {
      threw = e.status == VKJIT_ERR_NO_DEVICE;
    }
    CHECK(threw);
    threw = false;
    try {
      Var up(std::vector<float>{1.f, 2.f});
    } catch (const Error& e) {
      threw = e.status == VKJIT_ERR_NO_DEVICE;
    }
    CHECK(threw);
  }
  return 0;
}

static int device_part() {
  CHECK(vkjit_is_initialized());
  {  // src/main.rs: arange -> eval! -> dbg!
    Var x = arange(U32, 10);
    eval(x);
    CHECK(x.repr() == "Var(\"[0, 1, 2, 3, 4, 5, 6, 7, 8, 9]\")");
    CHECK((x.to_vec<uint32_t>() == std::vector<uint32_t>{0, 1, 2, 3, 4, 5, 6, 7, 8, 9}));
  }
  {  // vkjit-rust types.rs:214-228 `setattr`
    Var x(std::vector<float>{1.f, 2.f, 3.f});
    Var st = zeros(Struct({F32, F32}));
    st.setattr(x, 0);
    Var st_1 = st.getattr(0), st_2 = st.getattr(1);
    eval(st_1, st_2);
    CHECK((st_1.to_vec<float>() == std::vector<float>{1.f, 2.f, 3.f}));
    CHECK((st_2.to_vec<float>() == std::vector<float>{0.f, 0.f, 0.f}));
  }
  {  // vkjit-rust types.rs:230-242 `test_scatter`
    Var x(std::vector<float>{1.f, 2.f, 3.f});
    Var y(7.);
    Var idx = arange(U32, 3);
    y.scatter(x, idx);
    eval(y);
    CHECK((x.to_vec<float>() == std::vector<float>{7.f, 7.f, 7.f}));
  }
  {  // test.rs:10-20 linspace golden + config 1: z = arange * y + c with readback
    Var l = linspace(2.0f, 4.0f, 4);
    eval(l);
    CHECK((l.to_vec<float>() == std::vector<float>{2.f, 2.5f, 3.f, 3.5f}));
    const size_t n = 1 << 20;
    std::vector<float> ys(n);
    for (size_t k = 0; k < n; ++k) ys[k] = (float)((k * 2654435761u) >> 8 & 0xFFFFFF) * (1.0f / 16777216.0f);
    Var y(ys);
    Var z = arange(F32, n) * y + 0.5f;
    eval(z);
    const std::vector<float> got = z.to_vec<float>();
    CHECK(got.size() == n);
    for (size_t k = 0; k < n; ++k) {
      const volatile float prod = (float)k * ys[k];   // two roundings, as OpFMul + OpFAdd
      const float want = prod + 0.5f;
      if (std::memcmp(&want, &got[k], 4) != 0) {
        std::fprintf(stderr, "lane %zu: %g != %g\n", k, got[k], want);
        return 1;
      }
    }
    // extensions: reduction and stream compaction through the same front-end
    Var i = arange(U32, 1000);
    eval(i);
    CHECK(i.sum().to_vec<uint32_t>()[0] == 499500u);
    auto kept = compress(i, (i & 1u).eq(0u));
    CHECK(kept.second == 500 && kept.first.to_vec<uint32_t>()[2] == 4u);
  }
  sync();
  return 0;
}

int main(int argc, char** argv) {
  const bool device = argc > 1 && std::strcmp(argv[1], "--device") == 0;
  try {
    if (device) detail::check(vkjit_init(-1));  // fails loudly without a B200
    if (int rc = host_part()) return rc;
    if (device)
      if (int rc = device_part()) return rc;
  } catch (const Error& e) {
    std::fprintf(stderr, "vkjit::Error: %s\n", e.what());
    return 2;
  }
  std::printf("cpp client ok%s\n", device ? " (device)" : "");
  return 0;
}
