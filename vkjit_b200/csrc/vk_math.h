// vk_math.h — f32 exp / log / sin / cos, ONE implementation for both sides of the parity test.
//
// The reference has no transcendental ops at all (its op set is internal.rs:64-77), so for config 5
// (the Monte-Carlo mega-trace) there is no reference result to be within 1 ulp of: the CPU oracle is the
// specification.  Round 1 lowered these ops to CUDA's expf/logf/sinf/cosf on the device and to glibc on the host and
// could only promise "within 2 ulp of each other".  This header removes the gap: it is compiled by NVRTC into every
// generated kernel that uses a transcendental (the Makefile embeds it as text) and by g++ into the oracle
// (-ffp-contract=off), and it uses nothing but IEEE-754 binary32 add / sub / mul / fma, integer arithmetic and
// comparisons, each spelled out, so both compilers must produce the same bits:  GPU == oracle, bit for bit,
// for every input including denormals, huge arguments, +-inf and NaN (NaN results are the one pattern 0x7fc00000).
//
// Accuracy against the exact result (exhaustive over all 2^32 inputs, tools/vk_math_ulp.cpp, record in
// profiles/r02_vk_math_ulp.md): exp 0.89 ulp, log 0.92 ulp, sin 0.96 ulp, cos 0.97 ulp at worst (full range, huge
// arguments included) — every function is within 1 ulp.
//
// No #includes on the device side: NVRTC sees this text after the generator's typedef prelude.
#ifndef VK_MATH_H
#define VK_MATH_H

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define VKM_FN __device__ __forceinline__
#define VKM_SLOW_FN __device__ __noinline__   /* rarely taken paths: one copy per kernel, not one per call site */
#define VKM_TABLE __device__ __constant__ const
VKM_FN float vkm_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VKM_FN float vkm_mul(float a, float b) { return __fmul_rn(a, b); }
VKM_FN float vkm_add(float a, float b) { return __fadd_rn(a, b); }
VKM_FN float vkm_sub(float a, float b) { return __fsub_rn(a, b); }
VKM_FN unsigned int vkm_bits(float x) { return __float_as_uint(x); }
VKM_FN float vkm_float(unsigned int w) { return __uint_as_float(w); }
VKM_FN int vkm_clz64(unsigned long long v) { return __clzll((long long)v); }
VKM_FN float vkm_i2f(int v) { return __int2float_rn(v); }
#else
#include <string.h>
#define VKM_FN static inline
#define VKM_SLOW_FN static inline
#define VKM_TABLE static const
// -ffp-contract=off: a * b + c below is never fused by the compiler; the only fused operations are the explicit ones
VKM_FN float vkm_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
VKM_FN float vkm_mul(float a, float b) { return a * b; }
VKM_FN float vkm_add(float a, float b) { return a + b; }
VKM_FN float vkm_sub(float a, float b) { return a - b; }
VKM_FN unsigned int vkm_bits(float x) { unsigned int w; memcpy(&w, &x, 4); return w; }
VKM_FN float vkm_float(unsigned int w) { float x; memcpy(&x, &w, 4); return x; }
VKM_FN int vkm_clz64(unsigned long long v) { return __builtin_clzll(v); }
VKM_FN float vkm_i2f(int v) { return (float)v; }
#endif

// minimax coefficients (tools/fit_vk_math.py; each is an exact f32 value)
#define VKM_LOG_C3 0.3333333134651184f
#define VKM_LOG_C4 -0.2500081956386566f
#define VKM_LOG_C5 0.2000124305486679f
#define VKM_LOG_C6 -0.1662341058254242f
#define VKM_LOG_C7 0.14201541244983673f
#define VKM_LOG_C8 -0.13159553706645966f
#define VKM_LOG_C9 0.12762467563152313f
#define VKM_LOG_C10 -0.07636900246143341f
#define VKM_SIN_C3 -0.1666666716337204f
#define VKM_SIN_C5 0.008333379402756691f
#define VKM_SIN_C7 -0.00019853086268994957f
#define VKM_SIN_C9 2.832952077369555e-06f
#define VKM_COS_C4 0.04166664555668831f
#define VKM_COS_C6 -0.0013887310633435845f
#define VKM_COS_C8 2.4432552891084924e-05f

#define VKM_NAN 0x7fc00000u
#define VKM_INF 0x7f800000u
#define VKM_MAGIC 12582912.0f      /* 1.5 * 2^23: adding it rounds to the nearest integer (ties to even) */
#define VKM_MAGIC_BITS 0x4b400000u

// ---------------------------------------------------------------------------------------------------------------
// exp:  x = j ln2 + r, |r| <= ln2 / 2;  e^r by a degree-6 minimax polynomial in Horner form (tools/fit_vk_math.py).
// |x| < 87: the result is a normal number and the scaling by 2^j is an integer add to the exponent field.  The rest
// (NaN, overflow, results in the subnormal range) is one out-of-line copy per kernel.
// ---------------------------------------------------------------------------------------------------------------
VKM_FN float vkm_exp_poly(float r) {
  float q = 0.001382041140459478f;
  q = vkm_fma(q, r, 0.008368677459657192f);
  q = vkm_fma(q, r, 0.04166826978325844f);
  q = vkm_fma(q, r, 0.1666652113199234f);
  q = vkm_fma(q, r, 0.4999999403953552f);
  q = vkm_fma(q, r, 1.0f);
  return vkm_fma(q, r, 1.0f);
}
VKM_SLOW_FN float vkm_exp_slow(float x) {
  const unsigned int ax = vkm_bits(x) & 0x7fffffffu;
  if (ax > VKM_INF) return vkm_float(VKM_NAN);
  if (x > 88.72283935546875f) return vkm_float(VKM_INF);   // e^x >= 2^128 - 2^103 rounds to +inf
  if (x < -103.97208404541016f) return 0.0f;               // e^x <= 2^-150 rounds to +0
  const float t = vkm_fma(x, 1.44269502162933349609375f, VKM_MAGIC);
  const float j = vkm_sub(t, VKM_MAGIC);
  const int n = (int)(vkm_bits(t) - VKM_MAGIC_BITS);
  float r = vkm_fma(j, -0.693145751953125f, x);
  r = vkm_fma(j, -1.42860682030941723212e-6f, r);
  const float p = vkm_exp_poly(r);
  const int n1 = n >> 1, n2 = n - n1;                      // arithmetic shift (both compilers)
  // (p 2^n1) 2^n2: the second multiply is the only rounding when the result is subnormal
  return vkm_mul(vkm_mul(p, vkm_float((unsigned int)(n1 + 127) << 23)), vkm_float((unsigned int)(n2 + 127) << 23));
}
// |x| < 87 (the caller has checked, or the code generator has PROVED it from the ranges of the trace: program.cpp)
VKM_FN float vk_expf_fast(float x) {
  const float t = vkm_fma(x, 1.44269502162933349609375f, VKM_MAGIC);
  const float j = vkm_sub(t, VKM_MAGIC);
  float r = vkm_fma(j, -0.693145751953125f, x);            // ln2 high part: 16 bits, j * hi is exact
  r = vkm_fma(j, -1.42860682030941723212e-6f, r);
  // the low bits of t hold j (the bits of the magic constant vanish in the shift)
  return vkm_float(vkm_bits(vkm_exp_poly(r)) + (vkm_bits(t) << 23));
}
VKM_FN float vk_expf(float x) {
  if ((vkm_bits(x) & 0x7fffffffu) >= 0x42ae0000u) return vkm_exp_slow(x);   // |x| >= 87, inf, NaN
  return vk_expf_fast(x);
}

// ---------------------------------------------------------------------------------------------------------------
// log:  x = m 2^e with m in [sqrt(1/2), sqrt(2)), f = m - 1 (exact);  log1p(f) = f - f^2/2 + f^3 P(f) (degree-10
// minimax);  result = e ln2_hi + f (compensated) + (e ln2_lo + f^2 (f P - 1/2)).  Zero, subnormal, negative, inf and
// NaN arguments take the out-of-line path.
// ---------------------------------------------------------------------------------------------------------------
VKM_FN float vkm_log_core(unsigned int ix, int e) {
  const unsigned int d = ix - 0x3f3504f3u;                 // bits of sqrt(1/2)
  e += (int)d >> 23;                                       // arithmetic shift (both compilers)
  const float f = vkm_sub(vkm_float((d & 0x007fffffu) + 0x3f3504f3u), 1.0f);
  float q = VKM_LOG_C10;
  q = vkm_fma(q, f, VKM_LOG_C9);
  q = vkm_fma(q, f, VKM_LOG_C8);
  q = vkm_fma(q, f, VKM_LOG_C7);
  q = vkm_fma(q, f, VKM_LOG_C6);
  q = vkm_fma(q, f, VKM_LOG_C5);
  q = vkm_fma(q, f, VKM_LOG_C4);
  q = vkm_fma(q, f, VKM_LOG_C3);
  const float small = vkm_mul(vkm_mul(f, f), vkm_fma(f, q, -0.5f));
  const float fe = vkm_i2f(e);
  const float hi = vkm_mul(fe, 0.693145751953125f);        // exact: 16-bit constant times an 8-bit integer
  const float s = vkm_add(hi, f);                          // e != 0: |hi| >= 0.69 > |f|, Fast2Sum gives the exact error;
  const float serr = vkm_add(vkm_sub(hi, s), f);           // e == 0: s = f, serr = 0
  return vkm_add(s, vkm_add(serr, vkm_fma(fe, 1.42860682030941723212e-6f, small)));
}
VKM_SLOW_FN float vkm_log_slow(float x) {
  const unsigned int ix = vkm_bits(x);
  if ((ix << 1) == 0u) return vkm_float(0xff800000u);      // log(+-0) = -inf
  if (ix > VKM_INF) return vkm_float(VKM_NAN);             // negative, -inf or NaN
  if (ix == VKM_INF) return x;
  return vkm_log_core(vkm_bits(vkm_mul(x, 8388608.0f)), -23);   // subnormal: scale by 2^23 (exact)
}
VKM_FN float vk_logf_fast(float x) { return vkm_log_core(vkm_bits(x), 0); }   // x is a positive normal number
VKM_FN float vk_logf(float x) {
  const unsigned int ix = vkm_bits(x);
  if (ix - 0x00800000u >= 0x7f000000u) return vkm_log_slow(x);   // not a positive normal number
  return vkm_log_core(ix, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// sin / cos:  x = n pi/2 + r, |r| <= pi/4 (+ a hair).  |x| <= 105615: three-constant Cody-Waite with fma (the first
// product is exact: the leading constant has 21 bits and n < 2^17); larger: Payne-Hanek with 96 bits of 2/pi selected
// by the exponent, all in integer arithmetic.  sin r = r + r z S(z), cos r = 1 - z/2 + z^2 C(z), z = r^2.
// ---------------------------------------------------------------------------------------------------------------
VKM_TABLE unsigned int vkm_two_over_pi[8] = {0x00000000u, 0xa2f9836eu, 0x4e441529u, 0xfc2757d1u,
                                              0xf534ddc0u, 0xdb629599u, 0x3c439041u, 0xfe5163abu};

// Fast path, 0 < |x| <= 105615: three-constant Cody-Waite.  Returns n (only n mod 4 matters); *r_out + *lo_out = reduced
// argument (|lo| <= ulp(r) / 2).
VKM_FN unsigned int vkm_trig_reduce(float x, float* r_out, float* lo_out) {
  const float t = vkm_fma(x, 0.63661977236758138243f, VKM_MAGIC);
  const float j = vkm_sub(t, VKM_MAGIC);
  const float r1 = vkm_fma(j, -1.57079601287841796875f, x);   // pi/2, leading 21 bits: the product is exact, and so is r1
  // r1 - j c2 as an unevaluated sum: p + pe = j c2 exactly; Fast2Sum is exact because r1 is a multiple of ulp(p)
  // (r1 is a multiple of 2^-24, |p| < 2^-4)
  const float p = vkm_mul(j, 3.139164732601784635335207e-07f);
  const float pe = vkm_fma(j, 3.139164732601784635335207e-07f, -p);
  const float r = vkm_sub(r1, p);
  const float se = vkm_sub(vkm_sub(r1, r), p);
  *r_out = r;
  *lo_out = vkm_fma(j, -5.390302953474238392694851e-15f, vkm_sub(se, pe));
  return vkm_bits(t);                                      // the low bits of t hold n (the magic constant's are zero)
}

// sin(r + lo) = sin r + lo cos r = r + (r z S(z) + lo) up to lo z / 2
VKM_FN float vkm_sin_poly(float r, float lo) {
  const float z = vkm_mul(r, r);
  float s = VKM_SIN_C9;
  s = vkm_fma(s, z, VKM_SIN_C7);
  s = vkm_fma(s, z, VKM_SIN_C5);
  s = vkm_fma(s, z, VKM_SIN_C3);
  return vkm_add(r, vkm_fma(r, vkm_mul(z, s), lo));
}
// cos(r + lo) = cos r - lo sin r = 1 + (z (z C(z) - 1/2) - lo r)
VKM_FN float vkm_cos_poly(float r, float lo) {
  const float z = vkm_mul(r, r);
  float c = VKM_COS_C8;
  c = vkm_fma(c, z, VKM_COS_C6);
  c = vkm_fma(c, z, VKM_COS_C4);
  // (r + lo)^2 = z + (ze + 2 r lo): the part of the square that z's rounding dropped goes into the small term
  const float ze = vkm_fma(r, r, -z);
  const float w = vkm_fma(r, lo, vkm_mul(0.5f, ze));
  return vkm_add(1.0f, vkm_fma(z, vkm_fma(z, c, -0.5f), -w));
}
// n mod 4 selects:  sin(x) = s, c, -s, -c   cos(x) = c, -s, -c, s
VKM_FN float vkm_pick_sin(unsigned int n, float s, float c) {
  return vkm_float(vkm_bits((n & 1u) ? c : s) ^ ((n << 30) & 0x80000000u));
}
VKM_FN float vkm_pick_cos(unsigned int n, float s, float c) {
  return vkm_float(vkm_bits((n & 1u) ? s : c) ^ (((n + 1u) << 30) & 0x80000000u));
}

// Everything the fast path does not take: +-0 (sin(-0) = -0), inf and NaN (NaN), and |x| > 105615, reduced by
// Payne-Hanek: |x| = M 2^E with M the 24-bit significand as an integer; x (2/pi) mod 4 only needs the bits of 2/pi
// from position E - 1 on (higher ones contribute multiples of 4): take 96 of them, multiply by M mod 2^96; bits 95..94
// of the product are the quadrant, the rest the fraction of a quadrant.  One out-of-line copy per kernel.
VKM_SLOW_FN float vkm_sincos_slow(float x, int want_cos) {
  const unsigned int ix = vkm_bits(x), ax = ix & 0x7fffffffu;
  if (ax == 0u) return want_cos ? 1.0f : x;
  if (ax >= VKM_INF) return vkm_float(VKM_NAN);
  const int E = (int)(ax >> 23) - 150;                     // >= -7 here
  const unsigned int M = (ax & 0x007fffffu) | 0x00800000u;
  const int s = E - 2 + 32;                                // first needed bit, counted from the MSB of table word 0
  const int wi = s >> 5, sh = s & 31;
  const unsigned int t0 = vkm_two_over_pi[wi], t1 = vkm_two_over_pi[wi + 1], t2 = vkm_two_over_pi[wi + 2], t3 = vkm_two_over_pi[wi + 3];
  const unsigned int w2 = sh ? (t0 << sh) | (t1 >> (32 - sh)) : t0;
  const unsigned int w1 = sh ? (t1 << sh) | (t2 >> (32 - sh)) : t1;
  const unsigned int w0 = sh ? (t2 << sh) | (t3 >> (32 - sh)) : t2;
  const unsigned long long a0 = (unsigned long long)M * w0;
  const unsigned long long a1 = (unsigned long long)M * w1 + (a0 >> 32);
  const unsigned long long a2 = (unsigned long long)M * w2 + (a1 >> 32);
  const unsigned int P0 = (unsigned int)a0, P1 = (unsigned int)a1, P2 = (unsigned int)a2;
  unsigned int n = P2 >> 30;
  unsigned long long frac = ((unsigned long long)P2 << 34) | ((unsigned long long)P1 << 2) | (unsigned long long)(P0 >> 30);
  bool neg = false;
  if (frac >> 63) { frac = 0ull - frac; n += 1u; neg = true; }   // nearest quadrant: fraction in [-1/2, 1/2)
  float r = 0.0f, lo = 0.0f;
  if (frac != 0ull) {
    const int lz = vkm_clz64(frac);
    const unsigned long long mag = frac << lz;
    const float fh = vkm_mul(vkm_i2f((int)(mag >> 40)), vkm_float((unsigned int)(127 - 24 - lz) << 23));              // top 24 bits
    const float fl = vkm_mul(vkm_i2f((int)((mag >> 16) & 0xffffffu)), vkm_float((unsigned int)(127 - 48 - lz) << 23));  // next 24
    // (fh + fl) * pi/2 with pi/2 = hi + lo
    const float t = vkm_fma(fh, -4.371138828673793e-08f, vkm_mul(fl, 1.57079637050628662109375f));
    r = vkm_fma(fh, 1.57079637050628662109375f, t);
    lo = vkm_add(vkm_fma(fh, 1.57079637050628662109375f, -r), t);   // what the rounding of r dropped
  }
  if (neg != (bool)(ix >> 31)) { r = vkm_float(vkm_bits(r) ^ 0x80000000u); lo = vkm_float(vkm_bits(lo) ^ 0x80000000u); }
  if (ix >> 31) n = 0u - n;                                // sin/cos of -x from those of x
  const float sv = vkm_sin_poly(r, lo), cv = vkm_cos_poly(r, lo);
  return want_cos ? vkm_pick_cos(n, sv, cv) : vkm_pick_sin(n, sv, cv);
}

// fast-path test: 0 < |x| <= 105615  (one unsigned compare: |x| = 0 wraps around)
VKM_FN bool vkm_trig_fast(float x) { return (vkm_bits(x) & 0x7fffffffu) - 1u < 0x47ce4780u; }

// The unchecked forms: +0 <= x <= 105615 (x = +0 comes out right — r = lo = +0, sin = +0, cos = 1 — only -0 needs the
// slow path's sign rule, and a negative x is fine too; the generator only proves the non-negative case).
VKM_FN void vk_sincosf_fast(float x, float* sin_out, float* cos_out) {
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  const float s = vkm_sin_poly(r, lo), c = vkm_cos_poly(r, lo);
  *sin_out = vkm_pick_sin(n, s, c);
  *cos_out = vkm_pick_cos(n, s, c);
}
VKM_FN float vk_sinf_fast(float x) {
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  return vkm_pick_sin(n, vkm_sin_poly(r, lo), vkm_cos_poly(r, lo));
}
VKM_FN float vk_cosf_fast(float x) {
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  return vkm_pick_cos(n, vkm_sin_poly(r, lo), vkm_cos_poly(r, lo));
}
VKM_FN void vk_sincosf(float x, float* sin_out, float* cos_out) {
  if (!vkm_trig_fast(x)) { *sin_out = vkm_sincos_slow(x, 0); *cos_out = vkm_sincos_slow(x, 1); return; }
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  const float s = vkm_sin_poly(r, lo), c = vkm_cos_poly(r, lo);
  *sin_out = vkm_pick_sin(n, s, c);
  *cos_out = vkm_pick_cos(n, s, c);
}
VKM_FN float vk_sinf(float x) {
  if (!vkm_trig_fast(x)) return vkm_sincos_slow(x, 0);
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  return vkm_pick_sin(n, vkm_sin_poly(r, lo), vkm_cos_poly(r, lo));
}
VKM_FN float vk_cosf(float x) {
  if (!vkm_trig_fast(x)) return vkm_sincos_slow(x, 1);
  float r, lo;
  const unsigned int n = vkm_trig_reduce(x, &r, &lo);
  return vkm_pick_cos(n, vkm_sin_poly(r, lo), vkm_cos_poly(r, lo));
}

#endif  // VK_MATH_H
