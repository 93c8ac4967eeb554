// Exhaustive accuracy check of vk_math.h: every one of the 2^32 f32 bit patterns through vk_expf / vk_logf / vk_sinf /
// vk_cosf (and vk_sincosf == the two separate calls), against the f64 libm value.  Error in units of the f32 ulp of the
// true result.    g++ -O2 -ffp-contract=off -pthread tools/vk_math_ulp.cpp -o /tmp/vk_math_ulp && /tmp/vk_math_ulp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../vk_math.h"

static double ulp_of(double truth) {
  if (truth == 0.0 || !std::isfinite(truth)) return 0.0;
  int e;
  frexp(fabs(truth), &e);          // |truth| = m 2^e, m in [0.5, 1)
  int ue = e - 24;
  if (ue < -149) ue = -149;        // subnormal range
  return ldexp(1.0, ue);
}

struct Res { double max_ulp = 0; uint32_t at = 0; uint64_t wrong_special = 0, gt_half = 0, n = 0; double sum = 0; };

static uint64_t g_stride = 1;
template <class F, class G>
static Res sweep(F mine, G truth, int threads, uint32_t lo_bits, uint64_t count) {
  std::vector<Res> part(threads);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t)
    th.emplace_back([&, t] {
      Res r;
      for (uint64_t i0 = t; i0 * g_stride < count; i0 += threads) {
        const uint64_t i = i0 * g_stride + (i0 * 2654435761ull >> 7) % g_stride;
        if (i >= count) break;
        const uint32_t w = lo_bits + (uint32_t)i;
        float x; memcpy(&x, &w, 4);
        const float got = mine(x);
        const double want = truth((double)x);
        if (std::isnan(want)) { if (!std::isnan(got)) ++r.wrong_special; continue; }
        if (std::isnan(got)) { ++r.wrong_special; continue; }
        const float want_f = (float)want;
        if (std::isinf(want_f) || want_f == 0.0f || std::isinf(got)) {   // overflow / underflow / exact zero: compare after rounding
          if (got != want_f && !(fabs((double)got - want) <= ulp_of(want == 0 ? 1e-45 : want))) ++r.wrong_special;
          if (want == 0.0 && (std::signbit(got) != std::signbit(want))) ++r.wrong_special;
          continue;
        }
        const double u = ulp_of(want);
        const double err = fabs((double)got - want) / u;
        r.sum += err; ++r.n;
        if (err > 0.5) ++r.gt_half;
        if (err > r.max_ulp) { r.max_ulp = err; r.at = w; }
      }
      part[t] = r;
    });
  for (auto& t : th) t.join();
  Res r;
  for (auto& p : part) {
    if (p.max_ulp > r.max_ulp) { r.max_ulp = p.max_ulp; r.at = p.at; }
    r.wrong_special += p.wrong_special; r.gt_half += p.gt_half; r.n += p.n; r.sum += p.sum;
  }
  return r;
}

int main(int argc, char** argv) {
  const int T = std::max(1u, std::thread::hardware_concurrency());
  const uint64_t all = 1ull << 32;
  if (argc > 1) g_stride = strtoull(argv[1], nullptr, 10);   // > 1: a pseudo-random 1/stride sample instead of every input
  auto report = [](const char* name, const Res& r) {
    printf("%-28s max %.4f ulp at 0x%08x   mean %.4f ulp   > 0.5 ulp (not correctly rounded): %.3f %%   special-case mismatches: %llu   (%llu finite results)\n",
           name, r.max_ulp, r.at, r.sum / (double)r.n, 100.0 * (double)r.gt_half / (double)r.n, (unsigned long long)r.wrong_special, (unsigned long long)r.n);
    fflush(stdout);
  };
  report("vk_expf  all inputs", sweep([](float x) { return vk_expf(x); }, [](double x) { return exp(x); }, T, 0, all));
  report("vk_logf  all inputs", sweep([](float x) { return vk_logf(x); }, [](double x) { return log(x); }, T, 0, all));
  // |x| <= 105615 (fast path) and the rest (Payne-Hanek) separately
  report("vk_sinf  |x| <= 105615", sweep([](float x) { return fabsf(x) <= 105615.0f ? vk_sinf(x) : sinf(x); }, [](double x) { return sin(x); }, T, 0, all));
  report("vk_cosf  |x| <= 105615", sweep([](float x) { return fabsf(x) <= 105615.0f ? vk_cosf(x) : cosf(x); }, [](double x) { return cos(x); }, T, 0, all));
  report("vk_sinf  all inputs", sweep([](float x) { return vk_sinf(x); }, [](double x) { return sin(x); }, T, 0, all));
  report("vk_cosf  all inputs", sweep([](float x) { return vk_cosf(x); }, [](double x) { return cos(x); }, T, 0, all));
  // sincos must be the two separate calls, bit for bit
  {
    std::vector<uint64_t> bad(T, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        for (uint64_t i = t * g_stride; i < all; i += T * g_stride) {
          const uint32_t w = (uint32_t)i;
          float x; memcpy(&x, &w, 4);
          float s, c;
          vk_sincosf(x, &s, &c);
          const float s2 = vk_sinf(x), c2 = vk_cosf(x);
          if (memcmp(&s, &s2, 4) || memcmp(&c, &c2, 4)) ++bad[t];
        }
      });
    for (auto& t : th) t.join();
    uint64_t b = 0;
    for (auto v : bad) b += v;
    printf("vk_sincosf vs vk_sinf/vk_cosf: %llu differing bit patterns over 2^32 inputs\n", (unsigned long long)b);
  }
  return 0;
}
