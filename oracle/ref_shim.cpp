/* oracle/ref_shim.cpp -- TEST INFRASTRUCTURE.  C entry points over the two Eigen-free reference
 * translation units that can be compiled where they lie (sqaodc/common/Random.cpp,
 * sqaodc/cpu/Dot_SIMD.cpp).  Used by tests/test_oracle_rng.py to pin the oracle's MT19937 stream
 * and AVX2 dot product against the reference's own object code.  Built only when /root/reference
 * exists (oracle/Makefile target `ref`); output goes to oracle/_ref/ (git-ignored). */
#include <sqaodc/common/Random.h>
#include <sqaodc/cpu/Dot_SIMD.h>
#include <stdint.h>

extern "C" void ref_mt_stream(unsigned long long seed, int n, uint32_t *out) {
    sqaod::Random r;
    r.seed(seed);
    for (int i = 0; i < n; ++i) out[i] = (uint32_t)r.randInt32();
}
extern "C" void ref_mt_reals(unsigned long long seed, int n, float *f, double *d) {
    sqaod::Random r;
    r.seed(seed);
    for (int i = 0; i < n; ++i) f[i] = r.random<float>();
    r.seed(seed);
    for (int i = 0; i < n; ++i) d[i] = r.random<double>();
}
extern "C" float ref_dot_f32(const float *a, const float *b, int n) { return sqaod_cpu::dot_simd(a, b, n); }
extern "C" double ref_dot_f64(const double *a, const double *b, int n) { return sqaod_cpu::dot_simd(a, b, n); }
