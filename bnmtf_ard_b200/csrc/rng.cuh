// Counter-based RNG (Philox4x32-10) and the device samplers / moments that
// replace code/models/distributions/*.py of the reference.
//
// Distributional contract: the reference draws through numpy's
// MT19937 and Chopin's table sampler (rtnorm.py); its stream cannot be matched,
// so these are exact samplers of the same densities, KS-tested in
// tests/test_gpu_distributions.py.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace bnmtf {

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
  uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
  uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  uint32_t k[2] = {k0, k1};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
}

// Purposes (stream tags) -- part of the counter so that streams never collide.
enum : uint32_t { RNG_TN_U = 1, RNG_TN_V = 2, RNG_LAMBDA = 3, RNG_TAU = 4, RNG_INIT_U = 5, RNG_INIT_V = 6, RNG_API = 7, RNG_S = 8, RNG_INIT_S = 9 };

// One stream = (seed, purpose, index, k, iteration); successive calls bump a
// draw counter.  Results depend only on these identifiers, never on the launch
// geometry or on the number of GPUs.
struct Rng {
  uint32_t k0, k1;       // key = seed
  uint32_t c0, c1, c2;   // index (row), k | purpose<<24, iteration
  uint32_t draw;         // bumped per Philox call
  uint32_t buf[4];
  int have;              // 64-bit words left in buf (0..2)

  __device__ __forceinline__ Rng(uint64_t seed, uint32_t purpose, uint64_t index, uint32_t k, uint32_t iteration)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)index),
        c1((k & 0xFFFFFFu) | (purpose << 24)), c2(iteration ^ ((uint32_t)(index >> 32) * 0x9E3779B1u)), draw(0), have(0) {}

  __device__ __forceinline__ uint64_t next64() {
    if (have == 0) {
      uint32_t c[4] = {c0, c1, c2, draw++};
      philox4x32_10(c, k0, k1);
      buf[0] = c[0]; buf[1] = c[1]; buf[2] = c[2]; buf[3] = c[3];
      have = 2;
    }
    --have;
    return ((uint64_t)buf[2 * have + 1] << 32) | buf[2 * have];
  }
  // uniform on the open interval (0,1), 53 random bits
  __device__ __forceinline__ double u01() { return ((double)(next64() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  __device__ __forceinline__ double normal() { return normcdfinv(u01()); }
};

// ---- exponential_draw(lambda): Exp with RATE lambda (exponential.py:7-9) ----
__device__ __forceinline__ double exp_draw(double lambda, Rng& rng) { return -log(rng.u01()) / lambda; }

// ---- gamma_draw(alpha,beta): shape alpha, RATE beta (gamma.py:11-14) --------
// Marsaglia & Tsang (2000); alpha < 1 boosted by U^(1/alpha).
static __device__ __noinline__ double gamma_draw(double alpha, double beta, Rng& rng) {
  double boost = 1.0;
  if (alpha < 1.0) { boost = pow(rng.u01(), 1.0 / alpha); alpha += 1.0; }
  const double d = alpha - 1.0 / 3.0, c = rsqrt(9.0 * d);
  for (int it = 0; it < 1000; ++it) {
    double x = rng.normal();
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = rng.u01();
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v / beta;
  }
  return boost * d / beta;  // unreachable in practice (acceptance > 0.95 per trial)
}

// ---- TN_draw(mu,tau): N(mu,1/tau) truncated to [0,inf) ----------------------
// truncated_normal.py:37-44, truncated_normal_vector.py:37-50, rtnorm.py:16-218.
// Semantics kept: tau == 0 -> 0; a draw that is negative / inf / NaN -> 0.
// Method (exact for the target density p(x) ~ exp(-tau (x-mu)^2/2) 1[x>=0]):
//   a = -mu sqrt(tau) is the standardised lower bound.
//   a <= 4 : inverse CDF in the upper tail,  P(Z >= z) = u * P(Z >= a).
//   a  > 4 : Robert (1995) translated-exponential rejection; the excess z - a
//            is returned directly so that mu + sigma z never cancels.
static __device__ __noinline__ double tn_draw(double mu, double tau, Rng& rng) {
  if (tau == 0.0) return 0.0;
  const double sigma = 1.0 / sqrt(tau);
  const double a = -mu / sigma;
  double d;
  if (!(a > 4.0)) {
    const double q = 0.5 * erfc(a * 0.70710678118654752440);  // P(Z >= a)
    const double z = -normcdfinv(rng.u01() * q);
    d = mu + sigma * z;
  } else {
    const double alpha = 0.5 * (a + sqrt(a * a + 4.0));
    double x = 0.0;
    for (int it = 0; it < 1000; ++it) {
      x = -log(rng.u01()) / alpha;             // z = a + x
      const double t = (a + x) - alpha;
      if (log(rng.u01()) <= -0.5 * t * t) break;
    }
    d = sigma * x;
  }
  return (d >= 0.0 && isfinite(d)) ? d : 0.0;
}

// Same draw with the inputs the sweeps have at hand -- num = -lambda + tau*b and
// prec = tau*s, i.e. mu = num/prec -- and with the first uniform already drawn (the
// Philox rounds run before the row sums arrive).  No division or square root on the
// critical path: sigma = rsqrt(prec), a = -mu/sigma = -num*sigma, and the draw is
// sigma*(z - a), which also never cancels.  mu_out returns num/prec (to 1-2 ulp).
static __device__ __noinline__ double tn_draw_pre(double num, double prec, double u1, Rng& rng, double& mu_out) {
  if (prec == 0.0) { mu_out = num / prec; return 0.0; }
  const double sigma = rsqrt(prec);
  mu_out = num * sigma * sigma;
  const double a = -num * sigma;
  double x;   // z - a >= 0
  if (!(a > 4.0)) {
    const double q = 0.5 * erfc(a * 0.70710678118654752440);
    x = -normcdfinv(u1 * q) - a;
  } else {
    const double alpha = 0.5 * (a + sqrt(a * a + 4.0));
    double u = u1;
    x = 0.0;
    for (int it = 0; it < 1000; ++it) {
      x = -log(u) / alpha;
      const double t = (a + x) - alpha;
      if (log(rng.u01()) <= -0.5 * t * t) break;
      u = rng.u01();
    }
  }
  const double d = sigma * x;
  return (d >= 0.0 && isfinite(d)) ? d : 0.0;
}

// ---- truncated-normal draw with the random material prepared ahead ------------
// The conditional (num, prec) of a Gibbs update arrives at the end of a latency chain; the
// inverse-CDF draw above then costs erfc + normcdfinv on that chain.  Everything random can be
// prepared before the conditional is known: two standard normals (Box-Muller), two Exp(1)
// variates e and two l = -2 log u.  With a = -mu/sigma the standardised bound:
//   a <= A0 : plain rejection, accept z if z >= a            (acceptance Phi(-a) >= 0.67)
//   a >  A0 : Robert (1995): x = e/alpha, alpha = (a + sqrt(a^2+4))/2, accept if l >= (a+x-alpha)^2
//             (valid for any a; acceptance >= 0.67 beyond A0 = -0.45, -> 1 for large a)
// two proposals each; only when both are rejected (<= 11 %, typically ~0.1 %) the draw falls back to the
// inverse CDF with a fresh uniform.  Every branch returns an exact TN(mu, 1/prec) variate and
// the branch taken depends on a only, so the mixture is exact as well.
constexpr int TN_PRE_DOUBLES = 6;
constexpr double TN_FAST_A0 = -0.45;
struct TnPre { double z1, z2, e1, l1, e2, l2; };

__device__ __forceinline__ TnPre tn_pre_material(Rng& rng) {
  TnPre m;
  const double u1 = rng.u01(), u2 = rng.u01(), u3 = rng.u01(), u4 = rng.u01();
  const double g1 = -log(u1), g2 = -log(u2), g3 = -log(u3), g4 = -log(u4);
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  const double r = sqrt(2.0 * g1);
  m.z1 = r * cs; m.z2 = r * sn;       // Box-Muller: independent standard normals (used when a <= A0)
  m.e1 = g1; m.l1 = 2.0 * g2;         // Exp(1) and -2 log u pairs (used when a > A0)
  m.e2 = g3; m.l2 = 2.0 * g4;
  return m;
}

// out-of-line version for kernels whose registers are busy when the material is prepared
static __device__ __noinline__ void tn_pre_material_store(uint64_t seed, uint32_t purpose, uint64_t index, uint32_t k, uint32_t iteration,
                                                   double* dst) {
  Rng rng(seed, purpose, index, k, iteration);
  const TnPre m = tn_pre_material(rng);
  dst[0] = m.z1; dst[1] = m.z2; dst[2] = m.e1; dst[3] = m.l1; dst[4] = m.e2; dst[5] = m.l2;
}
__device__ __forceinline__ TnPre tn_pre_load(const double* src) {
  TnPre m;
  m.z1 = src[0]; m.z2 = src[1]; m.e1 = src[2]; m.l1 = src[3]; m.e2 = src[4]; m.l2 = src[5];
  return m;
}

// rng_fb: the same stream as the material, positioned after it (draw = 2)
__device__ __forceinline__ double tn_draw_fast(double num, double prec, const TnPre& m, Rng& rng_fb, double& mu_out,
                                               bool* fell_back = nullptr) {
  if (prec == 0.0) { mu_out = num / prec; return 0.0; }
  const double sigma = rsqrt(prec);
  mu_out = num * sigma * sigma;
  const double a = -num * sigma;
  double x;      // z - a >= 0
  bool ok;
  if (!(a > TN_FAST_A0)) {
    ok = (m.z1 >= a);
    x = m.z1 - a;
    if (!ok) { ok = (m.z2 >= a); x = m.z2 - a; }
  } else {
    // alpha = (a + sqrt(a^2 + 4)) / 2 and 1 / alpha = (sqrt(a^2 + 4) - a) / 2: no division on the chain (the
    // difference loses ~a^2 ulp, so very large bounds keep the division)
    const double sq = sqrt(a * a + 4.0);
    const double alpha = 0.5 * (a + sq);
    const double ia = a < 256.0 ? 0.5 * (sq - a) : 1.0 / alpha, am = a - alpha;
    x = m.e1 * ia;
    double t = am + x;
    ok = (m.l1 >= t * t);
    if (!ok) { x = m.e2 * ia; t = am + x; ok = (m.l2 >= t * t); }
  }
  if (fell_back) *fell_back = !ok;
  if (!ok) return tn_draw_pre(num, prec, rng_fb.u01(), rng_fb, mu_out);
  const double d = sigma * x;
  return (d >= 0.0 && isfinite(d)) ? d : 0.0;
}

// Out-of-line fallback of tn_draw_lazy: opens the row's Philox stream behind the prepared material (draw = 2) and
// draws by the inverse CDF.  Everything it needs travels by value, so the caller keeps its state in registers.
static __device__ __noinline__ double2 tn_draw_fallback(double num, double prec, uint64_t seed, uint32_t purpose, uint64_t index,
                                                 uint32_t k, uint32_t iteration) {
  Rng rng(seed, purpose, index, k, iteration);
  rng.draw = 2;                      // Philox blocks 0 and 1 of this stream made the material
  double mu;
  const double x = tn_draw_pre(num, prec, rng.u01(), rng, mu);
  return make_double2(x, mu);
}

// tn_draw_fast for the cluster sweep's control warp, where the conditional sits on the per-column dependency chain:
// the same draws bit for bit, but the fallback stream is only built when both proposals were rejected and neither it nor
// the mean are ever addressed (tn_draw_fast hands `Rng&` / `double&` to an out-of-line function, which pins both in local
// memory: five stores and a dependent load per conditional).
__device__ __forceinline__ double tn_draw_lazy(double num, double prec, const TnPre& m, uint64_t seed, uint32_t purpose,
                                               uint64_t index, uint32_t k, uint32_t iteration, double& mu_out) {
  if (prec == 0.0) { mu_out = num / prec; return 0.0; }
  const double sigma = rsqrt(prec);
  double mu = num * sigma * sigma;
  const double a = -num * sigma;
  double x;      // z - a >= 0
  bool ok;
  if (!(a > TN_FAST_A0)) {
    ok = (m.z1 >= a);
    x = m.z1 - a;
    if (!ok) { ok = (m.z2 >= a); x = m.z2 - a; }
  } else {
    const double sq = sqrt(a * a + 4.0);
    const double alpha = 0.5 * (a + sq);
    const double ia = a < 256.0 ? 0.5 * (sq - a) : 1.0 / alpha, am = a - alpha;
    x = m.e1 * ia;
    double t = am + x;
    ok = (m.l1 >= t * t);
    if (!ok) { x = m.e2 * ia; t = am + x; ok = (m.l2 >= t * t); }
  }
  double d = sigma * x;
  d = (d >= 0.0 && isfinite(d)) ? d : 0.0;
  if (!ok) {
    const double2 fb = tn_draw_fallback(num, prec, seed, purpose, index, k, iteration);
    d = fb.x; mu = fb.y;
  }
  mu_out = mu;
  return d;
}

// ---- TN_vector_expectation / TN_vector_variance -----------------------------
// truncated_normal_vector.py:53-73 (same operation order as the reference so
// that the results agree to rounding): sigma = 1/sqrt(tau); x = -mu/sigma;
// lambda(x) = pdf(x) / (0.5 erfc(x/sqrt2)); mean = mu + sigma lambda;
// var = sigma^2 (1 - lambda (lambda - x)); if mu < -30 sigma: mean = 1/(|mu| tau),
// var = mean^2; anything negative / inf / NaN -> 0.
static __device__ __noinline__ void tn_moments(double mu, double tau, double& e, double& v) {
  const double sigma = 1.0 / sqrt(tau);
  if (mu < -30.0 * sigma) {
    e = 1.0 / (fabs(mu) * tau);
    v = e * e;
  } else {
    const double x = -mu / sigma;
    const double pdf = exp(-0.5 * x * x) / 2.5066282746310002;  // sqrt(2 pi)
    const double lambdax = pdf / (0.5 * erfc(x / 1.4142135623730951));
    e = mu + sigma * lambdax;
    v = (sigma * sigma) * (1.0 - lambdax * (lambdax - x));
  }
  if (!(e >= 0.0 && isfinite(e))) e = 0.0;
  if (!(v >= 0.0 && isfinite(v))) v = 0.0;
}

// log(0.5 erfc(x)) as the reference's ELBO evaluates it (bnmf_vb.py:222,225; bnmtf_vb.py:286-292): scipy.special.erfc is
// cephes' erfc, which returns exactly 0 once x * x exceeds MAXLOG = 709.78... (x > 26.64) instead of a denormal, so the
// reference's ELBO is -inf from there on.  CUDA's erfc keeps returning denormals up to x = 27.2; the cut is mirrored so
// that the ELBO turns -inf in the same iteration as the reference's (tests/test_gpu_configs.py, GDSC matrix).
__device__ __forceinline__ double log_half_erfc_ref(double x) {
  const double c = (x > 0.0 && x * x > 7.09782712893383996843e2) ? 0.0 : erfc(x);
  return log(0.5 * c);
}

// digamma for x > 0 (scipy.special.psi in gamma.py:22-24): recurrence up to
// x >= 10, then the asymptotic series (error < 1e-16 relative there).
__device__ inline double digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0
                   + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

}  // namespace bnmtf
