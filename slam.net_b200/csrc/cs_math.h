// cs_math.h — scalar numerics shared by every kernel (device) and by the host-side
// candidate-table utility of the C-ABI library.  Header-only, no dependencies.
//
// Everything here is built from IEEE-754 basic operations (+ - * / sqrt fma, int ops) so the
// same source gives bit-identical results from nvcc (-fmad=false) and g++ (-ffp-contract=off).
// That is what makes the Philox production mode reproducible on the CPU and lets the device
// compute cosf/sinf itself in verification mode.
//
// What each piece replaces in the reference (paths relative to /root/reference):
//   cs_cvt_i32      (int)float casts           CoreSLAM/CoreSLAMProcessor.cs:240-241, 505-506, 521-530
//   cs_cosf/cs_sinf MathF.Cos / MathF.Sin      CoreSLAM/CoreSLAMProcessor.cs:234-235, 501-502
//   cs_normalize_angle  MathEx.NormalizeAngle  BaseSLAM/MathEx.cs:116-138
//   cs_philox4x32_10 + cs_gauss3   Redzen ZigguratGaussianSampler + Queue<float> pre-buffering
//                                              CoreSLAM/CoreSLAMProcessor.cs:136-137, 599-612, 633-638
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CS_HD __host__ __device__ __forceinline__
#else
#define CS_HD static inline
#endif

// ---------------------------------------------------------------------------------------------
// bit casts
// ---------------------------------------------------------------------------------------------
CS_HD uint32_t cs_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } v; v.f = f; return v.u;
#endif
}
CS_HD float cs_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } v; v.u = u; return v.f;
#endif
}
CS_HD uint64_t cs_d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  union { double d; uint64_t u; } v; v.d = d; return v.u;
#endif
}
CS_HD double cs_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { double d; uint64_t u; } v; v.u = u; return v.d;
#endif
}
CS_HD double cs_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
CS_HD double cs_sqrt(double a) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(a);
#else
  return __builtin_sqrt(a);
#endif
}

// ---------------------------------------------------------------------------------------------
// (int)f as RyuJIT compiles it on x64 (.NET 6): cvttss2si — truncate toward zero, and the
// "integer indefinite" 0x80000000 for NaN and for anything outside [-2^31, 2^31).
// CUDA's cvt.rzi saturates and maps NaN to 0, so the edge cases are patched here.
// ---------------------------------------------------------------------------------------------
CS_HD int32_t cs_cvt_i32(float f) {
#if defined(__CUDA_ARCH__)
  // fmaxf(NaN, a) == a on the device, so NaN joins the "too negative" family; 2^31 and above
  // are sent there explicitly.  Everything else is the plain truncating convert.
  int32_t v = __float2int_rz(f);
  return (f >= 2147483648.0f || !(f == f)) ? (int32_t)0x80000000 : v;
#else
  if (!(f == f) || f >= 2147483648.0f || f < -2147483648.0f) return (int32_t)0x80000000;
  return (int32_t)f;
#endif
}

// ---------------------------------------------------------------------------------------------
// cosf / sinf bit-identical to glibc 2.39 x86-64 (the libm .NET's MathF.Cos/Sin call on the
// B200 host).  The algorithm is the published double-precision polynomial scheme (ARM
// optimized-routines sinf/cosf, adopted by glibc 2.28+): quadrant reduction by hpi_inv*2^24,
// degree-7/8 minimax polynomials evaluated in double, one final rounding to float.  glibc picks
// its FMA build on any AVX2+FMA host; the fused operations are written out explicitly below
// so contraction settings cannot change the result.  tests/test_trig.py checks this against the
// host libm (2^32 inputs exhaustively on the CPU build, a dense sweep on the device build).
// ---------------------------------------------------------------------------------------------
struct cs_sincos_tab {
  double sign[4];
  double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};

CS_HD uint32_t cs_abstop12(float x) { return (cs_f2u(x) >> 20) & 0x7ffu; }

// n even -> sine polynomial, n odd -> cosine polynomial; neg selects the negated cosine set.
CS_HD float cs_sincos_poly(double x, double x2, int neg, int n) {
  const double c0 = neg ? -0x1p0 : 0x1p0;
  const double c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
  const double c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
  const double c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
  const double c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3;
  const double s2 = 0x1.1107605230bc4p-7;
  const double s3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double t1 = cs_fma(x2, s3, s2);
    double x7 = x3 * x2;
    double s = cs_fma(x3, s1, x);
    return (float)cs_fma(x7, t1, s);
  } else {
    double x4 = x2 * x2;
    double t2 = cs_fma(x2, c4, c3);
    double t1 = cs_fma(x2, c1, c0);
    double x6 = x4 * x2;
    double c = cs_fma(x4, c2, t1);
    return (float)cs_fma(x6, t2, c);
  }
}

CS_HD double cs_reduce_fast(double x, int* np) {
  const double hpi_inv = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
  const double hpi = 0x1.921FB54442D18p0;        // pi/2
  double r = x * hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return cs_fma(-(double)n, hpi, x);
}

// 4/pi to 192 bits, 8 new bits per entry
CS_HD uint32_t cs_inv_pio4(int i) {
  switch (i) {
    case 0: return 0xa2u; case 1: return 0xa2f9u; case 2: return 0xa2f983u; case 3: return 0xa2f9836eu;
    case 4: return 0xf9836e4eu; case 5: return 0x836e4e44u; case 6: return 0x6e4e4415u; case 7: return 0x4e441529u;
    case 8: return 0x441529fcu; case 9: return 0x1529fc27u; case 10: return 0x29fc2757u; case 11: return 0xfc2757d1u;
    case 12: return 0x2757d1f5u; case 13: return 0x57d1f534u; case 14: return 0xd1f534ddu; case 15: return 0xf534ddc0u;
    case 16: return 0x34ddc0dbu; case 17: return 0xddc0db62u; case 18: return 0xc0db6295u; case 19: return 0xdb629599u;
    case 20: return 0x6295993cu; case 21: return 0x95993c43u; case 22: return 0x993c4390u; default: return 0x3c439041u;
  }
}

CS_HD double cs_reduce_large(uint32_t xi, int* np) {
  int base = (int)((xi >> 26) & 15u);
  int shift = (int)((xi >> 23) & 7u);
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffffu) | 0x800000u;
  xi <<= shift;
  res0 = (uint64_t)(uint32_t)(xi * cs_inv_pio4(base));
  res1 = (uint64_t)xi * cs_inv_pio4(base + 4);
  res2 = (uint64_t)xi * cs_inv_pio4(base + 8);
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * 0x1.921FB54442D18p-62;
}

// which = 0 -> sinf, 1 -> cosf
CS_HD float cs_sincosf_one(float y, int which) {
  double x = (double)y;
  int n;
  if (cs_abstop12(y) < cs_abstop12(0x1.921FB6p-1f)) {
    if (cs_abstop12(y) < cs_abstop12(0x1p-12f)) return which ? 1.0f : y;
    return cs_sincos_poly(x, x * x, 0, which);
  } else if (cs_abstop12(y) < cs_abstop12(120.0f)) {
    x = cs_reduce_fast(x, &n);
    double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return cs_sincos_poly(x * s, x * x, (n & 2) ? 1 : 0, n ^ which);
  } else if (cs_abstop12(y) < 0x7f8u) {
    uint32_t xi = cs_f2u(y);
    int sign = (int)(xi >> 31);
    x = cs_reduce_large(xi, &n);
    int q = (n + sign) & 3;
    double s = (q == 1 || q == 2) ? -1.0 : 1.0;
    return cs_sincos_poly(x * s, x * x, (q & 2) ? 1 : 0, n ^ which);
  }
  return cs_u2f(0x7fc00000u);  // inf / NaN -> NaN (libm raises "invalid")
}
CS_HD float cs_sinf(float y) { return cs_sincosf_one(y, 0); }
CS_HD float cs_cosf(float y) { return cs_sincosf_one(y, 1); }

// ---------------------------------------------------------------------------------------------
// fmodf for finite x and y = 2*pi (exact: the remainder of two floats is representable).
// Generic shift-subtract on the integer significands, so it is exact for any finite operands.
// ---------------------------------------------------------------------------------------------
CS_HD float cs_fmodf(float x, float y) {
  uint32_t ux = cs_f2u(x), uy = cs_f2u(y);
  uint32_t sx = ux & 0x80000000u;
  uint32_t ax = ux & 0x7fffffffu, ay = uy & 0x7fffffffu;
  if (ay == 0 || ax >= 0x7f800000u || ay > 0x7f800000u) return cs_u2f(0x7fc00000u);
  if (ax < ay) return x;
  if (ax == ay) return cs_u2f(sx);
  int ex = (int)(ax >> 23), ey = (int)(ay >> 23);
  uint32_t mx, my;
  if (ex == 0) { mx = ax; ex = 1; while (!(mx & 0x800000u)) { mx <<= 1; ex--; } } else mx = (ax & 0x7fffffu) | 0x800000u;
  if (ey == 0) { my = ay; ey = 1; while (!(my & 0x800000u)) { my <<= 1; ey--; } } else my = (ay & 0x7fffffu) | 0x800000u;
  for (; ex > ey; ex--) {
    if (mx >= my) mx -= my;
    mx <<= 1;
  }
  if (mx >= my) mx -= my;
  if (mx == 0) return cs_u2f(sx);
  while (!(mx & 0x800000u)) { mx <<= 1; ex--; }
  uint32_t r;
  if (ex > 0) r = (mx & 0x7fffffu) | ((uint32_t)ex << 23);
  else r = mx >> (1 - ex);
  return cs_u2f(r | sx);
}

// MathEx.NormalizeAngle (BaseSLAM/MathEx.cs:116-138): pi2 = MathF.PI * 2.0f (float),
// a = ((angle % pi2) + pi2) % pi2; if (a > MathF.PI) a -= 2.0f * MathF.PI.
CS_HD float cs_normalize_angle(float angle) {
  const float pi = 3.14159274f;  // MathF.PI
  const float pi2 = pi * 2.0f;
  float a = cs_fmodf(cs_fmodf(angle, pi2) + pi2, pi2);
  if (a > pi) a -= 2.0f * pi;
  return a;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — counter-based, so candidate i of scan s is a pure
// function of (seed, s, i): no queues, no pre-buffering, no cross-thread sampler races.
// ---------------------------------------------------------------------------------------------
CS_HD uint32_t cs_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

CS_HD void cs_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t out[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = cs_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = cs_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// natural log of u in (0,1) (normal double), ~1e-13 relative: ln(m*2^e) = e ln2 + 2 atanh(z),
// z = (m-1)/(m+1) with m in [sqrt(1/2), sqrt(2)).  Basic operations only.
CS_HD double cs_log01(double u) {
  uint64_t b = cs_d2u(u);
  int e = (int)((b >> 52) & 0x7ffu) - 1023;
  double m = cs_u2d((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);  // [1,2)
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  double z = (m - 1.0) / (m + 1.0);
  double z2 = z * z;
  double p = 1.0 / 15.0;
  p = p * z2 + 1.0 / 13.0;
  p = p * z2 + 1.0 / 11.0;
  p = p * z2 + 1.0 / 9.0;
  p = p * z2 + 1.0 / 7.0;
  p = p * z2 + 1.0 / 5.0;
  p = p * z2 + 1.0 / 3.0;
  p = p * z2 + 1.0;
  return (double)e * 0.6931471805599453 + 2.0 * z * p;
}

// Box-Muller pair from two 32-bit words.
CS_HD void cs_box_muller(uint32_t w0, uint32_t w1, float* z0, float* z1) {
  double u1 = ((double)w0 + 0.5) * 0x1p-32;  // (0,1)
  double u2 = ((double)w1 + 0.5) * 0x1p-32;
  double r = cs_sqrt(-2.0 * cs_log01(u1));
  float a = (float)(6.283185307179586 * u2);
  *z0 = (float)(r * (double)cs_cosf(a));
  *z1 = (float)(r * (double)cs_sinf(a));
}

// The three N(0, sigma) deviates of Monte-Carlo candidate `index` (0-based among the T*I random
// candidates) of scan `scan`, in the reference's draw order X, Y, Theta
// (CoreSLAM/CoreSLAMProcessor.cs:633-638).
CS_HD void cs_gauss3(uint64_t seed, uint32_t scan, uint32_t index, float sigma_xy, float sigma_theta,
                     float out[3]) {
  uint32_t w[4];
  cs_philox4x32_10(index, scan, 0x434f5245u /*"CORE"*/, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  float zx, zy, zt, unused;
  cs_box_muller(w[0], w[1], &zx, &zy);
  cs_box_muller(w[2], w[3], &zt, &unused);
  out[0] = zx * sigma_xy;
  out[1] = zy * sigma_xy;
  out[2] = zt * sigma_theta;
}
