// hem_math.cuh — device arithmetic whose RESULT BITS are part of the contract.
//
// The gates of the reference HEM step (radius test, colour gate, KLD gate,
// parent flags, validity filter) decide cluster membership, so they must give
// the same answer as the reference's x86 build on the same inputs.  Everything
// here therefore spells out the reference's operation order with the
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/... are never contracted
// into FMAs, whatever -fmad says) and cites the reference line it follows
// (paths relative to the reference's src/cpp_ext/).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hem {

// --------------------------------------------------------------------------
// exact FP32 helpers
// --------------------------------------------------------------------------
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }

struct Sym3 { float e00, e01, e02, e11, e12, e22; };

// det(smat3), include/vec.hpp:863-866:
//   -e02*e02*e11 + 2*e01*e02*e12 - e00*e12*e12 - e01*e01*e22 + e00*e11*e22
__device__ __forceinline__ float det_sym3(const Sym3& c)
{
    float t0 = fm(fm(-c.e02, c.e02), c.e11);
    float t1 = fm(fm(fm(2.0f, c.e01), c.e02), c.e12);
    float t2 = fm(fm(c.e00, c.e12), c.e12);
    float t3 = fm(fm(c.e01, c.e01), c.e22);
    float t4 = fm(fm(c.e00, c.e11), c.e22);
    return fa(fs(fs(fa(t0, t1), t2), t3), t4);
}

// inverse(smat3), include/vec.hpp:868-872 with operator/ (:510-514):
// adjugate elements times (1.0f / det).
__device__ __forceinline__ Sym3 inverse_sym3(const Sym3& c, float det)
{
    Sym3 r;
    r.e00 = fs(fm(c.e11, c.e22), fm(c.e12, c.e12));
    r.e01 = fs(fm(c.e02, c.e12), fm(c.e01, c.e22));
    r.e02 = fs(fm(c.e01, c.e12), fm(c.e02, c.e11));
    r.e11 = fs(fm(c.e00, c.e22), fm(c.e02, c.e02));
    r.e12 = fs(fm(c.e02, c.e01), fm(c.e00, c.e12));
    r.e22 = fs(fm(c.e00, c.e11), fm(c.e01, c.e01));
    float inv_s = fd(1.0f, det);
    r.e00 = fm(r.e00, inv_s); r.e01 = fm(r.e01, inv_s); r.e02 = fm(r.e02, inv_s);
    r.e11 = fm(r.e11, inv_s); r.e12 = fm(r.e12, inv_s); r.e22 = fm(r.e22, inv_s);
    return r;
}

// squaredDist, include/vec.hpp:158-166 (dx*dx + dy*dy + dz*dz, left to right)
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = fs(ax, bx), dy = fs(ay, by), dz = fs(az, bz);
    return fa(fa(fm(dx, dx), fm(dy, dy)), fm(dz, dz));
}

// --------------------------------------------------------------------------
// logf with glibc 2.27+ result bits (the reference's KLD calls std::log(float),
// include/gaussian.hpp:108, which binds to glibc logf on the oracle host).
// Algorithm: Szabolcs Nagy's logf from ARM optimized-routines as shipped in
// glibc (sysdeps/ieee754/flt-32/e_logf.c, LOGF_TABLE_BITS = 4, degree-3
// polynomial, double internals).  The restatement was checked on the build host
// against glibc logf for ALL 2,139,095,039 positive finite floats, with and
// without FMA contraction: 0 mismatches (DESIGN.md "logf").
// --------------------------------------------------------------------------
__device__ __constant__ double c_logf_tab[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2 }, { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1,  0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },
};

__device__ __forceinline__ float logf_glibc(float x)
{
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0u) return __uint_as_float(0xff800000u);          // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                                 // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return __uint_as_float(0x7fc00000u);
        ix = __float_as_uint(fm(x, 8388608.0f));                         // subnormal: scale by 2^23
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = c_logf_tab[i][0], logc = c_logf_tab[i][1];
    double z = (double)__uint_as_float(iz);
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, 0x1.62e42fefa39efp-1, logc);
    double r2 = __dmul_rn(r, r);
    double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = fma(-0x1.00ea348b88334p-2, r2, y);
    y = fma(y, r2, __dadd_rn(y0, r));
    return (float)y;
}

// --------------------------------------------------------------------------
// largest eigenvalue, smat3::eigenvalues, include/vec.hpp:736-768: c0,c1,c2 in
// FP32, trigonometric closed form in FP64, result cast to FP32.  CUDA's double
// atan2/sin/cos are not bit-identical to glibc's; a last-bit difference of the
// double survives the cast to float with probability ~2^-29.
// --------------------------------------------------------------------------
__device__ __forceinline__ float eig_max_sym3(const Sym3& m, double* emin_out = nullptr, double* emax_out = nullptr)
{
    const double inv3 = 0.33333333333333333333333333333333;
    const double root3 = 1.7320508075688772935274463415059;
    float c0f = fs(fs(fs(fa(fm(fm(m.e00, m.e11), m.e22), fm(fm(fm(2.0f, m.e01), m.e02), m.e12)),
                         fm(fm(m.e00, m.e12), m.e12)), fm(fm(m.e11, m.e02), m.e02)), fm(fm(m.e22, m.e01), m.e01));
    float c1f = fs(fa(fs(fa(fs(fm(m.e00, m.e11), fm(m.e01, m.e01)), fm(m.e00, m.e22)), fm(m.e02, m.e02)),
                      fm(m.e11, m.e22)), fm(m.e12, m.e12));
    float c2f = fa(fa(m.e00, m.e11), m.e22);
    double c0 = c0f, c1 = c1f, c2 = c2f;
    double c2Div3 = __dmul_rn(c2, inv3);
    double aDiv3 = __dsub_rn(__dmul_rn(c1, inv3), __dmul_rn(c2Div3, c2Div3));
    if (aDiv3 > 0.0) aDiv3 = 0.0;
    double mbDiv2 = __dsub_rn(__dadd_rn(__dmul_rn(0.5, c0), __dmul_rn(__dmul_rn(c2Div3, c2Div3), c2Div3)),
                              __dmul_rn(__dmul_rn(0.5, c2Div3), c1));
    double q = __dadd_rn(__dmul_rn(mbDiv2, mbDiv2), __dmul_rn(__dmul_rn(aDiv3, aDiv3), aDiv3));
    if (q > 0.0) q = 0.0;
    double magnitude = sqrt(-aDiv3);
    double angle = __dmul_rn(atan2(sqrt(-q), mbDiv2), inv3);
    if (angle != angle) angle = 0.0;
    double sn, cs;
    sincos(angle, &sn, &cs);
    double e0 = __dadd_rn(c2Div3, __dmul_rn(__dmul_rn(2.0, magnitude), cs));
    double e1 = __dsub_rn(c2Div3, __dmul_rn(magnitude, __dadd_rn(cs, __dmul_rn(root3, sn))));
    double e2 = __dsub_rn(c2Div3, __dmul_rn(magnitude, __dsub_rn(cs, __dmul_rn(root3, sn))));
    double h;
    if (e2 < e1) { h = e1; e1 = e2; e2 = h; }
    if (e1 < e0) { h = e0; e0 = e1; e1 = h; }
    if (e2 < e1) { h = e1; e1 = e2; e2 = h; }
    if (emin_out) *emin_out = e0;
    if (emax_out) *emax_out = e2;
    return (float)e2;
}

// --------------------------------------------------------------------------
// Certificates for the Mahalanobis pre-test of the KLD gate (DESIGN.md "smd pre-test").
//
// For symmetric positive definite P (parent) and C (child), exact arithmetic gives
//     tr(P^-1 C) - 3 - ln(det C / det P) >= 0            (AM-GM on the eigenvalues of P^-1 C),
// so KLD = 0.5 (smd + tr - 3 - ln) > kthr as soon as smd > 2 kthr.  The reference evaluates the
// formula in FP32; the pre-test may only replace that evaluation where the FP32 values provably
// stay close to the exact ones.  Both certificates are MEASURED per component in FP64 (no
// worst-case bounds):
//   child  : C finite and positive definite (Sylvester, FP64), |det_f - det| <= 1e-3 det
//   parent : P positive definite, |det_f - det| <= 1e-3 det, and with ip = the FP32 inverse the
//            gate uses, E = ip - P^-1:  eta = ||E||_F * lambda_max(P) + 2e-6 * kappa_ub(P) <= 0.01,
//            kappa_ub = lambda_max^3 / det >= kappa (lambda_max itself is accurate: it defines the radius)
// Then |tr(E C)| <= ||E||_2 tr(C) <= eta tr(P^-1 C), the FP32 dot products add <= 2e-6 kappa
// relative, |ln_f - ln| <= 2.1e-3, hence tr_f - 3 - ln_f >= 3 ln(1 - eta) - 2.1e-3 >= -0.033 and
// the FP32 KLD exceeds kthr whenever smd_f > 2 kthr + 0.5 (kthr <= 1e4; the margin also covers the
// three FP32 additions of the formula).
// --------------------------------------------------------------------------
__device__ __forceinline__ double det_sym3_d(const Sym3& c)
{
    const double e00 = c.e00, e01 = c.e01, e02 = c.e02, e11 = c.e11, e12 = c.e12, e22 = c.e22;
    return e00 * (e11 * e22 - e12 * e12) - e01 * (e01 * e22 - e12 * e02) + e02 * (e01 * e12 - e11 * e02);
}

__device__ __forceinline__ bool child_cov_certified(const Sym3& c, float det_f)
{
    const double e00 = c.e00, e01 = c.e01, e11 = c.e11;
    const double m2 = e00 * e11 - e01 * e01;
    const double dt = det_sym3_d(c);
    // NaN / inf anywhere makes one of the comparisons false
    const bool fin = (fabs(dt) < 1.7e308) && (fabsf(c.e22) < 3.0e38f) && (fabsf(c.e02) < 3.0e38f) && (fabsf(c.e12) < 3.0e38f);
    return fin && e00 > 0.0 && m2 > 0.0 && dt > 0.0 && fabs((double)det_f - dt) <= 1e-3 * dt;
}

__device__ __forceinline__ bool parent_cov_certified(const Sym3& c, float det_f, double emin, double emax)
{
    (void)emin;     // the closed form's smallest root is unreliable for ill-conditioned input: not used
    if (!(emax > 0.0) || !(emax < 1.7e308) || !child_cov_certified(c, det_f)) return false;
    const double dt = det_sym3_d(c);
    emax *= 1.001;
    const Sym3 ip = inverse_sym3(c, det_f);           // exactly what k_gate computes
    const double e00 = c.e00, e01 = c.e01, e02 = c.e02, e11 = c.e11, e12 = c.e12, e22 = c.e22;
    const double id = 1.0 / dt;
    const double d00 = (e11 * e22 - e12 * e12) * id - ip.e00, d01 = (e02 * e12 - e01 * e22) * id - ip.e01;
    const double d02 = (e01 * e12 - e02 * e11) * id - ip.e02, d11 = (e00 * e22 - e02 * e02) * id - ip.e11;
    const double d12 = (e02 * e01 - e00 * e12) * id - ip.e12, d22 = (e00 * e11 - e01 * e01) * id - ip.e22;
    const double fro = sqrt(d00 * d00 + d11 * d11 + d22 * d22 + 2.0 * (d01 * d01 + d02 * d02 + d12 * d12));
    const double eta = fro * emax + 2e-6 * (emax * emax * emax / dt);
    return eta <= 0.01;
}

// --------------------------------------------------------------------------
// Counter-based parent flags.  The reference draws rand01() from unseeded libc
// rand (include/base.hpp:44-56) in component order; both this path and the
// oracle replace the source by Philox4x32-10(key = seed, counter = k) keyed by
// the running pre-filter component index k, keeping the reference's
// conversion float(u32) / 2^32 < 1.0f / hemReduction (mixture.cpp:257-259,330).
// --------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t k)
{
    uint32_t c0 = (uint32_t)k, c1 = (uint32_t)(k >> 32), c2 = 0, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

__device__ __forceinline__ bool parent_flag(uint64_t seed, uint64_t k, float parent_probability)
{
    return fd(__uint2float_rn(philox_u32(seed, k)), 4294967296.0f) < parent_probability;
}

// clamp, include/base.hpp:24-27 (select-based fminf/fmaxf: NaN/inf -> hi, <= lo -> lo)
__device__ __forceinline__ float clamp_ref(float f, float lo, float hi)
{
    float m = f < hi ? f : hi;
    return lo > m ? lo : m;
}

} // namespace hem
