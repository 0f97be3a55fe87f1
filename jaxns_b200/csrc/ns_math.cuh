// Device-side arithmetic primitives: partitionable Threefry-2x32 (jax.random), uniform / normal
// bit recipes, XLA's f64 erf_inv polynomial, the Cephes/TFP ndtri rational, jnp.logaddexp.
// References: jax/_src/prng.py + random.py (third party, restated from the published algorithm),
// /root/reference/src/jaxns/internals/mixed_precision.py:11-15 (forces partitionable threefry + x64).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nsb {

struct Key {
    uint32_t a, b;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(x, x, r);
#else
    return (x << r) | (x >> (32 - r));
#endif
}

// Threefry-2x32, 20 rounds.
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t &x0, uint32_t &x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define NSB_TF_ROUND(r) \
    x0 += x1;           \
    x1 = rotl32(x1, r); \
    x1 ^= x0;
    x0 += k0;
    x1 += k1;
    NSB_TF_ROUND(13) NSB_TF_ROUND(15) NSB_TF_ROUND(26) NSB_TF_ROUND(6)
    x0 += k1;
    x1 += k2 + 1u;
    NSB_TF_ROUND(17) NSB_TF_ROUND(29) NSB_TF_ROUND(16) NSB_TF_ROUND(24)
    x0 += k2;
    x1 += k0 + 2u;
    NSB_TF_ROUND(13) NSB_TF_ROUND(15) NSB_TF_ROUND(26) NSB_TF_ROUND(6)
    x0 += k0;
    x1 += k1 + 3u;
    NSB_TF_ROUND(17) NSB_TF_ROUND(29) NSB_TF_ROUND(16) NSB_TF_ROUND(24)
    x0 += k1;
    x1 += k2 + 4u;
    NSB_TF_ROUND(13) NSB_TF_ROUND(15) NSB_TF_ROUND(26) NSB_TF_ROUND(6)
    x0 += k2;
    x1 += k0 + 5u;
#undef NSB_TF_ROUND
}

// One shared copy of the block function for device code: the sampler kernels call it from dozens of
// sites and their hot loops have to stay inside the instruction cache.
__device__ __noinline__ uint2 threefry_block(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    threefry2x32(k0, k1, x0, x1);
    return make_uint2(x0, x1);
}

// jax.random.split(key, n)[i]  (_threefry_split_foldlike): threefry(key; hi(i), lo(i)).
__host__ __device__ __forceinline__ Key split_child(Key k, uint64_t i) {
#ifdef __CUDA_ARCH__
    const uint2 r = threefry_block(k.a, k.b, (uint32_t) (i >> 32), (uint32_t) i);
    return Key{r.x, r.y};
#else
    uint32_t x0 = (uint32_t) (i >> 32), x1 = (uint32_t) i;
    threefry2x32(k.a, k.b, x0, x1);
    return Key{x0, x1};
#endif
}

// element i of a 64-bit random_bits draw (_threefry_random_bits_partitionable).
__host__ __device__ __forceinline__ uint64_t bits64(Key k, uint64_t i) {
#ifdef __CUDA_ARCH__
    const uint2 r = threefry_block(k.a, k.b, (uint32_t) (i >> 32), (uint32_t) i);
    return ((uint64_t) r.x << 32) | (uint64_t) r.y;
#else
    uint32_t x0 = (uint32_t) (i >> 32), x1 = (uint32_t) i;
    threefry2x32(k.a, k.b, x0, x1);
    return ((uint64_t) x0 << 32) | (uint64_t) x1;
#endif
}

// jax.random.uniform f64 (_uniform): mantissa fill, minus one, affine, max(lo, .).
__device__ __forceinline__ double bits_to_unit(uint64_t bits) {
    return __longlong_as_double((long long) ((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
}

__device__ __forceinline__ double uniform_lohi(uint64_t bits, double lo, double hi) {
    double v = bits_to_unit(bits) * (hi - lo) + lo;
    return fmax(lo, v);
}

__device__ __forceinline__ double uniform01(Key k, uint64_t i) {
    return fmax(0.0, bits_to_unit(bits64(k, i)));
}

// XLA ErfInv f64 (Giles): piecewise polynomial in w = -log1p(-x*x).  The central piece (w < 6.25,
// |x| < 0.99903) is inlined with its coefficients in constant memory; the two tail pieces are one
// shared out-of-line function.
__constant__ double kErfInvA[23] = {
    -3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18, 1.115787767802518096e-17,
    -1.333171662854620906e-16, 2.0972767875968561637e-17, 6.6376381343583238325e-15, -4.0545662729752068639e-14,
    -8.1519341976054721522e-14, 2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
    1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08, 4.2347877827932403518e-07,
    -1.3654692000834678645e-06, -1.3882523362786468719e-05, 0.0001867342080340571352, -0.00074070253416626697512,
    -0.0060336708714301490533, 0.24015818242558961693, 1.6536545626831027356};
__constant__ double kErfInvB[19] = {
    2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07, 1.8239629214389227755e-08,
    1.5027403968909827627e-06, -4.013867526981545969e-06, 2.9234449089955446044e-06, 1.2475304481671778723e-05,
    -4.7318229009055733981e-05, 6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
    0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066, -0.0037512085075692412107,
    0.005370914553590063617, 1.0052589676941592334, 3.0838856104922207635};
__constant__ double kErfInvC[17] = {
    -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09, -3.7894654401267369937e-09,
    7.6157012080783393804e-09, -1.4960026627149240478e-08, 2.9147953450901080826e-08, -6.7711997758452339498e-08,
    2.2900482228026654717e-07, -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
    7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026, 1.0103004648645343977,
    4.8499064014085844221};

__device__ __noinline__ double erfinv_tail(double x, double w) {
    double p;
    if (w < 16.0) {
        w = sqrt(w) - 3.25;
        p = kErfInvB[0];
#pragma unroll
        for (int i = 1; i < 19; ++i) p = kErfInvB[i] + p * w;
    } else {
        w = sqrt(w) - 5.0;
        p = kErfInvC[0];
#pragma unroll
        for (int i = 1; i < 17; ++i) p = kErfInvC[i] + p * w;
    }
    const double r = p * x;
    return (fabs(x) == 1.0) ? x * __longlong_as_double(0x7FF0000000000000ll) : r;
}

// log(x) for the quantile's tail (0 < x <= 0.075; any normal positive double is handled, everything else goes to the
// library log).  x = 2^e m, m in [1, 2); c_i = 1 + (2 i + 1) / 128 is the centre of m's 1/64-wide cell, so
// u = m / c_i - 1 has |u| <= 2^-7 and log1p(u) needs a degree-9 Taylor polynomial: log x = e ln 2 + log c_i + log1p(u).
// One table load + ~10 dependent FMAs (~115 cycles) where the library routine has ~290 cycles of dependent latency
// (profiles/r1/microbench_lat.txt); relative error <= 3e-16 for |log x| >= 2.5.  -DNSB_FAST_LOG=0 keeps the library log.
#ifndef NSB_FAST_LOG
#define NSB_FAST_LOG 1
#endif
__device__ const double2 kLogTab[64] = {  // {1 / c_i, log c_i}
    {0x1.fc07f01fc07f0p-1, 0x1.fe02a6b106789p-8}, {0x1.f44659e4a4271p-1, 0x1.7b91b07d5b11bp-6},
    {0x1.ecc07b301ecc0p-1, 0x1.39e87b9febd60p-5}, {0x1.e573ac901e574p-1, 0x1.b42dd711971bfp-5},
    {0x1.de5d6e3f8868ap-1, 0x1.16536eea37ae1p-4}, {0x1.d77b654b82c34p-1, 0x1.51b073f06183fp-4},
    {0x1.d0cb58f6ec074p-1, 0x1.8c345d6319b21p-4}, {0x1.ca4b3055ee191p-1, 0x1.c5e548f5bc743p-4},
    {0x1.c3f8f01c3f8f0p-1, 0x1.fec9131dbeabbp-4}, {0x1.bdd2b899406f7p-1, 0x1.1b72ad52f67a0p-3},
    {0x1.b7d6c3dda338bp-1, 0x1.371fc201e8f74p-3}, {0x1.b2036406c80d9p-1, 0x1.526e5e3a1b438p-3},
    {0x1.ac5701ac5701bp-1, 0x1.6d60fe719d21dp-3}, {0x1.a6d01a6d01a6dp-1, 0x1.87fa06520c911p-3},
    {0x1.a16d3f97a4b02p-1, 0x1.a23bc1fe2b563p-3}, {0x1.9c2d14ee4a102p-1, 0x1.bc286742d8cd6p-3},
    {0x1.970e4f80cb872p-1, 0x1.d5c216b4fbb91p-3}, {0x1.920fb49d0e229p-1, 0x1.ef0adcbdc5936p-3},
    {0x1.8d3018d3018d3p-1, 0x1.0402594b4d041p-2}, {0x1.886e5f0abb04ap-1, 0x1.1058bf9ae4ad5p-2},
    {0x1.83c977ab2beddp-1, 0x1.1c898c16999fbp-2}, {0x1.7f405fd017f40p-1, 0x1.2895a13de86a3p-2},
    {0x1.7ad2208e0ecc3p-1, 0x1.347dd9a987d55p-2}, {0x1.767dce434a9b1p-1, 0x1.404308686a7e4p-2},
    {0x1.724287f46debcp-1, 0x1.4be5f957778a1p-2}, {0x1.6e1f76b4337c7p-1, 0x1.5767717455a6cp-2},
    {0x1.6a13cd1537290p-1, 0x1.62c82f2b9c795p-2}, {0x1.661ec6a5122f9p-1, 0x1.6e08eaa2ba1e4p-2},
    {0x1.623fa77016240p-1, 0x1.792a55fdd47a2p-2}, {0x1.5e75bb8d015e7p-1, 0x1.842d1da1e8b17p-2},
    {0x1.5ac056b015ac0p-1, 0x1.8f11e873662c7p-2}, {0x1.571ed3c506b3ap-1, 0x1.99d958117e08bp-2},
    {0x1.5390948f40febp-1, 0x1.a484090e5bb0ap-2}, {0x1.5015015015015p-1, 0x1.af1293247786bp-2},
    {0x1.4cab88725af6ep-1, 0x1.b9858969310fbp-2}, {0x1.49539e3b2d067p-1, 0x1.c3dd7a7cdad4dp-2},
    {0x1.460cbc7f5cf9ap-1, 0x1.ce1af0b85f3ebp-2}, {0x1.42d6625d51f87p-1, 0x1.d83e7258a2f3ep-2},
    {0x1.3fb013fb013fbp-1, 0x1.e24881a7c6c26p-2}, {0x1.3c995a47babe7p-1, 0x1.ec399d2468cc0p-2},
    {0x1.3991c2c187f63p-1, 0x1.f6123fa7028acp-2}, {0x1.3698df3de0748p-1, 0x1.ffd2e0857f498p-2},
    {0x1.33ae45b57bcb2p-1, 0x1.04bdf9da926d2p-1}, {0x1.30d190130d190p-1, 0x1.0986f4f573521p-1},
    {0x1.2e025c04b8097p-1, 0x1.0e44985d1cc8cp-1}, {0x1.2b404ad012b40p-1, 0x1.12f719593efbcp-1},
    {0x1.288b01288b013p-1, 0x1.179eabbd899a1p-1}, {0x1.25e22708092f1p-1, 0x1.1c3b81f713c25p-1},
    {0x1.23456789abcdfp-1, 0x1.20cdcd192ab6ep-1}, {0x1.20b470c67c0d9p-1, 0x1.2555bce98f7cbp-1},
    {0x1.1e2ef3b3fb874p-1, 0x1.29d37fec2b08bp-1}, {0x1.1bb4a4046ed29p-1, 0x1.2e47436e40268p-1},
    {0x1.19453808ca29cp-1, 0x1.32b1339121d71p-1}, {0x1.16e0689427379p-1, 0x1.37117b54747b6p-1},
    {0x1.1485f0e0acd3bp-1, 0x1.3b68449fffc23p-1}, {0x1.12358e75d3033p-1, 0x1.3fb5b84d16f42p-1},
    {0x1.0fef010fef011p-1, 0x1.43f9fe2f9ce67p-1}, {0x1.0db20a88f4696p-1, 0x1.48353d1ea88dfp-1},
    {0x1.0b7e6ec259dc8p-1, 0x1.4c679afccee3ap-1}, {0x1.0953f39010954p-1, 0x1.50913cc01686bp-1},
    {0x1.073260a47f7c6p-1, 0x1.54b2467999498p-1}, {0x1.05197f7d73404p-1, 0x1.58cadb5cd7989p-1},
    {0x1.03091b51f5e1ap-1, 0x1.5cdb1dc6c1765p-1}, {0x1.0101010101010p-1, 0x1.60e32f44788d9p-1},
};

__device__ __forceinline__ double tail_log(double x) {
#if defined(NSB_EXACT_MATH) || !NSB_FAST_LOG
    return log(x);
#else
    const int hi = __double2hiint(x), lo = __double2loint(x);
    if ((unsigned) (hi - 0x00100000) >= 0x7fe00000u) return log(x);  // zero, denormal, negative, inf, NaN
    const int e = (hi >> 20) - 1023;
    const double2 tc = __ldg(&kLogTab[(hi >> 14) & 63]);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double u = fma(m, tc.x, -1.0);
    // log1p(u) = u + u^2 (-1/2 + u/3 - u^2/4 + u^3/5 - u^4/6 + u^5/7 - u^6/8 + u^7/9), Estrin
    const double u2 = u * u, u4 = u2 * u2;
    const double p01 = fma(u, 1.0 / 3.0, -0.5), p23 = fma(u, 0.2, -0.25), p45 = fma(u, 1.0 / 7.0, -1.0 / 6.0),
                 p67 = fma(u, 1.0 / 9.0, -0.125);
    const double q = fma(fma(p67, u2, p45), u4, fma(p23, u2, p01));
    const double r = fma(u2, q, u);
    return fma((double) e, 0.6931471805599453094, tc.y + r);
#endif
}

// The same log inside erfinv (one per normal deviate of the chain-stream generator).  -DNSB_FAST_LOG_GEN=1.
#ifndef NSB_FAST_LOG_GEN
#define NSB_FAST_LOG_GEN 1
#endif

__device__ __forceinline__ double erfinv_xla(double x) {
    // w = -log1p(-x*x).  Evaluated as -log(1 - x*x): w only enters additively (w - 3.125, sqrt(w) - c), so the
    // absolute error of the plain log (~1e-16) is what matters and libdevice's log is ~4x cheaper than log1p.
#if NSB_FAST_LOG_GEN
    const double w0 = -tail_log(fma(x, -x, 1.0));
#else
    const double w0 = -log(fma(x, -x, 1.0));
#endif
    if (w0 < 6.25) {
        const double w = w0 - 3.125;
        double p = kErfInvA[0];
#pragma unroll
        for (int i = 1; i < 23; ++i) p = kErfInvA[i] + p * w;
        return p * x;
    }
    return erfinv_tail(x, w0);
}

// jax.random.normal f64 (_normal_real) from one 64-bit draw.
__device__ __forceinline__ double normal_from_bits(uint64_t bits) {
    const double lo = -0.99999999999999988897769753748;  // nextafter(-1, 0)
    double u = uniform_lohi(bits, lo, 1.0);
    return 1.4142135623730951 * erfinv_xla(u);
}

// Newton steps on the MUFU reciprocal seed before the quotient and its residual correction (experiments: 1 or 2).
// One step suffices: the seed is good to ~2^-20, one step gives 2^-40, and the residual correction of the quotient
// squares that again.  Measured (profiles/r2/regcap_ab_r2.txt): quantiles bit-identical to the two-step version over
// 2.1e6 points from 1e-300 to 1 - 1e-16, a config-2 run 95.2 -> 94.1 ms.
#ifndef NSB_DIV_NEWTON
#define NSB_DIV_NEWTON 1
#endif

// The IEEE division of fast_div's fallback as a real call: inlined, the compiler evaluates the ~25-instruction
// division sequence unconditionally and selects afterwards; out of line it costs one never-taken branch.
__device__ __noinline__ double slow_div(double a, double b) { return a / b; }

// a / b for the chains' critical path: MUFU reciprocal seed + two Newton steps + one residual
// correction (8 instructions, ~75 cycles) instead of the IEEE division sequence (20 instructions with
// its slow-path check, ~125 cycles).  Result within 1 ulp of a / b; non-finite intermediates (b = 0,
// inf, denormal) fall back to the IEEE division.
__device__ __forceinline__ double fast_div(double a, double b) {
#ifdef NSB_EXACT_MATH
    return a / b;  // parity build: IEEE division (see cephes_ndtri below)
#endif
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    if (NSB_DIV_NEWTON >= 2) r = fma(fma(-b, r, 1.0), r, r);
    double q = a * r;
    q = fma(fma(-b, q, a), r, q);
    return (q - q == 0.0) ? q : slow_div(a, b);  // q - q != 0 for NaN / inf
}

// Same without the fallback, for denominators known to be finite, normal and non-zero (the AS241
// denominators are polynomials bounded away from zero on their intervals).
__device__ __forceinline__ double fast_div_finite(double a, double b) {
#ifdef NSB_EXACT_MATH
    return a / b;
#endif
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    if (NSB_DIV_NEWTON >= 2) r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// Normal quantile for the prior transform (tfd.Normal.quantile -> tfp special_math.ndtri).
// The reference evaluates Cephes' ndtri: central rational + a tail that needs two dependent logs, a
// sqrt and two divisions (~1000 cycles of latency on B200, and with 32 lanes per chain some lane is
// almost always in the tail).  This is 49 % of a chain's critical path, so the device code uses
// Wichura's AS241 (PPND16) instead: same double-precision accuracy (it agrees with Cephes/scipy to
// <= 1.1e-15 relative over (1e-300, 1 - 1e-16), tests/test_gpu_parity.py), but its tail is one log, one
// sqrt and one division, and its central region is wider (|p - 0.5| <= 0.425).  Central part is
// branch-free; the tail runs only if some lane of the group needs it (warp-uniform vote), so its log
// overlaps the central division instead of following it.
__constant__ double kPpndA[8] = {3.3871328727963666080, 1.3314166789178437745E+2, 1.9715909503065514427E+3,
                                 1.3731693765509461125E+4, 4.5921953931549871457E+4, 6.7265770927008700853E+4,
                                 3.3430575583588128105E+4, 2.5090809287301226727E+3};
__constant__ double kPpndB[8] = {1.0, 4.2313330701600911252E+1, 6.8718700749205790830E+2, 5.3941960214247511077E+3,
                                 2.1213794301586595867E+4, 3.9307895800092710610E+4, 2.8729085735721942674E+4,
                                 5.2264952788528545610E+3};
__constant__ double kPpndC[8] = {1.42343711074968357734, 4.63033784615654529590, 5.76949722146069140550,
                                 3.64784832476320460504, 1.27045825245236838258, 2.41780725177450611770E-1,
                                 2.27238449892691845833E-2, 7.74545014278341407640E-4};
__constant__ double kPpndD[8] = {1.0, 2.05319162663775882187, 1.67638483018380384940, 6.89767334985100004550E-1,
                                 1.48103976427480074590E-1, 1.51986665636164571966E-2, 5.47593808499534494600E-4,
                                 1.05075007164441684324E-9};
__constant__ double kPpndE[8] = {6.65790464350110377720, 5.46378491116411436990, 1.78482653991729133580,
                                 2.96560571828504891230E-1, 2.65321895265761230930E-2, 1.24266094738807843860E-3,
                                 2.71155556874348757815E-5, 2.01033439929228813265E-7};
__constant__ double kPpndF[8] = {1.0, 5.99832206555887937690E-1, 1.36929880922735805310E-1, 1.48753612908506148525E-2,
                                 7.86869131145613259100E-4, 1.84631831751005468180E-5, 1.42151175831644588870E-7,
                                 2.04426310338993978564E-15};

// Degree-7 polynomial.  -DNSB_ESTRIN=1: Estrin's scheme (dependent depth 4 FMAs instead of 7, two extra multiplies).
#ifndef NSB_ESTRIN
#define NSB_ESTRIN 0
#endif
__device__ __forceinline__ double horner8(const double *c, double x) {
#if NSB_ESTRIN
    const double x2 = x * x, x4 = x2 * x2;
    const double p01 = fma(c[1], x, c[0]), p23 = fma(c[3], x, c[2]), p45 = fma(c[5], x, c[4]), p67 = fma(c[7], x, c[6]);
    const double q0 = fma(p23, x2, p01), q1 = fma(p67, x2, p45);
    return fma(q1, x4, q0);
#else
    double r = c[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) r = fma(r, x, c[i]);
    return r;
#endif
}

#ifdef NSB_EXACT_MATH
// -DNSB_EXACT_MATH (parity build, tests/test_gpu_exact_math.py): the reference's own quantile -- Cephes ndtri as
// tfp.math.ndtri evaluates it (central rational in (p - 1/2)^2 for exp(-2) < p < 1 - exp(-2), two tail rationals in
// 1 / sqrt(-2 log p)) -- and IEEE divisions everywhere, instead of AS241 and the reciprocal-Newton division.  What
// still differs from a CPU evaluation of the same formulas: libdevice's log / exp / log1p (<= 1 ulp, like glibc's,
// but not the same bits), FMA contraction and the order of the 32-lane sums.
__constant__ double kCephesP0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                                    1.39312609387279679503E1, -1.23916583867381258016E0};
__constant__ double kCephesQ0[9] = {1.0, 1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
                                    -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                                    1.59056225126211695515E1, -1.18331621121330003142E0};
__constant__ double kCephesP1[9] = {4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
                                    4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
                                    -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
__constant__ double kCephesQ1[9] = {1.0, 1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
                                    1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
                                    -3.80806407691578277194E-2, -9.33259480895457427372E-4};
__constant__ double kCephesP2[9] = {3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
                                    1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
                                    3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
__constant__ double kCephesQ2[9] = {1.0, 6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
                                    2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                                    2.89247864745380683936E-6, 6.79019408009981274425E-9};

__device__ __noinline__ double cephes_ndtri(double p) {
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    if (p == 0.0) return -kInf;
    if (p == 1.0) return kInf;
    if (!(p > 0.0 && p < 1.0)) return __longlong_as_double(0x7FF8000000000000ll);
    const bool upper = p > 0.8646647167633873;  // 1 - exp(-2)
    const double q = upper ? 1.0 - p : p;
    double x;
    if (q > 0.1353352832366127) {  // exp(-2)
        const double w = q - 0.5, ww = w * w;
        double num = kCephesP0[0], den = kCephesQ0[0];
        for (int i = 1; i < 5; ++i) num = num * ww + kCephesP0[i];
        for (int i = 1; i < 9; ++i) den = den * ww + kCephesQ0[i];
        x = w + w * ww * (num / den);
        x *= -2.5066282746310002;  // -sqrt(2 pi)
    } else {
        const double z = sqrt(-2.0 * log(q));
        const double first = z - log(z) / z;
        const double rz = 1.0 / z;
        const double *P = (z >= 8.0) ? kCephesP2 : kCephesP1;
        const double *Q = (z >= 8.0) ? kCephesQ2 : kCephesQ1;
        double num = P[0], den = Q[0];
        for (int i = 1; i < 9; ++i) num = num * rz + P[i];
        for (int i = 1; i < 9; ++i) den = den * rz + Q[i];
        x = first - num / den / z;
    }
    return upper ? x : -x;
}
#endif

// `mask` = lanes that execute this call together (the chain's lane group).
__device__ __forceinline__ double ndtri(double p, unsigned mask) {
#ifdef NSB_EXACT_MATH
    (void) mask;
    return cephes_ndtri(p);
#endif
    const double q = p - 0.5;
    const double r = fma(-q, q, 0.180625);
    double x = fast_div_finite(q * horner8(kPpndA, r), horner8(kPpndB, r));
    const bool tail = !(fabs(q) <= 0.425);  // also true for NaN
    if (__any_sync(mask, tail)) {
        const double kInf = __longlong_as_double(0x7FF0000000000000ll);
        double pp = (q < 0.0) ? p : 1.0 - p;
        pp = tail ? pp : 0.05;  // keep the non-tail lanes on the fast paths of log / sqrt
        const double rr = sqrt(-tail_log(pp));
        const double a = rr - 1.6, b = rr - 5.0;
        double v = fast_div_finite(horner8(kPpndC, a), horner8(kPpndD, a));
        if (__any_sync(mask, rr > 5.0)) {  // far tail (p < 1.4e-11): rare, warp-uniform
            const double v2 = fast_div_finite(horner8(kPpndE, b), horner8(kPpndF, b));
            v = (rr <= 5.0) ? v : v2;
        }
        v = (q < 0.0) ? -v : v;
        if (p == 0.0) v = -kInf;
        if (p == 1.0) v = kInf;
        if (!(p >= 0.0 && p <= 1.0)) v = __longlong_as_double(0x7FF8000000000000ll);
        x = tail ? v : x;
    }
    return x;
}

// P quantiles at once: the P central rationals are straight-line code (their dependent chains interleave), and
// ONE warp vote decides whether the near-tail piece runs -- for the whole batch, again as straight-line code --
// instead of P votes that serialise the proposals.  (Late in a run every lane of a chain sits in the central
// region, so the tail must stay skippable.)  Same operations per element as ndtri(): bit-identical results.
template <int P>
__device__ __forceinline__ void ndtri_batch(const double (&p)[P], unsigned mask, double (&x)[P]) {
#ifdef NSB_EXACT_MATH
    (void) mask;
#pragma unroll
    for (int i = 0; i < P; ++i) x[i] = cephes_ndtri(p[i]);
    return;
#endif
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    double q[P];
    bool tail[P], any_tail = false;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        q[i] = p[i] - 0.5;
        tail[i] = !(fabs(q[i]) <= 0.425);
        any_tail |= tail[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const double r = fma(-q[i], q[i], 0.180625);
        x[i] = fast_div_finite(q[i] * horner8(kPpndA, r), horner8(kPpndB, r));
    }
    if (__any_sync(mask, any_tail)) {
        double rr[P], v[P];
        bool far = false;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double pp = (q[i] < 0.0) ? p[i] : 1.0 - p[i];
            pp = tail[i] ? pp : 0.05;  // keep the non-tail lanes on the fast paths of log / sqrt
            rr[i] = sqrt(-tail_log(pp));
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const double a = rr[i] - 1.6;
            v[i] = fast_div_finite(horner8(kPpndC, a), horner8(kPpndD, a));
            far |= rr[i] > 5.0;
        }
        if (__any_sync(mask, far)) {  // far tail (p < 1.4e-11): rare, warp-uniform
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const double b = rr[i] - 5.0;
                const double v2 = fast_div_finite(horner8(kPpndE, b), horner8(kPpndF, b));
                v[i] = (rr[i] <= 5.0) ? v[i] : v2;
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double w = (q[i] < 0.0) ? -v[i] : v[i];
            if (p[i] == 0.0) w = -kInf;
            if (p[i] == 1.0) w = kInf;
            if (!(p[i] >= 0.0 && p[i] <= 1.0)) w = __longlong_as_double(0x7FF8000000000000ll);
            x[i] = tail[i] ? w : x[i];
        }
    }
}

// K quantiles per lane for warps whose 32 lanes are convergent (the round-synchronous DMMA slice kernel keeps 8
// dimensions of one proposal in every lane).  The K central rationals are straight-line code; tail values are
// sparse (late in a run a few of the warp's 32 K values per round), so they are drained through a vote loop
// that evaluates ONE pending value per lane per trip instead of a K-wide tail block.  Same operations per element
// as ndtri(): bit-identical results.
#ifdef NSB_PROFILE
__device__ unsigned long long g_tail_trips;
#endif
template <int K>
__device__ __forceinline__ void ndtri_multi(const double (&p)[K], double (&x)[K]) {
#ifdef NSB_EXACT_MATH
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = cephes_ndtri(p[i]);
    return;
#endif
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    unsigned pend = 0;
    {
        // coefficient-major order: the 2 K Horner chains and the K divisions advance together, so a lone warp on
        // its SM sub-partition issues an independent FP64 instruction every slot instead of waiting 8 cycles for
        // its own last result (value-major source order was scheduled as K back-to-back dependent chains)
        double q[K], r[K], na[K], nb[K], rc[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            q[i] = p[i] - 0.5;
            r[i] = fma(-q[i], q[i], 0.180625);
            na[i] = kPpndA[7];
            nb[i] = kPpndB[7];
            pend |= (!(fabs(q[i]) <= 0.425)) ? (1u << i) : 0u;  // also true for NaN
        }
#pragma unroll
        for (int c = 6; c >= 0; --c) {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                na[i] = fma(na[i], r[i], kPpndA[c]);
                nb[i] = fma(nb[i], r[i], kPpndB[c]);
            }
        }
        // fast_div_finite(q * A(r), B(r)), the K quotients interleaved
#pragma unroll
        for (int i = 0; i < K; ++i) {
            na[i] = q[i] * na[i];
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc[i]) : "d"(nb[i]));
        }
#pragma unroll
        for (int i = 0; i < K; ++i) rc[i] = fma(fma(-nb[i], rc[i], 1.0), rc[i], rc[i]);
#pragma unroll
        for (int i = 0; i < K; ++i) rc[i] = fma(fma(-nb[i], rc[i], 1.0), rc[i], rc[i]);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const double qq = na[i] * rc[i];
            x[i] = fma(fma(-nb[i], qq, na[i]), rc[i], qq);
        }
    }
    while (__any_sync(0xFFFFFFFFu, pend != 0)) {
#ifdef NSB_PROFILE
        if (threadIdx.x == 0 && blockIdx.x == 0) g_tail_trips += 1;
#endif
        const int i0 = __ffs((int) pend) - 1;  // -1: nothing pending in this lane
        double pv = 0.5;
#pragma unroll
        for (int i = 0; i < K; ++i) pv = (i == i0) ? p[i] : pv;
        const bool tail = pend != 0;
        const double q = pv - 0.5;
        double pp = (q < 0.0) ? pv : 1.0 - pv;
        pp = tail ? pp : 0.05;  // idle lanes stay on the fast paths of log / sqrt
        const double rr = sqrt(-tail_log(pp));
        const double a = rr - 1.6, b = rr - 5.0;
        double v = fast_div_finite(horner8(kPpndC, a), horner8(kPpndD, a));
        if (__any_sync(0xFFFFFFFFu, rr > 5.0)) {  // far tail (p < 1.4e-11): rare, warp-uniform
            const double v2 = fast_div_finite(horner8(kPpndE, b), horner8(kPpndF, b));
            v = (rr <= 5.0) ? v : v2;
        }
        v = (q < 0.0) ? -v : v;
        if (pv == 0.0) v = -kInf;
        if (pv == 1.0) v = kInf;
        if (!(pv >= 0.0 && pv <= 1.0)) v = __longlong_as_double(0x7FF8000000000000ll);
#pragma unroll
        for (int i = 0; i < K; ++i) x[i] = (i == i0) ? v : x[i];
        pend &= pend - 1u;
    }
}

// log1p(e) for e in [0, 1] with one log and one division (Kahan): libdevice's log1p costs ~300
// instructions, this ~70, and it is accurate to an ulp or two on that range.
__device__ __forceinline__ double log1p_unit(double e) {
    const double u = 1.0 + e;
    const double d = u - 1.0;
    return (d == 0.0) ? e : log(u) * (e / d);
}

// jnp.logaddexp: amax + log1p(exp(-|delta|)), NaN delta (inf - inf) -> a + b
__device__ __noinline__ double logaddexp(double a, double b) {
    double amax = fmax(a, b);
    double delta = a - b;
    if (delta != delta) return a + b;  // inf - inf (same sign) or NaN input
    return amax + log1p_unit(exp(-fabs(delta)));
}

// Order-preserving map f64 -> u64 for sorting with jnp.argsort semantics: -0 == +0, NaN last.
__host__ __device__ __forceinline__ uint64_t sort_key_f64(double x) {
    if (x != x) return 0xFFFFFFFFFFFFFFFFull;  // all NaNs equal, after +inf
    if (x == 0.0) x = 0.0;                     // canonical zero
    uint64_t u;
#ifdef __CUDA_ARCH__
    u = (uint64_t) __double_as_longlong(x);
#else
    __builtin_memcpy(&u, &x, 8);
#endif
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

}  // namespace nsb
