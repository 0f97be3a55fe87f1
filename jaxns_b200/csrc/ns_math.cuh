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

__device__ __forceinline__ double erfinv_xla(double x) {
    // w = -log1p(-x*x).  Evaluated as -log(1 - x*x): w only enters additively (w - 3.125, sqrt(w) - c), so the
    // absolute error of the plain log (~1e-16) is what matters and libdevice's log is ~4x cheaper than log1p.
    const double w0 = -log(fma(x, -x, 1.0));
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

// tfp special_math.ndtri == Cephes ndtri (Normal.quantile in the prior transform).  Central piece
// (exp(-2) < p < 1 - exp(-2)) inline and branch-free so batched proposals interleave; the tails and
// the p in {0, 1, NaN} edge cases are one shared out-of-line function.
__constant__ double kNdtriP0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                                   1.39312609387279679503E1, -1.23916583867381258016E0};
__constant__ double kNdtriQ0[8] = {1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
                                   -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                                   1.59056225126211695515E1, -1.18331621121330003142E0};
__constant__ double kNdtriP1[9] = {4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
                                   4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
                                   -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
__constant__ double kNdtriQ1[8] = {1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
                                   1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
                                   -3.80806407691578277194E-2, -9.33259480895457427372E-4};
__constant__ double kNdtriP2[9] = {3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
                                   1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
                                   3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
__constant__ double kNdtriQ2[8] = {6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
                                   2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                                   2.89247864745380683936E-6, 6.79019408009981274425E-9};

__device__ __noinline__ double ndtri_tail(double p) {
    const double kInf = __longlong_as_double(0x7FF0000000000000ll);
    if (p == 0.0) return -kInf;
    if (p == 1.0) return kInf;
    if (!(p > 0.0 && p < 1.0)) return __longlong_as_double(0x7FF8000000000000ll);
    const bool upper = p > 0.8646647167633873;  // -expm1(-2)
    const double q = upper ? 1.0 - p : p;
    const double z = sqrt(-2.0 * log(q));
    const double rz = 1.0 / z;
    const double first = z - log(z) * rz;
    double num, den = 1.0;
    if (z >= 8.0) {
        num = kNdtriP2[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) num = num * rz + kNdtriP2[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) den = den * rz + kNdtriQ2[i];
    } else {
        num = kNdtriP1[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) num = num * rz + kNdtriP1[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) den = den * rz + kNdtriQ1[i];
    }
    const double x = first - (num / den) * rz;
    return upper ? x : -x;
}

__device__ __forceinline__ double ndtri(double p) {
    const bool upper = p > 0.8646647167633873;
    const double q = upper ? 1.0 - p : p;
    if (q > 0.1353352832366127) {  // exp(-2): implies 0 < p < 1
        const double w = q - 0.5;
        const double ww = w * w;
        double num = kNdtriP0[0];
#pragma unroll
        for (int i = 1; i < 5; ++i) num = num * ww + kNdtriP0[i];
        double den = 1.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) den = den * ww + kNdtriQ0[i];
        double x = w + w * ww * (num / den);
        x *= -2.5066282746310002;  // -sqrt(2 pi)
        return upper ? x : -x;
    }
    return ndtri_tail(p);
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
