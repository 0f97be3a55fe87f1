/*
 * ns_oracle.c -- CPU restatement of the jaxns 2.6.9 static nested-sampling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under jaxns_b200/ may include, link or call this
 * file; it is the checker for the CUDA path (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).
 *
 * Parity status: the reference (pure Python on JAX + TFP) cannot be imported in the
 * build container (no jax / jaxlib / tfp wheels), so per-chain trajectories, the word
 * order of 64-bit draws and XLA's f64 erf_inv rounding are "parity unpinned" against a
 * live jaxns run.  What IS pinned: Threefry-2x32 against the Random123 KATs and the
 * published legacy jax.random.split(PRNGKey(0)) words, the partitionable split/counter
 * layout + uniform/normal recipe against the key(42) values printed in the JAX
 * documentation's PRNG tutorial, ndtri against scipy's Cephes ndtri, erf_inv
 * against scipy.special.erfinv, tree counts against the reference's golden vectors
 * (src/jaxns/internals/tests/test_tree_structure.py:19-70) and the log-space
 * identities of src/jaxns/internals/tests/test_log_semiring.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/jaxns unless noted).  Third-party arithmetic restated from the
 * published algorithms: jax (>=0.6, unpinned) jax/_src/prng.py + random.py
 * (Threefry-2x32-20, partitionable split / random_bits, uniform, normal), XLA
 * ErfInv f64 (Giles' polynomial), tfp_nightly special_math.ndtri (Cephes ndtri).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------
 * Threefry-2x32-20 (jax/_src/prng.py threefry2x32; Random123).
 * ---------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void o_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o0, uint32_t *o1) {
    static const int R0[4] = {13, 15, 26, 6};
    static const int R1[4] = {17, 29, 16, 24};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    x0 += ks[0];
    x1 += ks[1];
    for (int g = 0; g < 5; ++g) {
        const int *R = (g & 1) ? R1 : R0;
        for (int i = 0; i < 4; ++i) {
            x0 += x1;
            x1 = rotl32(x1, R[i]);
            x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t) (g + 1);
    }
    *o0 = x0;
    *o1 = x1;
}

/* jax.random.split(key, n)[i] under jax_threefry_partitionable=True
 * (_threefry_split_foldlike): child i = threefry(key; hi(i), lo(i)).
 * jaxns forces that flag: internals/mixed_precision.py:11-15. */
static inline void split_child(const uint32_t key[2], uint64_t i, uint32_t out[2]) {
    o_threefry2x32(key[0], key[1], (uint32_t) (i >> 32), (uint32_t) i, &out[0], &out[1]);
}

/* element i of a 64-bit random_bits draw (_threefry_random_bits_partitionable). */
static inline uint64_t bits64(const uint32_t key[2], uint64_t i) {
    uint32_t a, b;
    o_threefry2x32(key[0], key[1], (uint32_t) (i >> 32), (uint32_t) i, &a, &b);
    return ((uint64_t) a << 32) | (uint64_t) b;
}

/* jax.random.uniform f64 (_uniform): mantissa fill, -1, scale, max(lo, .). */
static inline double bits_to_unit(uint64_t bits) {
    uint64_t fb = (bits >> 12) | 0x3FF0000000000000ull;
    double f;
    memcpy(&f, &fb, 8);
    return f - 1.0;
}

static inline double uniform_lohi(uint64_t bits, double lo, double hi) {
    double v = bits_to_unit(bits) * (hi - lo) + lo;
    return v > lo ? v : lo; /* lax.max(minval, .) */
}

static inline double uniform01(const uint32_t key[2], uint64_t i) {
    return uniform_lohi(bits64(key, i), 0.0, 1.0);
}

/* XLA ErfInv for f64 (xla/client/lib/math.cc ErfInv64 / chlo erf_inv): Giles' piecewise
 * polynomial in w = -log1p(-x*x). */
double o_erfinv(double x) {
    static const double A[23] = {
            -3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18,
            1.115787767802518096e-17, -1.333171662854620906e-16, 2.0972767875968561637e-17,
            6.6376381343583238325e-15, -4.0545662729752068639e-14, -8.1519341976054721522e-14,
            2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
            1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08,
            4.2347877827932403518e-07, -1.3654692000834678645e-06, -1.3882523362786468719e-05,
            0.0001867342080340571352, -0.00074070253416626697512, -0.0060336708714301490533,
            0.24015818242558961693, 1.6536545626831027356};
    static const double B[19] = {
            2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07,
            1.8239629214389227755e-08, 1.5027403968909827627e-06, -4.013867526981545969e-06,
            2.9234449089955446044e-06, 1.2475304481671778723e-05, -4.7318229009055733981e-05,
            6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
            0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066,
            -0.0037512085075692412107, 0.005370914553590063617, 1.0052589676941592334,
            3.0838856104922207635};
    static const double C[17] = {
            -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,
            -3.7894654401267369937e-09, 7.6157012080783393804e-09, -1.4960026627149240478e-08,
            2.9147953450901080826e-08, -6.7711997758452339498e-08, 2.2900482228026654717e-07,
            -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
            7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026,
            1.0103004648645343977, 4.8499064014085844221};
    if (fabs(x) == 1.0) return x * INFINITY;
    double w = -log1p(x * -x);
    double p;
    if (w < 6.25) {
        w = w - 3.125;
        p = A[0];
        for (int i = 1; i < 23; ++i) p = A[i] + p * w;
    } else if (w < 16.0) {
        w = sqrt(w) - 3.25;
        p = B[0];
        for (int i = 1; i < 19; ++i) p = B[i] + p * w;
    } else {
        w = sqrt(w) - 5.0;
        p = C[0];
        for (int i = 1; i < 17; ++i) p = C[i] + p * w;
    }
    return p * x;
}

/* jax.random.normal f64 (_normal_real): sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1)). */
static inline double normal_from_bits(uint64_t bits) {
    const double lo = -0.99999999999999988897769753748; /* nextafter(-1, 0) */
    double u = uniform_lohi(bits, lo, 1.0);
    return 1.4142135623730951 * o_erfinv(u);
}

/* tfp special_math.ndtri (Cephes ndtri): piecewise rational approximations.
 * Called from WrappedTFPDistribution._forward (framework/wrapped_tfp_distribution.py:77-84)
 * via Normal.quantile. */
double o_ndtri(double p) {
    static const double P0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1,
                                 -5.66762857469070293439E1, 1.39312609387279679503E1,
                                 -1.23916583867381258016E0};
    static const double Q0[9] = {1.0, 1.95448858338141759834E0, 4.67627912898881538453E0,
                                 8.63602421390890590575E1, -2.25462687854119370527E2,
                                 2.00260212380060660359E2, -8.20372256168333339912E1,
                                 1.59056225126211695515E1, -1.18331621121330003142E0};
    static const double P1[9] = {4.05544892305962419923E0, 3.15251094599893866154E1,
                                 5.71628192246421288162E1, 4.40805073893200834700E1,
                                 1.46849561928858024014E1, 2.18663306850790267539E0,
                                 -1.40256079171354495875E-1, -3.50424626827848203418E-2,
                                 -8.57456785154685413611E-4};
    static const double Q1[9] = {1.0, 1.57799883256466749731E1, 4.53907635128879210584E1,
                                 4.13172038254672030440E1, 1.50425385692907503408E1,
                                 2.50464946208309415979E0, -1.42182922854787788574E-1,
                                 -3.80806407691578277194E-2, -9.33259480895457427372E-4};
    static const double P2[9] = {3.23774891776946035970E0, 6.91522889068984211695E0,
                                 3.93881025292474443415E0, 1.33303460815807542389E0,
                                 2.01485389549179081538E-1, 1.23716634817820021358E-2,
                                 3.01581553508235416007E-4, 2.65806974686737550832E-6,
                                 6.23974539184983293730E-9};
    static const double Q2[9] = {1.0, 6.02427039364742014255E0, 3.67983563856160859403E0,
                                 1.37702099489081330271E0, 2.16236993594496635890E-1,
                                 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                                 2.89247864745380683936E-6, 6.79019408009981274425E-9};
    if (p == 0.0) return -INFINITY;
    if (p == 1.0) return INFINITY;
    if (!(p > 0.0 && p < 1.0)) return NAN;
    const double one_minus_em2 = 0.8646647167633873; /* -expm1(-2) */
    const double em2 = 0.1353352832366127;           /* exp(-2) */
    int upper = p > one_minus_em2;
    double q = upper ? 1.0 - p : p; /* maybe_complement_p */
    double x;
    if (q > em2) {
        double w = q - 0.5;
        double ww = w * w;
        double num = P0[0], den = Q0[0];
        for (int i = 1; i < 5; ++i) num = num * ww + P0[i];
        for (int i = 1; i < 9; ++i) den = den * ww + Q0[i];
        x = w + w * ww * (num / den);
        x *= -2.5066282746310002; /* -sqrt(2 pi) */
    } else {
        double z = sqrt(-2.0 * log(q));
        double first = z - log(z) / z;
        double rz = 1.0 / z;
        const double *P = (z >= 8.0) ? P2 : P1;
        const double *Q = (z >= 8.0) ? Q2 : Q1;
        double num = P[0], den = Q[0];
        for (int i = 1; i < 9; ++i) num = num * rz + P[i];
        for (int i = 1; i < 9; ++i) den = den * rz + Q[i];
        x = first - num / den / z;
    }
    return upper ? x : -x;
}

/* jnp.logaddexp (jax/_src/numpy/ufuncs.py): amax + log1p(exp(-|delta|)), nan-delta -> x1+x2. */
static inline double logaddexp_(double a, double b) {
    double amax = a > b ? a : b;
    double delta = a - b;
    if (isnan(delta)) return a + b;
    return amax + log1p(exp(-fabs(delta)));
}

double o_logaddexp(double a, double b) { return logaddexp_(a, b); }

/* ------------------------------------------------------------------------------------
 * Model: prior quantile transform + registered likelihood family.
 * Follows Model.forward (framework/model.py:167-176) -> compute_log_likelihood
 * (framework/ops.py:302-326; NaN -> -inf at :323-325).
 * ---------------------------------------------------------------------------------- */
enum { FAM_GAUSS_DENSE = 0, FAM_GAUSS_MIX_DIAG = 1, FAM_EGGBOX = 2, FAM_ROSENBROCK = 3, FAM_SHELLS = 4 };
enum { PRIOR_UNIFORM = 0, PRIOR_NORMAL = 1 };

typedef struct {
    int32_t family;
    int32_t D;
    int32_t prior_kind;
    int32_t K;             /* mixture components / shells */
    const double *prior_a; /* [D] low or loc */
    const double *prior_b; /* [D] (high-low) or scale */
    const double *params;  /* family-specific, see include/nsb200.h */
} OModel;

static void transform_(const OModel *m, const double *U, double *X) {
    for (int j = 0; j < m->D; ++j) {
        if (m->prior_kind == PRIOR_UNIFORM)
            X[j] = U[j] * m->prior_b[j] + m->prior_a[j]; /* tfd.Uniform.quantile */
        else
            X[j] = o_ndtri(U[j]) * m->prior_b[j] + m->prior_a[j]; /* tfd.Normal.quantile */
    }
}

static double loglik_(const OModel *m, const double *X) {
    const int D = m->D;
    const double *P = m->params;
    double r;
    switch (m->family) {
        case FAM_GAUSS_DENSE: {
            /* MultivariateNormalTriL(loc, scale_tril).log_prob(x) with host-precomputed
             * Linv = inv(scale_tril) (row-major, lower) and c = -sum log diag - D/2 log 2pi.
             * params = [c, mu[D], Linv[D*D]]. */
            double c = P[0];
            const double *mu = P + 1;
            const double *Linv = P + 1 + D;
            double q = 0.0;
            for (int i = 0; i < D; ++i) {
                double z = 0.0;
                for (int j = 0; j <= i; ++j) z += Linv[i * D + j] * (X[j] - mu[j]);
                q += z * z;
            }
            r = c - 0.5 * q;
            break;
        }
        case FAM_GAUSS_MIX_DIAG: {
            /* logaddexp over K diagonal Gaussians (benchmarks/difficult_problems/main.py:95-125).
             * params per component: [logc, mean[D], inv_sigma[D]]. */
            r = -INFINITY;
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (1 + 2 * D);
                double q = 0.0;
                for (int j = 0; j < D; ++j) {
                    double z = (X[j] - pk[1 + j]) * pk[1 + D + j];
                    q += z * z;
                }
                double g = pk[0] - 0.5 * q;
                r = (k == 0) ? g : logaddexp_(r, g);
            }
            break;
        }
        case FAM_EGGBOX: {
            /* docs/examples/egg_box.ipynb cell 2: (2 + prod cos(theta/2))^5 */
            double y = 1.0;
            for (int j = 0; j < D; ++j) y *= cos(0.5 * X[j]);
            y = 2.0 + y;
            double y2 = y * y;
            r = y2 * y2 * y;
            break;
        }
        case FAM_ROSENBROCK: {
            /* benchmarks/difficult_problems/main.py:69-92 */
            double y = 0.0;
            for (int i = 0; i < D - 1; ++i) {
                double a = X[i + 1] - X[i] * X[i];
                double b = 1.0 - X[i];
                y += 100.0 * (a * a) + b * b;
            }
            r = -y;
            break;
        }
        case FAM_SHELLS: {
            /* docs/examples/gaussian_shells.ipynb cell 2, K shells.
             * params per shell: [w, r, c[D]]. */
            r = -INFINITY;
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (2 + D);
                double w = pk[0], rad = pk[1];
                double s = 0.0;
                for (int j = 0; j < D; ++j) {
                    double dlt = X[j] - pk[2 + j];
                    s += dlt * dlt;
                }
                double e = sqrt(s) - rad;
                double g = -0.5 * (e * e) / (w * w) - log(sqrt(2.0 * M_PI * (w * w)));
                r = (k == 0) ? g : logaddexp_(r, g);
            }
            break;
        }
        default:
            r = NAN;
    }
    if (isnan(r)) r = -INFINITY; /* ops.py:323-325 */
    return r;
}

double o_forward(const OModel *m, const double *U, double *Xscratch) {
    transform_(m, U, Xscratch);
    return loglik_(m, Xscratch);
}

void o_forward_batch(const OModel *m, const double *U, int64_t n, double *out_logL, double *out_X) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double *X = out_X ? out_X + i * m->D : (double *) alloca(sizeof(double) * m->D);
        out_logL[i] = o_forward(m, U + i * m->D, X);
    }
}

/* d log L / dU at U: jax.grad(model.forward) (samplers/uni_slice_sampler.py:135) written out by hand for the five
 * families -- chain rule through the per-dimension quantile (Uniform: dX/dU = b; Normal: b sqrt(2 pi) exp(z^2/2),
 * z = ndtri(U)).  Used only by the gradient variants (o_set_gradient_flags). */
void o_grad_U(const OModel *m, const double *U, double *g) {
    const int D = m->D;
    const double *P = m->params;
    double *X = (double *) alloca(sizeof(double) * D);
    double *J = (double *) alloca(sizeof(double) * D);
    for (int j = 0; j < D; ++j) {
        if (m->prior_kind == PRIOR_UNIFORM) {
            X[j] = U[j] * m->prior_b[j] + m->prior_a[j];
            J[j] = m->prior_b[j];
        } else {
            double z = o_ndtri(U[j]);
            X[j] = z * m->prior_b[j] + m->prior_a[j];
            J[j] = m->prior_b[j] * sqrt(2.0 * M_PI) * exp(0.5 * z * z);
        }
        g[j] = 0.0;
    }
    switch (m->family) {
        case FAM_GAUSS_DENSE: { /* -Linv^T (Linv (x - mu)) */
            const double *mu = P + 1;
            const double *Linv = P + 1 + D;
            for (int i = 0; i < D; ++i) {
                double z = 0.0;
                for (int j = 0; j <= i; ++j) z += Linv[i * D + j] * (X[j] - mu[j]);
                for (int j = 0; j <= i; ++j) g[j] -= Linv[i * D + j] * z;
            }
            break;
        }
        case FAM_GAUSS_MIX_DIAG: { /* softmax-weighted component gradients */
            double *gk = (double *) alloca(sizeof(double) * m->K);
            double mx = -INFINITY;
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (1 + 2 * D);
                double q = 0.0;
                for (int j = 0; j < D; ++j) {
                    double z = (X[j] - pk[1 + j]) * pk[1 + D + j];
                    q += z * z;
                }
                gk[k] = pk[0] - 0.5 * q;
                if (gk[k] > mx) mx = gk[k];
            }
            double den = 0.0;
            for (int k = 0; k < m->K; ++k) den += exp(gk[k] - mx);
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (1 + 2 * D);
                double w = exp(gk[k] - mx) / den;
                for (int j = 0; j < D; ++j) g[j] -= w * (X[j] - pk[1 + j]) * pk[1 + D + j] * pk[1 + D + j];
            }
            break;
        }
        case FAM_EGGBOX: {
            double y = 1.0;
            for (int j = 0; j < D; ++j) y *= cos(0.5 * X[j]);
            double b = 2.0 + y;
            double b4 = (b * b) * (b * b);
            for (int j = 0; j < D; ++j) {
                double rest = 1.0;
                for (int i = 0; i < D; ++i)
                    if (i != j) rest *= cos(0.5 * X[i]);
                g[j] = 5.0 * b4 * rest * (-0.5 * sin(0.5 * X[j]));
            }
            break;
        }
        case FAM_ROSENBROCK: {
            for (int i = 0; i < D - 1; ++i) {
                double a = X[i + 1] - X[i] * X[i];
                double b = 1.0 - X[i];
                g[i] -= -400.0 * a * X[i] - 2.0 * b;
                g[i + 1] -= 200.0 * a;
            }
            break;
        }
        case FAM_SHELLS: {
            double *gk = (double *) alloca(sizeof(double) * m->K);
            double *ek = (double *) alloca(sizeof(double) * m->K);
            double *rk = (double *) alloca(sizeof(double) * m->K);
            double mx = -INFINITY;
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (2 + D);
                double w = pk[0], rad = pk[1], sq = 0.0;
                for (int j = 0; j < D; ++j) {
                    double dlt = X[j] - pk[2 + j];
                    sq += dlt * dlt;
                }
                rk[k] = sqrt(sq);
                ek[k] = rk[k] - rad;
                gk[k] = -0.5 * (ek[k] * ek[k]) / (w * w) - log(sqrt(2.0 * M_PI * (w * w)));
                if (gk[k] > mx) mx = gk[k];
            }
            double den = 0.0;
            for (int k = 0; k < m->K; ++k) den += exp(gk[k] - mx);
            for (int k = 0; k < m->K; ++k) {
                const double *pk = P + (size_t) k * (2 + D);
                double w = pk[0];
                double wt = exp(gk[k] - mx) / den;
                for (int j = 0; j < D; ++j) g[j] -= wt * (ek[k] / (w * w)) * (X[j] - pk[2 + j]) / rk[k];
            }
            break;
        }
        default:
            for (int j = 0; j < D; ++j) g[j] = NAN;
    }
    for (int j = 0; j < D; ++j) g[j] *= J[j];
}

/* Model.sample_U (framework/model.py:122-138) with the hidden split inside
 * Ctx.next_rng_key (framework/context.py:107-109): uniform(split(key,2)[1], (D,)). */
static void sample_U_(const uint32_t key[2], int D, double *U) {
    uint32_t k[2];
    split_child(key, 1, k);
    for (int j = 0; j < D; ++j) U[j] = uniform01(k, (uint64_t) j);
}

/* _single_uniform_sample (nested_samplers/common/uniform_sample.py:12-60). */
static void single_uniform_sample_(const OModel *m, const uint32_t key_in[2], double *U, double *logL,
                                   int64_t *nev) {
    double *X = (double *) alloca(sizeof(double) * m->D);
    uint32_t key[2], sk[2], tmp[2];
    split_child(key_in, 0, key);
    split_child(key_in, 1, sk);
    sample_U_(sk, m->D, U);
    *logL = o_forward(m, U, X);
    *nev = 1;
    while (*logL <= -INFINITY) {
        split_child(key, 1, sk);
        split_child(key, 0, tmp);
        key[0] = tmp[0];
        key[1] = tmp[1];
        sample_U_(sk, m->D, U);
        *logL = o_forward(m, U, X);
        *nev += 1;
    }
}

/* draw_uniform_samples over keys = split(sample_key, N) (common/initialisation.py:47-60). */
void o_init_batch(const OModel *m, const uint32_t sample_key[2], int64_t N, double *out_U, double *out_logL,
                  int64_t *out_nev) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < N; ++i) {
        uint32_t k[2];
        split_child(sample_key, (uint64_t) i, k);
        single_uniform_sample_(m, k, out_U + i * m->D, out_logL + i, out_nev + i);
    }
}

/* ------------------------------------------------------------------------------------
 * Seed choice: sample_uniformly_masked / resample_indicies (internals/random.py:55-60,78-84)
 * with cumulative_logsumexp (internals/log_semiring.py:51-92) restated through the table
 * c_1 = 0, c_{q+1} = logaddexp(c_q, 0) because the mask is a suffix of the sorted live set.
 * ---------------------------------------------------------------------------------- */
void o_seed_table(int64_t N, double *c) {
    double acc = -INFINITY;
    for (int64_t q = 0; q < N; ++q) {
        acc = logaddexp_(acc, 0.0);
        c[q] = acc; /* c[q] = c_{q+1} */
    }
}

/* Direct restatement (O(N) scan) used to validate the table route on small cases. */
int64_t o_seed_index_scan(const double *live_logL, int64_t N, double contour, double u) {
    double *cuml = (double *) malloc(sizeof(double) * N);
    double acc = -INFINITY;
    for (int64_t i = 0; i < N; ++i) {
        double lw = (live_logL[i] > contour) ? 0.0 : -INFINITY;
        acc = logaddexp_(acc, lw);
        cuml[i] = acc;
    }
    double log_r = cuml[N - 1] + log(1.0 - u);
    int64_t lo = 0, hi = N; /* searchsorted side='left' */
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (cuml[mid] < log_r) lo = mid + 1; else hi = mid;
    }
    free(cuml);
    return lo;
}

int64_t o_seed_index_table(const double *live_logL, int64_t N, const double *ctab, double contour, double u) {
    /* j0 = first index with log_L > contour (sorted ascending) */
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (live_logL[mid] > contour) hi = mid; else lo = mid + 1;
    }
    int64_t j0 = lo, nsat = N - j0;
    if (nsat == 0) return 0;
    double log_r = ctab[nsat - 1] + log(1.0 - u);
    lo = 0;
    hi = nsat;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (ctab[mid] < log_r) lo = mid + 1; else hi = mid;
    }
    int64_t idx = j0 + lo;
    /* log_r <= c_nsat always, so lo <= nsat-1; the -inf prefix is only hit when log_r = -inf */
    if (log_r == -INFINITY) idx = 0;
    return idx;
}

/* ------------------------------------------------------------------------------------
 * Slice chain: BaseAbstractMarkovSampler._get_sample (samplers/bases.py:63-75),
 * UniDimSliceSampler.get_seed_point / get_sample_from_seed / _new_proposal
 * (samplers/uni_slice_sampler.py:343-441, :114-273).
 * ---------------------------------------------------------------------------------- */
static void sample_direction_(const uint32_t key[2], int D, double *d) {
    /* _sample_direction (uni_slice_sampler.py:23-38) */
    if (D == 1) {
        d[0] = 1.0;
        return;
    }
    double s = 0.0;
    for (int j = 0; j < D; ++j) {
        d[j] = normal_from_bits(bits64(key, (uint64_t) j));
        s += d[j] * d[j];
    }
    double nrm = sqrt(s);
    for (int j = 0; j < D; ++j) d[j] /= nrm;
}

static void slice_bounds_(const double *U0, const double *d, int D, double *left, double *right) {
    /* _slice_bounds (uni_slice_sampler.py:41-64) */
    double t1r = INFINITY, t1l = -INFINITY, t0r = INFINITY, t0l = -INFINITY;
    for (int j = 0; j < D; ++j) {
        double t1 = (1.0 - U0[j]) / d[j];
        double t0 = -U0[j] / d[j];
        if (t1 >= 0.0 && t1 < t1r) t1r = t1;
        if (t1 <= 0.0 && t1 > t1l) t1l = t1;
        if (t0 >= 0.0 && t0 < t0r) t0r = t0;
        if (t0 <= 0.0 && t0 > t0l) t0l = t0;
    }
    *right = t0r < t1r ? t0r : t1r;
    *left = t0l > t1l ? t0l : t1l;
}

/* jnp.linspace(0.5, 1., S)[j] (jax/_src/numpy/lax_numpy.py _linspace): start*(1-step)+stop*step
 * with step = j/div, endpoint appended exactly. */
/* Test knob (tests/test_oracle_cpu.py::test_notebook_runs_*): a constant shrink factor instead of the 2.6.9
 * schedule -- the plain midpoint rule of the jaxns versions the example notebooks were run with.  < 0 = off. */
static double g_fixed_alpha = -1.0;
void o_set_fixed_alpha(double a) { g_fixed_alpha = a; }

/* Test knob: the gradient variants of the chain (uni_slice_sampler.py:202-214 gradient_slice = bit 0, :255-269
 * gradient_guided = bit 1).  Parity for these is UNPINNED against the reference (jax is not installed here and the
 * reference's tests hold no golden vectors for them): the oracle restates the published control flow. */
static int g_grad_flags = 0;
void o_set_gradient_flags(int f) { g_grad_flags = f; }

static double alpha_(int j, int S) {
    if (g_fixed_alpha >= 0.0) return g_fixed_alpha;
    if (S == 1) return 0.5;
    int div = S - 1;
    if (j == div) return 1.0;
    double step = (double) j / (double) div;
    return 0.5 * (1.0 - step) + 1.0 * step;
}

void o_slice_chain(const OModel *m, const uint32_t chain_key[2], double contour, const double *live_U,
                   const double *live_logL, int64_t N, const double *ctab, int S, int k, int midpoint,
                   double *out_U, double *out_logL, int64_t *out_nev, double *ph_U, double *ph_logL,
                   int64_t *out_seed_idx) {
    const int D = m->D;
    double *U0 = (double *) alloca(sizeof(double) * D);
    double *d = (double *) alloca(sizeof(double) * D);
    double *x = (double *) alloca(sizeof(double) * D);
    double *X = (double *) alloca(sizeof(double) * D);
    double *gr = (double *) alloca(sizeof(double) * D);
    uint32_t sample_key[2], seed_key[2], direction_key[2], sample_key2[2];
    split_child(chain_key, 0, sample_key); /* bases.py:64 */
    split_child(chain_key, 1, seed_key);
    double u = uniform01(seed_key, 0); /* random.py:59, shape (1,) */
    int64_t idx = o_seed_index_table(live_logL, N, ctab, contour, u);
    if (out_seed_idx) *out_seed_idx = idx;
    memcpy(U0, live_U + idx * D, sizeof(double) * D);
    double logL0 = live_logL[idx];
    split_child(sample_key, 0, direction_key); /* uni_slice_sampler.py:410 */
    split_child(sample_key, 1, sample_key2);
    sample_direction_(direction_key, D, d);
    int64_t nev = 0;
    for (int j = 0; j < S; ++j) {
        uint32_t slice_key[2], run_key[2], t_key[2], after_key[2], tmp[2];
        split_child(sample_key2, (uint64_t) j, slice_key); /* :420-423 */
        double alpha = alpha_(j, S);
        split_child(slice_key, 0, run_key); /* :201 (n_key = child 1 unused) */
        split_child(slice_key, 2, t_key);
        split_child(slice_key, 3, after_key);
        double left, right;
        if (g_grad_flags & 1) { /* climb the gradient (:202-214) */
            o_grad_U(m, U0, gr);
            nev += 1;
            double gs = 0.0;
            for (int q = 0; q < D; ++q) gs += gr[q] * gr[q];
            double gn = sqrt(gs);
            int mask = (gn == 0.0) || !isfinite(gn);
            if (!mask)
                for (int q = 0; q < D; ++q) d[q] = gr[q] / gn;
            slice_bounds_(U0, d, D, &left, &right);
            if (!mask) left = 0.0;
        } else {
            slice_bounds_(U0, d, D, &left, &right);
        }
        double uu = uniform01(t_key, 0);
        double t = left + uu * (right - left); /* :83-85 */
        for (int q = 0; q < D; ++q) x[q] = U0[q] + t * d[q];
        double logL = o_forward(m, x, X);
        int64_t ne = 1;
        for (;;) {
            int sat = logL > contour;
            int lesser = (logL0 == contour) && (logL == contour); /* :160-166 */
            if (sat || lesser) break;
            split_child(run_key, 1, t_key); /* :169 (child 2 = shrink_key unused) */
            split_child(run_key, 0, tmp);
            run_key[0] = tmp[0];
            run_key[1] = tmp[1];
            if (t < 0.0) left = t; /* :92-111 */
            if (t > 0.0) right = t;
            if (midpoint) {
                if (t < 0.0) left = alpha * left;
                if (t > 0.0) right = alpha * right;
            }
            uu = uniform01(t_key, 0);
            t = left + uu * (right - left);
            for (int q = 0; q < D; ++q) x[q] = U0[q] + t * d[q];
            logL = o_forward(m, x, X);
            ne += 1;
        }
        memcpy(U0, x, sizeof(double) * D);
        logL0 = logL;
        nev += ne;
        if (g_grad_flags & 2) { /* Householder reflection about the gradient at the accepted point (:255-269) */
            uint32_t after_key1[2];
            split_child(after_key, 0, after_key1);
            o_grad_U(m, U0, gr);
            nev += 1;
            double gs = 0.0, dot = 0.0, rs = 0.0;
            for (int q = 0; q < D; ++q) gs += gr[q] * gr[q];
            double gn = sqrt(gs);
            int mask = (gn < 1e-10) || !isfinite(gn);
            for (int q = 0; q < D; ++q) dot += d[q] * (gr[q] / gn);
            for (int q = 0; q < D; ++q) {
                x[q] = d[q] - 2.0 * dot * (gr[q] / gn);
                rs += x[q] * x[q];
            }
            double rn = sqrt(rs);
            if (mask)
                sample_direction_(after_key1, D, d);
            else
                for (int q = 0; q < D; ++q) d[q] = x[q] / rn;
        } else {
            sample_direction_(after_key, D, d); /* :272 */
        }
        /* phantom capture: cumulative_samples[-(k+1):-1] (:430-433) */
        if (k > 0 && j >= S - 1 - k && j < S - 1) {
            int slot = j - (S - 1 - k);
            memcpy(ph_U + (size_t) slot * D, U0, sizeof(double) * D);
            ph_logL[slot] = logL0;
        }
    }
    memcpy(out_U, U0, sizeof(double) * D);
    *out_logL = logL0;
    *out_nev = nev;
}

/* get_samples (nested_samplers/sharded/sharded_static.py:88-129): keys = split(key, m);
 * chains [chain_begin, chain_end) evaluated here (PartitionSpec('shard') = contiguous blocks). */
void o_slice_batch(const OModel *m, const uint32_t key[2], double contour, const double *live_U,
                   const double *live_logL, int64_t N, const double *ctab, int S, int k, int midpoint,
                   int64_t chain_begin, int64_t chain_end, double *out_U, double *out_logL,
                   int64_t *out_nev, double *ph_U, double *ph_logL, int64_t *out_seed_idx) {
    const int D = m->D;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = chain_begin; i < chain_end; ++i) {
        uint32_t ck[2];
        split_child(key, (uint64_t) i, ck);
        int64_t o = i - chain_begin;
        o_slice_chain(m, ck, contour, live_U, live_logL, N, ctab, S, k, midpoint, out_U + o * D,
                      out_logL + o, out_nev + o, ph_U ? ph_U + (size_t) o * k * D : NULL,
                      ph_logL ? ph_logL + (size_t) o * k : NULL, out_seed_idx ? out_seed_idx + o : NULL);
    }
}

/* UniformSampler._get_sample (samplers/uniform_samplers.py:42-85), max_likelihood_evals = 100. */
void o_uniform_batch(const OModel *m, const uint32_t key[2], double contour, int64_t chain_begin,
                     int64_t chain_end, double *out_U, double *out_logL, int64_t *out_nev) {
    const int D = m->D;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = chain_begin; i < chain_end; ++i) {
        double *X = (double *) alloca(sizeof(double) * D);
        uint32_t ck[2], k2[2], sk[2], tmp[2];
        split_child(key, (uint64_t) i, ck);
        int64_t o = i - chain_begin;
        double *U = out_U + o * D;
        split_child(ck, 0, k2);
        split_child(ck, 1, sk);
        sample_U_(sk, D, U);
        double logL = o_forward(m, U, X);
        int64_t ne = 1;
        while (!((logL > contour) || (logL == contour) || (ne >= 100))) {
            split_child(k2, 1, sk);
            split_child(k2, 0, tmp);
            k2[0] = tmp[0];
            k2[1] = tmp[1];
            sample_U_(sk, D, U);
            logL = o_forward(m, U, X);
            ne += 1;
        }
        out_logL[o] = logL;
        out_nev[o] = ne;
    }
}

/* ------------------------------------------------------------------------------------
 * Evidence recurrences: _update_evidence_calc_op (internals/shrinkage_statistics.py:43-94),
 * serial, exactly in the reference's operation order.
 * state layout (EvidenceCalculation field order, common/types.py:12-24):
 *   [log_L, log_X, log_X2, log_Z, log_ZX, log_Z2, log_dZ, log_dZ2]
 * ---------------------------------------------------------------------------------- */
void o_evidence_scan(double st[8], const double *logL, const double *nlive, int64_t M, double *per_sample) {
    const double log2_ = log(2.0), loghalf = log(0.5);
    for (int64_t i = 0; i < M; ++i) {
        double n = nlive[i];
        double ln = log(n), lnp1 = log(n + 1.0), lnp2 = log(n + 2.0);
        double midL = loghalf + logaddexp_(logL[i], st[0]);
        double T = -logaddexp_(0.0, -ln);
        double t = -lnp1;
        double T2 = -logaddexp_(0.0, log2_ - ln);
        double t2 = log2_ - lnp1 - lnp2;
        double tT = -logaddexp_(0.0, -ln) - lnp2;
        double lX = st[1], lX2 = st[2], lZ = st[3], lZX = st[4], lZ2 = st[5], ldZ2 = st[7];
        double dZ = lX + t + midL;
        double nX = lX + T;
        double nX2 = lX2 + T2;
        double nZ = logaddexp_(lZ, dZ);
        double nZX = logaddexp_(lZX + T, lX2 + tT + midL);
        double x2t2m2 = lX2 + t2 + 2.0 * midL;
        double nZ2 = logaddexp_(logaddexp_(lZ2, log2_ + lZX + t + midL), x2t2m2);
        double ndZ2 = logaddexp_(ldZ2, x2t2m2);
        st[0] = logL[i];
        st[1] = nX;
        st[2] = nX2;
        st[3] = nZ;
        st[4] = nZX;
        st[5] = nZ2;
        st[6] = dZ;
        st[7] = ndZ2;
        if (per_sample) memcpy(per_sample + i * 8, st, sizeof(double) * 8);
    }
}

/* sample_evidence (utils.py:433-476): S stochastic simulations of the shrinkage.  Simulation s uses
 * key_s = split(key, S)[s] (:475) and per-sample keys split(key_s, M)[i] (:470); the scan body (:450-465)
 * draws log T = log(uniform(key_i, ())) / n_i, next_X = X * T, dZ = (X - next_X) * L, next_Z = Z + dZ, with
 * LogSpace arithmetic (internals/log_semiring.py:28-48,132-178): X - next_X = signed_logaddexp(+,-). */
static inline double log_sub_(double la, double lb) { /* log(exp(la) - exp(lb)), la >= lb */
    const double amax = la > lb ? la : lb;
    const double delta = -fabs(lb - la);
    if (delta != delta) return la + lb;
    return amax + log1p(-exp(delta));
}

void o_sample_evidence(const uint32_t key[2], const double *nlive, const double *logL, int64_t M, int64_t S,
                       double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        uint32_t ks[2];
        split_child(key, (uint64_t) s, ks);
        double log_Z = -INFINITY, log_X = 0.0;
        for (int64_t i = 0; i < M; ++i) {
            uint32_t ki[2];
            split_child(ks, (uint64_t) i, ki);
            const double log_T = log(uniform01(ki, 0)) / nlive[i];
            const double next_X = log_X + log_T;
            const double dZ = log_sub_(log_X, next_X) + logL[i];
            log_Z = logaddexp_(log_Z, dZ);
            log_X = next_X;
        }
        out[s] = log_Z;
    }
}

int o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void o_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
