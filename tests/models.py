"""Model zoo shared by the tests: the same synthetic problems built for the product (jaxns_b200) and
for the oracle, from identical host arrays."""
import numpy as np


def product_models():
    import jaxns_b200 as j
    from jaxns_b200 import distributions as tfpd, likelihoods as lk

    def gauss(D, mu=15.0, rho=0.99):
        cov = np.full((D, D), rho) + (1 - rho) * np.eye(D)

        def prior_model():
            x = yield j.Prior(tfpd.MultivariateNormalTriL(loc=np.zeros(D), scale_tril=np.eye(D)), name='x')
            return x

        return j.Model(prior_model, lk.DenseGaussianLikelihood(np.full(D, mu), covariance_matrix=cov))

    def eggbox(D=2):
        def prior_model():
            x = yield j.Prior(tfpd.Uniform(low=np.zeros(D), high=10 * np.pi * np.ones(D)), name='theta')
            return x

        return j.Model(prior_model, lk.EggBoxLikelihood())

    def rosenbrock(D=10):
        def prior_model():
            z = yield j.Prior(tfpd.Uniform(low=-5 * np.ones(D), high=5 * np.ones(D)), name='z')
            return z

        return j.Model(prior_model, lk.RosenbrockLikelihood())

    def shells(D=2):
        c1 = np.zeros(D)
        c2 = np.zeros(D)
        c1[1 if D > 1 else 0] = -3.0
        c2[1 if D > 1 else 0] = 3.0

        def prior_model():
            x = yield j.Prior(tfpd.Uniform(low=-6. * np.ones(D), high=6. * np.ones(D)), name='theta')
            return x

        return j.Model(prior_model, lk.GaussianShellsLikelihood([c1, c2], [2.0, 2.0], [0.1, 0.1]))

    def mixture(D=100):
        m1 = np.zeros(D)
        m2 = np.zeros(D)
        m1[:2] = 6.0
        m2[:2] = 2.5

        def prior_model():
            z = yield j.Prior(tfpd.Uniform(low=-4. * np.ones(D), high=8. * np.ones(D)), name='z')
            return z

        return j.Model(prior_model, lk.GaussianMixtureLikelihood([m1, m2], [0.08, 0.8]))

    return dict(gauss=gauss, eggbox=eggbox, rosenbrock=rosenbrock, shells=shells, mixture=mixture)


def to_oracle(model, o):
    """Oracle model from the product model's packed host arrays (identical inputs on both sides)."""
    fam, D, pk, K, a, b, params = model.host_arrays()
    return o.OModel(fam, D, pk, a, b, params, K)
