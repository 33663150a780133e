"""gpflow.likelihoods stand-ins on the DSDGP path: Gaussian and SwitchedLikelihood of Gaussians
(dgplib/dsdgp.py:117, :194; tests/test_multitask_dsdgp.py:57).  The arithmetic runs in csrc/likelihood.cu."""
from .params import Param, Parameterized, ParamList


class Likelihood(Parameterized):
    pass


class Gaussian(Likelihood):
    def __init__(self, variance=1.0, name=None):
        super().__init__(name)
        self.variance = Param(variance, positive=True)


class SwitchedLikelihood(Likelihood):
    """Row-wise selection among likelihoods by the integer in the last column of Y."""

    def __init__(self, likelihood_list, name=None):
        super().__init__(name)
        for lik in likelihood_list:
            if not isinstance(lik, Gaussian):
                raise NotImplementedError("the B200 path supports SwitchedLikelihood over Gaussian likelihoods only")
        self.likelihood_list = ParamList(likelihood_list)
        self.num_likelihoods = len(likelihood_list)
