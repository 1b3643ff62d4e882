"""gibbssampler_b200 -- B200-native constrained-realization + C_l-sampling Gibbs step.

Drop-in for the sampler classes of Gabriel-Ducrocq/GibbsSampler (same class names, constructor
arguments and sample()/run() signatures); all arithmetic runs in hand-written sm_100a kernels
behind the C ABI of include/gibbs_b200.h.  There is no CPU fallback.
"""
from ._lib import GibbsB200Error, GS_ALM_COMPLEX, GS_ALM_REAL  # noqa: F401

__version__ = "0.1.0"
