"""mcmctoffitting_b200 -- B200-native ``lnprob`` path for neutron time-of-flight MCMC fitting.

One hot path of gcrich/mcmcTOFfitting rebuilt for sm_100a: the per-walker forward model and
log-likelihood that emcee evaluates on every step.  The product is ``libtofgpu.so`` (CUDA kernels
behind the C ABI of ``include/tofgpu.h``); this package is the thin Python host mirroring the
reference's ``lnprob`` / ``pool=`` interface.  There is no CPU implementation here.
"""
from . import config
from .config import ModelConfig
from .model import TofModel
from .lnprob import TofLnProb, BatchedPool, make_lnprob
from ._lib import TofError
from . import ppc, templates, shapes, dataio

__all__ = ["config", "ModelConfig", "TofModel", "TofLnProb", "BatchedPool", "make_lnprob", "TofError", "ppc", "templates", "shapes", "dataio"]
