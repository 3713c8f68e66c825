"""pyipn_b200 -- B200-native drop-in for pyipn's RFF-intensity -> binned-Poisson log-likelihood hot path.

Importing the package never touches the GPU; the first compute call loads ``libipn_b200.so`` (built in-tree by
``python -m pyipn_b200._build``) and creates a context on ``cuda:LOCAL_RANK``.  There is no CPU fallback.
"""
from ._lib import Context, IpnError, default_context, load_library  # noqa: F401
from .correlation import Correlator, correlate  # noqa: F401
from .fit import Fit, _expeced_rate, _expeced_rate_multiscale, ppc_generator  # noqa: F401
from .lightcurve import BinnedLightCurve, LightCurve  # noqa: F401
from .likelihood import loglik_delay_grid, loglik_sky_grid, sky_delays, time_delay  # noqa: F401
from .params import fold_draws  # noqa: F401
from .rff import RFF, RFF_multiscale  # noqa: F401
from .sky import (calculate_annulus, calculate_distance_and_norm, credible_regions, grb_phi_theta, localize_GRB,  # noqa: F401
                  theta_from_time_delay)
from .universe import Detector, GRB, Universe  # noqa: F401

__version__ = "0.1.0"
