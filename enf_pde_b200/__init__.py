"""enf_pde_b200 -- B200-native ENF steerable cross-attention (forward + backward) behind the
reference's `EquivariantCrossAttentionNeF.init/apply` API.  See DESIGN.md / INTEGRATION.md."""
from .invariant import (BaseInvariant, get_ca_invariant, get_sa_invariant, Ponita2D, RelativePositionND, NormRelativePositionND,   # noqa: F401
                        AbsolutePositionND, RelativePosition2DPeriodic, PonitaPos2D,
                        RelativePositionPolarPeriodic, RelativeLatitudePeriodic, BallInvariant, BallLatInvariant)
from .nef import EquivariantCrossAttentionNeF, params_to_leaves, leaves_to_params, last_launch_counts   # noqa: F401
from .latents import init_latents, make_coords   # noqa: F401
from .ode import PonitaODEGen, MLPODE, solve_latent_ode   # noqa: F401
from .meta import inner_loop, outer_step_gradients   # noqa: F401
from ._lib import EnfLibraryError, load as load_library   # noqa: F401
