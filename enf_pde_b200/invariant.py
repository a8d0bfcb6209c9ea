"""Host-side mirror of enf/steerable_attention/invariant (names, dims, factory).

Only metadata lives here: the arithmetic of every invariant and of its gaussian window is inside
the CUDA kernels (per-query / per-latent records, enf_pde_b200/csrc/enf_stages.cu).  The classes keep
the reference's attribute names (`dim`, `num_x_pos_dims`, `num_z_pos_dims`, `num_z_ori_dims`,
`is_periodic`) so host code written against the reference reads the same.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class BaseInvariant:
    """enf/steerable_attention/invariant/_base_invariant.py:5-23"""
    invariant_type: str
    dim: int
    num_x_pos_dims: int
    num_x_ori_dims: int
    num_z_pos_dims: int
    num_z_ori_dims: int
    is_periodic: bool = False

    @property
    def pose_dim(self):
        """width of the RAW latent pose p (positions + orientation angles)."""
        return self.num_z_pos_dims + self.num_z_ori_dims


def RelativePositionND(num_dims):            # rel_pos.py:5-24
    return BaseInvariant("rel_pos", num_dims, num_dims, 0, num_dims, 0)


def NormRelativePositionND(num_dims):        # norm_rel_pos.py:6-22
    return BaseInvariant("norm_rel_pos", 1, num_dims, 0, num_dims, 0)


def AbsolutePositionND(num_dims):            # abs_pos.py:6-25
    return BaseInvariant("abs_pos", num_dims, num_dims, 0, num_dims, 0)


def RelativePosition2DPeriodic(num_dims=2):  # rel_pos_periodic.py:6-33
    assert num_dims == 2, "RelativePosition2DPeriodic currently only supports 2D input."
    return BaseInvariant("rel_pos_periodic", 2 * num_dims, num_dims, 0, num_dims, 0, True)


def PonitaPos2D():                           # ponita.py:6-18
    return BaseInvariant("ponita", 2, 2, 0, 2, 1)


def Ponita2D():                              # ponita.py:46-61 (self-attention variant: the queries carry an orientation too)
    return BaseInvariant("ponita", 3, 2, 1, 2, 1)


def RelativePositionPolarPeriodic():         # polar_periodic.py:6-33
    return BaseInvariant("polar_periodic", 1, 2, 0, 2, 0, True)


def RelativeLatitudePeriodic():              # spherical_longitude.py:6-32
    return BaseInvariant("latitude_periodic", 4, 2, 0, 2, 0, True)


def BallInvariant():                         # ball.py:7-34
    return BaseInvariant("ball", 5, 3, 0, 4, 0)


def BallLatInvariant():                      # ball_lat.py:7-34
    return BaseInvariant("ball_lat", 6, 3, 0, 4, 0)


def get_ca_invariant(cfg) -> BaseInvariant:
    """enf/steerable_attention/invariant/__init__.py:47-78 (cfg needs .invariant_type and .num_in)."""
    t = cfg.invariant_type
    if t == "norm_rel_pos":
        return NormRelativePositionND(num_dims=cfg.num_in)
    if t == "rel_pos":
        return RelativePositionND(num_dims=cfg.num_in)
    if t == "rel_pos_periodic":
        assert cfg.num_in == 2, "RelativePosition2DPeriodic currently only supports 2D input."
        return RelativePosition2DPeriodic(num_dims=cfg.num_in)
    if t == "ponita":
        assert cfg.num_in == 2, "Ponita2D currently only supports 2D input."
        return PonitaPos2D()
    if t == "abs_pos":
        return AbsolutePositionND(num_dims=cfg.num_in)
    if t == "polar_periodic":
        return RelativePositionPolarPeriodic()
    if t == "latitude_periodic":
        return RelativeLatitudePeriodic()
    if t == "ball":
        return BallInvariant()
    if t == "ball_lat":
        return BallLatInvariant()
    raise ValueError(f"Unknown invariant type: {t}.")


def get_sa_invariant(cfg) -> BaseInvariant:
    """enf/steerable_attention/invariant/__init__.py:13-45: the self-attention variants (only `ponita` differs: Ponita2D)."""
    if cfg.invariant_type == "ponita":
        assert cfg.num_in == 2, "Ponita2D currently only supports 2D input."
        return Ponita2D()
    return get_ca_invariant(cfg)
