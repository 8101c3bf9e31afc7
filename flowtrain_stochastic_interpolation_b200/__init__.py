"""flowtrain-b200: the B200-native (sm_100a) hot path of flowtrain_stochastic_interpolation —
the Unet3D velocity field v_theta(x_t, t), the interpolant construction, the fixed-grid ODE/SDE
samplers and the categorical embed/decode — behind flowtrain's own Python API.

Importing this package loads ``csrc/libftb.so`` (built by ``__graft_entry__.build()``); there is
no PyTorch or CPU fallback for any of the math.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .interpolation import (BaseInterpolant, EncDecInterpolant, LinearInterpolant, MirrorInterpolant,
                            SBDMInterpolant, StochasticInterpolator, TrigInterpolant)
from .solvers import (ODEFlowSolver, ODEOneSidedDenoisingSolver, SDEOneSidedDenoisingSolver,
                      integrate_fixed, odeSol_RK4)
from .boreholes import (conditioning_frontend, draw_boreholes, jittered_grid_points, make_boreholes_mask,
                        make_combined_mask, make_surface_mask)
from .data import DevicePrefetcher, SyntheticGeoStreamingDataset, get_data_loader
from .ensemble import EnsembleVotes
from .task import (EMAShadow, Geo3DStochInterp, Geo3DStochInterpCond, lightning_checkpoint,
                   load_model_with_ema_option, decode, ema_update_, embed, flow_loss,
                   simplex_embedding)
from .training import BucketAllReduce, CondFlowTrainer, FlowTrainer, flatten_parameters
from .unet3d import Unet3D, Unet3DCond

__all__ = [
    "Unet3D", "Unet3DCond", "StochasticInterpolator", "BaseInterpolant", "LinearInterpolant", "TrigInterpolant",
    "EncDecInterpolant", "SBDMInterpolant", "MirrorInterpolant", "ODEFlowSolver",
    "ODEOneSidedDenoisingSolver", "SDEOneSidedDenoisingSolver", "odeSol_RK4", "integrate_fixed",
    "Geo3DStochInterp", "Geo3DStochInterpCond", "EMAShadow", "FlowTrainer", "CondFlowTrainer", "EnsembleVotes",
    "make_boreholes_mask", "make_surface_mask", "make_combined_mask", "conditioning_frontend", "draw_boreholes",
    "SyntheticGeoStreamingDataset", "DevicePrefetcher", "get_data_loader",
    "jittered_grid_points", "load_model_with_ema_option", "lightning_checkpoint", "BucketAllReduce", "flatten_parameters", "embed", "decode", "flow_loss", "ema_update_", "simplex_embedding",
]
