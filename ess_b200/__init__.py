"""ess_b200 -- B200-native (sm_100a) implementation of the ESS hot path (uzh-rpg/ess):
frozen recurrent E2VID event encoder unrolled over voxel-grid windows, SemSegE2VID decoder
forward/backward and the Dice + cross-entropy task loss, behind the reference's nn.Module surface.
"""
from ._lib import LIB_PATH, lib                                    # noqa: F401
from .e2vid import E2VIDRecurrent                                  # noqa: F401
from .loss import TaskLoss                                         # noqa: F401
from .metrics import MetricsSemseg                                 # noqa: F401
from .reconstructor import ImageReconstructor                      # noqa: F401
from .semseg import SemSegE2VID                                    # noqa: F401
from .style_encoder import StyleEncoderE2VID                       # noqa: F401
from .uda_losses import L1Loss, symJSDivLoss                       # noqa: F401

__all__ = ['E2VIDRecurrent', 'SemSegE2VID', 'StyleEncoderE2VID', 'L1Loss', 'symJSDivLoss', 'TaskLoss', 'MetricsSemseg', 'ImageReconstructor', 'lib', 'LIB_PATH']
