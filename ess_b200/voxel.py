"""Events -> voxel grid on the GPU (reference: DSEC/dataset/representations.py:9-55 `VoxelGrid`,
datasets/data_util.py:54-126 `generate_voxel_grid`)."""
import torch

from . import ops
from ._lib import call


class VoxelGrid:
    """DSEC trilinear voxel grid; same constructor and `convert(x, y, pol, time)` as the reference class."""

    def __init__(self, channels, height, width, normalize):
        if normalize:
            raise NotImplementedError('VoxelGrid(normalize=True) is not built (settings_DSEC.yaml: normalize_event False)')
        self.nb_channels, self.height, self.width, self.normalize = channels, height, width, normalize

    def convert(self, x, y, pol, time):
        assert x.shape == y.shape == pol.shape == time.shape
        assert x.ndim == 1
        with ops.on_device_of(x):
            ops.require_cuda(x, y, pol, time)
            grid = torch.empty((self.nb_channels, self.height, self.width), device=x.device, dtype=torch.float32)
            call('essb_voxel_grid_dsec', ops._p(x.float().contiguous()), ops._p(y.float().contiguous()),
                 ops._p(pol.float().contiguous()), ops._p(time.float().contiguous()), x.numel(), self.nb_channels,
                 self.height, self.width, ops._p(grid), ops._stream())
        return grid


def generate_voxel_grid(events, shape, nr_temporal_bins, separate_pol=True):
    """DDD17 voxel grid from an [N, 4] float64 CUDA tensor of rows [x, y, t, polarity]."""
    height, width = shape
    assert events.shape[1] == 4 and nr_temporal_bins > 0 and width > 0 and height > 0
    with ops.on_device_of(events):
        ev = events.double().contiguous()
        ch = nr_temporal_bins * (2 if separate_pol else 1)
        grid = torch.empty((ch, height, width), device=ev.device, dtype=torch.float32)
        call('essb_voxel_grid_ddd17', ops._p(ev), ev.shape[0], nr_temporal_bins, height, width, int(separate_pol),
             ops._p(grid), ops._stream())
    return grid
