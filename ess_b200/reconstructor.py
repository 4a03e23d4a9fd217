"""`ImageReconstructor` equivalent (reference: e2vid/image_reconstructor.py:18-163, default options)
and the fused T-window unroll.

`update_reconstruction(event_tensor) -> (img, states, latent)` keeps the reference's per-window
contract (normalise non-zeros over the whole batch tensor -> reflect-pad to a multiple of
2^num_encoders -> model -> carry states in `last_states_for_each_channel['grayscale']`), all under
no_grad.  Unlike the reference there is no host synchronisation: the `if num_nonzeros > 0` test
(e2vid/utils/inference_utils.py:100) is a device-side select and the CudaTimer syncs are gone.

`unroll(data, T, C)` is the fused form of the trainer loop (training/ess_supervised_trainer.py:126-130):
statistics of all T windows in ONE launch, then T encoder steps, with the E2VID image decoder run on
the last window only (the trainers overwrite `img_fake` every iteration) -- contract B of SURVEY.md s8d.
`stats_reduce_fn` lets a data-parallel caller all-reduce the [T,3] statistics (global-batch semantics).
"""
import math
import os

import torch

from . import ops


def crop_padding(height, width, num_encoders):
    """CropParameters.__init__ (e2vid/utils/inference_utils.py:311-330): (left, right, top, bottom)."""
    f = 2 ** num_encoders
    hc, wc = int(f * math.ceil(height / f)), int(f * math.ceil(width / f))
    return (math.ceil(0.5 * (wc - width)), math.floor(0.5 * (wc - width)),
            math.ceil(0.5 * (hc - height)), math.floor(0.5 * (hc - height)))


class ImageReconstructor:
    def __init__(self, model, height, width, num_bins, device=None, options=None, augmentation=False,
                 standardization=False):
        if augmentation:
            raise NotImplementedError('augmentation=True is a CPU/PIL path in the reference; not on the hot path')
        self.model = model
        self.device = device
        self.height, self.width, self.num_bins = height, width, num_bins
        self.standardization = standardization
        self.no_recurrent = bool(getattr(options, 'no_recurrent', False))
        self.no_normalize = bool(getattr(options, 'no_normalize', False))
        if getattr(options, 'color', False):
            raise NotImplementedError('color reconstruction (per-Bayer-channel states) is not built')
        # EventPreprocessor options (e2vid/utils/inference_utils.py:73-93)
        self.flip = bool(getattr(options, 'flip', False))
        self.hot_pixels = None          # int32 [n, 2] (x, y) on the device, or None
        hp_file = getattr(options, 'hot_pixels_file', None)
        if hp_file:
            import numpy as np
            try:
                self.set_hot_pixels(np.loadtxt(hp_file, delimiter=',').astype(np.int64).reshape(-1, 2))
            except IOError:
                print('WARNING: could not load hot pixels file: {}'.format(hp_file))   # as the reference
        self.last_states_for_each_channel = {'grayscale': None}
        self.stats_reduce_fn = None

    def set_hot_pixels(self, xy):
        """(x, y) pixel locations zeroed in every event tensor before normalisation (inference_utils.py:88-89)."""
        xy = torch.as_tensor(xy, dtype=torch.int32).reshape(-1, 2)
        self.hot_pixels = xy.to(self.device if self.device is not None else 'cuda').contiguous() if xy.numel() else None

    def _step(self, window, stats_row, states, with_image, want_head=True):
        """normalise + reflect-pad + layout/precision conversion (one kernel) -> model."""
        B, C, H, W = window.shape
        left, right, top, bottom = crop_padding(H, W, self.model.num_encoders)
        Hp, Wp = H + top + bottom, W + left + right
        buf = self.model.head_planes_buffer(B, Hp, Wp, window.device)
        if buf is not None:      # tensor-core head conv: write its bf16 hi/lo operand planes directly
            ops.event_prepare_planes(window, stats_row, not self.no_normalize, Hp, Wp, top, left, buf, flip=self.flip)
            return self.model.forward_planes(buf, Hp, Wp, states, with_image=with_image, want_head=want_head)
        x = ops.event_prepare(window, stats_row, not self.no_normalize, Hp, Wp, top, left, (C + 7) // 8 * 8, flip=self.flip)
        return self.model.forward_nhwc(x, states, with_image=with_image)

    @staticmethod
    def _per_sample_contiguous(t):
        B, C, H, W = t.shape
        if t.stride(1) == H * W and t.stride(2) == W and t.stride(3) == 1:
            return t
        return t.contiguous()

    def update_reconstruction(self, event_tensor, event_tensor_id=None, stamp=None, with_image=True):
        ev = event_tensor if event_tensor.is_cuda else event_tensor.to(self.device)
        with torch.no_grad(), ops.on_device_of(ev):
            ev = self._per_sample_contiguous(ev.float())
            B, C, H, W = ev.shape
            if self.hot_pixels is not None:
                ops.zero_pixels(ev, self.hot_pixels)       # in place, like the reference
            stats = None
            if not self.no_normalize:
                stats = ops.event_stats(ev, 1, C)
                if self.stats_reduce_fn is not None:
                    stats = self.stats_reduce_fn(stats)
            img, states, latent = self._step(ev, stats, self.last_states_for_each_channel['grayscale'], with_image)
            self.last_states_for_each_channel['grayscale'] = None if self.no_recurrent else states
            if self.standardization and img is not None:                 # image_reconstructor.py:129-134
                b, h, w = img.size(0), img.size(2), img.size(3)
                out = img.reshape(b, -1)
                out = out - out.min(1, keepdim=True)[0]
                out = out / out.max(1, keepdim=True)[0]
                img = out.view(b, 1, h, w)
        return img, states, latent

    def unroll(self, data, num_windows, channels, image_on_last_only=True, graph=None):
        """data [B, T*C, H, W] -> (img, states, latent) of the last window; resets the state first.

        graph=True (default: env ESS_B200_GRAPH, off): the T encoder steps (~15 launches each) are captured ONCE per
        input shape / input address into a CUDA graph and replayed -- the step is issued by one cudaGraphLaunch instead
        of ~300 launches, which removes the host-side issue time (14 ms at DSEC size, more than the device time at the
        DDD17 size) and the few-us device gaps between dependent kernels.  The statistics kernel (and its data-parallel
        all-reduce) stay outside the graph and feed it through a fixed buffer.  Graph semantics: the returned tensors
        live in the graph's memory pool and are OVERWRITTEN by the next graphed unroll of this reconstructor -- consume
        them (decoder forward/backward) before the next call, as every trainer does; clone() what must survive."""
        if graph is None:
            graph = os.environ.get('ESS_B200_GRAPH', '0') == '1'
        with torch.no_grad(), ops.on_device_of(data):
            data = self._per_sample_contiguous(data.float())
            self.last_states_for_each_channel = {'grayscale': None}
            if self.hot_pixels is not None:
                ops.zero_pixels(data, self.hot_pixels)     # all T*C channels at once
            stats = None
            if not self.no_normalize:
                stats = ops.event_stats(data, num_windows, channels)
                if self.stats_reduce_fn is not None:
                    stats = self.stats_reduce_fn(stats)
            if graph:
                img, states, latent = self._unroll_graphed(data, stats, num_windows, channels, image_on_last_only)
            else:
                img, states, latent = self._unroll_loop(data, stats, num_windows, channels, image_on_last_only)
            self.last_states_for_each_channel['grayscale'] = states
        return img, states, latent

    def _unroll_loop(self, data, stats, num_windows, channels, image_on_last_only):
        states = None
        img = latent = None
        for i in range(num_windows):
            win = data[:, i * channels:(i + 1) * channels]
            need_img = (i == num_windows - 1) or not image_on_last_only
            img, states, latent = self._step(win, stats[i] if stats is not None else None, states, need_img,
                                             want_head=(i == num_windows - 1))
            if self.no_recurrent:
                states = None
        return img, states, latent

    # ------------------------------------------------------------------------------------ CUDA graph of the unroll
    MAX_ADDRESS_GRAPHS = 4

    def _unroll_graphed(self, data, stats, num_windows, channels, image_on_last_only):
        """Replay (capturing on first use) the graph of _unroll_loop for this input.  Graphs are keyed by everything
        their recorded launches depend on: input address / shape / strides, the window split, the reconstructor's
        options, the model's parameter versions and mode.  Up to MAX_ADDRESS_GRAPHS graphs read the caller's tensor in
        place (a training loop's batches come back at the same one or two allocator addresses); further addresses
        share one graph that reads a reconstructor-owned copy of the input (one extra device-to-device copy)."""
        dev = data.device
        base = (tuple(data.shape), tuple(data.stride()), num_windows, channels, bool(image_on_last_only), self.flip,
                self.no_normalize, self.no_recurrent, self.standardization, self.model._key(),
                os.environ.get('ESS_B200_STACK', '1'), str(dev))
        G = self.__dict__.setdefault('_graphs', {})
        if G and next(iter(G))[1:] != base:      # another shape / new weights: drop the stale graphs and their memory
            G.clear()
            self.__dict__.pop('_graph_static_in', None)
        key = (data.data_ptr(),) + base
        ent = G.get(key)
        src = data
        if ent is None and sum(1 for k in G if k[0] != 'static') >= self.MAX_ADDRESS_GRAPHS:
            key = ('static',) + base
            ent = G.get(key)
            buf = self.__dict__.get('_graph_static_in')
            if buf is None or buf.shape != data.shape or buf.device != dev:
                buf = torch.empty(data.shape, device=dev, dtype=torch.float32)
                self._graph_static_in = buf
            buf.copy_(data)
            src = buf
        if ent is None:
            cur = torch.cuda.current_stream(dev)
            stream = self.__dict__.get('_graph_stream')
            if stream is None or stream.device != dev:
                stream = self._graph_stream = torch.cuda.Stream(device=dev)
                self._graph_pool = torch.cuda.graph_pool_handle()
            st_buf = stats.clone() if stats is not None else None
            # one eager pass on the capture stream first: every lazily created buffer (packed weights, scheduler slots
            # of that stream, zero-bordered input planes, row-stacked plane buffers) must exist before the capture
            stream.wait_stream(cur)
            with torch.cuda.stream(stream):
                self._unroll_loop(src, st_buf, num_windows, channels, image_on_last_only)
            cur.wait_stream(stream)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._graph_pool, stream=stream):
                out = self._unroll_loop(src, st_buf, num_windows, channels, image_on_last_only)
            ent = (g, st_buf, out)
            G[key] = ent
        g, st_buf, out = ent
        if st_buf is not None:
            st_buf.copy_(stats)
        g.replay()
        return out
