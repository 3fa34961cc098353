"""Drop-in for the reference's grad_cam_videos module (pt/grad_cam_videos.py): GradCamVideo with the
same constructor and call surface.  The reference hooks the target layer, backpropagates the class
score through the whole network (weight gradients included) and finishes on the host in numpy/cv2
(:64-142).  Here: one native forward that keeps the target activation, the head's backward only
(the gradient w.r.t. Mixed_5c needs nothing else), and ONE fused kernel for
channel weights -> weighted sum -> ReLU -> bilinear upsample -> temporal repeat -> normalisation.
"""
import numpy as np
import os

import torch

try:
    from interpreting_video_features_b200 import _lib, ops
except ImportError:
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from interpreting_video_features_b200 import _lib, ops


class GradCamVideo:
    def __init__(self, model, target_layer_names, class_dict, use_cuda, input_spatial_size=224,
                 normalizePerFrame=False, archType="I3D"):
        self.model = model
        self.archType = archType
        self.cuda = use_cuda
        self.normalizePerFrame = normalizePerFrame
        if isinstance(input_spatial_size, int):
            self.input_spatial_size = (input_spatial_size, input_spatial_size)
        elif len(input_spatial_size) == 1:
            self.input_spatial_size = (input_spatial_size[0], input_spatial_size[0])
        else:
            self.input_spatial_size = tuple(input_spatial_size)
        self.class_dict = class_dict
        self.target_layer_names = list(target_layer_names)
        if not use_cuda:
            raise _lib.IvfError("GradCamVideo: the native path runs on a B200 only (use_cuda=False has no "
                                "CPU fallback)")
        self.model = model.cuda()

    def __call__(self, input, index=None):
        """input [1,3,T,H,W]; returns (cam float32 numpy [T', H, W] in [0,1], output [1,classes])."""
        cams, output = self.batched(input, None if index is None else [int(index)])
        return cams[0], output

    def batched(self, input, indices=None):
        """Any batch size: cams [B,T',H,W] (numpy) and outputs [B,classes] (device tensor)."""
        x = input if input.dtype == torch.uint8 else input.float()  # uint8 frames are converted on the device
        x = x.cuda(non_blocking=True).contiguous()
        if self.archType == "I3D":
            cam, out = self._i3d(x, indices)
        elif self.archType == "CLSTM":
            cam, out = self._clstm(x, indices)
        else:
            raise ValueError("archType must be 'I3D' or 'CLSTM'")
        # pinned destination from torch's caching host allocator (a fresh block per call: the numpy array the caller
        # keeps owns it), one asynchronous D2H, one stream synchronise - not a pageable bounce
        host = torch.empty(cam.shape, dtype=cam.dtype, pin_memory=True)
        host.copy_(cam, non_blocking=True)
        torch.cuda.current_stream(cam.device).synchronize()
        return host.numpy(), out

    def stream(self, clips, indices=None, batch=32):
        """Grad-CAM over many clips with the copies hidden behind the compute: clips [N,3,T,H,W] on the HOST
        (uint8 frames or fp32; pinned memory makes the copies asynchronous), indices None (arg-max class per clip)
        or a list of N classes.  Chunks of `batch` clips go through three streams - H2D of chunk i+1, forward +
        head' + fused CAM kernel of chunk i, D2H of chunk i-1 - over double-buffered device staging; the maps land
        in ONE pinned host array.  Returns (cams float32 numpy [N,T',H,W], outputs [N,classes] host tensor).
        I3D only (the ConvLSTM maps go through batched())."""
        if self.archType != "I3D":
            raise _lib.IvfError("GradCamVideo.stream serves archType 'I3D'")
        if clips.is_cuda:
            raise _lib.IvfError("GradCamVideo.stream takes host clips (device-resident clips: use batched())")
        n = clips.shape[0]
        dev = next(self.model.parameters()).device
        w_out, h_out = self.input_spatial_size
        eng = self.model._engine(clips[:1], batch=batch)
        act = eng.acts[self.target_layer_names[-1]]
        t_out = act.d * (clips.shape[2] // act.d)
        cams = torch.empty((n, t_out, h_out, w_out), dtype=torch.float32, pin_memory=True)
        outs = torch.empty((n, eng.num_classes), dtype=torch.float32, pin_memory=True)
        st = getattr(self, "_pipe", None)
        key = (batch, tuple(clips.shape[1:]), clips.dtype, str(dev), h_out, w_out)
        if st is None or st["key"] != key:
            st = self._pipe = dict(key=key, copy_in=torch.cuda.Stream(dev), compute=torch.cuda.Stream(dev),
                                   copy_out=torch.cuda.Stream(dev),
                                   stage=[torch.empty((batch,) + tuple(clips.shape[1:]), dtype=clips.dtype, device=dev)
                                          for _ in range(2)],
                                   cam=[torch.empty((batch, t_out, h_out, w_out), dtype=torch.float32, device=dev)
                                        for _ in range(2)],
                                   probs=[torch.empty((batch, eng.num_classes), dtype=torch.float32, device=dev)
                                          for _ in range(2)])
        if getattr(eng, "_fwd_graph", None) is None:  # capture the forward before the streams start to interleave
            eng.forward_graphed()
            torch.cuda.synchronize(dev)
        idx_dev = None
        if indices is not None:  # all classes uploaded once (a per-chunk pageable copy would stall the host)
            pad = (-n) % batch
            idx_host = torch.as_tensor([int(v) for v in indices] + [int(indices[-1])] * pad, dtype=torch.int32)
            idx_dev = idx_host.to(dev)
        consumed, drained = [None, None], [None, None]
        for c0 in range(0, n, batch):
            k = (c0 // batch) & 1
            m = min(batch, n - c0)
            with torch.cuda.stream(st["copy_in"]):
                if consumed[k] is not None:
                    st["copy_in"].wait_event(consumed[k])  # the engine has taken the previous content of stage[k]
                st["stage"][k][:m].copy_(clips[c0:c0 + m], non_blocking=True)
                if m < batch:  # ragged tail: repeat the last clip, dropped on the way out
                    st["stage"][k][m:].copy_(st["stage"][k][m - 1:m].expand(batch - m, *clips.shape[1:]))
                arrived = torch.cuda.Event()
                arrived.record(st["copy_in"])
            with torch.cuda.stream(st["compute"]):
                st["compute"].wait_event(arrived)
                if drained[k] is not None:
                    st["compute"].wait_event(drained[k])  # cam[k] / probs[k] have been copied out
                eng.set_input(st["stage"][k])
                consumed[k] = torch.cuda.Event()
                consumed[k].record(st["compute"])
                idx = None if idx_dev is None else idx_dev[c0:c0 + batch]
                _, _, out = eng.gradcam(idx, (h_out, w_out), self.normalizePerFrame, cam=st["cam"][k],
                                        layer=self.target_layer_names[-1])
                st["probs"][k].copy_(out)
                done = torch.cuda.Event()
                done.record(st["compute"])
            with torch.cuda.stream(st["copy_out"]):
                st["copy_out"].wait_event(done)
                cams[c0:c0 + m].copy_(st["cam"][k][:m], non_blocking=True)
                outs[c0:c0 + m].copy_(st["probs"][k][:m], non_blocking=True)
                drained[k] = torch.cuda.Event()
                drained[k].record(st["copy_out"])
        st["copy_out"].synchronize()
        return cams.numpy(), outs

    def _i3d(self, x, indices):
        name = self.target_layer_names[-1]
        eng = self.model._engine(x)
        if name not in eng.acts:
            raise _lib.IvfError("Grad-CAM target layer %r is not an endpoint of the model (%s)" % (name, list(eng.acts)))
        eng.set_input(x)
        w_out, h_out = self.input_spatial_size  # cv2 dsize = (width, height), pt/grad_cam_videos.py:119-120
        cam, _, out = eng.gradcam(indices, (h_out, w_out), self.normalizePerFrame,
                                  graphed=os.environ.get("IVF_GRADCAM_GRAPH", "1") != "0", layer=name)
        return cam, out

    def _clstm(self, x, indices):
        cam, out = self.model.native_gradcam(x, indices, self.input_spatial_size, self.normalizePerFrame)
        return cam, out
