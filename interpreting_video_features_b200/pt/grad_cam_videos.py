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

    def _i3d(self, x, indices):
        name = self.target_layer_names[-1]
        eng = self.model._engine(x)
        if name not in eng.acts or name != "Mixed_5c":
            raise _lib.IvfError("native Grad-CAM targets 'Mixed_5c' (the layer the reference drivers use, "
                                "pt/FindMasksComparison_I3D_smth.py:258); got %r" % name)
        eng.set_input(x)
        w_out, h_out = self.input_spatial_size  # cv2 dsize = (width, height), pt/grad_cam_videos.py:119-120
        cam, _, out = eng.gradcam(indices, (h_out, w_out), self.normalizePerFrame,
                                  graphed=os.environ.get("IVF_GRADCAM_GRAPH", "1") != "0")
        return cam, out

    def _clstm(self, x, indices):
        cam, out = self.model.native_gradcam(x, indices, self.input_spatial_size, self.normalizePerFrame)
        return cam, out
