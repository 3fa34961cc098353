"""Drop-in for the reference's models.I3D_doubled (pt/models/I3D_doubled.py): same Model signature,
module names and state-dict keys; computation by the libivf sm_100a kernels (see _i3d_native.py).
Head pool [2,7,7] = the 224x224, 16-frame Something-Something geometry (:313-314)."""
from ._i3d_native import I3DBase, InceptionModule, MaxPool3dSamePadding, Unit3D  # noqa: F401


class Model(I3DBase):
    def __init__(self, num_classes=400, spatial_squeeze=True, final_endpoint='Logits', name='inception_i3d',
                 in_channels=3, dropout_keep_prob=0.5, last_stride=1, stride_mod_layers=[], softMax=False,
                 lastRelu=None):
        super().__init__()
        self._init_i3d(num_classes, spatial_squeeze, final_endpoint, name, in_channels, dropout_keep_prob,
                       last_stride, stride_mod_layers, softMax, lastRelu, pool_time=2, pool_hw=(7, 7))
