"""Drop-in for the reference's models.I3D_doubled_kth (pt/models/I3D_doubled_kth.py): identical to
I3D_doubled except the `finalTimeLength` argument and the [finalTimeLength,4,5] head pool for
120x160 KTH frames (:177,203,302,306-307)."""
from ._i3d_native import I3DBase, InceptionModule, MaxPool3dSamePadding, Unit3D  # noqa: F401


class Model(I3DBase):
    def __init__(self, num_classes=400, spatial_squeeze=True, final_endpoint='Logits', name='inception_i3d',
                 in_channels=3, dropout_keep_prob=0.5, last_stride=1, stride_mod_layers=[], finalTimeLength=2,
                 softMax=False, lastRelu=None):
        super().__init__()
        self.finalTimeLength = finalTimeLength
        self._init_i3d(num_classes, spatial_squeeze, final_endpoint, name, in_channels, dropout_keep_prob,
                       last_stride, stride_mod_layers, softMax, lastRelu, pool_time=finalTimeLength,
                       pool_hw=(4, 5))
