"""ivf-b200: B200-native temporal-mask search and Grad-CAM (the interpretation hot path of
interpreting-video-features).  `pt/` mirrors the reference's video_features_pytorch/ module names
(add it to sys.path for a drop-in `import mask`, `from models import I3D_doubled`, ...);
`search` holds the batched fast path; `_lib`/`ops`/`engine` bind the sm_100a kernels of libivf.so."""
import os

PT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pt")
__all__ = ["PT_DIR"]
