"""Repaired drop-in for the reference's FindMasksComparison_I3D_KTH.py (I3D or ConvLSTM on KTH Actions): same command
line, config files, find_masks(...) signature, result dicts, ClassScore files, folder names and pickles
(pt/FindMasksComparison_I3D_KTH.py:45-380), computed on the native path.  Repairs: the mask module's real
function names (bug 4), no shadowing of `mask` (bug 1), the ConvLSTM is explained at its own layer with archType
"CLSTM" (bug 8), missing config keys (bug 3).

    python FindMasksComparison_I3D_KTH.py -c configs/config_i3d_kth.py -g 0 --use_cuda --checkpoint m.pth.tar --subDir out
"""
import importlib
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
try:
    from interpreting_video_features_b200 import drivers
except ImportError:
    sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))
    from interpreting_video_features_b200 import drivers

import utils  # noqa: E402

RESIZE_SIZE_WIDTH = 160
RESIZE_SIZE_HEIGHT = 120
ARGS = None

_ACTIONS = ("boxing", "handclapping", "handwaving", "jogging", "running", "walking")


def clips_of_interest(split_type):
    """The 24 test clips the reference inspects (KTH.py:154-203): per action, scenarios d1..d4 of two subjects."""
    if split_type == "original":
        still, moving = ("person17", "person18"), ("person24", "person25")
    else:
        still, moving = ("person07", "person08"), ("person09", "person10")
    out = []
    for k, action in enumerate(_ACTIONS):
        a, b = still if k < 3 else moving
        out += [[a, action, "d1", "_1"], [a, action, "d2", "_1"], [b, action, "d3", "_1"], [b, action, "d4", "_1"]]
    return out


def find_masks(dat_loader, model, config, lam1, lam2, N, ita, maskType="gradient", temporalMaskType="freeze",
               classOI=None, verbose=True, maxMaskLength=None, doGradCam=False, runTempMask=True, backend=None,
               sub_dir=None, out_root=".", viz=None, select_all=False):
    """pt/FindMasksComparison_I3D_KTH.py:126-380.  dat_loader yields (input [B,3,T,H,W], target [B], label tags [B])."""
    sub_dir = sub_dir if sub_dir is not None else (ARGS.subDir if ARGS is not None else "run")
    if backend is None:
        model.eval()
        arch = "I3D" if "I3D" in config["conv_model"] else "CLSTM"
        backend = drivers.NativeBackend(model, arch, (RESIZE_SIZE_WIDTH, RESIZE_SIZE_HEIGHT),
                                        micro_batch=getattr(ARGS, "microBatch", None) or 8)
    coi = clips_of_interest(config["splitType"])
    masks, tm_all, gc_all = [], [], []
    for i, (inp, target, label) in enumerate(dat_loader):
        if i % 50 == 0:
            print("on idx: ", i)
        tags = [str(t).strip() for t in label]
        selected = [b for b, tag in enumerate(tags) if select_all or any(all(part in tag for part in c) for c in coi)]
        tm, gc, ms = drivers.process_batch(backend, inp, target, tags, selected, config["gradCamType"], lam1, lam2, N,
                                           temporalMaskType, sub_dir, run_temp_mask=runTempMask, do_grad_cam=doGradCam,
                                           out_root=out_root, video_id_cast=str, viz=viz, verbose=verbose)
        tm_all += tm
        gc_all += gc
        masks += ms
    results = os.path.join(out_root, "results")
    drivers.dump_results(tm_all, gc_all,
                         os.path.join(results, "I3d_KTH_allTimeMaskResults_original_" + sub_dir + ".p"),
                         os.path.join(results, "I3d_KTH_allGradCamResults_original_" + sub_dir + ".p"))
    return masks


def build_model(config, args, device, device_ids):
    cnn_def = importlib.import_module(config["conv_model"])
    if "I3D" in config["conv_model"]:
        model = cnn_def.Model(config["num_classes"], last_stride=1, stride_mod_layers=args.mod_stride_layers or "",
                              softMax=1, finalTimeLength=4,
                              dropout_keep_prob=args.dropout if args.dropout is not None else 0.5)
    else:
        model = cnn_def.Model(num_classes=6, nb_lstm_units=config["clstm_hidden"], channels=3, conv_kernel_size=(5, 5),
                              top_layer=True, avg_pool=False, batch_normalization=config["batch_norm"],
                              lstm_layers=config["clstm_layers"], step=32, dropout=config["dropout"],
                              conv_stride=config["conv_stride"], image_size=(160, 120),
                              effective_step=[7, 15, 23, 31])
    model = torch.nn.DataParallel(model, device_ids[:1]).to(device)
    if args.fp32:
        model.module.set_mode("fp32")
    if args.checkpoint and os.path.isfile(args.checkpoint):
        print(" > Loading checkpoint '{}'".format(args.checkpoint))
        checkpoint = torch.load(args.checkpoint, map_location="cpu")
        model.load_state_dict(checkpoint["state_dict"])
    else:
        print(" !#! No checkpoint found at '{}'".format(args.checkpoint))
    return model


def main(argv=None):
    global ARGS
    ARGS = args = utils.load_args(argv)
    config = utils.merged_config(args)
    device, device_ids = utils.setup_cuda_devices(args)
    if device.type != "cuda":
        raise SystemExit("the native path needs --use_cuda (there is no CPU fallback)")
    print(" > Using device: {}".format(device.type))
    print(" > Active GPU ids: {}".format(device_ids))
    if config["input_mode"] != "jpg":
        raise ValueError("Please provide a valid input mode")
    from data_loader_kth import KTHImLoader
    model = build_model(config, args, device, device_ids)
    val_data = KTHImLoader(config["data_folder"] + "/test", clip_size=config["clip_size"], get_item_id=True,
                           as_uint8=True)
    val_loader = torch.utils.data.DataLoader(val_data, batch_size=config["batch_size"], shuffle=bool(config["shuffle"]),
                                             num_workers=config["num_workers"], pin_memory=True, drop_last=True)
    lam1 = args.lam1 if args.lam1 is not None else 0.02
    lam2 = args.lam2 if args.lam2 is not None else 0.04
    N = args.optIter if args.optIter is not None else 100
    viz = None
    if not args.noViz:
        import visualisation
        viz = visualisation.driver_hook(RESIZE_SIZE_WIDTH, RESIZE_SIZE_HEIGHT)
    return find_masks(val_loader, model, config, lam1, lam2, N, 1, "central", config["maskPerturbType"], classOI=None,
                      doGradCam=True, runTempMask=True, viz=viz)


if __name__ == "__main__":
    main()
