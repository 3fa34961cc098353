"""Repaired drop-in for the reference's FindMasksComparison_I3D_smth.py: same command line (utils.load_args), same
config-dict files, same find_masks(...) signature, result dicts, ClassScore*case*.txt files, folder names and
pickles (pt/FindMasksComparison_I3D_smth.py:52-315) - computed on the native path (interpreting_video_features_b200.
drivers / search).  Nothing runs at import time (the reference parses sys.argv on import, :29).

    python FindMasksComparison_I3D_smth.py -c configs/config_i3d_smth.py -g 0 --use_cuda --checkpoint model.pth.tar \\
        --subDir out --subsetFile clips_of_interest.csv [--lam1 .01 --lam2 .02 --optIter 300 --gradCamType guessed -msl ""]
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

import utils  # noqa: E402  (the drop-in next to this file)

RESIZE_SIZE_WIDTH = 224
RESIZE_SIZE_HEIGHT = 224
ARGS = None  # set by main(); find_masks falls back to sub-directory "run" without it


def clips_of_interest(class_oi):
    """The csv the reference reads with pandas (:139): one column per class number, the video ids of interest below.
    Returns {class number (str): set of ids (int)}; None selects every clip (the `or classOI is None` of :173)."""
    if class_oi is None:
        return None
    import pandas as pd
    df = pd.read_csv(class_oi)
    return {str(col): set(int(v) for v in df[col].dropna()) for col in df.keys()}


def find_masks(dat_loader, model, hyper_params, lam1, lam2, N, maskType="gradient", temporalMaskType="freeze",
               classOI=None, verbose=True, maxMaskLength=None, doGradCam=False, runTempMask=True, backend=None,
               sub_dir=None, out_root=".", viz=None):
    """pt/FindMasksComparison_I3D_smth.py:125-315.  dat_loader yields (sequence [B,3,T,H,W], label [B], video_id [B]);
    returns the list of final masks and writes the result pickles."""
    sub_dir = sub_dir if sub_dir is not None else (ARGS.subDir if ARGS is not None else "run")
    if backend is None:
        model.eval()
        backend = drivers.NativeBackend(model, "I3D", (RESIZE_SIZE_WIDTH, RESIZE_SIZE_HEIGHT),
                                        micro_batch=getattr(ARGS, "microBatch", None) or 8)
    wanted = clips_of_interest(classOI)
    masks, tm_all, gc_all = [], [], []
    for i, (sequence, label, video_id) in enumerate(dat_loader):
        if i % 50 == 0:
            print("on idx: ", i)
        ids = [str(v) for v in video_id]  # bug 9: the batch's id list is never clobbered
        selected = []
        for b in range(len(ids)):
            cls = str(int(label[b]))
            if wanted is None or (cls in wanted and int(ids[b]) in wanted[cls]):
                selected.append(b)
        tm, gc, ms = drivers.process_batch(backend, sequence, label, ids, selected, hyper_params["gradCamType"], lam1,
                                           lam2, N, temporalMaskType, sub_dir, run_temp_mask=runTempMask,
                                           do_grad_cam=doGradCam, out_root=out_root, video_id_cast=int, viz=viz,
                                           verbose=verbose)
        tm_all += tm
        gc_all += gc
        masks += ms
    results = os.path.join(out_root, "results")
    tag = os.path.basename(str(classOI)) if classOI is not None else "None"
    drivers.dump_results(tm_all, gc_all,
                         os.path.join(results, "allTimeMaskResults_" + sub_dir + "_" + tag + "_" + ".p"),
                         os.path.join(results, "allGradCamResults_" + sub_dir + "_" + tag + "_" + ".p"))
    return masks


def build_model(config, args, device, device_ids):
    cnn_def = importlib.import_module(config["conv_model"])
    model = cnn_def.Model(config["num_classes"], last_stride=1, stride_mod_layers=args.mod_stride_layers or "",
                          softMax=1)
    model = torch.nn.DataParallel(model, device_ids[:1]).to(device)  # one process per GPU (torchrun shards clips)
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
    from data_loader_jpg import ImLoader
    model = build_model(config, args, device, device_ids)
    val_data = ImLoader(config["data_folder"] + "/validation/", get_item_id=True, clip_size=config["clip_size"],
                        as_uint8=True)
    val_loader = torch.utils.data.DataLoader(val_data, batch_size=config["batch_size"], shuffle=False,
                                             num_workers=config["num_workers"], pin_memory=True, drop_last=True)
    lam1 = args.lam1 if args.lam1 is not None else 0.01
    lam2 = args.lam2 if args.lam2 is not None else 0.02
    N = args.optIter if args.optIter is not None else 300
    viz = None
    if not args.noViz:
        import visualisation
        viz = visualisation.driver_hook(RESIZE_SIZE_WIDTH, RESIZE_SIZE_HEIGHT)
    return find_masks(val_loader, model, config, lam1, lam2, N, "central", config["maskPerturbType"],
                      classOI=args.subsetFile, doGradCam=True, runTempMask=True, viz=viz)


if __name__ == "__main__":
    main()
