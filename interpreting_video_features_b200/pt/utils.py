"""Drop-in for the reference's utils module (pt/utils.py): the command line of the FindMasksComparison drivers
(`load_args`, same flag names and short forms, :12-91), config loading by file path (`load_module`, :118-125),
device selection (`setup_cuda_devices`, :137-142) and the DataParallel checkpoint-key helper (:94-104).

Repairs: `--subsetFile` exists (bug 2: the smth driver reads args.subsetFile), `--gradCamType`,
`--maskPerturbType` and `--splitType` given on the command line override the config (bug 3: the drivers read them
from the config only, where they are missing), `--mod_stride_layers` defaults to "" (bug 5)."""
import argparse
import importlib.util
import os
import sys
from collections import OrderedDict

import torch

# (long flag, short flag, type or "flag", help) - the reference's option table
_OPTIONS = (
    ("config", "c", str, "config file path (a .py file defining `config = {...}`)"),
    ("eval_only", "e", "flag", "evaluate trained model on validation data"),
    ("resume", "r", "flag", "resume training from a given checkpoint"),
    ("gpus", "g", str, "GPU ids to use, comma separated"),
    ("use_cuda", None, "flag", "use GPUs"),
    ("iteration", "i", str, "suffix for model"),
    ("learning_rate", "lr", float, "initial lr"),
    ("batch_size", "bs", int, "batch size"),
    ("optimizer", "opt", str, "optimizer to use"),
    ("weight_decay", "wd", float, "weight decay"),
    ("shuffle", "sfl", int, "shuffle batch"),
    ("batch_norm", "bn", int, "use batch_norm or not"),
    ("subDir", "sd", str, "subdirectory to save figs to"),
    ("dataDir", "dd", str, "directory containing input data"),
    ("checkpoint", "chp", str, "checkpoint to read model from"),
    ("train", "tr", "flag", "use train data instead of validation"),
    ("lam1", "l1", float, "L1 loss coeff"),
    ("lam2", "l2", float, "TV loss coeff"),
    ("maskInitType", "mi", str, "how to find mask"),
    ("optIter", "opti", int, "iterations of the mask optimisation"),
    ("optRuns", "optr", int, "number of optimisation runs per mask"),
    ("classOI", "coi", int, "only consider this class"),
    ("clstm_hidden", "chu", int, "number of hidden units clstm"),
    ("clstm_layers", "chl", int, "number of hidden layers clstm"),
    ("conv_stride", "ccs", int, "conv stride in clstm"),
    ("final_temp_time", "ftt", int, "final temporal frame time in i3d"),
    ("last_stride", "ls", int, "stride on last pooling for I3D"),
    ("mod_stride_layers", "msl", str, "which I3D layers to modify stride on"),
    ("momentum", "mom", float, "optimizer momentum"),
    ("dropout", "drop", float, "dropout for CLSTM layers"),
    ("num_workers", "nwork", int, "number of workers for dataloader"),
    ("soft_max", "sm", int, "soft max at the end of I3D"),
    ("last_relu", "lact", str, "activation function at the last logit of I3D"),
    ("use_sequence", "ues", int, "send the entire sequence from the last CLSTM to the FC"),
    ("gradCamType", "gct", str, "Grad-CAM / temporal mask on 'guessed' or correct classes"),
    ("splitType", "kths", str, "kth split type"),
    # repairs
    ("subsetFile", "sf", str, "csv of clips of interest: one column per class, video ids below (smth driver)"),
    ("maskPerturbType", "mpt", str, "temporal perturbation of the mask search: freeze | reverse"),
    ("microBatch", "mb", int, "clips per launch sequence of the batched native search (default 8)"),
    ("fp32", None, "flag", "run the 1e-4 fp32 kernels instead of the bf16 tensor-core path"),
    ("noViz", None, "flag", "skip the image/GIF output (result pickles and class-score files only)"),
)


def build_parser():
    parser = argparse.ArgumentParser(description="temporal-mask search + Grad-CAM (B200-native drop-in)")
    for long, short, kind, text in _OPTIONS:
        names = ["--" + long] + (["-" + short] if short else [])
        if kind == "flag":
            parser.add_argument(*names, action="store_true", help=text)
        else:
            parser.add_argument(*names, type=kind, help=text)
    parser.set_defaults(mod_stride_layers="", subDir="run", gpus="0", checkpoint="")
    return parser


def load_args(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else list(argv)
    if not argv:
        parser.print_help()
        sys.exit(1)
    return parser.parse_args(argv)


def remove_module_from_checkpoint_state_dict(state_dict):
    """Strip the 'module.' prefix nn.DataParallel adds to parameter names (only where present)."""
    return OrderedDict(((k[7:] if k.startswith("module.") else k), v) for k, v in state_dict.items())


def load_module(module_path_and_name):
    """Import a config module from a file path; returns the module (its `config` dict is what the drivers read)."""
    name = os.path.splitext(os.path.basename(module_path_and_name))[0]
    spec = importlib.util.spec_from_file_location(name, module_path_and_name)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def merged_config(args):
    """The config dict of args.config with the command-line overrides the drivers need merged in."""
    config = dict(load_module(args.config).config)
    for key in ("gradCamType", "maskPerturbType", "splitType", "batch_size", "num_workers", "clstm_hidden",
                "clstm_layers", "conv_stride", "dropout"):
        val = getattr(args, key, None)
        if val is not None:
            config[key] = val
    config.setdefault("gradCamType", "guessed")
    config.setdefault("maskPerturbType", "freeze")
    config.setdefault("splitType", "original")
    if args.dataDir:
        config["data_folder"] = args.dataDir
    return config


def setup_cuda_devices(args):
    """(device, device_ids) from --use_cuda / --gpus.  The native path has no CPU fallback: without --use_cuda the
    drivers stop with a clear message instead of running on the host."""
    device = torch.device("cuda" if args.use_cuda else "cpu")
    device_ids = [int(i) for i in str(args.gpus).split(",")] if device.type == "cuda" else []
    if device.type == "cuda" and device_ids:
        device = torch.device("cuda", device_ids[0])
    return device, device_ids
