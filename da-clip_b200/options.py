"""Option files of the restoration scripts (config/daclip-sde/options.py:18-120 of the reference): `parse(path,
is_train)` reads the YAML the reference's `test.py -opt options/test.yml` takes and normalises it the same way (dataset
phase / scale / distortion / data_type, expanded paths, results_root / log under the repository root), and
`dict_to_nonedict` wraps it so that missing keys read as None - the object `create_model(opt)`, `IRSDE(**opt["sde"])`
and the driver are written against.  Inference only: the training-side entries (experiments_root, resume checks) are
produced when `is_train` is set so that the dictionary has the reference's shape, nothing here consumes them.
"""
import os
import os.path as osp

import yaml


class NoneDict(dict):
    """dict whose missing keys read as None (options.py:111-113)."""

    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict({k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt


def parse(opt_path, is_train=False, root=None, set_visible_devices=False):
    """root: what `path.root` (and results_root / log below it) is derived from; the reference uses the checkout that holds
    its options.py, here the default is the current directory.  set_visible_devices: export CUDA_VISIBLE_DEVICES from
    gpu_ids like the reference does (off by default: the one-process-per-GPU launcher owns that variable)."""
    with open(opt_path) as f:
        opt = yaml.safe_load(f)
    if set_visible_devices:
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in opt["gpu_ids"])
    opt["is_train"] = is_train
    scale = 1
    if opt.get("distortion") == "sr":
        scale = opt["degradation"]["scale"]
        opt["network_G"]["setting"]["upscale"] = scale
    for phase, ds in (opt.get("datasets") or {}).items():
        ds["phase"] = phase.split("_")[0]
        ds["scale"] = scale
        ds["distortion"] = opt.get("distortion")
        lmdb = False
        for key in ("dataroot_GT", "dataroot_LQ"):
            if ds.get(key) is not None:
                ds[key] = osp.expanduser(ds[key])
                lmdb = lmdb or ds[key].endswith("lmdb")
        ds["data_type"] = "lmdb" if lmdb else "img"
        if ds["mode"].endswith("mc"):
            ds["data_type"] = "mc"
            ds["mode"] = ds["mode"].replace("_mc", "")
    paths = opt.setdefault("path", {})
    for key, p in list(paths.items()):
        if p and key != "strict_load":
            paths[key] = osp.expanduser(p)
    paths["root"] = osp.abspath(root or os.getcwd())
    config_dir = "daclip-sde"
    if is_train:
        exp = osp.join(paths["root"], "experiments", config_dir, opt["name"])
        paths.update(experiments_root=exp, models=osp.join(exp, "models"),
                     training_state=osp.join(exp, "training_state"), log=exp, val_images=osp.join(exp, "val_images"))
    else:
        res = osp.join(paths["root"], "results", config_dir)
        paths["results_root"] = osp.join(res, opt["name"])
        paths["log"] = osp.join(res, opt["name"])
    return opt
