"""Copies the reference's own hot-path modules, UNMODIFIED, from /root/reference into the git-ignored baseline/_ref/
so that they travel to the GPU box (gpurun ships the repo snapshot, /root/reference does not exist there) and
`bench.py --impl reference` / the `cpu_baseline` leg can time THE REFERENCE ITSELF on the host cores (kind "reference")
instead of the oracle port.  TEST / BASELINE INFRASTRUCTURE - see oracle/__init__.py: nothing under da-clip_b200/ imports
these files, and they never enter git history (baseline/_ref/ is in .gitignore).

    python oracle/vendor_reference.py          # run in the build container; __graft_entry__.build() calls it too

Files: utils/sde_utils.py (IRSDE), config/daclip-sde/models/modules/{DenoisingUNet_arch,
module_util,attention}.py (ConditionalUNet), and the open_clip/ package the inference path imports (DaCLIP, tokenizer,
model configs).  `load_reference()` imports them without the reference's package __init__ files, which pull in cv2 /
lmdb / lpips that the hot path does not need.
"""
import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/universal-image-restoration"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [("utils/sde_utils.py", "sde_utils.py"),
         ("config/daclip-sde/models/modules/DenoisingUNet_arch.py", "sde_modules/DenoisingUNet_arch.py"),
         ("config/daclip-sde/models/modules/module_util.py", "sde_modules/module_util.py"),
         ("config/daclip-sde/models/modules/attention.py", "sde_modules/attention.py")]


def vendor():
    if not os.path.isdir(SRC):
        return False
    for rel, out in FILES:
        dst = os.path.join(DST, out)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    shutil.copytree(os.path.join(SRC, "open_clip"), os.path.join(DST, "open_clip"), dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns("__pycache__"))
    return True


def available():
    return all(os.path.exists(os.path.join(DST, out)) for _, out in FILES)


def load_reference():
    """(IRSDE class, ConditionalUNet class) of the vendored reference files."""
    if not available():
        raise FileNotFoundError(f"{DST} is empty: run oracle/vendor_reference.py where /root/reference exists")
    spec = importlib.util.spec_from_file_location("_ref_sde_utils", os.path.join(DST, "sde_utils.py"))
    sde_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sde_utils)
    pkg = types.ModuleType("_ref_sde_modules")             # a bare package: the files import each other relatively
    pkg.__path__ = [os.path.join(DST, "sde_modules")]
    sys.modules["_ref_sde_modules"] = pkg
    arch = importlib.import_module("_ref_sde_modules.DenoisingUNet_arch")
    return sde_utils.IRSDE, arch.ConditionalUNet


def load_open_clip():
    """The vendored open_clip package (needs the `ftfy` stub: only fix_text is used, identity for ASCII prompts)."""
    if not os.path.isdir(os.path.join(DST, "open_clip")):
        raise FileNotFoundError(f"{DST}/open_clip is missing")
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    if DST not in sys.path:
        sys.path.insert(0, DST)
    return importlib.import_module("open_clip")


if __name__ == "__main__":
    print("vendored" if vendor() else "no /root/reference here: nothing copied", "->", DST)
