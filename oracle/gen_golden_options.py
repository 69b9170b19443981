"""Generates tests/golden/options_test.yml (the reference's inference option file, byte for byte) and
tests/golden/options_test.json (what the REFERENCE's own options.parse + dict_to_nonedict make of it, with the
checkout-dependent path entries dropped).  Build container only; outputs are committed.  TEST INFRASTRUCTURE.

    python oracle/gen_golden_options.py
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/universal-image-restoration"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(REF, "config", "daclip-sde"))
sys.path.insert(0, REF)


def main():
    src = os.path.join(REF, "config", "daclip-sde", "options", "test.yml")
    shutil.copyfile(src, os.path.join(GOLD, "options_test.yml"))
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    import options as ref_options                    # the reference's config/daclip-sde/options.py
    opt = ref_options.dict_to_nonedict(ref_options.parse(src, is_train=False))
    if saved is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = saved
    plain = json.loads(json.dumps(opt))
    for k in ("root", "results_root", "log"):        # depend on where the reference checkout lives
        plain["path"].pop(k)
    missing = {"suffix": opt["suffix"], "crop_border": opt["crop_border"], "no_such_key": opt["no_such_key"],
               "path.strict_load": opt["path"]["strict_load"]}
    with open(os.path.join(GOLD, "options_test.json"), "w") as f:
        json.dump({"opt": plain, "missing_keys_read_as": missing}, f, indent=1, sort_keys=True)
    print("wrote options_test.yml / options_test.json")


if __name__ == "__main__":
    main()
