"""TEST INFRASTRUCTURE.  Recipe that stages the UNMODIFIED reference's hot-path modules under oracle/_ref/ so the
reference itself (not only the restatement in grl_oracle.py) can run on the GPU box, where /root/reference does not
exist.  oracle/_ref/ is git-ignored (never part of the history) but travels with the gpurun snapshot.

    python oracle/make_ref.py          # copies the files listed below, byte for byte, from /root/reference

Files: models/networks/grl.py, models/common/{mixed_attn_block_efficient, mixed_attn_block, ops,
swin_v1_block, upsample, resblock}.py with their package __init__.py files, and utils/utils_image.py (tensor_round /
shave / rgb2ycbcr used by the reference's PSNR).  Only bench.py's reference arm / cpu_baseline leg, tests and
make_golden*.py import it, through oracle/_ref_import.py (which also installs the stand-ins for timm / fairscale /
omegaconf helpers the model files import).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GRL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "models/__init__.py", "models/common/__init__.py", "models/networks/__init__.py",
    "models/networks/grl.py", "models/common/mixed_attn_block_efficient.py", "models/common/mixed_attn_block.py",
    "models/common/ops.py", "models/common/swin_v1_block.py", "models/common/upsample.py", "models/common/resblock.py",
    "utils/utils_image.py",
]


def stage():
    if not os.path.isfile(os.path.join(SRC, FILES[3])):
        return False
    digests = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            digests[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, sha256=digests), f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage()
    print("oracle/_ref staged from", SRC if ok else "(reference not present: nothing done)")
    sys.exit(0)
