"""Recipe for oracle/_ref: the reference's OWN stage-2 modules, staged so they exist on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing in compress-robust-vqa_b200/ imports it.

The reference is pure Python (no C/C++ to compile), so "building" it means staging the files its stage-2 path
imports.  This script imports that path from the read-only reference tree under the shims of
tests/golden/ref_shims.py, asks the interpreter which files under the tree were loaded (31 files, 0.8 MB:
masking/maskers*.py, masking/sparsity_control.py, hg_transformers/modeling_lxmert.py and what it pulls in,
hg_transformers/mask_trainer_{VQA,Robust_VQA}.py, vqa_debias_loss_functions.py, classifier.py, root optimization.py)
and copies exactly those, unmodified, to oracle/_ref/ with a MANIFEST.json of their SHA-256.  oracle/_ref/ is listed
in .gitignore (reference sources never enter the history) and NOT in .gpurunignore, so the copy travels to the GPU box
like a built .so.  Consumers: bench.py --impl reference (kind "reference") and --impl torch-gpu (the same-box bar,
SURVEY.md section 8(d)), both through oracle/ref_runner.py.

    python oracle/build_ref.py            # no-op with a message when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_ROOT = os.environ.get("CRVQA_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

_PROBE = r"""
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "tests", "golden"))
import make_golden as mg
mg.load_reference()
import importlib
importlib.import_module("hg_transformers.modeling_visualbert")
importlib.import_module("hg_transformers.configuration_visualbert")
importlib.import_module("hg_transformers.mask_trainer_visualBERT_VQA")
ref = os.path.realpath(%(ref)r) + os.sep
files = sorted({os.path.relpath(os.path.realpath(m.__file__), ref) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(ref)})
ns = sorted({os.path.relpath(os.path.realpath(p), ref) for m in list(sys.modules.values())
             if getattr(m, "__file__", None) is None and hasattr(m, "__path__")
             for p in list(m.__path__) if os.path.realpath(p).startswith(ref)})
print("FILES=" + json.dumps(files))
print("NSDIRS=" + json.dumps(ns))
"""


def build(quiet=False):
    if not os.path.isdir(os.path.join(REF_ROOT, "masking")):
        if not quiet:
            print(f"oracle/_ref: reference tree not found at {REF_ROOT}; keeping whatever is staged")
        return os.path.isfile(os.path.join(OUT, "MANIFEST.json"))
    env = dict(os.environ, CRVQA_REFERENCE_ROOT=REF_ROOT, WANDB_MODE="disabled", WANDB_SILENT="true")
    res = subprocess.run([sys.executable, "-c", _PROBE % {"root": ROOT, "ref": REF_ROOT}], env=env,
                         capture_output=True, text=True)
    line = [l for l in res.stdout.splitlines() if l.startswith("FILES=")]
    if res.returncode != 0 or not line:
        raise RuntimeError("could not import the reference's stage-2 path:\n" + res.stderr[-2000:])
    files = json.loads(line[0][6:])
    nsdirs = json.loads([l for l in res.stdout.splitlines() if l.startswith("NSDIRS=")][0][7:])
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    manifest = {"reference_root": REF_ROOT, "files": {}}
    for rel in files:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    for rel in nsdirs:          # namespace packages (no __init__.py): `import utils` only needs the directory
        os.makedirs(os.path.join(OUT, rel), exist_ok=True)
        open(os.path.join(OUT, rel, ".keep"), "w").close()
    manifest["namespace_dirs"] = nsdirs
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print(f"oracle/_ref: staged {len(files)} unmodified reference files from {REF_ROOT}")
    return True


if __name__ == "__main__":
    build()
