#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference Python package, copied from /root/reference so that it can travel to the
GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).  Test / measurement infrastructure only:

  * `bench.py --impl reference` times the reference's own get_prediction (src/eval_prepare_model.py:118-121) on the host
    cores (cpu_baseline.kind = "reference") and, with --ref-device cuda, its eager PyTorch path on the B200;
  * nothing in skeletondiffusion_b200/ imports it.

Only the importable package `src/` (pure Python, no build step) is copied, byte for byte; the three packages it imports that are
not installed here (denoising_diffusion_pytorch, hydra, omegaconf) are stubbed by tests/golden/_stubs.

    python oracle/make_ref.py            # in the build container, where /root/reference exists
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SKELDIFF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make(verbose: bool = True) -> bool:
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        if verbose:
            print(f"make_ref: {src} not found (GPU box?): keeping whatever oracle/_ref/ holds")
        return os.path.isdir(os.path.join(DST, "src"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n, h = 0, hashlib.sha256()
    for root, _dirs, files in sorted(os.walk(src)):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, os.path.relpath(s, REF))
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            h.update(open(s, "rb").read())
            n += 1
    with open(os.path.join(DST, "MANIFEST"), "w") as fh:
        fh.write(f"copied {n} .py files of {src} unmodified; sha256 of their concatenation {h.hexdigest()}\n")
    if verbose:
        print(f"make_ref: copied {n} files into {DST}")
    return True


def add_to_path() -> bool:
    """Put oracle/_ref and the import stubs on sys.path; False when the reference copy is absent."""
    if not os.path.isdir(os.path.join(DST, "src")):
        return False
    stubs = os.path.join(os.path.dirname(HERE), "tests", "golden", "_stubs")
    for p in (DST, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
