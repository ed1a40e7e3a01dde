"""Stage the UNMODIFIED reference for the CPU baseline arm (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

The reference is pure Python over ATen: nothing to compile.  ``/root/reference`` does not exist on the GPU box, so
``bench.py --impl reference`` cannot import it from there.  This recipe copies the three package directories the two
hot paths live in -- byte for byte, no edits --

    human_diffusion/improved_diffusion/   (UNetModel, GaussianDiffusion, SpacedDiffusion, script_util)
    human_diffusion/NeRF/                 (Renderer, fields)
    recon_NeRF/lib/                       (Renderer of the reconstruction side, if_nerf_data_utils)

into ``oracle/_ref/`` -- which is git-ignored (the reference's sources never enter this repository's history) but
NOT gpurun-ignored, so it travels with the snapshot exactly like the built ``.so`` files.  A manifest with the
sha256 of every staged file is written next to them; ``verify()`` re-hashes, so a modified copy is detected.

    python -m oracle.build_ref            # run by __graft_entry__.build() when /root/reference is present

``bench.py --impl reference`` then imports it through ``oracle/ref_shims.py`` with ``HUMANLIFF_REF=oracle/_ref``
(``cpu_baseline.kind == "reference"``); if ``oracle/_ref`` is absent it falls back to the oracle port and says so.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("HUMANLIFF_REF_SRC", "/root/reference")
PACKAGES = ("human_diffusion/improved_diffusion", "human_diffusion/NeRF", "recon_NeRF/lib")
MANIFEST = os.path.join(DEST, "MANIFEST.json")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(src=SRC, dest=DEST):
    """Copy the packages (``*.py`` only) and write the manifest.  Returns the manifest dict, or None when the
    reference tree is not present (the GPU box: the staged copy shipped with the snapshot is used as is)."""
    if not os.path.isdir(os.path.join(src, PACKAGES[0])):
        return None
    files = {}
    for pkg in PACKAGES:
        for root, _, names in os.walk(os.path.join(src, pkg)):
            for n in sorted(names):
                if not n.endswith(".py"):
                    continue
                sp = os.path.join(root, n)
                rel = os.path.relpath(sp, src)
                dp = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(dp), exist_ok=True)
                shutil.copyfile(sp, dp)
                files[rel] = _sha(dp)
    man = {"source": src, "packages": list(PACKAGES), "files": files}
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    return man


def available(dest=DEST):
    return os.path.exists(os.path.join(dest, "MANIFEST.json"))


def verify(dest=DEST):
    """True iff every staged file still hashes to its manifest entry (the copy is unmodified)."""
    if not available(dest):
        return False
    man = json.load(open(os.path.join(dest, "MANIFEST.json")))
    return all(os.path.exists(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == h
               for rel, h in man["files"].items())


if __name__ == "__main__":
    m = stage()
    if m is None:
        print("reference tree not found at", SRC, "-- nothing staged", file=sys.stderr)
        sys.exit(0 if available() else 1)
    print("staged %d files into %s" % (len(m["files"]), DEST))
