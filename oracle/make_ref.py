"""Vendor the UNMODIFIED reference Python tree into oracle/_ref/  --  TEST / BASELINE INFRASTRUCTURE.

    python oracle/make_ref.py            # needs /root/reference (the build container); idempotent

The reference is plain Python (no setup.py, nothing to compile), and /root/reference does not exist on the GPU box.
``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the copy travels
with the snapshot.  It is used only as
  * the reference arm of bench.py (``--impl reference``, ``cpu_baseline.kind == "reference"``): the reference's own
    modules timed on the box's host cores, and
  * the checker of the drop-in tests (tests/test_dropin_*.py): the reference's own ``lib/core/function.py`` driving the
    mirror, a checkpoint written by the reference modules loaded into the mirror.
Nothing under vae-2_b200/ reads it.  Files are byte-identical copies; MANIFEST lists their sha256.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VAE2_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "lib")):
        print("make_ref: %s not present; keeping %s as it is" % (REF, DST))
        return 0 if os.path.isdir(os.path.join(DST, "lib")) else 1
    lines = []
    for sub in ("lib", "tools"):
        for dirpath, dirnames, filenames in os.walk(os.path.join(REF, sub)):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "sync_bn")]   # sync_bn: dead code, JIT C++
            for f in filenames:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(dirpath, f)
                rel = os.path.relpath(src, REF)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                lines.append("%s  %s" % (hashlib.sha256(open(src, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST"), "w") as fh:
        fh.write("\n".join(sorted(lines)) + "\n")
    print("make_ref: %d files -> %s" % (len(lines), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
