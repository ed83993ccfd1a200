"""TEST / BASELINE INFRASTRUCTURE -- recipe that stages the UNMODIFIED reference next to the oracle.

    python oracle/vendor_ref.py            # /root/reference -> oracle/_ref   (git-ignored, NOT gpurun-ignored)

The GPU box receives /root/repo only, so the reference arm of bench.py (``--impl reference``: the reference's own
``LatentDynamics.predict`` on the host cores, and the same-box PyTorch-eager-on-B200 bar) needs a copy that travels with the
snapshot.  ``oracle/_ref`` is listed in .gitignore: no reference source ever enters the history; this script and
``oracle/ref_shims.py`` (the five import shims, no file of the reference is edited) are what is committed.
``__graft_entry__.build()`` runs it whenever /root/reference is present."""
import os
import shutil
import sys

SRC = os.environ.get("LNS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
KEEP_EXT = (".py", ".yml", ".yaml", ".md", "")  # sources, configs, README / LICENSE; no .pyc, no assets


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "modules")):
        if verbose:
            print(f"vendor_ref: {SRC} not present; keeping {DST} as it is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(os.path.join(DST, "modules"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "assets", ".git")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if os.path.splitext(f)[1] not in KEEP_EXT or f.startswith("."):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            shutil.copy2(os.path.join(root, f), os.path.join(DST, rel, f))
            n += 1
    if verbose:
        print(f"vendor_ref: staged {n} files of the unmodified reference under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
