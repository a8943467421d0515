"""Snapshot of the UNMODIFIED reference sources needed to RUN its hot path, written to oracle/_ref/.

TEST / BENCH INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container but not on the GPU box; the
reference is pure Python, so the files its `models.py` imports (pytorch/*.py, pytorch/models_2020/**, utils/*.py,
utils/gammatone/**) are copied verbatim to `oracle/_ref/` -- a directory that is git-ignored (no reference source ever
enters the repository history) but travels to the GPU box with the working tree, like the built `.so`.  There
`bench.py --impl reference` and the `cpu_baseline` leg import the real `Cnn_9layers_*` modules through
`oracle/ref_import.py` (same shims as here) and time THEM on the host cores (`cpu_baseline.kind = "reference"`);
without the snapshot they fall back to the oracle port (`kind = "port"`).

    python oracle/snapshot_ref.py            # run by __graft_entry__.build() whenever /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SED_REFERENCE_SOURCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
WANTED = (("pytorch", (".py",)), ("utils", (".py",)))


def snapshot():
    if not os.path.isfile(os.path.join(SRC, "pytorch", "models.py")):
        return None
    n = 0
    for top, exts in WANTED:
        for root, dirs, files in os.walk(os.path.join(SRC, top)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            rel = os.path.relpath(root, SRC)
            for f in files:
                if f.endswith(exts):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                    n += 1
    with open(os.path.join(DST, "SNAPSHOT.txt"), "w") as fh:
        fh.write("verbatim copy of %d reference source files from %s (not part of the repository)\n" % (n, SRC))
    return n


if __name__ == "__main__":
    n = snapshot()
    print("reference tree not present; nothing copied" if n is None else "copied %d files to %s" % (n, DST))
    sys.exit(0)
