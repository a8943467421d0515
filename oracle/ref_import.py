"""Import the UNMODIFIED reference (`/root/reference/pytorch/{stft,models}.py`) in the build container.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  `/root/reference` exists only in
the build container (not on the GPU box), so this module is used solely by `oracle/gen_golden.py`
and by the `not gpu` tests that pin `oracle/sed_oracle.py` against the real reference (they skip
when the reference tree is absent).  Recipe from SURVEY.md section 8(c).
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "_shims")


def _find_root():
    """The reference tree where it lies (build container), else the verbatim snapshot oracle/snapshot_ref.py wrote
    to oracle/_ref (git-ignored; travels to the GPU box with the working tree)."""
    for root in (os.environ.get("SED_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if root and os.path.isfile(os.path.join(root, "pytorch", "models.py")):
            return root
    return "/root/reference"


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pytorch", "models.py"))


def load():
    """Returns (ref_stft_module, ref_models_module)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("models", "stft", "librosa", "matplotlib")}
    try:
        for k in saved_mods:
            sys.modules.pop(k, None)
        sys.path[:0] = [_SHIMS, os.path.join(REF_ROOT, "pytorch"), os.path.join(REF_ROOT, "utils")]
        ref_stft = importlib.import_module("stft")
        ref_models = importlib.import_module("models")
    finally:
        sys.path[:] = saved_path
    # keep the reference modules reachable only through the returned handles
    for k in ("models", "stft"):
        sys.modules.pop(k, None)
        if saved_mods[k] is not None:
            sys.modules[k] = saved_mods[k]
    return ref_stft, ref_models


def load_utilities():
    """Returns the reference's utils/utilities.py module (merge, avg_merge, ...) and utils/vad.py."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    saved_path = list(sys.path)
    names = ("utilities", "vad", "config", "librosa", "matplotlib", "sed_eval", "h5py", "prettytable")
    saved_mods = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            sys.modules.pop(k, None)
        sys.path[:0] = [_SHIMS, os.path.join(REF_ROOT, "utils")]
        ref_utilities = importlib.import_module("utilities")
        ref_vad = importlib.import_module("vad")
    finally:
        sys.path[:] = saved_path
    for k in names:
        sys.modules.pop(k, None)
        if saved_mods[k] is not None:
            sys.modules[k] = saved_mods[k]
    return ref_utilities, ref_vad
