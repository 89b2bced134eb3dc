#!/usr/bin/env python3
"""Recipe for oracle/_ref/: the UNMODIFIED reference script, where bench.py's CPU legs can run it on the GPU box.

    python oracle/fetch_ref.py            (also called by __graft_entry__.build() when /root/reference exists)

The reference is one pure-Python file; there is nothing to compile.  It is copied byte for byte from where it lies
under /root/reference into oracle/_ref/ - git-ignored (reference sources never enter the history), not gpurun-ignored
(so it travels to the GPU box like the built .so files) - after its sha256 has been checked against the one SURVEY.md
section 8c records.  Test infrastructure: only bench.py's `cpu_baseline` / `--impl reference` legs import it, as the thing
that is timed on the host cores, never on the product path.  When oracle/_ref/ is absent those legs time the oracle's
literal-loop port instead and say so (kind "port")."""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference/src/TrigenicInteractionPredictor.py"
SHA256 = "898269709cbfcb2dd5ec01d501224ff63c2a6676406f8d15779af0f531f04290"
HERE = os.path.dirname(os.path.abspath(__file__))
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "TrigenicInteractionPredictor.py")


def fetch() -> str | None:
    """Returns the path of the copy, or None when the reference tree is not present (the GPU box)."""
    if not os.path.exists(SRC):
        return DST if os.path.exists(DST) else None
    with open(SRC, "rb") as fh:
        digest = hashlib.sha256(fh.read()).hexdigest()
    if digest != SHA256:
        raise RuntimeError("reference script changed: sha256 %s, expected %s" % (digest, SHA256))
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, DST)
    return DST


def load_reference_module():
    """Import the copied reference (or None).  The module's __main__ guard keeps the import free of side effects."""
    if not os.path.exists(DST):
        return None
    import importlib.util
    with open(DST, "rb") as fh:
        if hashlib.sha256(fh.read()).hexdigest() != SHA256:
            return None
    spec = importlib.util.spec_from_file_location("TrigenicInteractionPredictor_reference", DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(fetch())
    sys.exit(0)
