"""Loader of oracle/_ref/ (the UNMODIFIED reference modules copied by oracle/make_ref.sh).

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.  The reference
modules are flat top-level scripts (`import config`, `import utils`), so they are imported with oracle/_ref at the
front of sys.path and the path is restored afterwards; `config.DEVICE` is forced to "cpu" because
utils.create_padding_mask moves its mask to that global device (reference utils.py:70)."""
import hashlib
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("decoder.py", "utils.py", "config.py")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def verify_manifest() -> bool:
    """True iff every file still has the sha256 recorded when it was copied from /root/reference."""
    man = os.path.join(REF_DIR, "MANIFEST.sha256")
    if not (available() and os.path.exists(man)):
        return False
    want = dict(reversed(line.split()) for line in open(man) if line.strip())
    for f in FILES:
        with open(os.path.join(REF_DIR, f), "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != want.get(f):
                return False
    return True


def load():
    """(decoder module, config module) of the unmodified reference, or None when oracle/_ref is absent."""
    if not available():
        return None
    saved = {k: sys.modules.get(k) for k in ("config", "utils", "decoder")}
    sys.path.insert(0, REF_DIR)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        cfg = importlib.import_module("config")
        cfg.DEVICE = "cpu"
        dec = importlib.import_module("decoder")
        return dec, cfg
    finally:
        sys.path.remove(REF_DIR)
        for k, v in saved.items():          # leave the caller's own top-level modules of the same names alone
            if v is not None:
                sys.modules[k] = v
