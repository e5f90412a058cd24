"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference implementation of the hot path
where the GPU box can find it.

    python -m oracle.build_ref        (also run by __graft_entry__.build() when /root/reference is present)

The reference is one Python file (`/root/reference/srdiff/model.py`) plus its `config.yaml`; "building" it means
copying those two files, byte for byte, into `oracle/_ref/srdiff/` -- a build output like a compiled `.so`:
git-ignored (never part of the history), not gpurun-ignored (it travels to the GPU box with the snapshot).
`bench.py --impl reference` / the `cpu_baseline` leg then time the real reference (`kind: "reference"`) instead of
the oracle port.  Nothing under the product package ever imports it.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SRDIFF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ("srdiff/model.py", "srdiff/config.yaml")


def build(verbose=True):
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        if verbose:
            print("oracle.build_ref: %s not present, nothing to do" % SRC)
        return False
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), d)
        if verbose:
            print("oracle/_ref/%s  sha256 %s" % (f, hashlib.sha256(open(d, "rb").read()).hexdigest()[:16]))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 0)
