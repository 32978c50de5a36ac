"""TEST INFRASTRUCTURE ONLY - Python (ctypes) front end of the two CPU checkers.

* :mod:`oracle.restatement` wraps ``oracle/libverticut_oracle.so`` - the plain-C
  restatement of the reference algorithm (``oracle/verticut_oracle.c``).
* :mod:`oracle.reference` wraps ``oracle/_ref/libverticut_ref.so`` - the
  reference's own sources compiled unmodified (``oracle/ref_driver.cc``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
``verticut_b200`` must never do so.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libverticut_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libverticut_ref.so")
REFERENCE_ROOT = os.environ.get("VERTICUT_REFERENCE", "/root/reference")


def build(verbose=False):
    """Compile the restatement, and the reference build when the reference tree is present."""
    cmd = ["make", "-C", HERE, "REF=" + REFERENCE_ROOT, "all"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref():
    return os.path.exists(REF_SO)
