"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Restates the reference's algorithms (haan6/fm-for-online-recommendation) on the CPU so the CUDA
product can be checked against them.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
package ``fm_for_online_recommendation_b200`` never does (tests/test_boundary.py enforces it).
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    """Compile oracle/libfm_oracle.so with gcc (idempotent). Returns the path."""
    so = os.path.join(_HERE, "libfm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fm_oracle.c", "oracle_math.h", "rsqrt14_table.inc")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfm_oracle.so"], stdout=subprocess.DEVNULL)
    return so
