"""Measured-error log of the parity tests (TEST INFRASTRUCTURE).

``check(err, gate, what)`` asserts ``err < gate`` like a bare assert would, and also records the measured value next to
its gate under the running test's id; ``dump(path)`` writes the table (tests/conftest.py does so at the end of a session
that recorded anything, into gpurun_out/parity_r2.md, from where it is copied to profiles/).  This is how a reader sees how
close each stage sits to its tolerance instead of only "passed"."""
from __future__ import annotations

import os

RECORDS = []


def _test_id():
    t = os.environ.get("PYTEST_CURRENT_TEST", "?")
    return t.split(" ")[0].replace("tests/", "")


def check(err, gate, what=""):
    err = float(err)
    RECORDS.append((_test_id(), what, err, float(gate), "<"))
    assert err < gate, (what, err, gate)


def check_min(val, floor, what=""):
    val = float(val)
    RECORDS.append((_test_id(), what, val, float(floor), ">="))
    assert val >= floor, (what, val, floor)


def dump(path):
    if not RECORDS:
        return
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        f.write("| test | quantity | measured | gate | margin |\n|---|---|---|---|---|\n")
        for tid, what, v, g, op in RECORDS:
            if op == "<":
                margin = f"{g / v:.1f}x" if v > 0 else "inf"
                f.write(f"| `{tid}` | {what} | {v:.3e} | < {g:.1e} | {margin} |\n")
            else:
                f.write(f"| `{tid}` | {what} | {v:.2f} | >= {g:.1f} | +{v - g:.1f} |\n")
