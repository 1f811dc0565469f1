"""Per-kernel SASS opcode histogram of libtocvp.so (no GPU needed): python tools/sass_hist.py > profiles/sass_r2.txt

Shows, for every kernel in the library, the instruction count and the Blackwell-specific opcodes that prove which hardware
path it uses (B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma (".2CTA" = cta_group::2),
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copies, SYNCS = mbarrier,
HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async, ACQBULK / PREEXIT = programmatic dependent launch."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "textocvp_b200", "libtocvp.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "F2FP.SATFINITE",
        "ACQBULK", "PREEXIT", "UTCBAR", "UCGABAR"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            kernels[cur][base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("F2FP") and "SATFINITE" in op:
                kernels[cur]["F2FP.SATFINITE"] += 1
    names = demangle(list(kernels))
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), {len(kernels)} kernels")
    print("# columns: instructions | " + " ".join(KEYS))
    tot = collections.Counter()
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("tocvp::", "")
        cols = " ".join(f"{key}={c[key]}" for key in KEYS if c[key])
        print(f"{short:<60s} {c['_total']:>6d} | {cols}")
        for key in KEYS:
            tot[key] += c[key]
    print("# totals: " + " ".join(f"{key}={tot[key]}" for key in KEYS))


if __name__ == "__main__":
    sys.exit(main())
