"""Per-kernel headline metrics and per-source-line stall samples of an .ncu-rep (needs ncu on PATH).  Dev tool.
usage: ncu_lines.py report.ncu-rep            -> one line per captured kernel
       ncu_lines.py report.ncu-rep IDX [MIN]  -> CUDA source lines of kernel IDX with >= MIN stall samples"""
import csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout

def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0

if len(sys.argv) == 2:
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, data = rows[0], rows[2:]
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread"]
    idx = [hdr.index(w) for w in want if w in hdr]
    print([hdr[i].split(".")[0] for i in idx])
    for k, r in enumerate(data):
        print(k, [r[i][:44] for i in idx])
else:
    k = int(sys.argv[2]); lo = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    rows = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k),
                                            "--launch-count", "1"]))))
    cur, hdr, tot = None, None, 0
    out = []
    for r in rows:
        if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            out.append((cur, r))
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[4]) for f, r in out if f and f.endswith(".cu"))
    print("samples on .cu lines:", tot)
    for f, r in out:
        if num(r[4]) >= lo:
            st = sorted(((int(num(r[i])), hdr[i][6:]) for i in stall), reverse=True)[:3]
            print(f[:14].ljust(14), r[0].rjust(4), str(int(num(r[4]))).rjust(5), r[7].rjust(9), r[1].strip()[:90].ljust(90), st)
