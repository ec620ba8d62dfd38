#!/usr/bin/env python
"""Extract the per-launch counters bench.py cross-references from one `ncu --set full --import-source on` report of the
lean kernel: DRAM traffic, warp instructions, fp64-pipe instructions.  Stamped with the hash of the kernel sources the
capture was taken from, so that bench.py can tell a stale file from a current one.
usage: ncu_counters.py <report.ncu-rep> <workload> <branches_per_launch> [out.json]"""
import csv, hashlib, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "gammaspikemodel_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inc")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    rep, workload, branches = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "kernel_counters.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    col = lambda name: vals[hdr.index(name)]
    unit = rows[1]
    def bytes_of(name):
        v, u = float(col(name).replace(",", "")), unit[hdr.index(name)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h2 = rows[hi]
    ci, si = h2.index("Instructions Executed"), h2.index("Source")
    tot = f64 = 0
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[ci].isdigit():
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip()).split()[0].split(".")[0]
        n = int(r[ci])
        tot += n
        if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
            f64 += n
    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    d = {"workload": workload, "kernel": col("Kernel Name") if "Kernel Name" in hdr else "gsm_lean_kernel",
         "csrc_sha": csrc_sha(), "branches_per_launch": branches,
         "dram_bytes_per_launch": rd + wr, "dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr,
         "warp_instructions_per_launch": tot, "fp64_warp_instructions_per_launch": f64,
         "thread_instr_per_branch": tot * 32 / branches, "fp64_thread_instr_per_branch": f64 * 32 / branches,
         "fp64_pipe_active_pct": float(col("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
         "issue_active_pct": float(col("smsp__issue_active.avg.pct_of_peak_sustained_active")),
         "gpu_time_ms": float(col("gpu__time_duration.sum")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit[hdr.index("gpu__time_duration.sum")].lower(), 1),
         "source": "ncu --set full --import-source on --clock-control none, %s" % os.path.basename(rep)}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))
