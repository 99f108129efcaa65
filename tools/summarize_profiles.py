"""Turns the ncu artefacts tools/profile_r0N.sh left in gpurun_out/ into profiles/r0N_ncu_summary_<tag>.md.
   python tools/summarize_profiles.py v1 [r02]   (needs `ncu` on PATH to read the .ncu-rep files)"""
import collections, csv, glob, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "v4"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
PFX = "" if rnd == "r01" else "r2_"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum"]


def launch_table(path, title):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if "asvgp" not in r[ki]:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["## %s; asvgp kernels only\n" % title, "| kernel | launches | total ms | avg us | share of asvgp time |", "|---|---|---|---|---|"]
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.2f | %.1f %% |" % (name, c, ns / 1e6, ns / c / 1e3, 100 * ns / tot))
    return "\n".join(out) + "\n"


def rep_row(path):
    # a capture too large to bring back from the GPU box is exported there: `ncu -i X.ncu-rep --page raw --csv > X.raw.csv`
    if path.endswith(".raw.csv"):
        txt = open(path).read()
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    cells = []
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            cells.append("%s %s" % (vals[i], units[i]))
        else:
            cells.append("n/a")
    return "| `%s` | %s |" % (os.path.basename(path)[5:].replace(".ncu-rep", "").replace(".raw.csv", ""), " | ".join(cells))


parts = ["# %s %s — ncu evidence (B200, `--clock-control none`)\n" % (rnd, tag),
         "Commands: `tools/profile_%s.sh a|b|c`" % rnd + " (launch lists: `ncu --metrics gpu__time_duration.sum`; kernels: `ncu --set full`, one "
         "launch each, after a plain run of the same command that exited 0); this file: `tools/summarize_profiles.py %s`.\n"
         "Per-launch times are cold-cache and serialised (compare shares, not absolutes). The bench line of the same build: "
         "`profiles/%s_bench_*.json`.\n" % (tag, rnd),
         "ncu serialises the launches: in a real 1-D step the Kuu chain (the 8-CTA launch of each `elbo_chains_cluster_kernel` pair, "
         "94 us here) runs on a side stream BESIDE `accum_1d_kernel`, so the step's critical path is accumulate + `chain_rows_kernel<.., 2>` "
         "+ the 16-CTA P-chain launch (51 us here) + `elbo_finalize_kernel`: accumulate share 267 / (267 + 10 + 51 + 8) = 79 % of that "
         "path against 0.262 / 0.339 = 77 % of the step in the bench's CUDA-event phases.\n"]
for f, title in ((PFX + "launches_1d.csv", "1-D bench (`bench.py --steps 2 --warmup 3 --no-2d`), first 400 launches"),
                 (PFX + "launches_2d.csv", "2-D bench (`bench.py --workload 2d --steps 1 --warmup 3`), first 600 / 900 launches"),
                 (PFX + "launches_binned_1d.csv", "1-D accumulate, 1e8 points in random order (`tools/binned_1d_only.py`)"),
                 (PFX + "launches_binned_2d.csv", "2-D accumulate, shuffled 1e4 x 1e4 raster (`tools/binned_2d_only.py`)")):
    p = os.path.join(OUT, f)
    if os.path.exists(p):
        parts.append(launch_table(p, title))
reps = sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep")) + glob.glob(os.path.join(OUT, "prof_*.raw.csv"))
              + (glob.glob("/tmp/reps/prof_*.ncu-rep") if rnd == "r01" else []))
if reps:
    parts.append("## `--set full` captures (one launch each)\n")
    parts.append("| kernel | " + " | ".join(m.split(".")[0] for m in METRICS) + " |")
    parts.append("|" + "---|" * (len(METRICS) + 1))
    for r in reps:
        parts.append(rep_row(r))
LIBSO = os.path.join(ROOT, "asvgp_b200", "lib", "libasvgp_sm100a.so")
PAT = r"UBLKCP[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|LDGSTS[.A-Z0-9]*|SYNCS[.A-Z0-9]*|DMMA[.A-Z0-9]*|LDG\.E\.ENL2\.256[.A-Z]*|REDG\.E\.ADD\.F64[.A-Z]*|ATOMS\.ADD|LD\.E\.[0-9]*\.STRONG\.SYS|ST\.E\.STRONG\.SYS|UCGABAR_[A-Z]*|ACQBULK|MAPA[.A-Z0-9]*"
sass = subprocess.run("cuobjdump -sass %s | grep -oE '%s' | sort | uniq -c" % (LIBSO, PAT), shell=True, capture_output=True, text=True).stdout
parts.append("\n## SASS evidence (`cuobjdump -sass asvgp_b200/lib/libasvgp_sm100a.so`)\n\n```\n%s```\n" % sass)
# the same mnemonics per kernel (only kernels that have any): tensor-core fp64 (DMMA), TMA bulk copies (UBLKCP) and their mbarriers
# (SYNCS), cp.async (LDGSTS), cluster barriers (UCGABAR_*), fp64 REDs
import re
dump = subprocess.run("cuobjdump -sass %s | c++filt" % LIBSO, shell=True, capture_output=True, text=True).stdout
per, cur = collections.OrderedDict(), None
for line in dump.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        cur = m.group(1).split("(")[0].replace("void ", "").replace("asvgp::", "")
        continue
    if cur is None:
        continue
    for tok in re.findall(PAT, line):
        key = tok.split(".")[0] if not tok.startswith("REDG") else "REDG.F64"
        per.setdefault(cur, collections.Counter())[key] += 1
keys = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UCGABAR_ARV", "UCGABAR_WAIT", "REDG.F64", "ATOMS"]
rows = ["| kernel | " + " | ".join(keys) + " |", "|---|" + "---|" * len(keys)]
for name, c in per.items():
    if any(c.get(k, 0) for k in keys[:6]):
        rows.append("| `%s` | %s |" % (name[:90], " | ".join(str(c.get(k, 0)) for k in keys)))
parts.append("Per kernel (kernels with tensor-core, TMA, cp.async or cluster-barrier instructions):\n\n" + "\n".join(rows) + "\n")
dst = os.path.join(ROOT, "profiles", "%s_ncu_summary_%s.md" % (rnd, tag))
open(dst, "w").write("\n".join(parts))
print(dst)
