"""Summaries of the ncu captures pulled back in gpurun_out/ -> profiles/ (tracked).
  r2_launches_bench.csv (gpu__time_duration per launch of `bench.py --steps 1 --warmup 1`)  -> profiles/r2_launches_bench.md
  r2_full_<kernel>.csv  (ncu --set full, --page raw)                                       -> profiles/r2_ncu_full_summary.md
and profiles/roofline_traffic.json (DRAM bytes per launch of the dominant kernels)."""
import csv, collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

def short(name):
    name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("tnad::", "").replace("void ", ""))
    return re.sub(r"^.*unnamed>::", "", name.strip())

# ---- launch list
rows = []
with open(os.path.join(G, "r2_launches_bench.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    k = short(r[ik])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
out = ["# ncu launch list of `bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-trg` (round 2)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 4000` (first 4000 launches: warm-up call + most of the timed call).",
       "Per-launch times under ncu are cold-cache and serialised: the SHARE of each kernel is what must agree with the",
       "CUDA-event family table of bench.py (`roofline.kernel_ms_instrumented_pass`), not the absolute time.", "",
       "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:80]}` | {n} | {us / 1e3:.2f} | {100 * us / tot:.1f} % | {us / n:.1f} |")
out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot / 1e3:.2f} | | |")
os.makedirs(P, exist_ok=True)
with open(os.path.join(P, "r2_launches_bench.md"), "w") as f:
    f.write("\n".join(out) + "\n")
import shutil
shutil.copy(os.path.join(G, "r2_launches_bench.csv"), os.path.join(P, "r2_launches_bench.csv"))
print("\n".join(out[:22]))

# ---- full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
full = ["# ncu --set full summaries (round 2)", "",
        "Command per kernel: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 3 -c 2 python tools/c4_trace.py 4 128 1 1`",
        "(energy + gradient at d=4, chi=128, maxit=1: the kernels of the n = 2048 eigen-decomposition and the step's contractions; the third and",
        "fourth launch of each kernel are captured).  Raw pages: gpurun_out/r2_full_<kernel>.csv (scratch).",
        "`gemm_tma_kernel`: captured from `tools/contract_bench.py 128 16 2` (-s 6 -c 6: four launches of `ibd,dcl->ibcl`, 2048 x 2048 x 128, and two of",
        "`ibcl,jkcb->ijlk`, 16384 x 256 x 256 -- the two dominant shapes of a ctmrgstep).", ""]
traffic = {}
for fn in sorted(os.listdir(G)):
    m = re.match(r"r2_full_(.+)\.csv$", fn)
    if not m:
        continue
    with open(os.path.join(G, fn)) as f:
        lines = [l for l in f if l.startswith('"')]
    if len(lines) < 3:
        continue
    rd = list(csv.reader(lines))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        name = short(r[hdr.index("Kernel Name")])
        full.append(f"## `{name}`  (capture file r2_full_{m.group(1)})")
        full.append("")
        full.append("| metric | value | unit |")
        full.append("|---|---:|---|")
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = r[i]
                full.append(f"| {w} | {r[i]} | {units[i]} |")
        full.append("")
        try:
            db = float(vals["dram__bytes_read.sum"].replace(",", "")) + float(vals["dram__bytes_write.sum"].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            db = float(vals["dram__bytes_read.sum"].replace(",", "")) * scale.get(ur, 1) + float(vals["dram__bytes_write.sum"].replace(",", "")) * scale.get(uw, 1)
            traffic.setdefault(m.group(1) + "_dram_bytes_per_launch", db)
        except Exception:
            pass
with open(os.path.join(P, "r2_ncu_full_summary.md"), "w") as f:
    f.write("\n".join(full) + "\n")
tp = os.path.join(P, "roofline_traffic.json")
old = {}
if os.path.exists(tp):
    with open(tp) as f:
        old = json.load(f)
old.update(traffic)
with open(tp, "w") as f:
    json.dump(old, f, indent=1)
print(json.dumps(traffic, indent=1))
