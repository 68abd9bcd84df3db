"""gpurun_out/{bench,launches,prof_*}_<tag> -> profiles/ (bench JSON, launch list, ncu summary markdown)."""
import csv, glob, json, os, shutil, subprocess, sys
from collections import defaultdict

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
bench = json.loads(open(os.path.join(go, f"bench_{tag}.json")).read().strip().splitlines()[-1])
shutil.copy(os.path.join(go, f"bench_{tag}.json"), os.path.join(pr, f"bench_{tag}.json"))
shutil.copy(os.path.join(go, f"launches_{tag}.csv"), os.path.join(pr, f"launches_{tag}.csv"))
out = [f"# ncu / bench summary, tag {tag}"]
fp = bench.get("fp32_storage", {})
out.append(f"bench.py --steps {bench['steps']} --warmup {bench['warmup']} (B200, 1 GPU): **{bench['value']:.1f} steps/s** "
           f"bf16-storage ({1e3 * bench['ms_per_step']:.1f} us/step), fp32-storage {fp.get('value', 0):.1f} steps/s; "
           f"e2e (host round trip every step) {bench['e2e']['value']:.1f} steps/s; CPU baseline "
           f"{bench.get('cpu_baseline', {}).get('value')} steps/s on {bench.get('cpu_baseline', {}).get('cores')} threads; "
           f"rel-L2 vs fp32 CPU oracle: fp32-storage {fp.get('rel_l2', [0])[0]:.2e}, bf16-storage {bench['rel_l2'][0]:.2e}; "
           f"train step {bench['train_step']['ms_per_step']:.2f} ms; clocks {bench['clocks']}.")
out.append("\n## per-kernel CUDA-event times inside bench.py (us, bf16 | fp32)")
for n, v in bench["kernels"].items():
    other = fp.get("kernels", {}).get(n, {}).get("mean_us")
    out.append(f"- {n}: {v['mean_us']:.1f} | {other if other is None else round(other, 1)}  (x{v['launches_per_step']} per step)")
for n, v in fp.get("kernels", {}).items():
    if n not in bench["kernels"]:
        out.append(f"- {n}: - | {v['mean_us']:.1f}  (fp32-storage path only, x{v['launches_per_step']} per step)")
rf = bench["roofline"]
out.append(f"\nroofline (bench.py): {rf}")
# launch list
rows = list(csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if l.startswith('"')))
hdr = rows[0]
iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[iname]].append(float(r[ival].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for k, v in agg.items() if "fno::" in k and "pack" not in k)
out.append("\n## ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)")
for k, v in agg.items():
    unit = 1e-3  # ns -> us
    share = 100 * sum(v) / tot if ("fno::" in k and "pack" not in k) else 0.0
    out.append(f"- {k[:70]}: n={len(v)}, mean {unit * sum(v) / len(v):.1f} us, share of step {share:.1f} %")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
traffic = {}
for rep in sorted(glob.glob(os.path.join(go, f"prof_*_{tag}.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units, v = r[0], r[1], r[2]
    name = v[h.index("Kernel Name")]
    out.append(f"\n## ncu --set full: {name[:80]} (bf16 storage, B=256)")
    for w in want:
        if w in h:
            out.append(f"- {w}: {v[h.index(w)]} {units[h.index(w)]}")
    def mb(metric):
        x, u = float(v[h.index(metric)].replace(",", "")), units[h.index(metric)]
        return x * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[u]
    t = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    short = name.split("(")[0].split("<")[0].replace("void ", "").split("::")[-1].strip()
    traffic[short] = (int(t), os.path.basename(rep))
    out.append(f"- DRAM traffic (read+write): {t / 1e6:.1f} Mbyte")
out.append(f"\nDRAM traffic per launch (bytes): { {k: v[0] for k, v in traffic.items()} }")
# bench.py reads roofline.traffic from here (keyed by kernel | activation storage | batch of the capture)
tj = os.path.join(pr, "ncu_traffic.json")
table = json.load(open(tj)) if os.path.exists(tj) else {}
for k, (t, rep) in traffic.items():
    table[f"{k}|bf16|256"] = {"dram_bytes": t, "source": f"ncu --set full, {rep}, tools/prof_driver.py --act bf16 (B=256), tag {tag}"}
json.dump(table, open(tj, "w"), indent=1, sort_keys=True)
open(os.path.join(pr, f"ncu_{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
