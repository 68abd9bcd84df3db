"""profiles/sass_<tag>.txt: per kernel of libcfdbench_b200.so the count of the SASS mnemonics that prove a Blackwell-native
path (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk),
UTMALDG (TMA tensor load), SYNCS (mbarrier), plus legacy HMMA (must be 0).  Runs without a GPU (cuobjdump)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "cfdbench_b200", "libcfdbench_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "STG.E.ENL2.256", "HMMA", "FFMA2", "MUFU.EX2"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in ln:
        continue
    counts[cur]["instructions"] += 1
    for k in keys:
        if re.search(r"\b" + re.escape(k) + r"\b", ln) or (k.endswith("256") and k in ln):
            counts[cur][k] += 1
            total[k] += 1
out = [f"# SASS mnemonic counts per kernel, {os.path.basename(so)} (cuobjdump -sass), tag {tag}", "",
       "| kernel | instr | " + " | ".join(keys) + " |", "|---|---|" + "---|" * len(keys)]
for k, c in counts.items():
    if c["instructions"] < 20:
        continue
    out.append(f"| {k[:90]} | {c['instructions']} | " + " | ".join(str(c[x]) for x in keys) + " |")
out.append("")
out.append("totals: " + ", ".join(f"{k} {total[k]}" for k in keys))
path = os.path.join(ROOT, "profiles", f"sass_{tag}.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
