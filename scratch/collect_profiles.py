"""Copies the artefacts of `scratch/final_profiles.sh <tag>` from gpurun_out/ into profiles/ under the names r1_final_*,
writes the text summaries, ncu_traffic.json and the per-step launch shares.  usage: python scratch/collect_profiles.py <tag>"""
import collections, csv, io, json, re, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"
KERNELS = (("tc3", "glin_tc3_kernel"), ("tc", "glin_tc_kernel"), ("ffma", "glin_gemm_f2_kernel"), ("step", "reverse_step_kernel"),
           ("attn", "node_attention_bulk_kernel"))
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures profiles/r1_final_*.ncu-rep "
                       "(B = 25600, N = 21); see profiles/README.md"}
for k, name in KERNELS:
    rep = f"profiles/r1_final_{k}.ncu-rep"
    shutil.copy(f"gpurun_out/{tag}_{k}.ncu-rep", rep)
    with open(f"profiles/r1_final_{k}.summary.txt", "w") as f:
        f.write(subprocess.run([sys.executable, "scratch/ncu_summary.py", rep], capture_output=True, text=True).stdout)
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    h, u, r = rows[0], rows[1], rows[-1]
    ci = {n: i for i, n in enumerate(h)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic[name] = int(sum(float(r[ci[m]]) * scale[u[ci[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
    print(name, traffic[name], r[ci["gpu__time_duration.sum"]], u[ci["gpu__time_duration.sum"]])
json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=2)
shutil.copy(f"gpurun_out/{tag}_bench.json", "profiles/r1_final_bench.json")
shutil.copy(f"gpurun_out/{tag}_launches.csv", "profiles/r1_final_launches.csv")

rows = []
for row in csv.DictReader([l for l in open("profiles/r1_final_launches.csv") if not l.startswith("==")]):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("sd::", ""), v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else v))
gru = [i for i, r in enumerate(rows) if r[0].startswith("gru_step_fused")]
runs, start, prev = [], gru[0], gru[0]
for i in gru[1:]:
    if i - prev > 3:
        runs.append((start, prev)); start = i
    prev = i
runs.append((start, prev))
runs = [r for r in runs if sum(1 for i in gru if r[0] <= i <= r[1]) >= 120]      # decodes (the encoder's runs are shorter)
lo, hi = runs[-2][1] + 1, runs[-1][1]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, us in rows[lo:hi + 1]:
    agg[k][0] += 1; agg[k][1] += us
tot = sum(v[1] for v in agg.values())
with open("profiles/r1_final_launches.summary.txt", "w") as f:
    f.write(f"# one pipeline step = encode + 10 x Denoiser/step + 120-frame decode: launches {lo}..{hi} of profiles/r1_final_launches.csv\n"
            f"# (ncu per-launch times are cold-cache and serialised: compare SHARES with the CUDA-event numbers, not absolutes)\n"
            f"# total {tot / 1e3:.1f} ms, {hi - lo + 1} launches\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        f.write(f"{us / 1e3:8.2f} ms {100 * us / tot:5.1f}%  n={n:4d} avg {us / n:8.1f} us  {k[:80]}\n")
print(open("profiles/r1_final_launches.summary.txt").read())
