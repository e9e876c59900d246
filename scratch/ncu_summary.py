"""Text summary of an ncu report (run where ncu is installed): python scratch/ncu_summary.py rep.ncu-rep > profiles/x.summary.txt

Per captured kernel: duration, DRAM bytes, utilisation of the main units, and the 25 most-sampled SASS instructions
with their two dominant stall reasons (needs --import-source on / -lineinfo for the source page)."""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
       "sm__warps_active.avg.pct_of_peak_sustained_active"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ci = {n: i for i, n in enumerate(hdr)}
    print(f"# {rep}")
    for r in body:
        print(f"\n## {r[ci['Kernel Name']][:110]}")
        for m in RAW:
            if m in ci:
                print(f"{m:75s} {r[ci[m]]:>18s} {units[ci[m]]}")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    heads = [i for i, r in enumerate(src) if "Source" in r and any("Sampl" in c for c in r)]
    for k, start in enumerate(heads):
        h = src[start]
        end = heads[k + 1] - 1 if k + 1 < len(heads) else len(src)
        body = [r for r in src[start + 1:end] if len(r) == len(h)]
        c = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        total = sum(int(r[c["# Samples"]] or 0) for r in body) or 1
        name = src[start - 1][1][:100] if start > 0 and len(src[start - 1]) > 1 else ""
        print(f"\n## hottest SASS, kernel {k}: {name}   ({total} samples, {len(body)} instructions)")
        agg = {s: sum(int(r[c[s]] or 0) for r in body) for s in stalls}
        print("stall totals: " + ", ".join(f"{s[6:]} {100 * v / total:.0f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for r in sorted(body, key=lambda r: -int(r[c["# Samples"]] or 0))[:25]:
            st = sorted(((s[6:], int(r[c[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
            n = int(r[c["# Samples"]] or 0)
            print(f"{100 * n / total:5.1f}%  {r[c['Source']].strip()[:70]:70s} {st}")
    # per CUDA source line (needs -lineinfo and --import-source on): samples, executed warp instructions, dominant stall reasons
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    heads = [i for i, r in enumerate(src) if r and r[0] == "Line No"]
    for k, start in enumerate(heads):
        h = src[start]
        end = heads[k + 1] - 2 if k + 1 < len(heads) else len(src)
        i_samp, i_inst = h.index("# Samples"), h.index("Instructions Executed")
        stalls = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        lines = [r for r in src[start + 1:end] if len(r) == len(h) and r[0].strip().isdigit()]
        total = sum(int(r[i_samp] or 0) for r in lines) or 1
        if total < 1000:                 # headers and helper files with a handful of samples
            continue
        name = src[start - 1][1][:100] if start > 0 and len(src[start - 1]) > 1 else ""
        print(f"\n## hottest source lines, kernel {k}: {name}   ({total} samples)")
        for r in sorted(lines, key=lambda r: -int(r[i_samp] or 0))[:30]:
            n = int(r[i_samp] or 0)
            st = sorted(((nm[6:], int(r[i] or 0)) for i, nm in stalls), key=lambda kv: -kv[1])[:2]
            print(f"{100 * n / total:5.1f}%  L{r[0]:>4s} {int(r[i_inst] or 0):>10d} inst  {r[1].strip()[:90]:90s} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
