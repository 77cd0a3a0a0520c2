"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.
  python profiles/summarize_ncu.py launches gpurun_out/r1_launches_04.csv > profiles/r1_launches_04.summary.txt
  python profiles/summarize_ncu.py full gpurun_out/r1_k1_v2.ncu-rep > profiles/r1_k1_v2.summary.txt
"""
import collections, csv, io, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        try:
            rows.append((x["Kernel Name"], float(x["Metric Value"].replace(",", "")), x.get("Grid Size", "")))
        except Exception:
            pass
    tot = sum(v for _, v, _ in rows)
    print(f"# {path}: {len(rows)} launches, {tot/1e3:.1f} us total (gpu__time_duration.sum, serialised, cold cache: compare SHARES)")
    agg = collections.OrderedDict()
    for n, v, g in rows:
        n = re.sub(r"\(.*", "", n)[:90]
        agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
    print("## by kernel")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v/1e3:10.1f} us {c:4d}x {100*v/tot:5.1f}%  {n}")
    print("## in launch order (this repo's kernels)")
    for n, v, g in rows:
        if any(k in n for k in ("casmvs", "tc::", "tc2::", "tc3::", "tma::", "tma8::", "tma2::")):
            print(f"{v/1e3:9.1f} us {g:>16s}  {re.sub(r'\(.*', '', n)[:80]}")

def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
    print(f"# {path} (ncu --set full --clock-control none)")
    for w in want:
        if w in idx:
            print(f"{w:72s} [{units[idx[w]]}] " + " | ".join(r[idx[w]][:48] for r in rows[2:]))
    print("## warp stall reasons (cycles per issued instruction, > 0.3)")
    for h in hdr:
        if "issue_stalled" in h and h.endswith(".ratio"):
            vals = [float(r[idx[h]]) for r in rows[2:]]
            if max(vals) > 0.3:
                print(f"{h.split('issue_stalled_')[1]:50s} " + " | ".join(f"{v:.2f}" for v in vals))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
