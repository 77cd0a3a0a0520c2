"""Per tcgen05 conv launch of one cfg2 step: time, tensor-pipe active %, DRAM throughput %, from

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
        -k regex:conv3d_tma --clock-control none --profile-from-start off --csv --log-file X.csv \
        python profiles/run_step.py

    python profiles/summarize_k2_pipe.py X.csv profiles/rN_k2_pipe.summary.txt profiles/k2_tensor_pipe.json
"""
import csv
import json
import re
import sys

src, out_txt, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
launch = {}
for r in rows:
    d = launch.setdefault(int(r[0]), {"name": r[4], "grid": r[8]})
    d[r[-3]] = float(r[-1].replace(",", ""))
T = "gpu__time_duration.sum"
P = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
D = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
tot = sum(d[T] for d in launch.values())
unit_ns = tot > 1e5                       # ncu reports ns unless told otherwise
scale = 1e-3 if unit_ns else 1.0
lines = ["# per tcgen05 conv launch of one cfg2 step (ncu, cold cache, --clock-control none): time, "
         "tensor-pipe active %, DRAM throughput %"]
for i in sorted(launch):
    d = launch[i]
    lines.append(f"{d[T]*scale:7.1f} us  tensor {d[P]:5.1f} %  dram {d[D]:5.1f} %   {d['grid']:>14s}  "
                 f"{re.sub(r'\(CUtensorMap_st.*', '', d['name'])[:70]}")
w = sum(d[T] * d[P] for d in launch.values()) / tot
lines.append(f"# total {tot*scale:.1f} us, time-weighted tensor pipe {w:.1f} %")
open(out_txt, "w").write("\n".join(lines) + "\n")
json.dump({"time_weighted_pct": round(w, 2), "launches": len(launch), "sum_us_cold": round(tot * scale, 1),
           "source": f"ncu --metrics sm__pipe_tensor_cycles_active... -k regex:conv3d_tma (one cfg2 step, "
                     f"profiles/run_step.py), {src}; includes the FeatureNet planar / 5x5 convs"},
          open(out_json, "w"))
print(lines[-1])
