#!/usr/bin/env python
"""Per-kernel DRAM traffic of one `ncu --set full` capture -> r1_traffic.json (what bench.py scales into roofline.traffic).
usage: traffic.py raw.csv log_bytes "source text" [more_raw.csv ...] > r2_traffic.json   (kernels missing from the first capture are taken from the later ones)"""
import csv
import json
import sys

log_bytes = int(sys.argv[2])
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
NAMES = {"tokenize_kernel": "tokenize", "token_kernel": "token", "iptrie_kernel": "iptrie", "exact_kernel": "strings", "scan_kernel": "scan"}


def val(r, k):
    return float(r[idx[k]]) * SCALE.get(units[idx[k]], 1.0)


out = {"source": sys.argv[3], "kernels": {}}
for path in [sys.argv[1]] + sys.argv[4:]:
  rows = list(csv.reader(open(path)))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  for r in rows[2:]:
      name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()  # "void tokenize_kernel<31>(...)" -> tokenize_kernel
      key = NAMES.get(name)
      if key is None or key in out["kernels"]:
          continue
      rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
      dur = float(r[idx["gpu__time_duration.sum"]])
      if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond"):
          dur /= 1e3
      elif units[idx["gpu__time_duration.sum"]] in ("ms", "msecond"):
          dur *= 1e3
      out["kernels"][key] = {
          "dram_bytes_read": rd, "dram_bytes_write": wr, "log_bytes": log_bytes, "dram_bytes_per_log_byte": (rd + wr) / log_bytes,
          "duration_us_under_ncu": dur, "warp_instructions": float(r[idx["smsp__inst_executed.sum"]]),
          "threads_per_instruction": float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
          "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
          "alu_pipe_active_pct": float(r[idx["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]]),
          "fma_pipe_active_pct": float(r[idx["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]]),
          "registers_per_thread": int(float(r[idx["launch__registers_per_thread"]])),
      }
print(json.dumps(out, indent=1))
