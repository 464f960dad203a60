#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report holding the kernels of ONE step in launch order.

    ncu --set full --clock-control none -k regex:"attention|gemm_tc|embed_kernel|layernorm|topk_kernel|mask_seen" \
        -c 13 -o step python bench.py --steps 1 --warmup 3 --no-cpu
    python tools/ncu_traffic.py step.ncu-rep "how it was captured" > profiles/traffic.json

Per stage: DRAM bytes (read + write) of that launch, its duration and the pipe utilisations bench.py's `roofline`
explanation refers to.  Stage names follow api.cu's pipeline order (EasyDGL, one block)."""
import csv
import json
import subprocess
import sys

ORDER = ["embed", "qkvt_gemm", "attention", "ao_gemm", "ln_att", "ff1_gemm", "ff2_gemm", "ln_ff", "tr_gemm", "ln_out",
         "logits_gemm", "mask_seen", "topk"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(path, source):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    u = dict(zip(hdr, units))

    def val(rec, key):
        return float(rec[key]) * UNIT.get(u.get(key, ""), 1.0)

    stages = {}
    for name, r in zip(ORDER, data):
        rec = dict(zip(hdr, r))
        stages[name] = {
            "kernel": rec.get("Kernel Name", "?")[:70],
            "dram_bytes_per_launch": val(rec, "dram__bytes_read.sum") + val(rec, "dram__bytes_write.sum"),
            "duration_ms": val(rec, "gpu__time_duration.sum"),
            "tensor_pipe_pct": float(rec.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) or 0),
            "issue_active_pct": float(rec.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) or 0),
            "xu_pipe_pct": float(rec.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0) or 0),
            "dram_throughput_pct": float(rec.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0) or 0),
        }
    json.dump({"source": source, "stages": stages}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none")
