#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into the markdown table committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/r1_dec_layer_v4.ncu-rep > profiles/r01_ncu_decode_layer_v4.md
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def short(n):
    return re.sub(r"\(.*", "", n.replace("void ", "").replace("unnamed>::", "").replace("<unnamed>::", ""))


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full: {path}\n")
    print("Captured with `ncu --profile-from-start off --set full --clock-control none --import-source on` over "
          "`tools/profile_step.py` (medium.en bf16, B = 256; decode step at length 224 / one encoder chunk of 32 utterances). "
          "ncu replays each kernel ~40x with caches flushed: durations are cold-cache.\n")
    print("| # | kernel | grid | " + " | ".join(t for _, t in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for n, d in enumerate(data):
        cells = []
        for c, _ in COLS:
            if c in idx:
                v, u = d[idx[c]], units[idx[c]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("n/a")
        print(f"| {n} | `{short(d[idx['Kernel Name']])}` | {d[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
