#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`, B200_PROFILING.md) of tools/profile_step.py
into the per-kernel table committed under profiles/: launches, mean duration, share of the region.
Times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event numbers, not absolutes.

    python tools/summarize_launches.py gpurun_out/r1_launches_v4.csv > profiles/r01_launches_v4.md
"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace("void ", "").replace("wb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return re.sub(r"\(.*", "", name)


def table(title, rows):
    tot = sum(r[3] for r in rows)
    agg = collections.OrderedDict()
    for name, grid, block, us in rows:
        a = agg.setdefault((name, grid, block), [0, 0.0])
        a[0] += 1
        a[1] += us
    out = [f"### {title}: {len(rows)} launches, {tot / 1e3:.3f} ms of kernel time", "",
           "| kernel | grid | block | launches | mean us | share |", "|---|---|---|---|---|---|"]
    for (name, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {grid} | {block} | {n} | {t / n:.2f} | {100 * t / tot:.1f} % |")
    return "\n".join(out) + "\n"


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    next(r)
    seq = [(short(d[4]), d[8], d[7], float(d[-1]) / 1e3) for d in r]
    print(f"# ncu launch list: {path}\n")
    print("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv` over "
          "`tools/profile_step.py --region both --steps 1` (medium.en bf16, B = 256, decode step at length 224, then one "
          "encoder chunk of 32 utterances).  Cold-cache, serialised launches: shares are the evidence, not absolutes.\n")
    emb = [i for i, s in enumerate(seq) if s[0].startswith("decoder_embed")]
    gre = [i for i, s in enumerate(seq) if s[0].startswith("greedy_step")]
    if emb and gre:
        print(table("decode step (decoder_embed .. greedy_step)", seq[emb[0]:gre[0] + 1]))
        layer = seq[emb[0] + 1:emb[0] + 12]
        print("One decoder layer, in launch order:\n")
        print("| kernel | grid | us |\n|---|---|---|")
        for name, grid, _, us in layer:
            print(f"| `{name}` | {grid} | {us:.2f} |")
        print()
        rest = seq[gre[0] + 1:]
        if rest:
            print(table("encoder chunk (stem + 24 layers + final LN + cross-K/V projection)", rest))
            print("First encoder layer, in launch order:\n")
            print("| kernel | grid | us |\n|---|---|---|")
            for name, grid, _, us in rest[3:10]:
                print(f"| `{name}` | {grid} | {us:.2f} |")
    else:
        print(table("all launches", seq))


if __name__ == "__main__":
    main(sys.argv[1])
