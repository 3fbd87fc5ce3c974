#!/usr/bin/env python
"""Census of the LAST train step in an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv` of
`bench.py --steps 1 --warmup 1`): launches and device time per kernel from the step's first kernel (the text / image
assembly) to the optimiser.  Model-build kernels (torch dtype casts at load time) precede it and are listed separately.
  python tools/launch_summary.py profiles/r02_launches_c2_b256.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    names = [(r[ki].split("(")[0][:64], float(r[vi].replace(",", "")) / 1e3) for r in rows[hi + 1:] if len(r) > vi]
    sgd = [i for i, (n, _) in enumerate(names) if "sgd_kernel" in n]
    # a step ends with its last sgd launch; the previous step's last sgd launch precedes its first kernel
    ends = [i for k, i in enumerate(sgd) if k + 1 == len(sgd) or sgd[k + 1] != i + 1]
    start = ends[-2] + 1 if len(ends) > 1 else 0
    step = names[start:ends[-1] + 1]
    print(f"# {path}: {len(names)} launches captured; last step = launches {start}..{ends[-1]} ({len(step)} launches)")
    build = [x for x in names[:start] if not ("mvlpt" in x[0] or "<unnamed>" in x[0])]
    for title, part in (("last step", step), ("library kernels before it that are not ours (model build: dtype casts, the "
                                               "LayerNorm-carry vectors; prompt load)", build)):
        cnt, tot = collections.Counter(n for n, _ in part), collections.defaultdict(float)
        for n, v in part:
            tot[n] += v
        total = sum(tot.values())
        print(f"## {title}: {len(part)} launches, {total:.1f} us (serialised, cold-cache ncu times: compare shares)")
        for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"{n:66s} {cnt[n]:5d} {v:10.1f} us {100 * v / max(total, 1e-9):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
