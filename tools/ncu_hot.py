#!/usr/bin/env python
"""Top stall-sample SASS instructions of each kernel in an .ncu-rep (source page):  python tools/ncu_hot.py rep [topN]"""
import csv, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
seen = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:60]; hdr = rows[i + 1]; i += 2
        body = []
        while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
            body.append(rows[i]); i += 1
        if name in seen: continue
        seen.add(name)
        c = {h: k for k, h in enumerate(hdr)}
        S = c["Warp Stall Sampling (All Samples)"]
        tot = sum(int(r[S]) for r in body if len(r) > S and r[S].isdigit())
        print(f"== {name}: {tot} samples, {len(body)} instrs")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[c[h]]) for r in body if len(r) > c[h] and r[c[h]].isdigit()) for h in stall_cols}
        print("   totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
        idx = sorted(range(len(body)), key=lambda k: -int(body[k][S]) if body[k][S].isdigit() else 0)[:top]
        for k in sorted(idx):
            r = body[k]
            st = {h[6:]: int(r[c[h]]) for h in stall_cols if r[c[h]].isdigit() and int(r[c[h]]) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"   #{k:5d} {int(r[S]):6d} ({100*int(r[S])/max(tot,1):4.1f}%) exec={r[c['Instructions Executed']]:>8s} {r[1].strip()[:70]:70s} {st}")
    else:
        i += 1

# bucketed view: python tools/ncu_hot.py rep N buckets
if len(sys.argv) > 3:
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1][:50]; hdr = rows[i + 1]; i += 2
            body = []
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                body.append(rows[i]); i += 1
            c = {h: k for k, h in enumerate(hdr)}
            S = c["Warp Stall Sampling (All Samples)"]
            B = 40
            print("==", name)
            for b0 in range(0, len(body), B):
                seg = body[b0:b0 + B]
                n = sum(int(r[S]) for r in seg if r[S].isdigit())
                ops = {}
                for r in seg:
                    op = r[1].strip().split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                top3 = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
                ex = seg[0][c["Instructions Executed"]]
                print(f"   [{b0:5d},{b0+B:5d}) samples={n:6d} exec0={ex:>8s} {top3}")
        else:
            i += 1
