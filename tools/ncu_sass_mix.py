#!/usr/bin/env python3
"""Instruction mix and hot regions of ONE kernel from an .ncu-rep captured with --set full --import-source on.

    python tools/ncu_sass_mix.py profiles/r01_gemm_topk_cg2_mc2.ncu-rep [--bucket 64]

Prints (1) warp-instructions and stall samples per opcode, (2) the same per bucket of consecutive SASS lines (a cheap
stand-in for "which warp role / loop"), (3) the single hottest lines.  Used for DESIGN.md section 6 (barrier polling
share of the tcgen05 kernel) and profiles/r01_ivfpq_scan_v1.md."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    bucket = int(sys.argv[sys.argv.index("--bucket") + 1]) if "--bucket" in sys.argv else 64
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr_i]
    ci = {n: i for i, n in enumerate(h)}
    data = []
    for r in rows[hdr_i + 1:]:
        try:
            data.append((r[ci["Source"]], int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])))
        except (ValueError, IndexError):
            pass
    tot = sum(d[1] for d in data) or 1
    ts = sum(d[2] for d in data) or 1
    print(f"{rows[0][1] if rows and len(rows[0]) > 1 else rep}: {len(data)} SASS lines, {tot} warp-instructions, {ts} samples")
    ops, smp = collections.Counter(), collections.Counter()
    for s, n, sm in data:
        t = s.split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        ops[op] += n
        smp[op] += sm
    print("\n-- by opcode")
    for op, n in ops.most_common(20):
        print(f"{op:12s} inst {100 * n / tot:5.1f}%   samples {100 * smp[op] / ts:5.1f}%")
    print(f"\n-- by bucket of {bucket} lines")
    for b in range(0, len(data), bucket):
        ch = data[b:b + bucket]
        n, s = sum(c[1] for c in ch), sum(c[2] for c in ch)
        if n > 0.01 * tot or s > 0.01 * ts:
            print(f"lines {b:5d}-{b + len(ch) - 1:5d}  inst {100 * n / tot:5.1f}%   samples {100 * s / ts:5.1f}%   e.g. {max(ch, key=lambda c: c[2])[0].strip()[:70]}")
    print("\n-- hottest lines by samples")
    for i, (s, n, sm) in sorted(enumerate(data), key=lambda t: -t[1][2])[:12]:
        print(f"line {i:5d}  samples {100 * sm / ts:5.1f}%  inst {100 * n / tot:5.1f}%  {s.strip()[:80]}")


if __name__ == "__main__":
    main()
