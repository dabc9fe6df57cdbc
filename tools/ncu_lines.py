#!/usr/bin/env python3
"""Per-source-line stall samples of one kernel from an `ncu --set full --import-source on` report.
  python tools/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur, agg = None, None, []
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            agg.append((cur, int(r[0]), r[1], int(r[6]) if r[6].isdigit() else 0, int(r[7]) if r[7].isdigit() else 0, r))
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, toti = sum(a[3] for a in agg), sum(a[4] for a in agg)
    tots = {h: sum(int(a[5][idx[h]]) for a in agg if a[5][idx[h]].isdigit()) for h in stall}
    print("samples %d, warp instructions %d" % (tot, toti))
    print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:10]))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        st = sorted(((h[6:], int(a[5][idx[h]])) for h in stall if a[5][idx[h]].isdigit() and int(a[5][idx[h]]) > 0), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% samples %5.1f%% inst  %s:%d  %s   %s" % (100.0 * a[3] / tot, 100.0 * a[4] / toti, a[0], a[1], a[2].strip()[:80], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
