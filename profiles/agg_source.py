"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = None
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0, ""])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr and len(r) > i_inst and r[2] == "-":
            try:
                key = (cur_file, int(r[0]))
                agg[key][0] += int(r[i_inst])
                agg[key][1] += int(r[i_samp])
                agg[key][2] = r[1].strip()[:100]
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[1] for v in agg.values())
    print("total warp instructions", tot, "samples", tots)
    by_file = collections.Counter()
    for (f, _), v in agg.items():
        by_file[f] += v[0]
    print({k: f"{v / tot * 100:.1f}%" for k, v in by_file.items()})
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f:16s}:{l:4d} inst {v[0] / tot * 100:5.1f}% samp {v[1] / tots * 100:5.1f}%  {v[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
