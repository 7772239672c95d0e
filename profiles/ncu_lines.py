"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per source line:
share of warp instructions, average active threads per instruction, share of stall samples."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    out = collections.defaultdict(lambda: [0, 0, 0, ""])
    cur_file, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] in ("Line No", "Address"):
            hdr = r
            continue
        if hdr is None or hdr[0] != "Line No" or not cur_file:
            continue
        d = dict(zip(hdr, r))
        try:
            line = int(d["Line No"])
            ie = int(d.get("Instructions Executed") or 0)
            te = int(d.get("Thread Instructions Executed") or 0)
            smp = int(d.get("# Samples") or 0)
        except ValueError:
            continue
        key = (cur_file.split("/")[-1], line)
        out[key][0] += ie
        out[key][1] += te
        out[key][2] += smp
        out[key][3] = (d.get("Source") or "")[:100]
    tot_ie = sum(v[0] for v in out.values()) or 1
    tot_te = sum(v[1] for v in out.values())
    tot_smp = sum(v[2] for v in out.values()) or 1
    print(f"total warp inst {tot_ie}  thread inst {tot_te}  avg threads/inst {tot_te / tot_ie:.2f}  samples {tot_smp}")
    for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]}:{k[1]:4d} inst={100 * v[0] / tot_ie:5.2f}% thr/inst={v[1] / max(v[0], 1):5.1f} "
              f"stall_samples={100 * v[2] / tot_smp:5.2f}%  {v[3].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
