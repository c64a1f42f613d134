"""Aggregates an ncu source page (ncu -i X.ncu-rep --page source --csv) per SASS function segment:
executed warp instructions, FP64-pipe instructions, stall samples.  usage: python profiles/srcpage.py X.csv"""
import csv
import sys
from collections import Counter, defaultdict


def main(path, detail=None):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    segs, cur = [], []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        sass = r[col["Source"]].strip()
        toks = sass.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        cur.append((addr - base, op, sass, int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0), r))
        if op.startswith("RET") or (op.startswith("EXIT") and not toks[0].startswith("@")):
            segs.append(cur)
            cur = []
    if cur:
        segs.append(cur)
    tot_i = sum(x[4] for s in segs for x in s)
    tot_s = sum(x[3] for s in segs for x in s)
    print(f"total executed {tot_i:,}  samples {tot_s:,}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for k, s in enumerate(segs):
        ex = sum(x[4] for x in s)
        sm = sum(x[3] for x in s)
        if ex == 0 and sm == 0:
            continue
        f64 = sum(x[4] for x in s if x[1].split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "FRND", "F2I", "I2F", "MUFU"))
        mov = sum(x[4] for x in s if x[1].split(".")[0] in ("MOV", "IMAD") and ".MOV" in x[1] or x[1] == "MOV")
        ops = Counter()
        for x in s:
            ops[x[1].split(".")[0]] += x[4]
        st = Counter()
        for x in s:
            for h in stall_cols:
                v = x[5][col[h]]
                if v and v != "0":
                    st[h] += int(v)
        top = " ".join(f"{o}:{100*c/ex:.0f}%" for o, c in ops.most_common(8)) if ex else ""
        tst = " ".join(f"{h[6:]}:{100*c/max(sm,1):.0f}%" for h, c in st.most_common(6))
        print(f"seg {k} [{s[0][0]:#x}-{s[-1][0]:#x}] {len(s)} instr  executed {100*ex/tot_i:.1f}%  samples {100*sm/max(tot_s,1):.1f}%  fp64-pipe {100*f64/max(ex,1):.0f}%  mov {100*mov/max(ex,1):.0f}%")
        print(f"      ops {top}\n      stalls {tst}")
    if detail is not None:
        s = segs[int(detail)]
        for x in s:
            print(f"{x[0]:#7x} {x[4]:>10} {x[3]:>6}  {x[2]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
