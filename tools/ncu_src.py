"""Summarise an `ncu --page source --csv` export: stall samples and executed instructions per SASS region, top stalled
instructions with their dominant stall reason, and the dynamic opcode mix.  Usage: python tools/ncu_src.py file.csv [kernel-substr] [chunk]"""
import collections
import csv
import sys


def blocks(path):
    rows = list(csv.reader(open(path)))
    out, cur, name, hdr = [], None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            name = r[1]
        elif r and r[0] == "Address":
            hdr = r
            cur = []
            out.append((name, hdr, cur))
        elif cur is not None and len(r) > 10:
            cur.append(r)
    return out


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    f = lambda x: int(x) if x else 0
    for name, hdr, data in blocks(path):
        if sub not in name:
            continue
        iS, iAll, iEx = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot, totex = sum(f(r[iAll]) for r in data), sum(f(r[iEx]) for r in data)
        print("==", name[:90], "samples", tot, "inst", totex, "sass", len(data))
        acc = accx = 0
        st = collections.Counter()
        for k, r in enumerate(data):
            acc += f(r[iAll]); accx += f(r[iEx])
            for i in sc:
                st[hdr[i]] += f(r[i])
            if (k + 1) % chunk == 0 or k == len(data) - 1:
                top = ", ".join(f"{n[6:]} {100 * v / max(acc, 1):.0f}%" for n, v in st.most_common(3))
                print(f"{k + 1:5d} samples {100 * acc / tot:5.1f}%  inst {100 * accx / totex:5.1f}%  [{top}]  {r[iS].strip()[:40]}")
                acc = accx = 0; st = collections.Counter()
        for k in sorted(range(len(data)), key=lambda k: -f(data[k][iAll]))[:25]:
            r = data[k]
            s2 = sorted(((f(r[i]), hdr[i][6:]) for i in sc), reverse=True)[:2]
            print(f"{k:5d} {r[iS].strip()[:58]:58s} {r[iAll]:>6s} {r[iEx]:>9s} {s2}")
        c = collections.Counter()
        for r in data:
            ops = [o for o in r[iS].strip().split() if not o.startswith("@")]
            if ops:
                c[ops[0].split(".")[0]] += f(r[iEx])
        print([(k, round(100 * v / totex, 1)) for k, v in c.most_common(22)])


if __name__ == "__main__":
    main()
