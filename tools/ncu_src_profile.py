"""Summarise an `ncu --page source --csv` export: stall samples per block of SASS instructions, with the dominant opcodes and
stall reasons of each block (tools/ only).  Usage: python tools/ncu_src_profile.py src.csv [block]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = rows[2:]
tot = sum(int(r[ci["# Samples"]]) for r in data)
texec = sum(int(r[ci["Instructions Executed"]]) for r in data)
print("instructions %d, samples %d, warp-instr executed %d" % (len(data), tot, texec))
for b0 in range(0, len(data), blk):
    seg = data[b0:b0 + blk]
    s = sum(int(r[ci["# Samples"]]) for r in seg)
    ex = sum(int(r[ci["Instructions Executed"]]) for r in seg)
    ops = collections.Counter(r[ci["Source"]].split()[0].split(".")[0] if not r[ci["Source"]].strip().startswith("@") else r[ci["Source"]].split()[1].split(".")[0] for r in seg)
    st = collections.Counter()
    for r in seg:
        for n in stall_cols:
            st[n[6:]] += int(r[ci[n]])
    exc = sum(int(r[ci["L1 Wavefronts Shared Excessive"]] or 0) for r in seg)
    print("%5d-%5d  samples %5.1f%%  exec %5.1f%%  ops %-40s stalls %-50s smem-excess %d" % (
        b0, b0 + len(seg), 100.0 * s / tot, 100.0 * ex / texec, " ".join("%s:%d" % kv for kv in ops.most_common(4)),
        " ".join("%s:%d" % kv for kv in st.most_common(4)), exc))
