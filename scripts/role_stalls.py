"""Per-role reading of an ncu source-page CSV of rollout_bwd2_kernel (warp-specialised: producer code, EXIT, consumer code, EXIT):
instructions per ODE-stage adjoint, share of the stall samples, stall reasons per issued instruction and the instructions that
hold most samples, separately for the producer and the consumer warp.
usage: ncu -i <rep> --page source --csv --launch-skip K --launch-count 1 > sass.csv ; python scripts/role_stalls.py sass.csv <CTAs> <stages>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ctas, stages = float(sys.argv[2]), float(sys.argv[3])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen, u = set(), []
for r in rows[2:]:
    if r and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); u.append(r)
E_, S_ = idx["Instructions Executed"], idx["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
hot = max(int(r[E_]) for r in u) * 0.2
# segments of hot code separated by EXIT: first = producer, second = consumer
exits = [i for i, r in enumerate(u) if r[1].split()[-1] == "EXIT" or " EXIT" in r[1]]
seg_of, seg, has_hot = {}, 0, False
for i, r in enumerate(u):
    if int(r[E_]) >= hot:
        has_hot = True
    seg_of[i] = seg
    if i in exits and has_hot:
        seg += 1; has_hot = False
names = {0: "producer", 1: "consumer"}
tot = sum(int(r[S_]) for r in u)
print(f"kernel: {u and rows[0][1][:80]}  total samples {tot}")
for sg in (0, 1):
    R = [(i, r) for i, r in enumerate(u) if seg_of[i] == sg]
    H = [(i, r) for i, r in R if int(r[E_]) >= hot]
    samples = sum(int(r[S_]) for _, r in R)
    instr = sum(int(r[E_]) for _, r in H) / ctas / stages
    st = collections.Counter()
    for _, r in H:
        for c in stall_cols:
            v = int(r[idx[c]] or 0)
            if v:
                st[c[6:]] += v
    sel = max(st["selected"], 1)
    ops = collections.Counter()
    for _, r in H:
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[E_]) / ctas / stages
    f64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD"))
    print(f"\n{names[sg]}: {100 * samples / tot:.1f} % of samples, {instr:.0f} warp instructions per stage ({f64:.0f} FP64)")
    print("  stalls per issued instruction: " + "  ".join(f"{k} {v / sel:.2f}" for k, v in st.most_common(9) if k != "selected"))
    print("  instruction mix per stage: " + "  ".join(f"{k} {v:.0f}" for k, v in ops.most_common(12)))
    print("  instructions holding most samples:")
    for i, r in sorted(H, key=lambda ir: -int(ir[1][S_]))[:8]:
        top = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[0]
        print(f"    {100 * int(r[S_]) / tot:5.2f} %  {r[1].strip()[:70]:70s} {top[1]}")
