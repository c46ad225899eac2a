"""Development aid: how serial is the FP64 instruction stream of each kernel in an object file?
Prints, per function, the share of FP64-pipe instructions that consume the result of the FP64
instruction issued immediately before them (a proxy for exposed DFMA latency with one warp per SMSP)."""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for blk in txt.split("Function : ")[1:]:
    name = blk.split("\n")[0]
    ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", blk)
    fp = [i for i in ins if re.search(r"\bD(FMA|ADD|MUL|SETP)\b", i)]
    dep = 0
    for a, b in zip(fp, fp[1:]):
        m = re.search(r"\bD(?:FMA|ADD|MUL)\s+(R\d+)", a)
        if not m:
            continue
        d = int(m.group(1)[1:])
        rest = b.split(",", 1)[1] if "," in b else ""
        srcs = [int(x) for x in re.findall(r"R(\d+)", rest)]
        if d in srcs:
            dep += 1
    n_mufu = sum("MUFU" in i for i in ins)
    n_l = sum(("LDL" in i or "STL" in i) for i in ins)
    print(f"{name[:70]:70s} instr {len(ins):6d} fp64 {len(fp):5d} back-to-back-dependent {dep:5d} ({100.0*dep/max(len(fp),1):.0f}%) mufu {n_mufu} local {n_l}")
