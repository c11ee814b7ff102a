"""profiles/r02_sass_hot.txt: for every kernel of the library, registers / spills (cuobjdump -res-usage) and counts of the
SASS mnemonics that show which hardware paths the code uses.  python scripts/sass_hot.py > profiles/r02_sass_hot.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pose_refine_b200", "libpose_refine_b200.so")
KEYS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "LDG.E.ENL2.256", "LDG.E.128", "LDG", "LDS.128", "STS.128", "UBLKCP", "SYNCS", "MEMBAR",
        "UCGABAR_ARV", "ST.E.ASYNC|STAS", "ATOMS", "ATOMG|RED", "REDUX", "SHFL", "MUFU.RCP", "BAR.SYNC", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in ln:
        usage[cur] = ln.strip()
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            funcs[cur].append(m.group(1))
names = demangle(list(funcs))
print("# SASS of pose_refine_b200/libpose_refine_b200.so (sm_100a), per kernel: resource usage and mnemonic counts")
print("# made by scripts/sass_hot.py; FFMA2/FMUL2/FADD2 = two-wide FP32, LDG.E.ENL2.256 = 32-byte gathers, UBLKCP = TMA bulk copy,")
print("# SYNCS = mbarrier ops, ST.E.ASYNC/STAS = st.async into a peer CTA's shared memory, UCGABAR = barrier.cluster\n")
for f, ins in funcs.items():
    if len(ins) < 40:
        continue
    short = re.sub(r"\(.*", "", names[f])
    print(f"== {short}   [{len(ins)} instructions]  {usage.get(f, '')}")
    row = []
    for k in KEYS:
        pat = re.compile(r"(^|\s)(@!?U?P\d+\s+)?(" + k.replace(".", r"\.") + r")(\.|\s|$)")
        n = sum(1 for i in ins if pat.search(i))
        if k == "FFMA":
            n = sum(1 for i in ins if re.search(r"(^|\s)FFMA(\.|\s)", i))
        if k == "LDG":
            n = sum(1 for i in ins if re.search(r"(^|\s)LDG\.", i))
        if n:
            row.append(f"{k}={n}")
    print("   " + "  ".join(row))
