"""Opcode census of the in-tree library (no GPU needed): per kernel, the SASS mnemonics that prove which hardware
paths the build uses - tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG), mbarrier (SYNCS), clusters, and
the gather's memory instructions.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3dahv_b200", "lib3dahv_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "ELECT",
         "LDS", "STS", "LDG", "STG", "ATOM", "RED", "SHFL", "HFMA2", "FFMA", "HMMA", "DFMA", "MUFU", "F2FP", "BAR", "ACQBULK", "CCTL",
         "MEMBAR", "FENCE", "ERRBAR", "NANOSLEEP")
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["TOTAL"] += 1
        base = op.split(".")[0]
        if base in WATCH:
            kernels[cur][base] += 1
            if base in ("LDS", "STS", "LDG", "STG", "UTCHMMA", "LDTM", "STTM", "SYNCS", "UTMALDG", "UTCBAR"):
                kernels[cur][op] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a; opcode counts are static instruction counts per kernel)")
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
print("# ELF images:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
for k, c in kernels.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
    print(f"\n{name}  [{c['TOTAL']} instructions]")
    items = [(op, n) for op, n in sorted(c.items()) if op != "TOTAL"]
    print("  " + "  ".join(f"{op}={n}" for op, n in items))
