"""SASS evidence for profiles/: per-kernel counts of the tensor-core / TMA / bulk-copy / cp.async / FP64 instructions in the
shipped library, plus the first occurrence of each tcgen05 / TMA mnemonic.  python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-tfhe_b200", "lib", "libtfhe_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WHOLE = ["UTCIMMA", "UTMALDG", "UBLKPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "LDG", "LDS",
         "STS", "REDG", "ATOMG", "BAR", "MEMBAR", "PRMT", "SHFL"]
KEY = ["UTCIMMA", "UTMALDG", "UBLKPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL"]
per, first, cur = collections.OrderedDict(), {}, None
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        per[cur][op] += 1
        if op in KEY[:10] and op not in first:
            first[op] = (cur, re.sub(r"\s+", " ", line.split("/*")[1].split("*/")[1]).strip() if "/*" in line else line.strip())
tot = collections.Counter()
for c in per.values(): tot.update(c)
print("# SASS evidence, round 2 — cuobjdump -sass %s (default build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)\n" % os.path.relpath(lib))
print("architectures in the fatbin:", ", ".join(archs), "\n")
print("whole library, instruction counts of interest:")
for k in WHOLE:
    if tot[k]: print("  %-10s %d" % (k, tot[k]))
print("\nfirst occurrence of each tensor-core / TMA / bulk-copy mnemonic:")
for k, (fn, ins) in first.items(): print("  %-10s %s\n             in %s" % (k, ins, fn))
print("\nper kernel (only kernels with tensor-core / TMA / bulk-copy / cp.async / FP64 instructions):\n")
for fn in sorted(per):
    parts = ["%s x%d" % (k, per[fn][k]) for k in KEY if per[fn][k]]
    if parts: print("  %s\n     %s\n" % (fn, ", ".join(parts)))
