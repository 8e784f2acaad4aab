"""SASS opcode histogram of the in-tree library (evidence for profiles/: TMA / mbarrier / FP64 opcodes per kernel family).
usage: python tools/sass_hist.py > profiles/r02_sass.md   (runs cuobjdump -sass on fft-implementation-in-c_b200/lib/libfft_b200.so)"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "fft-implementation-in-c_b200", "lib", "libfft_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fam = collections.defaultdict(collections.Counter)
nk = collections.Counter()
cur = None
archs = collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"<.*", "", d.replace("fftb200::", "").replace("void ", "")).split("(")[0]
        nk[cur] += 1
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        archs[m.group(1)] += 1
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        fam[cur][m.group(1).split(".")[0]] += 1
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "UTCATOMSWS", "ATOMG", "RED", "MEMBAR", "FENCE", "LDL", "STL"]
print("# r02 - SASS opcode histogram of fft-implementation-in-c_b200/lib/libfft_b200.so (tools/sass_hist.py, cuobjdump -sass)\n")
print("cubins by architecture:", dict(archs), "\n")
print("Instances = template instantiations of the kernel in the library; counts are static instructions summed over them. UBLKCP = 1-D bulk async copy, UTMALDG / UTMASTG = TMA tensor")
print("load / store, SYNCS = mbarrier operations, LDL / STL = local-memory (spill) traffic, ATOMG = global atomics (dependency counters of the fused kernel, tile hand-out counters of the ring")
print("kernels). No UTC*MMA / HMMA: FP64 Stockham butterflies have no tensor-core path on sm_100a. LDTM / STTM / UTCATOMSWS = tensor-memory loads / stores / allocation: only in the opt-in")
print("8192-point variant fft_pipe13t_kernel, which uses tensor memory as a parking space, not for MMA accumulators.\n")
print("| kernel | instances | " + " | ".join(KEY) + " | total |")
print("|---|---|" + "|".join(["---"] * (len(KEY) + 1)) + "|")
for k in sorted(fam, key=lambda k: -sum(fam[k].values())):
    c = fam[k]
    print("| `%s` | %d | " % (k, nk[k]) + " | ".join(str(c.get(o, 0)) for o in KEY) + " | %d |" % sum(c.values()))
tot = collections.Counter()
for c in fam.values():
    tot.update(c)
print("\nLibrary totals: " + ", ".join("%s %d" % (o, tot[o]) for o in KEY if tot[o]))
