"""Per-kernel SASS mnemonic counts of libgpbo.so (DMMA, UTMALDG, LDGSTS, SYNCS, MUFU, ...) and resource usage,
written as the table kept under profiles/ (evidence for which kernels use the FP64 tensor pipe, TMA and mbarriers)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gp-bayesopinf_b200/libgpbo.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in ln:
        usage[cur] = dict(kv.split(":") for kv in ln.split() if ":" in kv)
counts = collections.OrderedDict()
cur = None
WANT = ("DMMA", "UTMALDG", "LDGSTS", "SYNCS", "LDS", "STS", "DFMA", "MUFU", "BAR", "STL", "LDL")
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1).split(".")[0]
        if op in WANT:
            counts[cur][op] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass), registers / static shared bytes (cuobjdump --dump-resource-usage)")
print("# " + " ".join(f"{w:>8s}" for w in WANT) + "      REG  SHARED  kernel")
tot = collections.Counter()
for k, c in counts.items():
    u = usage.get(k, {})
    name = re.sub(r"^gpbo::", "", demangle(k)).split("(")[0]
    print("  " + " ".join(f"{c[w]:8d}" for w in WANT) + f"  {u.get('REG', '?'):>6s} {u.get('SHARED', '?'):>7s}  {name}")
    tot.update(c)
print("# total " + " ".join(f"{w}={tot[w]}" for w in WANT))
