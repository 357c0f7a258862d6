"""Per-kernel counts of the Blackwell-specific SASS in libgpujoin.so (cuobjdump -sass): TMA bulk copies (UBLKCP),
mbarrier traffic (SYNCS), async-proxy fences, shared atomics, plus registers / shared memory from -res-usage.
usage: python tools/sass_summary.py > profiles/<round>/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "icde2019-gpu-join_b200", "lib", "libgpujoin.so")
PAT = ["UBLKCP.S.G", "UBLKCP.G.S", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "SYNCS.EXCH", "FENCE.VIEW.ASYNC", "ATOMS", "ATOMG", "RED.",
       "LDG.E.128", "LDG.E.64", "STG.E.64", "STG.E.128", "BAR.SYNC", "MATCH", "VOTE", "UTMALDG", "UTCMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur and "/*" in ln:
            for p in PAT:
                if p in ln:
                    counts[cur][p] += 1
            counts[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", ln) else 0
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    dm = demangle(order)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(order)} kernels (sm_100a)")
    print("# columns: instructions, registers, static shared bytes, then non-zero counts of " + ", ".join(PAT))
    for f in sorted(order, key=lambda x: dm[x]):
        name = re.sub(r"^void gj::", "", dm[f])
        name = re.sub(r"\(.*\)$", "", name)
        c = counts[f]
        r = usage.get(f, ("?", "?"))
        extra = "  ".join(f"{p}={c[p]}" for p in PAT if c[p])
        print(f"{name:95s} ins={c['instructions']:5d} reg={r[0]} smem={r[1]}  {extra}")
    tot = collections.Counter()
    for f in order:
        tot.update(counts[f])
    print("# totals: " + "  ".join(f"{p}={tot[p]}" for p in PAT if tot[p]))


if __name__ == "__main__":
    sys.exit(main())
