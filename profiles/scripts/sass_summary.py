"""Per-kernel SASS mnemonic counts of the in-tree product library (cuobjdump -sass; no GPU needed) and the ptxas
resource lines of the same build.  Mnemonics per /opt/skills/guides/B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = bulk
copies (cp.async.bulk), SYNCS = mbarrier, LDGSTS = cp.async, LDG.E.128 = 128-bit global loads.

    python profiles/scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
PKG = next(p for p in ROOT.iterdir() if p.is_dir() and p.name.endswith("_b200"))
LIB = PKG / "libcrdpn_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG.E.128",
         "LDG.E.ENL2.256", "SHFL", "MUFU", "FFMA2", "HMMA", "ATOM", "RED", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (("." in w) and op.startswith(w)):
                per[cur][w] += 1
    names = demangle(list(per))
    print(f"# SASS summary of {LIB.name} (sm_100a), {len(per)} kernels; counts are static instructions")
    print("# columns: total | " + " ".join(WATCH))
    tc = []
    for k, c in per.items():
        short = re.sub(r"\(.*", "", names.get(k, k))
        short = re.sub(r"^void ", "", short)
        if len(short) > 90:
            short = short[:87] + "..."
        cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short:<92} total={c['total']:<6} {cols}")
        if c["UTCHMMA"] or c["UTCQMMA"]:
            tc.append(short)
    print(f"\n# kernels issuing tcgen05.mma (UTCHMMA/UTCQMMA): {len(tc)}")
    for t in tc:
        print("#   " + t)
    log = PKG / "build" / "ptxas.log"
    if log.exists():
        print("\n# ptxas -v resource lines with spills or > 200 registers (all others: no spills)")
        fn = None
        for line in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = m.group(1)
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and (int(m.group(1)) or int(m.group(2))):
                print(f"#   SPILL {fn}: {line.strip()}")
            m = re.search(r"Used (\d+) registers", line)
            if m and int(m.group(1)) > 200:
                print(f"#   REGS  {fn}: {line.strip()}")


if __name__ == "__main__":
    sys.exit(main())
