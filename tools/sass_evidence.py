#!/usr/bin/env python
"""Which Blackwell / SIMD / synchronisation instructions each kernel of libgofpid_b200.so contains (SASS mnemonics,
`cuobjdump -sass`): the tile-movement and byte-SIMD proof that belongs next to the ncu numbers.

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt

UBLKCP = cp.async.bulk (TMA, 1-D), SYNCS = mbarrier, VABSDIFF4 / VIMNMX / IDP = byte / 16-bit SIMD and dp4a,
REDUX / MATCH = warp-aggregated reductions, ATOMS / ATOMG / RED = shared / global atomics, MUFU.RCP + FFMA without
FCHK = the range-check-free division cores of the lean MOG2 kernel.  No UTC*MMA / HMMA is expected: nothing on this
path is a contraction."""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pygofpid_b200", "libgofpid_b200.so")
PAT = re.compile(r'\b(UBLKCP\S*|UTMA\S*|SYNCS\S*|VABSDIFF4\S*|VIMNMX3?\S*|IDP\S*|REDUX\S*|MATCH\S*|ATOMS\S*|ATOMG\S*|RED\.\S*|'
                 r'MUFU\.RCP|FCHK|SHFL\S*|BAR\.SYNC\S*|LDS\.128|STS\.128|LDG\.E\.128|STG\.E\.128|LDG\.E\.64|UTC\S*MMA|HMMA\S*)')
GROUP = ('SHFL', 'BAR', 'REDUX', 'MATCH', 'ATOMS', 'ATOMG', 'RED', 'SYNCS', 'UBLKCP', 'VIMNMX3', 'VIMNMX', 'IDP', 'VABSDIFF4')


def main():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for f in sorted(os.listdir(d)):
            if not f.endswith(".cubin"):
                continue
            out = subprocess.run(["cuobjdump", "-sass", os.path.join(d, f)], capture_output=True, text=True).stdout
            fn, cnt, tot = None, collections.OrderedDict(), {}
            for l in out.split("\n"):
                m = re.search(r"Function : (\S+)", l)
                if m:
                    fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                    cnt[fn], tot[fn] = collections.Counter(), 0
                    continue
                if fn and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
                    tot[fn] += 1
                    mm = PAT.search(l)
                    if mm:
                        k = mm.group(1)
                        for g in GROUP:
                            if k.startswith(g):
                                k = g
                                break
                        cnt[fn][k] += 1
            if cnt:
                print("== %s" % f)
            for fn, c in cnt.items():
                print("%-44s %5d SASS instr  %s" % (fn.replace("gofpid::", "").replace("void ", "")[:44], tot[fn],
                                                    "  ".join("%s x%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    main()
