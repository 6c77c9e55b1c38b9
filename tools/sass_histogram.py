"""SASS opcode histogram per kernel of libjmac_b200.so (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_histogram.py [lib.so] > profiles/sass_opcodes_rNN.md

Only the opcodes that identify the hardware path are listed per kernel: tcgen05 MMA (UTC*MMA), TMEM loads
(LDTM), TMA tensor loads (UTMALDG), bulk copies (UBLKCP), mbarrier waits (SYNCS), vector REDs, 128/256-bit
global accesses, shuffles; plus the total instruction count."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "RED", "ATOMG", "LDG.E.128", "LDG.E.ENL2.256",
        "STG.E.128", "STG.E.ENL2.256", "LDS.128", "SHFL", "HMMA", "FFMA", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jmac_b200", "libjmac_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            hist[kern]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                    hist[kern][k] += 1
    names = demangle(list(hist))
    print("# SASS opcode histogram per kernel (`cuobjdump -sass jmac_b200/libjmac_b200.so`, sm_100a)\n")
    print("Static instruction counts. `UTCHMMA` = tcgen05.mma (kind::f16 / kind::tf32), `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load,")
    print("`UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier try_wait / arrive, `UTCBAR` = tcgen05.commit.\n")
    used = [k for k in KEYS if any(h[k] for h in hist.values())]
    print("| kernel | instrs | " + " | ".join(used) + " |")
    print("|---|---|" + "---|" * len(used))
    for k, h in hist.items():
        nm = re.sub(r"\(.*", "", names.get(k, k))
        nm = nm.replace("jmac::", "").replace("(anonymous namespace)::", "")
        print(f"| `{nm}` | {h['_total']} | " + " | ".join(str(h[u]) if h[u] else "" for u in used) + " |")
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("| **all kernels** | " + str(tot["_total"]) + " | " + " | ".join(str(tot[u]) if tot[u] else "" for u in used) + " |")


if __name__ == "__main__":
    main()
