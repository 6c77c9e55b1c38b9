"""Per-edge dynamic instruction mix of a kernel from an ncu report (source page with SASS).
usage: python tools/sass_mix.py report.ncu-rep <kernel regex> <edges> [--top N]"""
import collections, csv, re, subprocess, sys


def main():
    rep, rx, edges = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    seen, data = set(), []
    for r in rows[hi + 1:]:
        if len(r) > iex and r[ia].startswith("0x") and r[ia] not in seen:
            seen.add(r[ia])
            data.append((int(r[ia], 16), int(r[iex] or 0), r[isrc].strip(), int(r[isamp] or 0)))
    base = data[0][0]
    tot, samp = sum(d[1] for d in data), sum(d[3] for d in data)
    print(f"{rx}: {tot / edges:.1f} warp-instructions per edge, {samp} samples")
    c = collections.Counter()
    for a, e, src, s in data:
        c[re.sub(r"^@!?U?P\d\s+", "", src).split()[0].split(".")[0]] += e / edges
    print("  " + ", ".join(f"{k} {v:.1f}" for k, v in c.most_common(30)))
    if "--top" in sys.argv:
        n = int(sys.argv[sys.argv.index("--top") + 1])
        for a, e, src, s in sorted(data, key=lambda d: -d[3])[:n]:
            print(f"   {100 * s / samp:5.1f}%  ex/edge={e / edges:5.2f}  +{a - base:04x}  {src[:100]}")


if __name__ == "__main__":
    main()
