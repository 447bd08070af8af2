#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS + stall samples) with nvdisasm line info of the shipped cubin,
so stall samples can be attributed to CUDA source lines.  Usage:
    python tools/sass_lines.py <report.ncu-rep> <kernel-substring> [top]
The library must be the build the report was captured with."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparsefactorization_b200", "_lib", "libpsf_b200.so")


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("psf_chain_sm100")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    table, cur_line, infn, idx = [], None, False, 0
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            infn = kernel_sub in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
    return table


def main(rep, kernel_sub, top=30):
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                                          text=True).stdout.splitlines()))
    h, d = rows[1], rows[2:]
    iS, iA = h.index("Warp Stall Sampling (All Samples)"), h.index("Address")
    cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    base = int(d[0][iA], 16)
    table = {off: (line, txt) for off, line, txt in line_table(kernel_sub)}
    src = open(os.path.join(ROOT, "sparsefactorization_b200", "csrc", "psf_chain_sm100.cu")).read().splitlines()
    per_line = collections.defaultdict(lambda: [0, collections.Counter()])
    total = 0
    for r in d:
        off = int(r[iA], 16) - base
        n = int(r[iS])
        total += n
        line = table.get(off, (None, ""))[0]
        per_line[line][0] += n
        for i in cols:
            if r[i] not in ("", "0"):
                per_line[line][1][h[i]] += int(r[i])
    print("total samples", total)
    for line, (n, reasons) in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
        text = src[line - 1].strip()[:90] if line else "?"
        rs = ", ".join(f"{k.replace('stall_', '')}:{v}" for k, v in reasons.most_common(3))
        print(f"{100 * n / total:5.1f}%  L{line}: {text:90s} [{rs}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
