"""Attribute executed warp instructions of an ncu report to CUDA source lines.

ncu's CSV export of the source page is SASS-only; this joins it (by instruction order) with the
line table nvdisasm prints for the SAME build of libmcdp_b200.so.

    python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep [kernel-mangled-substring] [top_n]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.environ.get("MCDP_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # MCDP_ROOT: a checkout of the captured build

def sass_lines(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"): continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, ("?", 0)
        for line in txt.splitlines():
            m = re.match(r"^\.text\.(\S+):", line)
            if m: cur_fn = m.group(1); continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and cur_fn and kernel_sub in cur_fn:
                out.append((int(m.group(1), 16), cur_line, m.group(2).strip()))
    return out

def main():
    rep = sys.argv[1]
    ksub = sys.argv[2] if len(sys.argv) > 2 else "sweep_kernelILi0ELb1"
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; data = rows[2:]
    ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    sl = sass_lines(os.path.join(ROOT, "mc_dagprop_b200", "libmcdp_b200.so"), ksub)
    if len(sl) != len(data):
        print(f"warning: instruction count mismatch sass={len(sl)} ncu={len(data)} (different build?)")
    n = min(len(sl), len(data))
    tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
    by = collections.Counter(); bys = collections.Counter(); byfile = collections.Counter()
    for i in range(n):
        key = sl[i][1]
        by[key] += int(data[i][ia]); bys[key] += int(data[i][isamp]); byfile[key[0]] += int(data[i][ia])
    print("total warp instr %.4g" % tot)
    print({k: f"{v/tot*100:.1f}%" for k, v in byfile.most_common()})
    cache = {}
    def text(f, l):
        for d in ("mc_dagprop_b200/csrc", "include"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in cache: cache[p] = open(p).read().splitlines()
                return cache[p][l - 1].strip()[:90] if 0 < l <= len(cache[p]) else ""
        return ""
    for (f, l), c in by.most_common(topn):
        print(f"{f[:20]:20s} {l:4d} {c/tot*100:5.1f}% inst {bys[(f,l)]/max(tots,1)*100:5.1f}% smp  {text(f,l)}")

if __name__ == "__main__":
    main()
