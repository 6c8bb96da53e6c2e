"""Summarise an .ncu-rep (raw page + source page) without a GPU: key metrics, opcode mix, hot regions."""
import collections, csv, io, json, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']

def main(rep, out_json=None):
    hdr, units, rows = raw(rep)
    vals = rows[0]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr) if h in KEYS}
    stalls = {h: vals[i] for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('per_issue_active.ratio') or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('.ratio')}
    for k in sorted(d): print(f"{k:75s} {d[k][0]:>20s} {d[k][1]}")
    top = sorted(((float(v), k) for k, v in stalls.items()), reverse=True)[:8]
    for v, k in top: print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; data = rows[2:]
    ia, isrc, isamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
    byop = collections.Counter(); bys = collections.Counter()
    for r in data:
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0]
        byop[op] += int(r[ia]); bys[op] += int(r[isamp])
    print(f"total warp instructions {tot:.4g}; samples {tots}")
    for op, c in byop.most_common(22): print(f"  {op:10s} {c/tot*100:6.2f}% inst   {bys[op]/max(tots,1)*100:6.2f}% samples")
    if out_json:
        json.dump({"metrics": d, "stalls": dict((k, v) for v, k in top), "opcode_mix_pct": {op: c/tot*100 for op, c in byop.most_common(25)}, "warp_instructions": tot}, open(out_json, "w"), indent=1)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
