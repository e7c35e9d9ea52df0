#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + instruction/stall share per source function.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [-k kernel_regex] [kernel_source_file ...]"""
import collections, csv, io, re, subprocess, sys

KFILTER = []

def run(args):
    return subprocess.run(["ncu", "-i"] + args + KFILTER, capture_output=True, text=True).stdout

def main():
    rep = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "-k":
        KFILTER.extend(["-k", "regex:" + sys.argv[3]])
        del sys.argv[2:4]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    d = {h: (u, v) for h, u, v in zip(raw[0], raw[1], raw[2])}
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__t_sectors_srcunit_tex_op_read.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
    print("== %s" % d.get("Kernel Name", ("", "?"))[1])
    for k in keys:
        if k in d:
            print("%-70s %-12s %s" % (k, d[k][0], d[k][1]))
    print("-- warp stall reasons (per issue-active cycle)")
    st = [(float(v[1]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for k, v in d.items() if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")]
    for v, k in sorted(st, reverse=True)[:9]:
        print("   %-22s %.3f" % (k, v))
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, agg = None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 8 and r[0] not in ("", "Line No", "Function Name") and r[2] == "-":
            try:
                agg.append((cur, int(r[0]), r[1].strip(), int(r[4]), int(r[7])))
            except ValueError:
                pass
    tot_s = sum(a[3] for a in agg) or 1
    tot_i = sum(a[4] for a in agg) or 1
    funcs = {}
    for path in sys.argv[2:]:
        name = path.split("/")[-1]
        fl = []
        for n, l in enumerate(open(path).read().split("\n"), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|LG_HD).*?\b([A-Za-z_0-9]+)\s*\(", l)
            if m and not l.startswith(" "):
                fl.append((n, m.group(1)))
        funcs[name] = fl
    def phase(f, l):
        if f in funcs:
            name = f
            for n, fn in funcs[f]:
                if l >= n:
                    name = fn
            return name
        return f
    ph = collections.defaultdict(lambda: [0, 0])
    for f, l, src, sm, ins in agg:
        k = phase(f, l)
        ph[k][0] += sm
        ph[k][1] += ins
    print("-- by source function: stall-sample share, warp-instructions (share)")
    for k, v in sorted(ph.items(), key=lambda kv: -kv[1][0]):
        if v[0] * 200 < tot_s and v[1] * 200 < tot_i:
            continue
        print("   %-28s samples %5.1f%%   inst %8.1fM (%4.1f%%)" % (k, 100 * v[0] / tot_s, v[1] / 1e6, 100 * v[1] / tot_i))
    print("-- top source lines by stall samples")
    for a in sorted(agg, key=lambda a: -a[3])[:25]:
        print("   %-22s %4d  s=%4.1f%%  i=%6.1fM  %s" % (a[0], a[1], 100 * a[3] / tot_s, a[4] / 1e6, a[2][:80]))

if __name__ == "__main__":
    main()
