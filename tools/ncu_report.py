#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report of the tile kernels: per captured launch the headline
metrics, the LSU-wavefront budget (the unit that bounds these kernels) and the per-phase instruction mix / stall reasons
(phases = code between BAR.SYNCs).  usage: python tools/ncu_report.py REPORT.ncu-rep PIXELS_PER_LAUNCH [launch indices]"""
import csv, io, re, subprocess, sys
from collections import Counter

HEAD = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(rep, pixels, which):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(src) if l.startswith('"Kernel Name"')]
    for w in (which or range(len(rows))):
        r = rows[w]
        col = {h: r[i] for i, h in enumerate(hdr)}
        print(f"== launch {w}: {col.get('Kernel Name', '')[:90]}  grid {col.get('Grid Size')}")
        for h in HEAD:
            if h in col:
                print(f"   {h:84s} {col[h]} {units[hdr.index(h)]}")
        lg = f(col.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "0"))
        sh = f(col.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "0"))
        if lg or sh:
            n_sm = 148
            print(f"   LSU wavefronts per pixel: shared {sh * n_sm / pixels:.2f} + global/local {lg * n_sm / pixels:.2f}")
        # source page of the same launch: the page does not list the launches in the order of the raw page, so take the
        # block whose executed-instruction total matches this launch's smsp__inst_executed.sum
        want = f(col.get("smsp__inst_executed.sum", "0"))
        best, best_err = None, None
        for bi in range(len(starts)):
            a, b = starts[bi], starts[bi + 1] if bi + 1 < len(starts) else len(src)
            blk = list(csv.reader(io.StringIO("\n".join(src[a:b]))))
            if len(blk) < 3 or "Instructions Executed" not in blk[1]:
                continue
            wi = blk[1].index("Instructions Executed")
            tot = sum(f(x[wi]) for x in blk[2:] if len(x) > wi)
            err = abs(tot - want)
            if best_err is None or err < best_err:
                best, best_err = (a, b), err
        if best is not None:
            a, b = best
            srows = list(csv.reader(io.StringIO("\n".join(src[a:b]))))
            sh_, data = srows[1], [x for x in srows[2:] if len(x) > 5]
            ix = {h: i for i, h in enumerate(sh_)}
            s_, smp, thr = ix["Source"], ix["# Samples"], ix["Thread Instructions Executed"]
            stalls = [h for h in sh_ if h.startswith("stall_") and "Not Issued" not in h]
            bars = [i for i, x in enumerate(data) if "BAR.SYNC" in x[s_]]
            bounds = [0] + bars + [len(data)]
            tot_s = sum(f(x[smp]) for x in data) or 1.0
            tot_i = sum(f(x[thr]) for x in data)
            print(f"   thread instructions {tot_i:.3e} = {tot_i / pixels:.1f} per pixel; {tot_s:.0f} stall samples")
            for k, (p, q) in enumerate(zip(bounds[:-1], bounds[1:])):
                seg = data[p:q]
                ti = sum(f(x[thr]) for x in seg)
                if ti / pixels < 0.5:
                    continue
                mix = Counter()
                for x in seg:
                    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", x[s_])
                    if m:
                        mix[m.group(1)] += f(x[thr])
                ss = sum(f(x[smp]) for x in seg)
                agg = sorted(((s[6:], sum(f(x[ix[s]]) for x in seg)) for s in stalls), key=lambda kv: -kv[1])[:5]
                print(f"   segment {k}: {ti / pixels:6.1f} instr/px, {100 * ss / tot_s:5.1f} % of stall samples; "
                      + " ".join(f"{n}={100 * v / max(ss, 1):.0f}%" for n, v in agg))
                print("       " + " ".join(f"{n} {v / pixels:.1f}" for n, v in mix.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), [int(x) for x in sys.argv[3:]])
