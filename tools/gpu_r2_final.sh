#!/bin/bash
# Final 1-GPU evidence pass of round 2 (B200_PROFILING.md recipe).
#   1. full GPU suite
#   2. bench lines: c2 (default, with the CPU arm's sample), c2 zipf, c3, c4 (with e2e), the reference arm
#   3. ncu launch lists of one step per config (shares) + --set full captures of the dominant kernels, reduced ON THE BOX
#      to csv summaries (multi-launch .ncu-rep files exceed gpurun's 64 MiB return limit)
# Numbers printed by bench.py under ncu are never bench values.
mkdir -p gpurun_out
R=/tmp/ncu_reports; mkdir -p $R
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_tests.log
timeout 900 python bench.py > gpurun_out/r2z_c2.json 2> gpurun_out/r2z_c2.err
timeout 600 python bench.py --dist zipf --no-cpu-baseline > gpurun_out/r2z_c2_zipf.json 2> gpurun_out/r2z_c2_zipf.err
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2z_c3.json 2> gpurun_out/r2z_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2z_c4.json 2> gpurun_out/r2z_c4.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2z_reference_c2.json 2> gpurun_out/r2z_reference_c2.err
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
for c in c2 c3 c4; do
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/r2_launches_$c.csv $B --config $c > gpurun_out/r2z_prof_$c.log 2>&1
done
M='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_subpipe_hmma_cycles_active|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|smsp__average_warps_issue_stalled_(long_scoreboard|barrier|math_pipe_throttle|mio_throttle|short_scoreboard|wait|no_instruction|membar|lg_throttle|dispatch|branch_resolving|sleeping)|l1tex__t_sector_hit_rate|smsp__issue_active.avg.pct|sm__cycles_active.avg |l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum|l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'
cap() {  # name, kernel regex, skip, count, config, extra flags
  timeout 900 $NCU --set full $6 -k regex:$2 -s $3 -c $4 -f -o $R/$1 $B --config $5 >> gpurun_out/r2z_prof_$5.log 2>&1
  ncu -i $R/$1.ncu-rep --page raw --csv > $R/$1_raw.csv 2>/dev/null
  python - $R/$1_raw.csv gpurun_out/$1_summary.csv "$M" <<'PY'
import csv, re, sys
src, dst, pat = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
rows = list(csv.reader(open(src)))
if len(rows) >= 3:
    hdr = rows[0]
    keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or pat.search(h)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] if i < len(r) else "" for i in keep])
PY
}
# (B2CTR_STEP_GRAPH=0: eager launches, so that -s / -c count this process' kernels in program order)
export B2CTR_STEP_GRAPH=0
cap r2_full_gather gather_uniform_fwd 4 1 c2 "--import-source on"
cap r2_full_scatter scatter_uniform_bwd 4 1 c2 "--import-source on"
cap r2_full_gemm gemm_planes_ws 27 9 c2 ""
cap r2_full_cin gemm_planes_ws 45 15 c3 ""
cap r2_full_att gemm_planes_ws 45 15 c4 ""
# per-instruction stall profile (source page, SASS view) of the first CIN launch of the captured step
ncu -i $R/r2_full_cin.ncu-rep --page source --csv 2>/dev/null | head -c 2500000 > gpurun_out/r2_full_cin_source.csv
du -sh gpurun_out; tail -3 gpurun_out/r2z_tests.log
