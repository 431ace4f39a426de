#!/bin/bash
# pass E: bench C3 / C4 on the final kernels, then ncu evidence (B200_PROFILING.md recipe): launch lists of one step per
# config + --set full captures of the dominant kernels, reduced ON THE BOX to csv summaries (the .ncu-rep files of
# multi-launch captures exceed gpurun's 64 MiB return limit).  Numbers printed by bench.py under ncu are never bench values.
mkdir -p gpurun_out
R=/tmp/ncu_reports; mkdir -p $R
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
for c in c2 c3 c4; do
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/r2_launches_$c.csv $B --config $c > gpurun_out/r2_prof_$c.log 2>&1
done
M='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_subpipe_hmma_cycles_active|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|smsp__average_warp.*stall|smsp__average_warps_issue_stalled_(long_scoreboard|barrier|math_pipe_throttle|mio_throttle|short_scoreboard|wait|no_instruction|membar|lg_throttle|dispatch)|l1tex__t_sector_hit_rate|smsp__issue_active.avg.pct|sm__cycles_active.avg '
cap() {  # name, kernel regex, skip, count, config, extra flags
  timeout 900 $NCU --set full $6 -k regex:$2 -s $3 -c $4 -f -o $R/$1 $B --config $5 >> gpurun_out/r2_prof_$5.log 2>&1
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
cap r2_full_gather gather_uniform_fwd 4 1 c2 "--import-source on"
cap r2_full_scatter scatter_uniform_bwd 4 1 c2 "--import-source on"
cap r2_full_gemm gemm_planes_ws 36 9 c2 ""
cap r2_full_cin gemm_planes_ws 40 10 c3 ""
cap r2_full_att gemm_planes_ws 30 8 c4 ""
# per-instruction stall profile of the generated-operand CIN kernel (source page of the first captured launch)
ncu -i $R/r2_full_cin.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/r2_full_cin_source.csv
cp $R/r2_full_gather.ncu-rep $R/r2_full_scatter.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
