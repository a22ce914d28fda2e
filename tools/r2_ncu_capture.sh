#!/bin/bash
# ncu --set full capture of one kernel of a bench.py workload, reduced ON THE GPU BOX to text summaries (the .ncu-rep files
# exceed what gpurun copies back): <tag> <kernel mangled substring, e.g. k_stepILi50ELi1> <launch-skip> <bench args...>
tag=$1; kern=$2; skip=$3; shift 3
rep=/tmp/$tag.ncu-rep
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip $skip --launch-count ${NCU_COUNT:-2} -f -o /tmp/$tag python bench.py "$@" --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_summary.py $rep > gpurun_out/${tag}_raw.txt 2>&1
src=topotrafficrl_b200/csrc/build; [ -d $src ] || src=topotrafficrl_b200/csrc/libttrl_b200.so   # the build directory does not travel to the GPU box
python tools/ncu_lines.py $rep $src $kern 60 > gpurun_out/${tag}_lines.txt 2>&1
ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum','gcc__raw_requests','sm__icc_requests','smsp__warp_issue_stalled_no_instruction','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
idx=[i for i,h in enumerate(hdr) if any(w in h for w in want)]
for r in rows:
    print(' | '.join(r[i] for i in idx))
" > gpurun_out/${tag}_metrics.txt
rm -f $rep
