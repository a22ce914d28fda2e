#!/bin/bash
# ncu --set full capture of one Q-network rollout kernel (tools/bench_qnet.py 131072), reduced on the GPU box to text:
# <tag> <kernel regex, e.g. k_qnet_ego_tc> <mangled substring for the source-line table>
tag=$1; kern=$2; sub=$3
rep=/tmp/$tag.ncu-rep
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$kern --launch-skip 4 --launch-count 1 -f -o /tmp/$tag python tools/bench_qnet.py 131072 > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_summary.py $rep > gpurun_out/${tag}_raw.txt 2>&1
src=topotrafficrl_b200/csrc/build; [ -d $src ] || src=topotrafficrl_b200/csrc/libttrl_b200.so
python tools/ncu_lines.py $rep $src $sub 70 > gpurun_out/${tag}_lines.txt 2>&1
ls -la $rep >> gpurun_out/${tag}_ncu.log; [ $(stat -c%s $rep) -lt 30000000 ] && cp $rep gpurun_out/; rm -f $rep
