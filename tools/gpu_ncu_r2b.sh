#!/bin/bash
# second ncu pass: the kernels the first window missed (one launch each), text summaries only; plus the bf16 parity table
mkdir -p gpurun_out
python tools/bf16_table.py > gpurun_out/bf16_table.log 2>&1; echo "table exit $?"
CMD="python bench.py --no-graph --steps 1 --warmup 1 --quick"
$CMD > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
i=0
for k in attn_step_fwd aux_loss colsum_vec attn_post_reg kd_token_loss_pipe; do
  ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -f -o /tmp/prof_r2_d$i $CMD > gpurun_out/ncu_r2_d$i.log 2>&1
  ncu -i /tmp/prof_r2_d$i.ncu-rep --page raw --csv > gpurun_out/r2_ncu_d${i}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_r2_d$i.ncu-rep --page details > gpurun_out/r2_ncu_d${i}_details.txt 2>/dev/null
  i=$((i+1))
done
ls gpurun_out/r2_ncu_d*raw.csv | wc -l
