#!/bin/bash
# round 2: ncu --set full of the kernels VERDICT r1 asked for (attention steps, LSTM-epilogue GEMM, aux loss, cell adjoint, column sums,
# vocabulary head) inside a real eagerly issued KD step; the same command runs plainly first.  The reports stay on the box (they exceed
# the 64 MiB return limit with sources); the raw metric tables and the details pages come back as text.
mkdir -p gpurun_out
CMD="python bench.py --no-graph --steps 1 --warmup 1 --quick"
$CMD > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
ncu --set full --clock-control none -k regex:'attn_step_fwd|attn_step_bwd|aux_loss|lstm_pointwise_bwd|colsum_vec|attn_post' -s 40 -c 14 -f -o /tmp/prof_r2_a $CMD > gpurun_out/ncu_r2_a.log 2>&1
echo "ncu A exit $?"
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, float, \(bool\)1>' -s 20 -c 2 -f -o /tmp/prof_r2_b $CMD > gpurun_out/ncu_r2_b.log 2>&1
echo "ncu B exit $?"
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)256, \(bool\)0, \(bool\)0, __nv_bfloat16' -s 3 -c 2 -f -o /tmp/prof_r2_c $CMD > gpurun_out/ncu_r2_c.log 2>&1
echo "ncu C exit $?"
for x in a b c; do
  ncu -i /tmp/prof_r2_$x.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${x}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_r2_$x.ncu-rep --page details > gpurun_out/r2_ncu_${x}_details.txt 2>/dev/null
done
ls -la gpurun_out/r2_ncu_* | awk '{print $5, $9}'
