#!/bin/bash
# round-2 closing pass: every GPU test, smoke(), the default bench line, the reference arm, the ncu launch list of an eager step and
# ncu --set full of the kernels that changed this round (LSTM-epilogue GEMM with the fp32 recurrent addend, merged [u | W_hh h] GEMM,
# the opt-in cluster recurrence kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/t_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_r2.json | head -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_r2_reference.json
tools/gpu_list.sh
CMD="python bench.py --no-graph --steps 1 --warmup 1 --quick"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, float, \(int\)5>' -s 21 -c 2 -f -o /tmp/prof_r2_e $CMD > gpurun_out/ncu_r2_e.log 2>&1; echo "ncu E exit $?"
B2C_CLUSTER=1 $CMD > gpurun_out/r2_cluster_plain.log 2>&1 && B2C_CLUSTER=1 ncu --set full --clock-control none --import-source on -k regex:'recur_cluster_fwd' -c 1 -f -o /tmp/prof_r2_f $CMD > gpurun_out/ncu_r2_f.log 2>&1; echo "ncu F exit $?"
for x in e f; do
  ncu -i /tmp/prof_r2_$x.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${x}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_r2_$x.ncu-rep --page details > gpurun_out/r2_ncu_${x}_details.txt 2>/dev/null
done
ls -la gpurun_out/r2_ncu_[ef]* gpurun_out/launches_final.csv | awk '{print $5, $9}'
exit 0
