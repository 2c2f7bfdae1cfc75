#!/bin/bash
# A/B of the SSR march variants on the GPU box: parity tests of the default build, per-stage times of every variant, and the
# lanes active per issued instruction (smsp__thread_inst_executed_per_inst_executed) of the march with and without lane refill.
python -m pytest tests/test_frame_parity.py -m gpu -q -x 2>&1 | tail -3
for v in "" _norefill _refill16 _refill4; do
  ALTHEA_CUDA_LIB=althea_b200/lib/libalthea_cuda$v.so python tools/stage_bench.py --iters 5 --views 2
done
for v in "" _norefill; do
  ALTHEA_CUDA_LIB=althea_b200/lib/libalthea_cuda$v.so ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"ssr_capture" -c 2 python tools/stage_bench.py --iters 1 --views 1 2>&1 | grep -E "ssr_capture|ratio|inst_executed|duration|warps_active|issue_active" > gpurun_out/ssr_refill_ncu$v.log
  cat gpurun_out/ssr_refill_ncu$v.log
done
