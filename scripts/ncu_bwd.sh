#!/bin/bash
# ncu capture of the tensor-core backward kernel inside one training step (one GPU; run under gpurun AFTER the plain
# command exited 0).   scripts/ncu_bwd.sh <tag>
# writes gpurun_out/<tag>.ncu-rep (--set full + shared-memory / tensor-core counters) and its raw CSV page.
set -e
tag=$1
export AHV_SAVE=1 AHV_TC_BWD=1
extra=l1tex__data_pipe_tc_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum,sm__sass_inst_executed_op_tmem_ldt.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
python scripts/train_step_bench.py > /dev/null
ncu --set full --metrics $extra --clock-control none --import-source on -k regex:score_bwd_tc -s 2 -c 1 -f -o gpurun_out/$tag python scripts/train_step_bench.py > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv
