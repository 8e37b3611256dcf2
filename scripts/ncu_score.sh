#!/bin/bash
# ncu capture of the scoring kernel at the bench size (one GPU; run under gpurun AFTER the plain command exited 0).
#   scripts/ncu_score.sh <tag> [f32|bf16] [variant.so]
# writes gpurun_out/<tag>.ncu-rep (--set full + the tensor-core operand-traffic and L2 counters), the raw CSV page and
# the source page CSV.
set -e
tag=$1; vol=${2:-f32}; lib=${3:-}
export AHV_VOL=$vol
[ -n "$lib" ] && export AHV_VARIANT_LIB=$lib
extra=l1tex__data_pipe_tc_wavefronts.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum,lts__t_sectors.sum,sm__sass_inst_executed_op_tmem_ldt.sum,sm__sass_inst_executed_op_tmem_stt.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max
python scripts/profile_target.py > /dev/null
ncu --set full --metrics $extra --clock-control none --import-source on -k regex:score_tc -s 3 -c 1 -f -o gpurun_out/$tag python scripts/profile_target.py > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/$tag.source.csv 2>/dev/null || true
