M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
run() { name=$1; shift; ncu --metrics $M --clock-control none -k regex:step_kernel -s 2 -c 1 --csv --log-file gpurun_out/ops_$name.csv python tools/prof_one.py "$@" > gpurun_out/ops_$name.log 2>&1; }
run cfg2 8 1000000 lrot+reg rk4 4
run euler_L8_lrot 8 1000000 lrot+reg euler 4
run cfg5_step 8 1000000 lrot+ddrx+reg euler 4
run cfg3 12 1000000 lrot+ddrx+reg euler 4
run cfg4 20 300000 lrot+ddrx+cdrx+reg euler 4
SFB_RNLM=1 run cfg2_rnlm 8 1000000 lrot+reg rk4 4
SFB_RNLM=1 run euler_L8_lrot_rnlm 8 1000000 lrot+reg euler 4
ncu --metrics $M --clock-control none -k regex:eij_kernel -s 2 -c 1 --csv --log-file gpurun_out/ops_eij.csv python tools/prof_eij.py > gpurun_out/ops_eij.log 2>&1
tail -n 3 gpurun_out/ops_*.csv | cut -c1-60
python tools/make_fp64_ops.py gpurun_out
