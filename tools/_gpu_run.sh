cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/launches_r01_v3.csv $CMD > gpurun_out/ncu_l.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"binned_kernel|smem_tma|combine_kernel|generic_list" -s 12 -c 4 -o gpurun_out/prof_r01_v3_full -f $CMD > gpurun_out/ncu_f.log 2>&1
echo rc=$?
