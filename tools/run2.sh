mkdir -p gpurun_out
export CLB200_XE_FC=${XEFC:-16} CLB200_XE_SLICES=${XESL:-2}
REPS=3 timeout 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:k_xengine_tma -s 1 -c 1 -f -o gpurun_out/xe_tma python tools/prof_one.py xengine > gpurun_out/ncu_xe.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_xe.log
