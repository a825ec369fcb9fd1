cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/r2_frame -f python tools/ncu_rs41.py > gpurun_out/r2_frame.log 2>&1; tail -1 gpurun_out/r2_frame.log
