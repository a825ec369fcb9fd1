cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:demod_pipe -s 2 -c 1 -o gpurun_out/r2_k1 -f python tools/ncu_rs41.py > gpurun_out/r2_k1.log 2>&1; tail -1 gpurun_out/r2_k1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/r2_frame -f python tools/ncu_rs41.py > gpurun_out/r2_frame.log 2>&1; tail -1 gpurun_out/r2_frame.log
echo "--- stalls"; for t in 0 1 2; do echo "--- type $t"; timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | grep -v "PW last" | cut -c1-100; done
echo "--- 146 channels, one per CTA"; timeout 60 python tools/stalls.py 0 146 2>&1 | tail -6 | grep -v "PW last" | cut -c1-100
