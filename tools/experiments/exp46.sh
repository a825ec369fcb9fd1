cd $GRAFT_REPO_ROOT
bash tools/sanitize.sh 2>&1 | grep -v "^=========     \|Host Frame" | tail -40
echo "=== chan memcheck"
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck_chan.log python -m pytest tests/test_channelizer.py -x -q -m gpu -k "rational or split or cutoffs" > gpurun_out/r2_memcheck_chan.out 2>&1; tail -2 gpurun_out/r2_memcheck_chan.log; tail -1 gpurun_out/r2_memcheck_chan.out
echo "=== stalls"
for t in 0 1 2; do echo "--- type $t"; timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | grep -v "PW last"; done
echo "--- 146 channels, one per CTA"; timeout 60 python tools/stalls.py 0 146 2>&1 | tail -6 | grep -v "PW last"
echo "=== ncu K1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:demod_pipe -s 2 -c 1 -o gpurun_out/r2_k1 -f python tools/ncu_rs41.py > gpurun_out/r2_k1.log 2>&1; tail -1 gpurun_out/r2_k1.log
