cd $GRAFT_REPO_ROOT
bash tools/sanitize.sh 2>&1 | grep -E "rc=|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" | head -20
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck_chan.log python -m pytest tests/test_channelizer.py -x -q -m gpu -k "rational or split or cutoffs" > gpurun_out/r2_memcheck_chan.out 2>&1; tail -1 gpurun_out/r2_memcheck_chan.log; tail -1 gpurun_out/r2_memcheck_chan.out
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck_mixed.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seven or afsk_frames or bch_error" > gpurun_out/r2_memcheck_mixed.out 2>&1; tail -1 gpurun_out/r2_memcheck_mixed.log; tail -1 gpurun_out/r2_memcheck_mixed.out
