cd $GRAFT_REPO_ROOT
timeout 100 tools/ubench/ubench3 | grep TM
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match_reference_fm or bits_soft or mixed or ragged or pipelined or iq_path or zero" 2>&1 | tail -4
for t in 0 1 2; do
  for m in CCCCCC CCDDCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | grep -v "PW last"
  done
done
SONDE_PW_MASK=CCCCCC timeout 300 ncu --set full --clock-control none --import-source on -k regex:demod_pipe -s 2 -c 1 -o gpurun_out/k1b -f python tools/prof_k1.py 0 > gpurun_out/k1b.log 2>&1; tail -2 gpurun_out/k1b.log
