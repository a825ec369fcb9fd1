cd $GRAFT_REPO_ROOT
timeout 100 tools/ubench/ubench3 | grep TM
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match_reference_fm or bits_soft or mixed or ragged or pipelined or iq_path or zero" 2>&1 | tail -4
for t in 0 1 2; do
  for m in CCCCCC CCDDCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6
  done
done
