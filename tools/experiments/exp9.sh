cd $GRAFT_REPO_ROOT
for a in "264 notma" "777" "1027" "776" "48000"; do
echo "=== RS41 $a"; timeout 60 python tools/dbg2.py 0 $a 2>&1 | tail -3
done
echo "=== M10 777"; timeout 60 python tools/dbg2.py 2 777 2>&1 | tail -3
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match_reference_fm or bits_soft or mixed or ragged or pipelined or iq_path or zero" 2>&1 | tail -8
for m in 00CCCC 00DDCC 00FFCC CCCCCC CCDDCC 00EECC; do
    echo "=== type 0 mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py 0 2>&1 | tail -6
done
for t in 1 2; do
  for m in 00DDCC CCDDCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6
  done
done
