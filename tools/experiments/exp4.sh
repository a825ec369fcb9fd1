cd $GRAFT_REPO_ROOT
for m in CCCC DDCC; do
  echo "=== mask $m fm 1024"; SONDE_PW_MASK=$m timeout 60 python tools/dbg1.py 0 1024 fm 2>&1 | tail -6
  echo "=== mask $m fm 48000"; SONDE_PW_MASK=$m timeout 60 python tools/dbg1.py 0 48000 fm 2>&1 | tail -6
done
echo "=== M10 fm 777"; timeout 60 python tools/dbg1.py 2 777 fm 2>&1 | tail -6
echo "=== DFM fm 4096"; timeout 60 python tools/dbg1.py 1 4096 fm 2>&1 | tail -6
echo "=== parity (GFSK subset)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match_reference_fm or bits_soft or mixed or ragged or pipelined or iq_path or zero" 2>&1 | tail -15
for t in 0 1 2; do
  for m in CCCC DDCC 0CCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6
  done
done
