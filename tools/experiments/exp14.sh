cd $GRAFT_REPO_ROOT
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frames_match_reference_fm or bits_soft or mixed or ragged or pipelined or iq_path or zero" 2>&1 | tail -3
for m in CCCCCC CCCDCC CCCECC CCCFCC CCDDCC; do
    echo "=== type 0 mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py 0 2>&1 | tail -6 | grep -v "PW last"
done
for t in 1 2; do for m in CCCDCC CCDDCC; do
    echo "=== type $t mask $m"
    SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $t 2>&1 | tail -6 | grep -v "PW last"
done; done
