cd $GRAFT_REPO_ROOT
echo "=== parity (GFSK subset)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "iq_path or mixed or ragged or pipelined or kernel_equals or zero or bits_soft" 2>&1 | tail -3
run() { for m in $2; do echo "=== type $1 mask $m"; SONDE_PW_MASK=$m timeout 60 python tools/stalls.py $1 2>&1 | tail -6 | grep -v "PW last\|rounds" | cut -c1-100; done; }
run 0 "CCCFCC CCEECC CCCDCC CCEFCC CCCECC"
run 1 "CCCFCC CCEECC"
run 2 "CCEECC CCEFCC CEEECC"
echo "--- 146 channels, one per CTA"; timeout 60 python tools/stalls.py 0 146 2>&1 | tail -6 | grep -v "PW last"
