cd $GRAFT_REPO_ROOT
export SONDE_FRAME_SERIAL=1
echo "=== A: two handles, RS41 511 + RS41 511"; timeout 200 python tools/mixprobe.py --split 0:511 0:511 2>&1 | grep -v Warn | tail -8
echo "=== A2: one handle RS41 146 alone"; timeout 200 python tools/mixprobe.py 0:146 2>&1 | grep -v Warn | tail -4
echo "=== E1: RS41 146 + IMET 146"; timeout 200 python tools/mixprobe.py 0:146 5:146 2>&1 | grep -v Warn | tail -6
echo "=== E2: RS41 146 + M10 146"; timeout 200 python tools/mixprobe.py 0:146 2:146 2>&1 | grep -v Warn | tail -6
echo "=== E3: RS41 511 + M10 511"; timeout 200 python tools/mixprobe.py 0:511 2:511 2>&1 | grep -v Warn | tail -6
echo "=== E4: RS41 146 + DFM 438"; timeout 200 python tools/mixprobe.py 0:146 1:438 2>&1 | grep -v Warn | tail -6
