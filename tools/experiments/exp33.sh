cd $GRAFT_REPO_ROOT
export SONDE_FRAME_SERIAL=1
for tp in 0 1; do
export SONDE_TPC_PAIRS=$tp
echo "=== pairs=$tp E3: RS41 511 + M10 511"; timeout 200 python tools/mixprobe.py 0:511 2:511 2>&1 | grep -v Warn | tail -3
echo "=== pairs=$tp E4: RS41 146 + DFM 438"; timeout 200 python tools/mixprobe.py 0:146 1:438 2>&1 | grep -v Warn | tail -3
echo "=== pairs=$tp homogeneous RS41 1024"; timeout 200 python tools/mixprobe.py 0:1024 2>&1 | grep -v Warn | tail -2
echo "=== pairs=$tp E5: RS41 292 + IMET 292"; timeout 200 python tools/mixprobe.py 0:292 5:292 2>&1 | grep -v Warn | tail -3
done
