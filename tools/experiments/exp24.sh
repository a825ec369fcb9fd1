cd $GRAFT_REPO_ROOT
python - <<'PY'
import numpy as np, subprocess
from sdrpp_radiosonde_b200 import synth
synth.make_fm(synth.default_spec(synth.M10, 12), 48000 * 4 + 517).astype(np.float32).tofile("/tmp/m10.raw")
PY
./sdrpp_radiosonde_b200/sonde_b200_batch -t m10 -c /tmp/b_ /tmp/m10.raw | cut -c1-80
oracle/_ref/sondedump_ref -q -t m10 -c /tmp/ref.csv /tmp/m10.raw 2>/dev/null
echo "--- batch"; cut -c1-60 /tmp/b_0.csv | head -30; wc -l /tmp/b_0.csv
echo "--- ref"; cut -c1-60 /tmp/ref.csv | head -30; wc -l /tmp/ref.csv
