#!/bin/bash
# Sanitizer evidence: compute-sanitizer on the kernels, ASan/UBSan on the C++ host layer.
# Run on a GPU box from the repo root; logs land in gpurun_out/ and are summarised into profiles/ by hand.
cd ${GRAFT_REPO_ROOT:-.}
export PYTHONPATH=.
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file $OUT/r2_${tool}_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$?"; tail -3 $OUT/r2_${tool}_smoke.log
done
# racecheck on the GFSK pipeline kernel alone (one RS41 channel, 4096-sample buffers, soft tap + bit tap)
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file $OUT/r2_racecheck_k1.log python tools/dbg1.py 0 4096 fm > $OUT/r2_racecheck_k1.out 2>&1
echo "racecheck rc=$?"; tail -6 $OUT/r2_racecheck_k1.log; tail -4 $OUT/r2_racecheck_k1.out
timeout 900 compute-sanitizer --tool memcheck --log-file $OUT/r2_memcheck_parity.log python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged or mixed_batch" > $OUT/r2_memcheck_parity.out 2>&1
echo "memcheck parity rc=$?"; tail -3 $OUT/r2_memcheck_parity.log; tail -2 $OUT/r2_memcheck_parity.out
# ASan + UBSan on the host layer: the channel bank (threads, backlog, two-deep pipeline) and the compat API
python - <<'PY'
import numpy as np
from sdrpp_radiosonde_b200 import synth
n = 48000 * 2
for c, t in enumerate([synth.RS41, synth.M10, synth.DFM09]):
    synth.make_iq(synth.default_spec(t, 60 + c), n).tofile(f"/tmp/san_iq{c}")
synth.make_fm(synth.default_spec(synth.RS41, 0), n).tofile("/tmp/san_fm")
PY
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g -O1 -std=c++17 -Iinclude"
g++ $SAN tests/cpp/host_bank_test.cpp -o /tmp/host_bank_asan -Lsdrpp_radiosonde_b200 -lsonde_b200 -Wl,-rpath,$PWD/sdrpp_radiosonde_b200 -lpthread 2> $OUT/r2_asan_build.log
g++ $SAN tests/cpp/host_block_test.cpp sdrpp_radiosonde_b200/csrc/compat.cpp -o /tmp/host_block_asan -Lsdrpp_radiosonde_b200 -lsonde_b200 -Wl,-rpath,$PWD/sdrpp_radiosonde_b200 -lpthread 2>> $OUT/r2_asan_build.log
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
timeout 600 /tmp/host_bank_asan 3 96000 5000 0 /tmp/san_iq0 2 /tmp/san_iq1 1 /tmp/san_iq2 > $OUT/r2_asan_host_bank.log 2>&1; echo "asan bank rc=$?"; tail -5 $OUT/r2_asan_host_bank.log
timeout 600 /tmp/host_block_asan /tmp/san_iq0 /tmp/san_fm 96000 4096 > $OUT/r2_asan_host_block.log 2>&1; echo "asan block rc=$?"; tail -4 $OUT/r2_asan_host_block.log
