cd $GRAFT_REPO_ROOT
echo "=== full gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
echo "=== bench default"; timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -1 gpurun_out/r2_bench.err
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; done
python - <<'PY'
import json
for f in ("r2_bench","r2_bench_c3","r2_bench_c4","r2_bench_c5"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["ms_per_step"],4), round(d["value"]), "kernel", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d.get("result",{}).get("auto"))
PY
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep -c "demod_pipe" gpurun_out/r2_launches.csv
