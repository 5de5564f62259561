# one GPU round: parity tests, bench, ncu launch list.  Usage: bash scripts/gpu_round.sh <tag> [bench args]
mkdir -p gpurun_out
TAG=${1:-x}; shift
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 1500 python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "peak_mem_gib", "clocks")}, d["e2e"], d["cpu_baseline"]); print("roofline", d["roofline"]); print("module", d.get("triplet_module_fwd")); print("device", d.get("device_kernels"))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share_of_step"]):
        print(f"  {k:24s} n/step={v['launches_per_step']:5.0f} ms={v['ms_per_launch']:8.3f} share={v['share_of_step']:.3f} "
              f"hbm={v.get('hbm_frac', 0):.3f} tc={v.get('tc_frac', 0):.4f}")
except Exception as e:
    print("bench json unreadable:", e)
PY
if [ -n "$NCU_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_$TAG.md; head -45 gpurun_out/launches_$TAG.md | cut -c1-160
fi
