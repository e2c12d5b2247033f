# One GPU-box job of round 2: the GPU test suite, then both bench arms.  usage: bash scripts/r2_job.sh <tag> [pytest-args]
TAG=${1:-r2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
free -g | head -2; nproc
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 "$@" ) > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_$TAG.log
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "e2e", "other_configs", "cpu_baseline")})
    print({k: v for k, v in d["roofline"].items() if not isinstance(v, str)})
except Exception as e:
    print("bench line unreadable:", e)
PY
