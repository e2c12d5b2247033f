# Multi-GPU job: torchrun bench at N GPUs (+ optionally N/2), the C++ cli_old twin on N GPUs.  usage: bash scripts/multi_job.sh <tag> <N> [also]
TAG=$1; N=$2; ALSO=$3
nvidia-smi -L | head -8
for n in $N $ALSO; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 ) > gpurun_out/bench_${TAG}_${n}gpu.json 2> gpurun_out/bench_${TAG}_${n}gpu.err; echo "bench N=$n rc=$?"
  grep -c "nranks" gpurun_out/bench_${TAG}_${n}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_${n}gpu.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "e2e", "e2e_torchrun", "other_configs", "f32_shading")})
except Exception as e:
    print("bench line unreadable:", e)
PY
done
RM_TRACE=1 bash scripts/cli_old_multi.sh $N 2>&1 | tail -40 > gpurun_out/cli_old_twin_${TAG}.log; grep -E "frame|Total|ready" gpurun_out/cli_old_twin_${TAG}.log | tail -14
