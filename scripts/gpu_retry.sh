# gpurun with retries while the pod answers busy (exit 3 / "transient").  usage: bash scripts/gpu_retry.sh <log> <timeout> [--gpus N] -- <command>
LOG=$1; TO=$2; shift 2
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
shift   # --
for i in $(seq 1 40); do
  gpurun $GP --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|no box\|busy" $LOG && ! grep -q "status=ok\|status=fail" $LOG; then sleep 90; continue; fi
  break
done
