# the C++ twin of cli_old on 1 and N GPUs of this box (stand-in dragon written as PLY first)
set -e
N=${1:-8}
python - <<'PY'
import sys; sys.path.insert(0, '.')
from raymond_b200 import fixtures as F
F.write_ply('/tmp/dragon_standin.ply', F.dragon_standin())
PY
g++ -O2 -std=c++17 -Iinclude examples/cli_old.cpp -Lraymond_b200 -lraymond_cuda -Wl,-rpath,$PWD/raymond_b200 -o /tmp/cli_old
for g in 1 $N; do
  /tmp/cli_old --mesh /tmp/dragon_standin.ply --out gpurun_out/cli_old_${g}gpu.png --width 1920 --height 1080 --spp 500 --gpus $g --repeat 4 2>&1 | grep -E "Total render time|cli_old" | sed "s/^/gpus=$g /"
done
