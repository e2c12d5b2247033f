# time the build variants under build_variants/*.so (made here with raymond_b200.build.build(defines=..., out=...)) on the GPU box
for lib in raymond_b200/libraymond_cuda.so build_variants/*.so; do
  RAYMOND_CUDA_LIB=$PWD/$lib python scripts/stage_times.py 32 f64 dragon 2>&1 | tail -1
done
