# time the build variants under build_variants/*.so (made here with raymond_b200.build.build(defines=..., out=...)) on the GPU box:
# GoldDragon 32-spp stage times (L2-resident regime) and the 4 M-triangle soup query (HBM regime)
for lib in raymond_b200/libraymond_cuda.so build_variants/*.so; do
  RAYMOND_CUDA_LIB=$PWD/$lib python scripts/stage_times.py 32 f64 dragon 2>&1 | tail -1
  echo "$(basename $lib) $(RAYMOND_CUDA_LIB=$PWD/$lib python scripts/c4_profile.py 4000000 cubic 2>&1 | tail -1 | cut -c150-420)"
done
