# The reference's own Kokkos-CUDA kernels (oracle/_ref/cuda/ppkMHD_cuda, sm_90 SASS + PTX, JIT-compiled for sm_100 at first
# launch) on the same B200, same ini family as bench.py: Orszag-Tang 3-D kt=1, 256^3, 25 steps (5 taken as warm-up by
# differencing two runs).
set -e
cd gpurun_out
for V in 0 1; do
for NS in 5 25; do
python - "$V" "$NS" <<'PY'
import sys
sys.path.insert(0, "..")
import bench
ini = bench.make_ini(256, 1, int(sys.argv[2])).replace("implementationVersion=0", "implementationVersion=" + sys.argv[1])
open("refcuda.ini", "w").write(ini)
PY
T0=$(date +%s.%N); LD_PRELOAD=../oracle/_ref/cuda/libcc_shim.so LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH ../oracle/_ref/cuda/ppkMHD_cuda refcuda.ini > refcuda_v${V}_n${NS}.log 2>&1 || true; T1=$(date +%s.%N); echo "wall $(echo "$T1 $T0" | awk "{print \$1-\$2}") s" >> refcuda_v${V}_n${NS}.log
echo "== v$V nsteps=$NS"; grep -i "total\|perf\|godunov\|wall\|error\|what" refcuda_v${V}_n${NS}.log | head -12
done
done
