python - <<'PY'
import sys, importlib, os, numpy as np
sys.path.insert(0,'.')
import bench
pkg = importlib.import_module("urlearning-cpp_b200")
wl = bench.make_bic_workload(pkg,1)
bench.fast_write_csv('/dev/shm/cfg3.csv', wl['codes'])
pkg.datagen.write_skeleton_matrix('/dev/shm/cfg3_skel.csv', wl['edges'], wl['p'])
PY
for t in 1 4 8; do
URLGPU_DEBUG_TIMING=1 ./urlearning-cpp_b200/score /dev/shm/cfg3.csv /dev/shm/out.pss -k /dev/shm/cfg3_skel.csv -f BIC -p 12 --prune -t $t --quiet > gpurun_out/r02r_score_t$t.out 2> gpurun_out/r02r_score_t$t.err
tail -1 gpurun_out/r02r_score_t$t.out
done
grep "urlgpu cube" gpurun_out/r02r_score_t1.err | awk '{pa+=$9; cpu+=$NF; r+=$(NF-9)} END {print "t1 sum plan+alloc ms", pa, "cpu total", cpu}'
grep "host waits" gpurun_out/r02r_score_t1.err
