# experiment driver (gpurun): GPU tests, then bench.py with alternative builds of the library
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { # label, lib variant, extra env
  local label=$1 v=$2; shift 2
  if [ -n "$v" ]; then export TRICO_B200_LIB=$PWD/trico_b200/lib/libtrico_b200_$v.so; else unset TRICO_B200_LIB; fi
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/var_$label.json 2> gpurun_out/var_$label.err || { echo "bench $label failed"; tail -5 gpurun_out/var_$label.err; return; }
  python - <<PY
import json
d=json.load(open("gpurun_out/var_$label.json"))
print("variant [$label]", d["value"], d["encode_gbs"], d["decode_gbs"], d["ratio"], " ".join("%.4f" % v["ms"] for v in d["roofline"]["all_kernels"].values()))
PY
}
