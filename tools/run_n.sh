#!/bin/bash
# tools/run_n.sh N PORT OUT.json [bench args...]: one torchrun bench line into OUT.json, one-line summary on stdout
N=$1; PORT=$2; OUT=$3; shift 3
if [ "$N" = "1" ]; then
  timeout 400 python bench.py --gpus 1 "$@" 2>>gpurun_out/run_n_err.log | tail -1 > "$OUT"
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" bench.py --gpus "$N" "$@" 2>>gpurun_out/run_n_err.log | tail -1 > "$OUT"
fi
python - "$OUT" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "n=%d %.1f frames/s %.3f ms/step e2e %.1f | %s | clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("parallelism"), d["clocks"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
