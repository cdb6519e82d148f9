#!/usr/bin/env bash
# Multi-GPU evidence on ONE box (gpurun --gpus 8): bench lines at N = 8 and 4 with the default pinned buffers and with
# transparent-huge-page pinned buffers (SPF_PIN=huge, same box A/B), plus the NVLS all-reduce parity test.
mkdir -p gpurun_out
run() {  # n tag env...
  local n=$1 tag=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n \
      bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/scale_${tag}.err | tail -1 > gpurun_out/scale_${tag}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${tag}.json").read())
    e = d["e2e"]
    print("${tag}", "value", d["value"], "ms", d["ms_per_step"], "e2e", e["value"], "copy GB/s per rank", e["h2d_copy_only_gbs_per_rank"],
          "all", e["h2d_copy_only_gbs_all_ranks"], "nodep", d["config"]["value_without_allreduce_dependency"])
except Exception as exc:
    print("${tag}", "FAILED", exc)
PY
}
run 8 n8 SPF_PIN=default
run 8 n8_huge SPF_PIN=huge
run 4 n4 SPF_PIN=default
run 4 n4_huge SPF_PIN=huge
python -m pytest tests/test_nvls_gpu.py -m gpu -q 2>&1 | tail -1
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu.txt
