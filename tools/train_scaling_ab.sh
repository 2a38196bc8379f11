#!/bin/bash
# Same-box A/B of the data-parallel training step on an 8-GPU box (run under gpurun --gpus 8):
#   a) 8 independent single-GPU replicas at once (no exchange: what "perfect" weak scaling of THIS box looks like, straggler included)
#   b) N = 8, bucketed exchange overlapped with the backward (exchange auto)   c) the same, one exchange after the backward
#   d) N = 8, NCCL buckets (ranks pinned to one device each)
OUT=gpurun_out/r2_train_scaling_ab.txt
: > $OUT
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i python bench.py --mode train --steps 30 --warmup 5 > gpurun_out/_rep$i.json 2>/dev/null &
done
wait
python - >> $OUT <<'PY'
import json
ms = [json.load(open(f"gpurun_out/_rep{i}.json"))["ms_per_step"] for i in range(8)]
print("a) 8 independent replicas, ms/step per GPU:", [round(m, 3) for m in ms], "max", round(max(ms), 3), "median", round(sorted(ms)[4], 3))
PY
run() { # tag, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --mode train --steps 30 --warmup 5 "${@:3}" > gpurun_out/_t.json 2>/dev/null
  python - "$2" >> $OUT <<'PY'
import json, sys
d = json.load(open("gpurun_out/_t.json"))
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "|", d["config"]["collective"][:60])
PY
}
run 29801 "b) N=8 exchange auto, overlapped buckets:"
run 29802 "c) N=8 exchange auto, one exchange after the backward:" --no-overlap
UB_BENCH_PIN=1 run 29803 "d) N=8 NCCL (pinned ranks), overlapped buckets:" --exchange nccl
UB_BENCH_PIN=1 run 29804 "e) N=8 NCCL (pinned ranks), one all-reduce after the backward:" --exchange nccl --no-overlap
cat $OUT
rm -f gpurun_out/_rep*.json gpurun_out/_t.json
