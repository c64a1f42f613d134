#!/bin/bash
# multi-GPU runs of round 2 (gpurun --gpus N -- bash profiles/scale_r02.sh N): weak scaling, strong scaling and, at
# N = 8, BASELINE.json configs[4] as written (131 072 poses per GPU); JSON lines into gpurun_out/
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
COMMON="--gpus $N --warmup 3 --no-cpu-baseline --no-other-configs --no-pipelined"
$RUN --master-port 29511 bench.py $COMMON --steps 10 > gpurun_out/r02_scale_n${N}_weak.json 2> gpurun_out/r02_scale_n${N}_weak.err
$RUN --master-port 29512 bench.py $COMMON --steps 10 --scaling strong > gpurun_out/r02_scale_n${N}_strong.json 2> gpurun_out/r02_scale_n${N}_strong.err
if [ "$N" = "8" ]; then
$RUN --master-port 29513 bench.py $COMMON --steps 5 --batch 131072 > gpurun_out/r02_scale_n8_config5.json 2> gpurun_out/r02_scale_n8_config5.err
fi
tail -n 2 gpurun_out/r02_scale_n${N}_*.err
for f in gpurun_out/r02_scale_n${N}_*.json; do python profiles/benchline.py < $f; done
