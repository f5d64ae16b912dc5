#!/bin/bash
# Config 5 scaling curve: 128k-token causal ring attention at N GPUs of one box (strong scaling; N=1 is the same
# workload as one causal launch).   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_ring_scale.sh r01e "8 4 2 1"'
# With a third argument "all" also the head-sharded config 4 (fwd+bwd) and the FLUX forward at 8 GPUs (weak scaling).
TAG=${1:-ring}
NS=${2:-"8 4 2 1"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 240 python bench.py --gpus 1 --workload ring128k --steps 4 --warmup 2 > $OUT/${TAG}_ring128k_n1.json 2> $OUT/${TAG}_ring128k_n1.err
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
        bench.py --gpus $N --workload ring128k --steps 8 --warmup 3 > $OUT/${TAG}_ring128k_n$N.json 2> $OUT/${TAG}_ring128k_n$N.err
  fi
  tail -1 $OUT/${TAG}_ring128k_n$N.json | cut -c1-260
done
if [ "$3" = "all" ]; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
      bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_flux_fwd_n8.json 2> $OUT/${TAG}_flux_fwd_n8.err
  tail -1 $OUT/${TAG}_flux_fwd_n8.json | cut -c1-260
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 \
      bench.py --gpus 8 --workload long_window --mode fwdbwd --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_c4_fwdbwd_n8.json 2> $OUT/${TAG}_c4_fwdbwd_n8.err
  tail -1 $OUT/${TAG}_c4_fwdbwd_n8.json | cut -c1-260
fi
for f in $OUT/${TAG}_*.err; do tail -n 2 $f; done | grep -v "^\*\|OMP_NUM" | tail -10
