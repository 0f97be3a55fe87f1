export NSB200_GEN_FENCE=0
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -5
for cfg in "1 0" "1 1" "0 0"; do set -- $cfg
  NSB200_P2P=$1 NSB200_GEN_FENCE=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_p2p$1_f$2.err | tee gpurun_out/bench_n2_p2p$1_f$2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('p2p=$1 fence=$2', 'value %.4g'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'strong', d['config']['strong_scaling'])"
done
