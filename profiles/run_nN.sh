# usage: run_nN.sh <N>   (torchrun bench at N GPUs, fused all-gather, default schedule)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_r2_n$N.err | tee gpurun_out/bench_r2_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', 'value %.4g'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'strong', d['config']['strong_scaling'])"
tail -3 gpurun_out/bench_r2_n$N.err
