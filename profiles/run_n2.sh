python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
bash profiles/run_nN.sh 2
NSB200_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_r2_n2_nccl.err | tee gpurun_out/bench_r2_n2_nccl.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2 NCCL host all-gather', 'value %.4g'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'])"
