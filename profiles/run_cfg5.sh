N=${1:-8}
NSB200_TRACE=${TRACE:-0} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 profiles/config5_run.py 100000 0 2>gpurun_out/config5_n$N.err | tee gpurun_out/config5_n$N.json
grep "trace" gpurun_out/config5_n$N.err | head -40
