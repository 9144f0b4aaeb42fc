mkdir -p gpurun_out/r01j
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r01j/pytest.log 2>&1; tail -4 gpurun_out/r01j/pytest.log
for n in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r01j/bench$n.json 2> gpurun_out/r01j/bench$n.err; cat gpurun_out/r01j/bench$n.json; tail -2 gpurun_out/r01j/bench$n.err
done
