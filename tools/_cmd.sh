mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01t_2gpu.out 2> gpurun_out/r01t_2gpu.err
echo "exit code $?"
wc -c gpurun_out/r01t_2gpu.out gpurun_out/r01t_2gpu.err
tail -5 gpurun_out/r01t_2gpu.err | cut -c1-300
head -c 300 gpurun_out/r01t_2gpu.out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r01t_ref2.out 2> gpurun_out/r01t_ref2.err
echo "exit code $?"; head -c 200 gpurun_out/r01t_ref2.out
