mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench15_n2.json 2> gpurun_out/bench15_n2.err
echo "rc=$?"; wc -c gpurun_out/bench15_n2.json gpurun_out/bench15_n2.err; tail -c 1500 gpurun_out/bench15_n2.err; head -c 600 gpurun_out/bench15_n2.json
