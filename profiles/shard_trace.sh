TR='python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541'
export BFILT_LIB=$PWD/exp_libs/lib_trace.so
BF_TRACE_FILE=gpurun_out/trace_g2 $TR profiles/run_shard_trace.py 100 2>&1 | grep -v OMP | tail -30
