# per-CTA trace of one step of the sharded filter (build the library with -DBF_EXP_TRACE into exp_libs/lib_trace.so)
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
export BFILT_LIB=$PWD/exp_libs/lib_trace.so
BF_TRACE_FILE=gpurun_out/trace_g$G $TR profiles/run_shard_trace.py 100 2>&1 | grep -v OMP | tail -30
for r in $(seq 0 $((G-1))); do echo "## rank $r"; python profiles/trace_summary.py gpurun_out/trace_g${G}_r$r.bin; done
