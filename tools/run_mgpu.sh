# usage: bash tools/run_mgpu.sh N [probe]  -- multi-GPU checks of the round (gather arms, host-link probe)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 300 $TR --master-port 29511 tools/xe_gather_test.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" | tail -8 | tee gpurun_out/xe_gather_n$N.txt
echo "== no PDL on the signal kernel"
CLB200_XE_PDL=0 timeout 300 $TR --master-port 29514 tools/xe_gather_test.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" | tail -8 | tee gpurun_out/xe_gather_inkernel_n$N.txt
if [ "$2" = "probe" ]; then
timeout 200 $TR --master-port 29512 tools/h2d_probe.py 2>&1 | tail -1 | tee gpurun_out/h2d_probe_n$N.json
fi
