# usage: bash tools/run_mgpu.sh N  -- fused-gather completion mechanisms, A/B
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in 0 2 1; do
echo "== CLB200_XE_GATHER_INKERNEL=$m"
CLB200_XE_GATHER_INKERNEL=$m timeout 200 $TR --master-port 2951$m tools/xe_gather_test.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm\|NCCL version\|^$" | tail -5
done
