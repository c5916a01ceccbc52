# block-cyclic Cholesky N=65536 on 8 GPUs: NCCL channel count x SMs kept free of the persistent GEMM CTAs
# (profiles/r02_bc_chol_8gpu.txt: 0/0 542 ms, 2/2 601, 4/4 519, 8/8 492 ...)
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/dist_chol.py 65536 ${BC_NB:-512} 2>&1 | grep "block-cyclic" | tail -1 | sed "s/^/[$*] /"; }
for cfg in ${BC_CONFIGS:-"0:0 4:4 8:8 2:2"}; do
  ch=${cfg%%:*}; rs=${cfg##*:}
  if [ "$ch" = "0" ]; then run NAB_BC_RESERVE=$rs; else run NCCL_MAX_NCHANNELS=$ch NAB_BC_RESERVE=$rs; fi
done
