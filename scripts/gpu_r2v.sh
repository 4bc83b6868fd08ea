# GPU call V (2 GPUs): which gradient differs between the 2-rank step and the single-process batch, and under which conv path
set -x
O=gpurun_out/r2v; mkdir -p $O
for cfg in "A" "B SSR_CONV_HALO=0" "C SSR_CONV_HALO=0 SSR_CONV_LEAN=0" "D SSR_CONV_HALO=0 SSR_CONV_LEAN=0 SSR_CONV_WSTAT=0"; do
  set -- $cfg; tag=$1; shift
  env "$@" timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/ddp_diag.py > $O/diag_$tag.log 2>&1
  echo "== $tag $@"; grep "ddp-vs-single" $O/diag_$tag.log | head -6
done
