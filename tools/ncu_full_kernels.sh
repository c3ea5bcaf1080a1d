# ncu --set full of a few launches of every hot kernel on the small profile workload; writes one raw CSV per kernel
# (reports are too big to ship from the GPU box).  usage: bash tools/ncu_full_kernels.sh <outdir>
OUT=${1:-gpurun_out/ncu_full}
mkdir -p $OUT
for k in gemm2_kernel gemm_bf16_tc_kernel attn_fwd2_tc_kernel attn_bwd_tc_kernel ln_fwd_kernel ln_bwd_kernel conv0_kernel conv0_bwd_accum_kernel adam_vec4_kernel posconv_tc_kernel suta_loss_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 4 -o $OUT/$k python tools/profile_step.py --utts 48 --seconds 6.5 --mode feature > $OUT/$k.log 2>&1
  ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/$k.csv 2>/dev/null
  rm -f $OUT/$k.ncu-rep
  echo "$k: $(wc -l < $OUT/$k.csv) csv lines"
done
