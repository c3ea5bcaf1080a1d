OUT=gpurun_out/ncu_full2
mkdir -p $OUT
for k in conv0_kernel conv0_bwd_accum_kernel adam_vec4_kernel posconv_tc_kernel suta_loss_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -o $OUT/$k python tools/profile_step.py --utts 48 --seconds 6.5 --mode feature > $OUT/$k.log 2>&1
  ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/$k.csv 2>/dev/null
  rm -f $OUT/$k.ncu-rep
  echo "$k: $(wc -l < $OUT/$k.csv) csv lines"
done
