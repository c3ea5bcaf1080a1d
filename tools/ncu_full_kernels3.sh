# ncu --set full of the lv60-family kernels (LayerNorm feature extractor): GELU(LayerNorm) forward / backward over the conv
# layers' rows, conv0 + bias, and the pre-LN encoder's LayerNorm variants.  One small batch (8 x 10 s, large_lv60, 1 step).
OUT=gpurun_out/ncu_full3
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ln_fwd_kernel|ln_bwd_kernel|conv0_bias' -c 40 -o $OUT/lv60_ln \
  python tools/profile_step.py --model large_lv60 --utts 8 --seconds 10 --steps 1 > $OUT/lv60_ln.log 2>&1
ncu -i $OUT/lv60_ln.ncu-rep --page raw --csv > $OUT/lv60_ln.csv 2>/dev/null
rm -f $OUT/lv60_ln.ncu-rep
echo "lv60_ln: $(wc -l < $OUT/lv60_ln.csv) csv lines"
