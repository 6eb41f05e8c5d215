#!/bin/bash
# profiles/r2_sass_hot_kernels.txt: SASS of the hot kernels of the bench step (cuobjdump, no GPU needed):
#   K1 fp_kernel<4, white epilogue>, K2 bp_kernel<4, barrier-update epilogue, TH=32>, K2 with the plain store epilogue,
#   the standalone K3 barrier_update_kernel<2,4,fast>
lib=whitomo_b200/libwhitomo_b200.so
out=${1:-profiles/r2_sass_hot_kernels.txt}
cuobjdump -sass $lib > /tmp/wt_all_sass.txt
: > $out
for pat in 'fp_kernelILi4ENS_8HFpWhiteILi2EEELb0' 'bp_kernelILi4ENS_18BpEpiBarrierUpdateILi2EEELb0ELi32' \
           'bp_kernelILi4ENS_10BpEpiStoreELb0ELi32' 'barrier_update_kernelILi2ELi4ELb1'; do
  awk -v pat="$pat" '/Function : /{f = index($0, pat) > 0} f{print}' /tmp/wt_all_sass.txt | sed 's/\/\* 0x[0-9a-f]* \*\///g; s/ *$//' | grep -v '^ *$' >> $out
done
for m in "Function :" UBLKCP UTMALDG LDS.128 SYNCS FFMA2 "F2I\|FRND"; do echo "$m: $(grep -c "$m" $out)"; done
# profiles/r2_sass_single_image_kernels.txt: the one-image kernels (packed two-line / two-pixel arithmetic: FADD2, FFMA2)
out2=${2:-profiles/r2_sass_single_image_kernels.txt}
: > $out2
for pat in 'fp_kernelILi1ENS_8HFpStoreELb0' 'bp_kernelILi1ENS_10BpEpiStoreELb0ELi32'; do
  awk -v pat="$pat" '/Function : /{f = index($0, pat) > 0} f{print}' /tmp/wt_all_sass.txt | sed 's/\/\* 0x[0-9a-f]* \*\///g; s/ *$//' | grep -v '^ *$' >> $out2
done
for m in "Function :" UTMALDG UBLKCP "LDS " FADD2 FFMA2 "F2I\|FRND"; do echo "$m: $(grep -c "$m" $out2)"; done
