#!/bin/bash
# profiles/r1_sass_hot_kernels.txt: SASS of the three hot kernels of the bench step (cuobjdump, no GPU needed)
lib=whitomo_b200/libwhitomo_b200.so
out=${1:-profiles/r1_sass_hot_kernels.txt}
: > $out
for f in _ZN2wt9fp_kernelILi4ENS_8HFpWhiteILi2EEELb0EEEvNS_8FpParamsET0_ \
         _ZN2wt9bp_kernelILi4ENS_10BpEpiStoreELb0ELi32EEEvNS_8BpParamsET0_ \
         _ZN2wt21barrier_update_kernelILi2ELi4ELb1EEEvPfPKfm16wt_update_paramsPKhPd; do
  cuobjdump -sass -fun $f $lib | sed -n '/Function :/,$p' >> $out
done
for m in UBLKCP LDS.128 SYNCS FFMA2 "F2I\|FRND"; do echo "$m: $(grep -c "$m" $out)"; done
