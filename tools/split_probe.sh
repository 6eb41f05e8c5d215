# development aid: A/B of the small-problem paths (cluster split, plan table, copy issue) on the C-side SIRT loop
echo "default:        $(python tools/sirt_probe.py 256 180 300 | cut -c40-200)"
echo "no plan table:  $(WT_NO_PLAN_TABLE=1 python tools/sirt_probe.py 256 180 300 | cut -c40-200)"
for bs in 1 2 4 8; do echo "bp_split=$bs $(WT_BP_SPLIT=$bs python tools/sirt_probe.py 256 180 300 | cut -c40-200)"; done
echo "512: $(python tools/sirt_probe.py 512 360 200 | cut -c40-200)"
echo "512 no table: $(WT_NO_PLAN_TABLE=1 python tools/sirt_probe.py 512 360 200 | cut -c40-200)"
