# ncu evidence for the kernels of the last widening of round 2 (run under gpurun)
set -u
O=gpurun_out
T=${1:-r4d}
python tools/profile_r4.py > $O/${T}_new_kernels.json 2> $O/${T}_new_kernels.err
ncu --set full --import-source on --clock-control none -k regex:"k_affine_resample_nn|k_head_direct_fwd|k_head_direct_bwd|k_stem_dx|k_transpose2d" -c 7 -o $O/${T}_new python tools/profile_r4.py once > /dev/null 2>&1
python tools/ncu_summary.py $O/${T}_new.ncu-rep > $O/${T}_ncu_new_kernels.txt 2>&1
rm -f $O/${T}_*.ncu-rep
