# ncu evidence for the kernels changed late in round 2 (run under gpurun)
set -u
O=gpurun_out
T=${1:-r3v}
python tools/profile_ops.py stemb 1 > /dev/null 2>&1   # plain GPU process first
ncu --set full --import-source on --clock-control none -k regex:k_stem_wgrad_pipe -s 1 -c 1 -o $O/${T}_stemb python tools/profile_ops.py stemb 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_head_bwd_wrow|k_head_bwd_h_gather|k_head_bwd_d_gather" -s 3 -c 3 -o $O/${T}_headb python tools/profile_head.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_spectral_core" -s 2 -c 1 -o $O/${T}_coref python tools/profile_ops.py chainf 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_spectral_core" -s 2 -c 1 -o $O/${T}_coreb python tools/profile_ops.py chainb 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_ncu.csv python tools/profile_step.py 2 > $O/${T}_step.log 2>&1
for f in stemb headb coref coreb; do python tools/ncu_summary.py $O/${T}_$f.ncu-rep > $O/${T}_ncu_$f.txt 2>&1; done
rm -f $O/${T}_*.ncu-rep
