# ncu evidence, round 2 (run under gpurun): launch list of one training step + full captures of the kernels that changed
set -u
O=gpurun_out
T=${1:-r3}
python tools/profile_ops.py pw48b 1 > /dev/null 2>&1   # a plain GPU process first: ncu wrapping the first process of a fresh box dies
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_ncu.csv python tools/profile_step.py 2 > $O/${T}_step.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_pwconv_bwd_split -s 2 -c 1 -o $O/${T}_pw48b python tools/profile_ops.py pw48b 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_analysis|k_spectral_core|k_tc_stream" -s 5 -c 5 -o $O/${T}_chainf python tools/profile_ops.py chainf 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_analysis|k_spectral_core|k_tc_stream" -s 10 -c 5 -o $O/${T}_chainb python tools/profile_ops.py chainb 1 > /dev/null 2>&1
for f in pw48b chainf chainb; do python tools/ncu_summary.py $O/${T}_$f.ncu-rep > $O/${T}_ncu_$f.txt 2>&1; done
rm -f $O/${T}_chainf.ncu-rep $O/${T}_chainb.ncu-rep
