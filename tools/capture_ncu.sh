# ncu evidence (end of round 1) (run under gpurun): launch list of one training step + full captures of the hot kernels
set -u
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1e_launches_ncu.csv python tools/profile_step.py 2 > $O/r1e_step.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_pwconv_bwd_tc -s 2 -c 1 -o $O/r1e_pw48b python tools/profile_ops.py pw48b 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_stream|k_dht_tail" -s 3 -c 3 -o $O/r1e_dhtf python tools/profile_ops.py dhtf 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_stream|k_dht_tail" -s 3 -c 3 -o $O/r1e_dhts python tools/profile_ops.py dhts 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_stream" -s 3 -c 2 -o $O/r1e_dhta python tools/profile_ops.py dhta 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tc_stream" -s 1 -c 1 -o $O/r1e_pw48f python tools/profile_ops.py pw48f 1 > /dev/null 2>&1
ls -la $O/*.ncu-rep
# summaries are what gets committed (the reports with source are 15-20 MB each; gpurun copies back <= 64 MiB)
for f in pw48b dhtf dhts dhta pw48f; do python tools/ncu_summary.py $O/r1e_$f.ncu-rep > $O/r1e_ncu_$f.txt 2>&1; done
rm -f $O/r1e_dhtf.ncu-rep $O/r1e_dhts.ncu-rep $O/r1e_dhta.ncu-rep
