set -u
echo "== regs"; HNO_TC_KERNEL=regs python tools/profile_ops.py dhtf dhts dhta pw48f 3
echo "== regs prof"; HNO_TC_KERNEL=regs HNO_TC_PROF=1 python tools/profile_ops.py dhtf dhts pw48f 1 2>&1 | grep -v "^\[tc_regs.*" | tail -5; HNO_TC_KERNEL=regs HNO_TC_PROF=1 python tools/profile_ops.py dhtf dhts pw48f 1 2>&1 | grep "tc_regs" | awk 'NR%6==0'
echo "== stream loader 0 (LDGSTS) on everything"; HNO_TC_LOADER=0 python tools/profile_ops.py dhtf dhts dhta pw48f 3
