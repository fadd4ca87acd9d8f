set -u
echo "== default"; python tools/profile_ops.py dhtf dhts dhta pw48f 3
for v in 1 2 3; do echo "== variant $v"; HNO_TC_VARIANT=$v python tools/profile_ops.py dhtf pw48f 3; done
echo "== loader TMA"; HNO_TC_LOADER=1 python tools/profile_ops.py dhtf dhts dhta pw48f 3
echo "== prof"; HNO_TC_PROF=1 python tools/profile_ops.py dhtf dhts dhta pw48f 1 2>&1 | tail -40
