#!/bin/bash
# Times every lib_variants/*.so (see build_variants.sh) with a short bench.py run; restores the stock library afterwards.
# A variant named *_nNNN is run with --streams NNN.  AB_ARGS: extra bench.py arguments (e.g. "--no-parity").
cd "$(dirname "$0")/.."
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
for v in /tmp/stock.so visual-odometry-rs_b200/lib_variants/*.so; do
  [ -e "$v" ] || continue
  [ "$v" != /tmp/stock.so ] && cp $v $LIB
  name=$(basename $v .so)
  extra=""
  if [[ "$name" =~ _n([0-9]+)$ ]]; then extra="--streams ${BASH_REMATCH[1]}"; fi
  envs=""
  if [[ "$name" =~ _generic$ ]]; then envs="VORS_NO_TILED=1"; fi  # dense keyframes through the generic (compacted) records
  env $envs python bench.py --no-cpu-baseline --steps ${AB_STEPS:-6} --warmup ${AB_WARMUP:-3} $extra ${AB_ARGS} > /tmp/ab.json 2> /tmp/ab.err; rc=$?
  [ -s /tmp/ab.json ] || { echo "$name: FAILED rc=$rc"; tail -3 /tmp/ab.err; continue; }
  python -c "
import json,sys; d=json.load(open('/tmp/ab.json')); r=d['roofline']; p=d.get('parity_in_run',{})
print('%-14s value %7.0f e2e %7.0f align_ms %.3f (per 296: %.3f) frac %.3f arms_diff %.1e failed %s parity ok=%s rad %.1e m %.1e rc=%s' % (sys.argv[1], d['value'], d['e2e']['value'], r['avg_launch_ms'], r['avg_launch_ms']*296/d['config']['streams_per_gpu'], r['frac'], d['config']['arms_max_abs_pose_diff'], d['config']['failed_alignments'], p.get('ok'), p.get('max_rad') or -1, p.get('max_m') or -1, sys.argv[2]))" $name $rc
done
cp /tmp/stock.so $LIB
