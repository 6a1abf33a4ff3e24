# A/B of launch-mechanism switches on C3 (1 GPU): programmatic dependent launch (none / release-on-exit /
# early trigger) and high-priority side branches.  Results of round 1: profiles/r1_pdl_ab.txt.
#   make -C bendy2d_b200/csrc OUT=../lib_pdl_early EXTRA="-DBENDY_PDL_EARLY"     # optional variant build
#   gpurun -- 'bash profiles/pdl_ab.sh'
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 100 python bench.py --no-scaling-ref --no-cpu-baseline --no-e2e > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - "$label" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], [round(x, 3) for x in d['ms_per_step_series'][:4]], d['roofline']['substep']['frac'])
PY
}
EARLY=$PWD/bendy2d_b200/lib_pdl_early/libbendy2d_b200.so
run pdl0 BENDY_PDL=0
run pdl1 BENDY_PDL=1
run pdl2 BENDY_PDL=2
run pdl2_prio BENDY_PDL=2 BENDY_SIDE_PRIORITY=1
[ -f "$EARLY" ] && run pdl1_early BENDY_PDL=1 BENDY2D_B200_LIB=$EARLY
[ -f "$EARLY" ] && run pdl2_early BENDY_PDL=2 BENDY2D_B200_LIB=$EARLY
