# A/B of library variants built side by side (DSB_LIB_TAG=<tag> [DSB_NVCC_EXTRA=...] python -m diffsol_b200.build):
#   TAGS="base d1 d2" bash tools/ab.sh
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 2 --batch ${BATCH:-1000000} --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   inst/s %.4g  ms %.2f  nli %d' % (d['instances_per_sec'], d['ms_per_step'], d['newton_iters_per_step']))"; }
for tag in $TAGS; do run DSB_LIB_TAG=$tag; done
