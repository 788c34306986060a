run() { echo "$@"; env "$@" python bench.py --steps 2 --warmup 2 --batch 500000 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   inst/s %.3g  ms %.2f' % (d['instances_per_sec'], d['ms_per_step']))"; }
run DSB_SCHED_MODE=0 DSB_QUORUM=12
run DSB_SCHED_MODE=1 DSB_QUORUM=8
run DSB_SCHED_MODE=1 DSB_QUORUM=12
run DSB_SCHED_MODE=1 DSB_QUORUM=16
run DSB_SCHED_MODE=1 DSB_QUORUM=12 DSB_Q_POST=33 DSB_POST_NUM=1 DSB_POST_DEN=4
run DSB_SCHED_MODE=1 DSB_QUORUM=12 DSB_Q_POST=33 DSB_POST_NUM=1 DSB_POST_DEN=3
run DSB_SCHED_MODE=1 DSB_QUORUM=12 DSB_Q_POST=33 DSB_POST_NUM=1 DSB_POST_DEN=2
run DSB_SCHED_MODE=1 DSB_QUORUM=12 DSB_Q_POST=33 DSB_POST_NUM=2 DSB_POST_DEN=3
run DSB_SCHED_MODE=1 DSB_QUORUM=12 DSB_Q_POST=8 DSB_POST_NUM=1 DSB_POST_DEN=2
