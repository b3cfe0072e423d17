cd /root/repo
run() { env "$@" timeout 100 python scripts/quick_time.py 300 $PREC 2>&1 | tail -1; }
PREC=f64
run A=1
run SJ_H_PIPE=1
run SJ_H_PIPE=1 SJ_ZCHUNK=32
PREC=f32
run SJ_H_PIPE=1
SJ_H_PIPE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
SJ_H_PIPE=1 SJ_NO_FAN=1 timeout 120 python scripts/trace_step.py 2>&1 | grep h_interior | tail -1
