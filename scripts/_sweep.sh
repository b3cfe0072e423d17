cd /root/repo
run() { env "$@" timeout 100 python scripts/quick_time.py 300 $PREC 2>&1 | tail -1; }
PREC=f64
run A=1
run SJ_PML_HALF=1
run SJ_PML_HALF=1 SJ_ZCHUNK=8
PREC=f32
run A=1
run SJ_PML_HALF=1
