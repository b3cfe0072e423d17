cd /root/repo
run() { env "$@" timeout 100 python scripts/quick_time.py 300 $PREC 2>&1 | tail -1; }
PREC=f64
run A=1
run SJ_N_AUX=9
run SJ_N_AUX=9 SJ_PML_PRIO=1
run SJ_PML_PRIO=1
run SJ_PML_PRIO=1 SJ_PML_LAST=1
