# knob sweep of the bench workload (run on the GPU box): bash scripts/_sweep.sh [f64|f32]
cd "$(dirname "$0")/.."
PREC=${1:-f64}
run() { env "$@" timeout 100 python scripts/quick_time.py 300 $PREC 2>&1 | tail -1; }
run A=1
run SJ_NO_GRAPH=1
run SJ_N_AUX=9
run SJ_N_AUX=10 SJ_PML_PRIO=1
run SJ_ZC_FACE=10
run SJ_ZC_FACE=5
run SJ_ZC_FACE=10 SJ_ZC_GEN=10
run SJ_ZC_FACE=10 SJ_ZC_GEN=5
run SJ_ZC_FACE=16 SJ_ZC_GEN=8
run SJ_ZCHUNK=12
run SJ_ZCHUNK=24
run SJ_PML_LAST=1
run SJ_ZC_FACE=10 SJ_N_AUX=9
