# quick GPU check: smoke + gpu tests + sparse benches (AA on / off)
(time python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest.log 2>&1
tail -4 gpurun_out/smoke.log; tail -12 gpurun_out/pytest.log
run() { name=$1; shift; env "$@" python bench.py --steps 200 --warmup 20 --no-cpu-baseline $EXTRA > gpurun_out/bench_$name.json 2>> gpurun_out/bench.err; }
: > gpurun_out/bench.err
EXTRA="--sparse" run cav_sparse_aa LBM3D_AA=1
EXTRA="--sparse" run cav_sparse_ab LBM3D_AA=0
EXTRA="--sparse --workload porous --size 384" run por384_sparse_aa LBM3D_AA=1
EXTRA="--sparse --workload porous --size 384" run por384_sparse_ab LBM3D_AA=0
EXTRA="--sparse --workload porous --size 512 --steps 100" run por512_sparse_aa LBM3D_AA=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*sparse_a*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s MLUPS %.0f  ms/step %.4f  frac %.4f  e2e %.0f  nf %d"%(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value",0), d["config"]["fluid_nodes"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/bench.err
