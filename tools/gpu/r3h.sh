set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.txt 2>&1
tail -5 gpurun_out/r3h_pytest.txt
timeout 900 python bench.py > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h_bench_n1.err
tail -c 300 gpurun_out/r3h_bench_n1.json; tail -3 gpurun_out/r3h_bench_n1.err
K='regex:^(encode|decode)'
cap() { n=$1; s=$2; c=$3; shift 3
  timeout 400 ncu --set full --clock-control none -k "$K" --launch-skip $s -c $c -f -o /tmp/r3h_$n python tools/prof_target.py "$@" > gpurun_out/r3h_$n.log 2>&1
  python tools/ncu_summary.py /tmp/r3h_$n.ncu-rep gpurun_out/r3h_$n.md
  ncu -i /tmp/r3h_$n.ncu-rep --page raw --csv > gpurun_out/r3h_$n.csv 2>/dev/null
}
cap f64_r8 1 2 1024 f64 8 1
cap f32_2d_r8 1 2 16384x16384 f32 8 1
cap f64_1d_r8 1 2 268435456x f64 8 1
python __graft_entry__.py smoke > gpurun_out/r3h_smoke.txt 2>&1; tail -1 gpurun_out/r3h_smoke.txt
