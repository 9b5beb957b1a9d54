set -x
mkdir -p gpurun_out
K='regex:^(encode|decode)'
cap() { n=$1; s=$2; c=$3; shift 3
  timeout 400 ncu --set full --clock-control none -k "$K" --launch-skip $s -c $c -f -o /tmp/r3s_$n python tools/prof_target.py "$@" > gpurun_out/r3s_$n.log 2>&1
  python tools/ncu_summary.py /tmp/r3s_$n.ncu-rep gpurun_out/r3s_$n.md
  ncu -i /tmp/r3s_$n.ncu-rep --page raw --csv > gpurun_out/r3s_$n.csv 2>/dev/null
}
cap f32_2d_r8 1 2 16384x16384 f32 8 1
cap f64_1d_r8 1 2 268435456x f64 8 1
