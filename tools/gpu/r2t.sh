set -x
mkdir -p gpurun_out
K='regex:^(encode|decode|scan_|compact|zero_new|clear_word|bitcopy|index_)'
cap() { # name, skip, count, args...
  n=$1; s=$2; c=$3; shift 3
  timeout 400 ncu --set full --clock-control none -k "$K" --launch-skip $s -c $c -f -o /tmp/r2t_$n python tools/prof_target.py "$@" > gpurun_out/r2t_$n.log 2>&1
  tail -1 gpurun_out/r2t_$n.log
  python tools/ncu_summary.py /tmp/r2t_$n.ncu-rep gpurun_out/r2t_$n.md
  ncu -i /tmp/r2t_$n.ncu-rep --page raw --csv > gpurun_out/r2t_$n.csv 2>/dev/null
}
cap f64_r8 1 2 1024 f64 8 1
cap f64_acc 0 24 512 f64 a1e-6 1
cap f32_r8 1 2 1024 f32 8 1
cap f32_2d_r8 1 2 16384x16384 f32 8 1
cap f64_4d_r8 1 2 64x64x64x64 f64 8 1
cap i32_rev 0 24 512 i32 rev 1
cap f64_1d_r8 1 2 268435456x f64 8 1
du -sh gpurun_out
