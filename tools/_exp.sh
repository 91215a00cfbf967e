for v in "" _mf1 _mf2; do
  echo "== variant '$v'"
  VBMP_LIB=$PWD/pyvbmp_b200/libvbmp_b200$v.so TK_N=4194304 TK_WHAT=e timeout 300 python tools/time_kernels.py 2>&1 | tail -2
done
VBMP_LIB=$PWD/pyvbmp_b200/libvbmp_b200_mf2.so timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mf2 in-step', d['ms_per_step'], json.dumps(d['roofline']['kernels_ms_per_step']), d['clocks'])"
