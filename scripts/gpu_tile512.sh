#!/bin/bash
# 256 x 512 pair tiles (gemm_impl 5): unit tests, forward parity, sustained probes against the 256 x 256 default, bench.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py gemm --impl 5 > gpurun_out/t5_gemm.log 2>&1
echo "gemm impl 5 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py forward --impl 5 > gpurun_out/t5_fwd.log 2>&1
echo "forward impl 5 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 5 --configs xlmr,tinyllama,mistral > gpurun_out/t5_fwd_big.log 2>&1
echo "forward big impl 5 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096;16384,4096,8192" > gpurun_out/t5_sustained.log 2>&1
echo "sustained exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --gemm-impl 5 --no-cpu-baseline > gpurun_out/t5_bench_i5.log 2>&1
echo "bench impl 5 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --gemm-impl 2 --no-cpu-baseline > gpurun_out/t5_bench_i2.log 2>&1
echo "bench impl 2 exit $?" >> gpurun_out/summary.txt
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second"
timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv --log-file gpurun_out/t5_ncu_i5_t2.csv \
  python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 5 --terms 2 > /dev/null 2>&1
echo "ncu impl 5 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -h '"kind": "sustained"' gpurun_out/t5_sustained.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%6d %-55s ms %7.3f  MHz %6.0f  W %5.0f' % (d['m'], d.get('label',d['impl']), d['ms'], d.get('sm_mhz',0), d.get('power_w',0)))
"
for f in gpurun_out/t5_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-200; done
