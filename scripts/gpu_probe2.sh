#!/bin/bash
# (1) pair de-duplication of the first encoder layer: parity + bit-exactness + bench with / without;
# (2) sustained energy / throughput probes of every operand format x cluster shape, with power sampling;
# (3) ncu L2 / tensor metrics of CTA pairs vs pairs-of-pairs in the fp8-correction format.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py forward --impl 2 > gpurun_out/p2_fwd_impl2.log 2>&1
echo "forward impl 2 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py forward --impl 3 > gpurun_out/p2_fwd_impl3.log 2>&1
echo "forward impl 3 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 0 --configs xlmr,tinyllama,mistral > gpurun_out/p2_fwd_big.log 2>&1
echo "forward big exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_native.py -x -q -k "dedup or golden or row_independence or shape_variants or end_to_end" > gpurun_out/p2_pytest.log 2>&1
echo "pytest subset exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096;16384,4096,8192" > gpurun_out/p2_sustained.log 2>&1
echo "sustained exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p2_bench_dedup1.log 2>&1
echo "bench dedup on exit $?" >> gpurun_out/summary.txt
ZETT_DEDUP_PAIRS=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p2_bench_dedup0.log 2>&1
echo "bench dedup off exit $?" >> gpurun_out/summary.txt
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second,lts__cycles_elapsed.avg.per_second,lts__t_sectors_srcunit_tex_op_read.sum"
for v in "2 3" "4 3" "2 2" "4 2" "2 1" "4 1"; do
  set -- $v
  timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv \
    --log-file gpurun_out/p2_ncu_i$1_t$2.csv python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl $1 --terms $2 > /dev/null 2>&1
  echo "ncu impl $1 terms $2 exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
grep -h '"kind": "sustained"' gpurun_out/p2_sustained.log | cut -c1-260
for f in gpurun_out/p2_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-200; done
tail -n 5 gpurun_out/p2_pytest.log
