#!/bin/bash
# gpu_s3d.sh : run every build/bin/s3d_* harness variant named in $VARIANTS with the option sets in $RUNS
# ("zsplit:promo" pairs), then (optionally) ncu metric passes for the (variant:zsplit:promo) triples in $NCU.
TAG=${TAG:-s3d}
L=${L:-512}
out=gpurun_out/${TAG}_results.jsonl; : > $out
for v in ${VARIANTS:-base}; do
  for r in ${RUNS:-0:3}; do
    zs=${r%%:*}; pr=${r#*:}
    timeout 300 build/bin/s3d_$v $L ${REPS:-10} $zs $pr ${CHECK:-1} ${FLAGS:-0} ${MODE:-1} > /tmp/o.txt 2> /tmp/e.txt; rc=$?
    echo "{\"variant\": \"$v\", \"rc\": $rc, \"res\": $(cat /tmp/o.txt | tail -n 1 | grep '^{' || echo null)}" >> $out
    [ $rc -ne 0 ] && tail -n 3 /tmp/e.txt
  done
done
cat $out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum
for n in $NCU; do
  IFS=: read v zs pr <<< "$n"
  timeout 300 build/bin/s3d_$v $L 1 $zs $pr 0 > /dev/null 2>&1 && \
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_stream3d --csv --log-file gpurun_out/${TAG}_ncu_${v}_${zs}_${pr}.csv \
      build/bin/s3d_$v $L 1 $zs $pr 0 > /dev/null 2>&1; echo "ncu $n rc=$?"
done
