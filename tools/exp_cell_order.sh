# A/B of the cell kernel's work order / multicast: DRAM bytes of one K=20 beam-step launch + bench value.
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cell_fwd --profile-from-start off -s 10 -c 1 --csv"
for mc in 1 0; do for o in 0 1; do
MVB_CELL_MULTICAST=$mc MVB_CELL_ORDER=$o ncu $M --log-file gpurun_out/ord_mc${mc}_o${o}.csv python tools/profile_step.py --workload c4 --global-batch 512 > /dev/null 2>&1
echo "mc=$mc order=$o"; tail -4 gpurun_out/ord_mc${mc}_o${o}.csv | cut -d, -f13-
done; done
